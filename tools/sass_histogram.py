"""SASS opcode histogram of libpysolv_b200.so per kernel family:
    cuobjdump -sass pysolvers_b200/libpysolv_b200.so | python tools/sass_histogram.py > profiles/roundN_sass_opcodes.txt
"""
import collections
import re
import sys

FAMILIES = ('pcg_mega_kernel', 'spmv_bulk_kernel', 'spmv_stream_kernel', 'spmv_vector_kernel', 'spmv_merge',
            'trsv_cta_kernel', 'trsv_solve_kernel', 'gmres_', 'amg_', 'tri_gemv', 'blockdiag', 'dist_', 'p2p_',
            'pcg_', 'stencil', 'bratu', 'csr_')
KEYS = ['UBLKCP', 'SYNCS', 'UCGABAR', 'NANOSLEEP', 'LDG', 'STG', 'LDS', 'STS', 'DMUL', 'DADD', 'DFMA', 'MEMBAR',
        'ATOM', 'ATOMG', 'RED', 'BAR', 'SHFL']
fam = None
tot = collections.Counter()
per = collections.defaultdict(collections.Counter)
nfun = collections.Counter()
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        fam = 'other'
        for key in FAMILIES:
            if key in name:
                fam = key
                break
        nfun[fam] += 1
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and fam:
        op = m.group(1)
        if op.startswith('UCGABAR'):
            op = 'UCGABAR'                      # UCGABAR_ARV / UCGABAR_WAIT
        tot[op] += 1
        per[fam][op] += 1
print('SASS opcode histogram of pysolvers_b200/libpysolv_b200.so (cuobjdump -sass, sm_100a; -fmad=false: DFMA only in')
print('division / sqrt / reciprocal sequences).  UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier, UCGABAR = cluster barrier.')
print()
print('%-20s%6s' % ('kernel family', 'fns') + ''.join('%10s' % k for k in KEYS))
for f in sorted(per):
    print('%-20s%6d' % (f[:20], nfun[f]) + ''.join('%10d' % per[f][k] for k in KEYS))
print('%-20s%6d' % ('TOTAL', sum(nfun.values())) + ''.join('%10d' % tot[k] for k in KEYS))
print()
print('tensor-core opcodes (HMMA / IMMA / DMMA / UTC*MMA):', sum(v for k, v in tot.items() if 'MMA' in k),
      '(fp64 sparse path: none, by design)')
