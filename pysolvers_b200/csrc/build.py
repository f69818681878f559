"""Build libpysolv_b200.so in-tree with nvcc for sm_100a (no GPU needed).

``python -m pysolvers_b200.csrc.build [--force]`` or ``build_native()`` from
``__graft_entry__.build()``.  Each .cu is compiled to an object in parallel and
the objects are linked into ``pysolvers_b200/libpysolv_b200.so``; a source is
recompiled only when it (or a header) is newer than its object.

Flags: ``-gencode arch=compute_100a,code=sm_100a`` (B200 only), ``-lineinfo``
(ncu source page), ``-fmad=false`` (reference-identical rounding, see
common.cuh), ``-Xptxas -v`` output is kept in ``build/ptxas.log``.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, 'libpysolv_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-fmad=false', '-Xcompiler', '-fPIC',
          '-Xptxas', '-v', '-I', os.path.join(ROOT, 'include')]


def _sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith('.cu'))


def _headers_mtime():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith('.cuh')]
    hs.append(os.path.join(ROOT, 'include', 'pysolv_b200.h'))
    hs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, hdr_mtime):
    obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
    path = os.path.join(HERE, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime)):
        return obj, '', False
    cmd = [NVCC] + ARCH + CFLAGS + ['-c', path, '-o', obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, p.stdout, p.stderr))
    return obj, p.stderr, True


def _source_hash():
    import hashlib
    h = hashlib.sha256()
    files = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE))
             if f.endswith(('.cu', '.cuh'))]
    files += [os.path.join(ROOT, 'include', 'pysolv_b200.h'), os.path.abspath(__file__)]
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, 'rb').read())
    return h.hexdigest()


def build_native(force=False, verbose=False):
    """Compile (if stale) and return the path of the shared library.

    Staleness is decided by a content hash of the sources stored beside the
    library (file times do not survive the snapshot to the GPU box), so a
    prebuilt .so that matches the sources is used as is."""
    stamp = LIB + '.srchash'
    digest = _source_hash()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        if verbose:
            print('up to date', LIB)
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    hdr = _headers_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, hdr), srcs))
    objs = [r[0] for r in results]
    rebuilt = any(r[2] for r in results)
    log = os.path.join(OBJ_DIR, 'ptxas.log')
    if rebuilt:
        with open(log, 'a' if not force else 'w') as f:
            for (_, err, did), s in zip(results, srcs):
                if did:
                    f.write('==== %s\n%s\n' % (s, err))
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-ldl']
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (p.stdout, p.stderr))
    with open(stamp, 'w') as f:
        f.write(digest + '\n')
    if verbose:
        print('built' if rebuilt else 'relinked', LIB)
    return LIB


if __name__ == '__main__':
    build_native(force='--force' in sys.argv, verbose=True)
