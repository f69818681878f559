"""Turn the ncu outputs of a gpurun call into the text/JSON summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/launches_r1.csv gpurun_out/prof_r1.ncu-rep profiles/round1
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__inst_executed.sum']


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        us = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1.0)
        agg.setdefault(name, []).append(us)
    tot = sum(sum(v) for v in agg.values())
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
        f.write('%-62s %6s %10s %8s\n' % ('kernel', 'n', 'mean_us', 'share'))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write('%-62s %6d %10.1f %7.1f%%\n' % (k[:62], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print(open(out).read())


def full(rep, out_txt, out_json):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    summary = []
    with open(out_txt, 'w') as f:
        f.write('# ncu --set full --clock-control none, one launch per kernel\n')
        for row in rows[2:]:
            name = row[hdr.index('Kernel Name')].split('(')[0]
            f.write('\n== %s\n' % name)
            d = {'kernel': name}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write('%-62s %16s %s\n' % (k, row[i], units[i]))
                    try:
                        d[k] = float(row[i].replace(',', ''))
                        d[k + '.unit'] = units[i]
                    except ValueError:
                        pass
            summary.append(d)
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    for d in summary:
        tot = 0.0
        for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            if k in d:
                tot += d[k] * mult.get(d.get(k + '.unit', 'byte'), 1.0)
        d['dram_traffic_bytes'] = tot
    json.dump(summary, open(out_json, 'w'), indent=1)
    print(open(out_txt).read())


if __name__ == '__main__':
    csv_path, rep, prefix = sys.argv[1:4]
    launches(csv_path, prefix + '_launches.txt')
    full(rep, prefix + '_ncu_full.txt', prefix + '_ncu_full.json')
