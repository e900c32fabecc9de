"""Turns ncu outputs into the text summaries kept under profiles/ (run here, on the files gpurun brings back).

    python tools/ncu_summaries.py launches <launches.csv> "<command profiled>"     -> per-kernel launch count, time, share
    python tools/ncu_summaries.py full <report.ncu-rep>                            -> key metrics of every captured launch
"""
import collections
import csv
import subprocess
import sys


def launches(path, command):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]
    name_i, val_i, metric_i, unit_i = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name'), H.index('Metric Unit')
    tot = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(H) or r[metric_i] != 'gpu__time_duration.sum':
            continue
        v = float(r[val_i].replace(',', ''))
        v = v / 1e3 if r[unit_i] in ('ns', 'nsecond') else (v * 1e3 if r[unit_i] in ('ms', 'msecond') else v)
        name = r[name_i].split('(')[0].replace('pnn::', '').replace('<unnamed>::', '')
        e = tot.setdefault(name, [0, 0.])
        e[0] += 1
        e[1] += v
    total = sum(v[1] for v in tot.values())
    print('ncu launch list (cold-cache, serialised) of: %s' % command)
    print('kernel, launches, total_us, share')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('%s, %d, %.1f, %.3f' % (k, v[0], v[1], v[1] / total))


WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor pipe (inst) %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active %'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__t_bytes.sum', 'L2 bytes'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'L1/TEX throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__issue_active.avg.pct', 'issue active %'),
]


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print('ncu --set full, %s' % path)
    for r in rows[2:]:
        print('---- %s' % r[hdr.index('Kernel Name')][:110])
        for metric, label in WANT:
            if metric in hdr:
                i = hdr.index(metric)
                print('  %-24s %s %s' % (label, r[i], units[i]))
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.3:
                    print('  stall %-18s %.2f warps per issue' % (h.split('issue_stalled_')[1].split('_per_')[0], v))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
    else:
        full(sys.argv[2])
