#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into a text file under profiles/.

    python scripts/ncu_summary.py <launches.csv> <prof.ncu-rep> <out.txt> [title]
"""
import collections
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'usecond': v, 'ms': v * 1e3, 'msecond': v * 1e3, 'nsecond': v / 1e3}.get(r[ui], v)
        name = r[ki].split('(')[0].replace('void ', '')
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f'{"kernel":60s} {"launches":>8s} {"total_us":>12s} {"avg_us":>10s} {"share":>7s}']
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'{k[:60]:60s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}')
    return '\n'.join(out)


def full_capture(path):
    r = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        return 'no rows in ' + path
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        out.append('kernel: ' + row[hdr.index('Kernel Name')][:100])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f'    {m:75s} {row[i]:>16s} {units[i]}')
    return '\n'.join(out)


if __name__ == '__main__':
    launches, rep, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else ''
    with open(dst, 'w') as f:
        f.write(title + '\n\n== launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised) ==\n')
        f.write(launch_list(launches) + '\n\n== full-set capture of the heaviest kernels (ncu --set full --clock-control none) ==\n')
        f.write(full_capture(rep) + '\n')
    print(open(dst).read())
