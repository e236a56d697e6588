#!/usr/bin/env python
"""Per-kernel summary of an `ncu --csv` launch list (one CSV row per launch and metric):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/launches.csv python bench.py --sentences 128 --steps 1 --warmup 1
    python scripts/ncu_summary.py gpurun_out/launches.csv > profiles/<round>_ncu_launch_summary.csv

The launches are serialised and cold-cache under ncu: compare each kernel's SHARE of the total with the live CUDA-event
shares in the bench line, not the absolute times."""
import csv
import sys
from collections import OrderedDict

UNIT = {'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    launches = OrderedDict()
    for r in rows:
        d = launches.setdefault(r['ID'], {'kernel': r['Kernel Name'], 'ms': 0.0, 'bytes': 0.0})
        v = float(r['Metric Value'].replace(',', ''))
        if r['Metric Name'] == 'gpu__time_duration.sum':
            d['ms'] = v * UNIT[r['Metric Unit']]
        elif r['Metric Name'].startswith('dram__bytes_'):
            d['bytes'] += v * BYTES[r['Metric Unit']]
    per = OrderedDict()
    for d in launches.values():
        k = per.setdefault(d['kernel'], {'n': 0, 'ms': 0.0, 'bytes': 0.0})
        k['n'] += 1; k['ms'] += d['ms']; k['bytes'] += d['bytes']
    total = sum(k['ms'] for k in per.values())
    w = csv.writer(sys.stdout)
    w.writerow(['kernel', 'launches', 'total_ms', 'share_pct', 'avg_ms', 'dram_GB', 'dram_GBps'])
    for name, k in sorted(per.items(), key=lambda kv: -kv[1]['ms']):
        w.writerow([name[:90], k['n'], '%.3f' % k['ms'], '%.1f' % (100.0 * k['ms'] / total), '%.4f' % (k['ms'] / k['n']),
                    '%.3f' % (k['bytes'] / 1e9), '%.0f' % (k['bytes'] / 1e9 / (k['ms'] / 1e3) if k['ms'] > 0 else 0.0)])


COLS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'launch__block_size', 'launch__cluster_dim_x', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_elapsed.avg.per_second',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def raw(path):
    """`ncu -i x.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_summary.py --raw raw.csv`: the columns the
    roofline discussion uses, one row per captured launch (second CSV row = units)."""
    with open(path, newline='') as f:
        rows = list(csv.reader(l for l in f if l.startswith('"')))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(c) for c in COLS if c in head]
    w = csv.writer(sys.stdout)
    w.writerow(['Kernel Name'] + ['%s [%s]' % (head[i], units[i]) for i in idx])
    k = head.index('Kernel Name')
    for r in body:
        w.writerow([r[k][:60]] + [r[i] for i in idx])


if __name__ == '__main__':
    if sys.argv[1] == '--raw':
        raw(sys.argv[2])
    else:
        main(sys.argv[1])
