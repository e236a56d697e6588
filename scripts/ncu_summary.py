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


if __name__ == '__main__':
    main(sys.argv[1])
