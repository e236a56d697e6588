#!/usr/bin/env python
"""C5 parity at full size (V = 50 000, 10 sweeps) against the chunked float64 oracle; prints one JSON line (kept under profiles/).
    python scripts/c5_parity_check.py [--V 50000 --k 8 --n 3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import c5_parity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--V', type=int, default=50000)
    ap.add_argument('--k', type=int, default=8)
    ap.add_argument('--n', type=int, default=3)
    ap.add_argument('--sweeps', type=int, default=10)
    a = ap.parse_args()
    print(json.dumps(c5_parity.c5_parity(V=a.V, layouts=tuple('p' * a.k for _ in range(a.n)), sweeps=a.sweeps)))


if __name__ == '__main__':
    main()
