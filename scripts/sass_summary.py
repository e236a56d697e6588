#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libmlbp.so (no GPU needed):  python scripts/sass_summary.py > profiles/<tag>_sass_summary.txt
UTCHMMA(.2CTA) = tcgen05.mma (cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,
SYNCS = mbarrier ops, FMUL2 / FFMA2 = packed fp32x2 math, D* = float64 pipe, STL / LDL = local-memory spills."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'macaronicusermodeling_b200', 'libmlbp.so')
COLS = ['UTCHMMA.2CTA', 'UTCHMMA', 'LDTM', 'UTMALDG.2D.2CTA', 'UTMALDG.2D', 'UTCBAR', 'UBLKCP', 'SYNCS', 'FMUL2', 'FFMA2', 'HADD2',
        'MUFU.EX2', 'MUFU.RCP', 'DADD', 'DFMA', 'DMUL', 'ATOM', 'RED', 'STG.E.128', 'STG.E.64', 'LDG.E.128', 'STL', 'LDL']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = {}
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
        if m and cur is not None:
            op = m.group(1)
            cur['instructions'] += 1
            for c in COLS:
                if op == c or op.startswith(c + '.') or (c in ('UTCHMMA', 'UTMALDG.2D') and op.startswith(c) and '2CTA' not in op):
                    if c in ('UTCHMMA', 'UTMALDG.2D') and '2CTA' in op:
                        continue
                    cur[c] += 1
    dem = subprocess.run(['cu++filt'] + list(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines() if kernels else []
    for k, d in zip(list(kernels), dem):
        depth, cut = 0, len(d)                                  # strip the trailing parameter list (balanced parentheses)
        for i in range(len(d) - 1, -1, -1):
            depth += (d[i] == ')') - (d[i] == '(')
            if depth == 0 and d[i] == '(':
                cut = i
                break
        names[k] = d[:cut] if d.endswith(')') else d
    print('# SASS mnemonic counts per kernel of macaronicusermodeling_b200/libmlbp.so (cuobjdump -sass, sm_100a), scripts/sass_summary.py')
    print('\t'.join(['kernel', 'instructions'] + COLS))
    for k, cnt in sorted(kernels.items(), key=lambda kv: -kv[1]['instructions']):
        print('\t'.join([names.get(k, k), str(cnt['instructions'])] + [str(cnt[c]) for c in COLS]))


if __name__ == '__main__':
    main()
