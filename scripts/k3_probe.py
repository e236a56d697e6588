#!/usr/bin/env python
"""K3 micro-benchmark: the streaming (two-read) and the resident (single-read, cluster) leave-one-out kernels on the
three group shapes of a C3 sweep (V = 10 000, 19 incoming pairwise messages per variable), timed with CUDA events
on inputs larger than L2.  Prints algorithmic GB/s (each input and output row counted once) and checks that both
variants write the same operand rows.   python scripts/k3_probe.py [--groups 2432] [--V 10000]"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaronicusermodeling_b200 import _lib                      # noqa: E402
from macaronicusermodeling_b200.engine import _p, round_up       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--groups', type=int, default=2432)
    ap.add_argument('--V', type=int, default=10000)
    ap.add_argument('--n', type=int, default=19)
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    lib = _lib.require_device()
    dev = torch.device('cuda', 0)
    V, n, G = a.V, a.n, a.groups
    ld = round_up(V, 64)
    g = torch.Generator(device=dev).manual_seed(1)
    U = torch.rand((G, ld), device=dev, generator=g) + 0.5
    D = torch.rand((G * n + 1, ld), device=dev, generator=g) + 0.5
    D[0].fill_(1.0)                                              # the constant-one row of the ABI
    A = torch.zeros((2, 2, 2 * G * n, ld), dtype=torch.float16, device=dev)  # [variant, hi/lo, rows, ld]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    peak = 6541.8
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peak = float(json.load(open(pk)).get('hbm_gbs', peak))
    # per-stage clock64 totals: only when the library was built with MLBP_EXTRA_NVCC_FLAGS=-DMLBP_K3_STAGE_TIMES
    dbg = torch.zeros((148 * 2 * 8, 6), dtype=torch.int64, device=dev)
    lib.mlbp_debug_k3_times.argtypes = [ctypes.c_void_p]
    lib.mlbp_debug_k3_times.restype = None
    lib.mlbp_debug_k3_times(ctypes.c_void_p(dbg.data_ptr()))
    out = {}
    for shape, (ng, outs, readers) in {'down (18 outputs)': (G, list(range(1, n)), 1), 'up-b (1 output)': (G, [5 % n], 1),
                                       'root (19 outputs, 128 groups)': (min(128, G), list(range(n)), 1),
                                       'last sweep down (18 outputs, 2 readers each)': (G, list(range(1, n)), 2)}.items():
        grp_u = np.arange(ng, dtype=np.int32)
        grp_off = (np.arange(ng + 1) * n).astype(np.int32)
        in_row = 1 + np.arange(ng * n, dtype=np.int32)
        dest_off = np.zeros(ng * n + 1, dtype=np.int32)
        has = np.zeros(n, dtype=np.int32); has[outs] = readers
        dest_off[1:] = np.cumsum(np.tile(has, ng))
        dest = np.arange(int(dest_off[-1]), dtype=np.int32)
        first = np.where(np.diff(dest_off) > 0, dest[np.minimum(dest_off[:-1], len(dest) - 1)], -1).astype(np.int32)
        t = lambda x: torch.from_numpy(x).to(dev)
        second = np.where(np.diff(dest_off) > 1, dest[np.minimum(dest_off[:-1] + 1, len(dest) - 1)], -1).astype(np.int32)
        d_u, d_off, d_in, d_doff, d_dest, d_first, d_second = t(grp_u), t(grp_off), t(in_row), t(dest_off), t(dest), t(first), t(second)
        nbytes = (ng + ng * n + len(dest)) * V * 4.0
        res = {}
        for variant, impl in (('streaming', '1'), ('resident', '2')):
            os.environ['MLBP_K3_IMPL'] = impl
            Ah, Al = A[0 if impl == '1' else 1, 0], A[0 if impl == '1' else 1, 1]
            call = lambda: _lib.check(lib.mlbp_var_to_factor(ng, _p(d_u), _p(d_off), _p(d_in), _p(d_doff), _p(d_dest), _p(d_first), _p(d_second), _p(U),
                                                             _p(D), ld, V, _p(Ah), _p(Al), n, 30.0, st))
            for _ in range(2):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            if variant == 'resident' and bool((dbg.sum(1) > 0).any()):
                dd = dbg[dbg.sum(1) > 0].double()
                print('   resident stage cycles per CTA (issue, fetch, wait, phase1, sync, phase2) mean:', [int(x) for x in dd.mean(0).tolist()], 'CTAs', dd.shape[0], 'groups/cluster', ng / max(dd.shape[0] / 8, 1))
                dbg.zero_()
            res[variant] = {'ms': ms, 'GB/s': nbytes / ms / 1e6, 'frac_of_hbm_peak': nbytes / ms / 1e6 / peak}
        nd = len(dest)
        x0 = A[0, 0, :nd, :V].float() + A[0, 1, :nd, :V].float()
        x1 = A[1, 0, :nd, :V].float() + A[1, 1, :nd, :V].float()
        # rows are scale-free up to the rounding of the row sum: compare after normalising
        x0 = x0 / x0.sum(1, keepdim=True); x1 = x1 / x1.sum(1, keepdim=True)
        res['max_diff_rel_to_row_max'] = float(((x0 - x1).abs() / x0.max(1, keepdim=True)[0]).max())
        out[shape] = res
        print(shape, json.dumps(res), flush=True)
    print(json.dumps({'k3_probe': out, 'V': V, 'n_in': n, 'groups': G, 'peak_gbs': peak}))


if __name__ == '__main__':
    main()
