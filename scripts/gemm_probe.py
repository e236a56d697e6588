"""Time the K4 message GEMM alone at a given size (CUDA events, L2 flushed by size: operands >> 126 MB)."""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaronicusermodeling_b200 import _lib, build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--M', type=int, default=16384)
    ap.add_argument('--V', type=int, default=10000)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--impl', type=int, nargs='+', default=[0])
    ap.add_argument('--b-times-uniform', action='store_true', help='B := B o U[0,1): the value distribution of a T o PMI plane')
    ap.add_argument('--b-sparse', type=float, default=0.0, help='with --table-like: fraction of B entries set to zero (a sparse T o PMI plane)')
    ap.add_argument('--split-pairs', type=int, nargs='+', default=[0], help='launch the GEMM in slices of this many 256-row M pairs (0 = one launch)')
    ap.add_argument('--table-like', action='store_true', help='B := exp(N(0, 0.5)) like a potential table, A := message-like')
    a = ap.parse_args()
    build.build()
    lib = _lib.require_device()
    V, M = a.V, a.M
    ld = (V + 63) // 64 * 64
    g = torch.Generator(device='cuda').manual_seed(1)
    Ah = (torch.rand((M, ld), device='cuda', generator=g) * 4).half()
    Al = (torch.rand((M, ld), device='cuda', generator=g) * 1e-3).half()
    Bh = (torch.rand((V, ld), device='cuda', generator=g) * 4).half()
    Bl = (torch.rand((V, ld), device='cuda', generator=g) * 1e-3).half()
    if a.table_like or a.b_times_uniform or a.b_sparse > 0:
        Bf = torch.exp(torch.randn((V, ld), device='cuda', generator=g) * 0.5) * 8.0
        if a.b_times_uniform:
            Bf = Bf * torch.rand((V, ld), device='cuda', generator=g)
        if a.b_sparse > 0:
            Bf = Bf * (torch.rand((V, ld), device='cuda', generator=g) >= a.b_sparse).float()
        Bh = Bf.half(); Bl = (Bf - Bh.float()).half()
        Af = torch.rand((M, ld), device='cuda', generator=g) * (2.0 ** 14 / V * 2)
        Ah = Af.half(); Al = (Af - Ah.float()).half()
    D = torch.empty((M, ld), dtype=torch.float32, device='cuda')
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    ref = (Ah[:64, :V].double() + Al[:64, :V].double()) @ (Bh[:, :V].double() + Bl[:, :V].double()).T
    ref -= Al[:64, :V].double() @ Bl[:, :V].double().T          # the kernel drops lo*lo by design
    for impl, split in [(i, sp) for i in a.impl for sp in a.split_pairs]:
        def run():
            step = split * 256 if split > 0 else M
            for r0 in range(0, M, step):
                n = min(step, M - r0)
                _lib.check(lib.mlbp_factor_to_var_gemm(P(Ah), P(Al), M, r0, n, P(Bh), P(Bl), V, ld, P(D), r0, ld, 1.0, impl, st))

        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        flops = 2.0 * M * V * V
        rel = ((D[:64, :V].double() - ref) / ref)
        print(json.dumps({'impl': impl, 'split_pairs': split, 'M': M, 'V': V, 'ms': round(ms, 4), 'algorithmic_tflops': round(flops / ms / 1e9, 1),
                          'executed_tflops': round(3 * flops / ms / 1e9, 1), 'rel_err_max': rel.abs().max().item(),
                          'rel_err_mean': rel.mean().item(), 'rel_err_std': rel.std().item(),
                          'timeout_code': lib.mlbp_gemm_barrier_timeout_code()}), flush=True)


if __name__ == '__main__':
    main()
