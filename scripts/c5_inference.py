"""BASELINE config C5 at full size: inference-only belief sweep, V = 50 000 candidates, 10 BP sweeps, k = 20.

The float64 oracle cannot hold V = 50k tables (20 GB each), so this script checks size-independent properties
(normalised finite beliefs, run-to-run determinism, label rank consistent with the belief) plus a spot check of the
tcgen05 GEMM against the CUDA-core float64-accumulate kernel on 64 rows at the full K = N = 50 000, and reports the
inference throughput.  Run on a B200:  python scripts/c5_inference.py [--V 50000 --sentences 64]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaronicusermodeling_b200 import _lib, build, synth  # noqa: E402
from macaronicusermodeling_b200.engine import Corpus, Engine, Model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--V', type=int, default=50000)
    ap.add_argument('--Vd', type=int, default=2000)
    ap.add_argument('--k', type=int, default=20)
    ap.add_argument('--sweeps', type=int, default=10)
    ap.add_argument('--sentences', type=int, default=64)
    a = ap.parse_args()
    build.build()
    t0 = time.time()
    rng = np.random.default_rng(1234)
    V, Vd = a.V, a.Vd
    # feature planes generated in float32 row blocks to bound host memory
    def plane(rows, cols):
        out = np.empty((rows, cols), dtype=np.float32)
        for r0 in range(0, rows, 4096):
            out[r0:r0 + 4096] = rng.random((min(4096, rows - r0), cols), dtype=np.float32)
        return out
    model = {'V': V, 'Vd': Vd, 'pmi': plane(V, V), 'pmi_w1': plane(V, V), 'ed': plane(V, Vd), 'ped': plane(V, Vd)}
    t_gen = time.time() - t0
    from macaronicusermodeling_b200.engine import Kernels
    k = Kernels()
    eng = Engine(Model.from_dict(model, k.device), kernels=k, workspace_bytes=40 << 30)
    small = {'V': V, 'Vd': Vd}
    sents = synth.make_corpus(small, a.sentences, k=a.k, g=0, seed=5)
    del model
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(synth.draw_roots(sents, a.sweeps, seed=3))
    te, td = [0.8, 0.5, -0.3], [1.0, -0.6, 0.5, 0.3, 0.4, -0.2]
    eng.set_theta(te, td, with_grad=False)
    torch.cuda.synchronize()

    def run():
        return eng.run_many(corpus, roots, a.sweeps, want_grad=False, want_marg=True)

    g, lp, t1, rk = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g2, lp2, t12, rk2 = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # properties
    r = eng.run(corpus.slice(0, 2), roots[:2], a.sweeps, want_grad=False, want_marg=True, want_beliefs=True)
    b = r.beliefs[:, :V].double()
    ok_norm = bool(((b.sum(dim=1) - 1.0).abs() < 1e-5).all().item())
    ok_finite = bool(torch.isfinite(b).all().item() and (b >= 0).all().item())
    lab = torch.from_numpy(corpus.var_label[:b.shape[0]].astype(np.int64)).to(b.device)
    bl = b.gather(1, lab[:, None])[:, 0]
    rank_chk = (b > bl[:, None]).sum(dim=1).to(torch.int32)
    ok_rank = bool((rank_chk == r.rank).all().item())
    ok_top1 = bool((b.argmax(dim=1).to(torch.int32) == r.top1).all().item())
    ok_det = bool((t1 == t12).all().item() and torch.allclose(lp, lp2, rtol=0, atol=0))
    # GEMM spot check at full K = N = V
    lib = _lib.load()
    ld = eng.ld
    M = 64
    A = torch.rand((2, M, ld), device='cuda')
    Ah = (A[0] * 4).half(); Al = (A[1] * 1e-3).half()
    D0 = torch.zeros((M, ld), dtype=torch.float32, device='cuda'); D1 = torch.zeros_like(D0)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for impl, D in ((0, D0), (1, D1)):
        _lib.check(lib.mlbp_factor_to_var_gemm(P(Ah), P(Al), M, 0, M, P(eng.plane(0, 0)), P(eng.plane(0, 1)), V, ld, P(D), 0, ld,
                                               1.0, impl, st))
    torch.cuda.synchronize()
    rel = ((D0[:, :V].double() - D1[:, :V].double()).abs().max() / D1[:, :V].double().abs().max()).item()
    out = {'config': 'C5 inference only: V=%d, Vd=%d, k=%d, %d sweeps, %d sentences' % (V, Vd, a.k, a.sweeps, a.sentences),
           'sentences_per_s': a.sentences / (ms / 1e3), 'ms': ms, 'host_feature_generation_s': t_gen,
           'beliefs_normalised': ok_norm, 'beliefs_finite_nonnegative': ok_finite, 'rank_consistent': ok_rank,
           'top1_consistent': ok_top1, 'deterministic': ok_det, 'gemm_tcgen05_vs_simt_rel_err_K50k': rel,
           'gpu_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30, 'gemm_rows': eng.gemm_rows}
    print(json.dumps(out))
    assert ok_norm and ok_finite and ok_rank and ok_top1 and ok_det and rel < 5e-6


if __name__ == '__main__':
    main()
