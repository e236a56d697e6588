#!/usr/bin/env python
"""Parity of the product path at the full C3 shape (V = 10 000, Vd = 2 000, k = 20, 3 sweeps) against the float64 oracle on a
sample of sentences: top-1 agreement, largest belief error, largest relative gradient error.  (tests/ check the same on two
sentences; this script is the larger one-off whose output is kept under profiles/.)   python scripts/c3_parity_check.py [--n 16]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaronicusermodeling_b200 import synth  # noqa: E402
from macaronicusermodeling_b200.engine import Corpus, Engine  # noqa: E402
from oracle import lbp_oracle as orc  # noqa: E402  (the checker)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=16)
    ap.add_argument('--gemm-impl', type=int, default=0, help='K4 variant (include/mlbp.h, csrc/gemm_tcgen05.cu)')
    ap.add_argument('--grad-terms', type=int, default=1, help='2 = three-pass gradient rows (for comparison)')
    ap.add_argument('--msg-passes', type=int, default=None, help='1 / 2 / 3 tensor-core passes on the message rows (default: Engine rule)')
    ap.add_argument('--theta', default='flat', choices=['flat', 'trained'],
                    help="trained = the theta bench.py's SGD reaches after ~12 steps (history weight 7.5: beliefs of 0.8 on the label, spiky messages)")
    a = ap.parse_args()
    model = synth.make_model(10000, 2000, seed=1234, dtype=np.float32)
    sents = synth.make_corpus(model, a.n, k=20, g=0, seed=4242)
    roots_pos = synth.draw_roots(sents, 3, seed=11)
    te, td = ([0.8, 0.5, -0.3], [1.0, -0.6, 0.5, 0.3, 0.4, -0.2]) if a.theta == 'flat' else \
        ([-0.003, 0.049, -0.3], [0.101, -0.047, 7.534, 0.3, 0.4, -0.2])
    eng = Engine(model, grad_a_terms=a.grad_terms, grad_b_terms=a.grad_terms, gemm_impl=a.gemm_impl, msg_passes=a.msg_passes)
    eng.set_theta(te, td)
    corpus = Corpus(sents)
    r = eng.run(corpus, corpus.roots_from_positions(roots_pos), 3, want_beliefs=True)
    B, T1, G, LP = (x.cpu().numpy() for x in (r.beliefs, r.top1, r.grad, r.logp))
    m64 = {k: (np.asarray(v, dtype=np.float64) if hasattr(v, 'dtype') else v) for k, v in model.items()}
    t0 = time.time()
    tb = orc.Tables(m64, te, td)
    off = corpus.var_off
    worst_b = worst_g = worst_lp = peak = 0.0
    flips = n_var = 0
    for i, s in enumerate(sents):
        o = orc.run_fast(tb, s, roots_pos[i], 3)
        b = B[off[i]:off[i + 1], :10000]
        worst_b = max(worst_b, float(np.abs(b - o['marginals']).max()))
        peak = max(peak, float(o['marginals'].max()))
        flips += int((T1[off[i]:off[i + 1]] != o['top1']).sum())
        n_var += off[i + 1] - off[i]
        ref = np.concatenate([o['g_ee_unreg'][0], o['g_ed_unreg'][0]])
        nz = np.abs(ref) > 1e-3
        worst_g = max(worst_g, float((np.abs(G[i] - ref)[nz] / np.abs(ref)[nz]).max()))
        worst_lp = max(worst_lp, abs(float(LP[i]) - o['logp']) / abs(o['logp']))
    print(json.dumps({'config': 'C3 shape, %d sentences, V=10000, k=20, 3 sweeps' % a.n, 'variables': int(n_var),
                      'theta': a.theta, 'largest_belief': peak, 'top1_mismatches': flips, 'max_abs_belief_error': worst_b, 'pass_stats': eng.pass_stats(), 'max_rel_gradient_error': worst_g,
                      'max_rel_logposterior_error': worst_lp, 'message_rows_passes': a.msg_passes, 'gradient_rows_passes': 3 if a.grad_terms == 2 else (1 if eng.grad_one_pass_ok else 2),
                      'oracle_seconds': time.time() - t0}))


if __name__ == '__main__':
    main()
