"""BASELINE config C5 (run-predictions.sh:12: inference-only belief sweep) at the vocabulary size BASELINE names: the CUDA
path against the float64 oracle's chunked evaluator (oracle.lbp_oracle.run_chunked never holds a V x V float64 table).
Shared by tests/test_gpu_c5.py and scripts/c5_parity_check.py."""
import time

import numpy as np

from macaronicusermodeling_b200 import synth
from macaronicusermodeling_b200.engine import Corpus, Engine, Kernels, Model
from oracle import lbp_oracle as orc   # the checker


def c5_parity(V=50000, Vd=500, layouts=('ppppppp', 'pppgpppp', 'ppppgp'), sweeps=10, seed=1234, workspace_gb=16,
              theta=([0.8, 0.5, -0.3], [1.0, -0.6, 0.5, 0.3, 0.4, -0.2]), engine_kw=None):
    """Returns a dict of worst-case errors; raises nothing (the caller asserts)."""
    import torch
    t0 = time.time()
    model = synth.make_model_large(V, Vd, seed=seed)
    t_gen = time.time() - t0
    sents = [synth.sentence_to_arrays(synth.make_sentence(model, lay, seed=900 + i, n_history=3)) for i, lay in enumerate(layouts)]
    roots_pos = synth.draw_roots(sents, sweeps, seed=17)
    te, td = theta
    k = Kernels()
    eng = Engine(Model.from_dict(model, k.device), kernels=k, workspace_bytes=int(workspace_gb) << 30, **(engine_kw or {}))
    eng.set_theta(te, td, with_grad=False)
    corpus = Corpus(sents)
    r = eng.run(corpus, corpus.roots_from_positions(roots_pos), sweeps, want_grad=False, want_marg=True, want_beliefs=True)
    torch.cuda.synchronize()
    B, T1, LP, RK = (x.cpu().numpy() for x in (r.beliefs, r.top1, r.logp, r.rank))
    t_gpu = time.time() - t0 - t_gen
    t1 = time.time()
    ref = orc.run_chunked(model, sents, te, td, roots_pos, sweeps, block=1024)
    t_orc = time.time() - t1
    off = corpus.var_off
    out = {'V': V, 'sweeps': sweeps, 'sentences': len(sents), 'variables': int(off[-1]), 'top1_mismatches': 0,
           'max_abs_belief_error': 0.0, 'max_rel_belief_error_at_top1': 0.0, 'max_rel_logposterior_error': 0.0,
           'rank_mismatches': 0, 'min_rel_top1_margin': 1.0,
           'seconds': {'features': t_gen, 'gpu_incl_upload': t_gpu, 'oracle': t_orc}, 'gemm_rows': int(r.stats['gemm_rows']),
           'levels': int(r.stats['levels'])}
    for i, o in enumerate(ref):
        b = B[off[i]:off[i + 1], :V].astype(np.float64)
        m = o['marginals']
        out['max_abs_belief_error'] = max(out['max_abs_belief_error'], float(np.abs(b - m).max()))
        out['top1_mismatches'] += int((T1[off[i]:off[i + 1]] != o['top1']).sum())
        bt = m[np.arange(len(m)), o['top1']]
        out['max_rel_belief_error_at_top1'] = max(out['max_rel_belief_error_at_top1'],
                                                  float(np.abs(b[np.arange(len(m)), o['top1']] / bt - 1.0).max()))
        srt = np.sort(m, axis=1)
        out['min_rel_top1_margin'] = min(out['min_rel_top1_margin'], float(((srt[:, -1] - srt[:, -2]) / srt[:, -1]).min()))
        out['max_rel_logposterior_error'] = max(out['max_rel_logposterior_error'], abs(float(LP[i]) - o['logp']) / abs(o['logp']))
        rk, ork = RK[off[i]:off[i + 1]], o['label_rank']
        out['rank_mismatches'] += int((~((rk == ork) | ((ork >= 50) & (rk >= 50)))).sum())
    return out
