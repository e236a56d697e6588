#!/usr/bin/env python
"""Diagnostic for the exact re-score (csrc/rescore.cu) on the two C3 test sentences (V = 10 000): for the variable with the
smallest top-1 margin, compare the belief ratio of the two leading candidates as the float64 oracle, the raw GPU beliefs and a
float64 torch re-evaluation of the last hop from the engine's own A rows / table planes see it."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaronicusermodeling_b200 import synth  # noqa: E402
from macaronicusermodeling_b200 import engine as E  # noqa: E402
from oracle import lbp_oracle as orc  # noqa: E402  (the checker)


def main():
    model = synth.make_model(10000, 2000, seed=1234, dtype=np.float32)
    sents = synth.make_corpus(model, 2, k=20, g=0, seed=77)
    roots_pos = synth.draw_roots(sents, 3, seed=8)
    te, td = [0.8, 0.5, -0.3], [1.0, -0.6, 0.5, 0.3, 0.4, -0.2]
    m64 = {k: (np.asarray(v, dtype=np.float64) if hasattr(v, 'dtype') else v) for k, v in model.items()}
    tb = orc.Tables(m64, te, td)
    ref = [orc.run_fast(tb, s, r, 3, want_grad=False) for s, r in zip(sents, roots_pos)]
    M = np.concatenate([o['marginals'] for o in ref])
    want = np.concatenate([o['top1'] for o in ref])
    srt = np.argsort(M, axis=1)
    a, b = srt[:, -1], srt[:, -2]
    margin = M[np.arange(len(M)), a] / M[np.arange(len(M)), b] - 1.0
    g = int(np.argmin(margin))
    out = {'variable': g, 'cands': [int(a[g]), int(b[g])], 'oracle_ratio_minus_1': float(margin[g])}
    corpus = E.Corpus(sents)
    roots = corpus.roots_from_positions(roots_pos)
    for name, kw in (('three', dict(msg_passes=3)), ('two_raw', dict(tau=0.0, tau_label=0.0)), ('two', dict())):
        eng = E.Engine(model, **kw)
        eng.set_theta(te, td)
        r = eng.run(corpus, roots, 3, want_beliefs=True)
        torch.cuda.synchronize()
        B = r.beliefs.double().cpu().numpy()
        t1 = r.top1.cpu().numpy()
        out[name] = {'belief_ratio_minus_1': float(B[g, a[g]] / B[g, b[g]] - 1.0), 'top1': int(t1[g]), 'mismatches': int((t1 != want).sum()),
                     'stats': eng.pass_stats()}
        if name != 'two':
            continue
        # float64 re-evaluation of the last hop for variable g from the engine's buffers
        blob = eng._blob_host.numpy()
        mo, mi, mu = int(blob[E.H_MARG_OFF]), int(blob[E.H_MARG_IN]), int(blob[E.H_MARG_U])
        off = blob[mo:mo + len(M) + 1]
        rows = blob[mi + off[g]: mi + off[g + 1]]
        nb, bo = int(blob[E.H_MSG_BLK_N]), int(blob[E.H_MSG_BLK_OFF])
        blocks = blob[bo:bo + 4 * nb].reshape(-1, 4)
        V = eng.V
        cand = torch.tensor([int(a[g]), int(b[g])], device='cuda')
        score = eng._U[int(blob[mu + g]), cand].double()
        score_d = score.clone()
        per = []
        for rr in rows:
            rr = int(rr)
            if rr < 0:
                continue
            ar = rr - E.D_CONST_ROWS
            t = [int(x[0]) for x in blocks if x[1] <= ar < x[1] + x[3]]
            drow = eng._D[rr, cand].double()
            if not t or rr < E.D_CONST_ROWS:
                score *= drow; score_d *= drow
                continue
            A = eng._A[0][ar, :V].double() + eng._A[1][ar, :V].double()
            Bp = eng.plane(t[0], 0)[cand, :V].double() + eng.plane(t[0], 1)[cand, :V].double()
            dot = Bp @ A
            per.append(float((drow[0] / drow[1]) / (dot[0] / dot[1]) - 1.0))
            score *= dot
            score_d *= drow
        out['torch_exact_last_hop_ratio_minus_1'] = float(score[0] / score[1] - 1.0)
        out['torch_stored_D_ratio_minus_1'] = float(score_d[0] / score_d[1] - 1.0)
        out['per_message_ratio_error_of_stored_D'] = per
        Uex = np.exp(td[0] * m64['ed'][[a[g], b[g]], sents[0].de[sents[0].predicted[g]] if g < 20 else 0]) if g < 20 else None
    print(json.dumps(out))


if __name__ == '__main__':
    main()
