"""Parity checks shared by the CPU tier (emulated kernels) and the GPU tier (real kernels through the C ABI).

Tolerances (BASELINE.json north_star): top-1 exact, beliefs <= 1e-4 max-abs in probability space, gradients
<= 1e-4 relative (with an absolute floor for the identically-zero components).  The checks here
are tighter than that: 1e-6 on beliefs, 2e-6 relative on log-posteriors -- 1e-5 where the engine runs its ONE-pass message
rows (V >= 8192: measured 1.3e-6 at C3 size; 3.2e-6 on the hostile cases of test_gpu_gates.py at V = 4608)."""
import json
import os

import numpy as np

from macaronicusermodeling_b200 import synth
from macaronicusermodeling_b200.engine import Corpus, Engine
from oracle import lbp_oracle as orc

BELIEF_ATOL = 1e-6
GRAD_RTOL, GRAD_ATOL = 1e-4, 2e-6


def load_fixture(path):
    z = np.load(path, allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'], 'ed': z['ed'],
             'ped': z['ped']}
    return z, model, json.loads(str(z['spec'])), synth.sentence_to_arrays(str(z['sentence']))


def run_engine(make_engine, model, sents, theta_ee, theta_ed, roots_pos, sweeps, beliefs=True, grad=True, **kw):
    eng = make_engine(model)
    eng.set_theta(theta_ee, theta_ed, with_grad=grad)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(roots_pos)
    return eng.run(corpus, roots, sweeps, want_grad=grad, want_marg=True, want_beliefs=beliefs, **kw), corpus


def check_fixture(make_engine, path):
    z, model, spec, sent = load_fixture(path)
    r, corpus = run_engine(make_engine, model, [sent], z['theta_ee'], z['theta_ed'], [list(z['roots'])], spec['sweeps'],
                           approx_inference=spec.get('approx_inference', False), approx_beliefs=spec.get('approx_beliefs', False))
    V = model['V']
    b = r.beliefs.cpu().numpy()[:, :V]
    assert np.abs(b - z['marginals']).max() < BELIEF_ATOL
    if os.path.basename(path) != 'graph_zeros.npz':          # theta = 0: every belief ties
        np.testing.assert_array_equal(r.top1.cpu().numpy(), z['top1'])
    np.testing.assert_allclose(r.logp.cpu().numpy()[0], float(z['logp']), rtol=2e-6)
    g = r.grad.cpu().numpy()[0]
    np.testing.assert_allclose(g[:3], z['g_ee_unreg'][0], rtol=GRAD_RTOL, atol=GRAD_ATOL)
    np.testing.assert_allclose(g[3:], z['g_ed_unreg'][0], rtol=GRAD_RTOL, atol=GRAD_ATOL)


def check_against_oracle(make_engine, model, sents, roots, te, td, sweeps=3, belief_atol=BELIEF_ATOL):
    r, corpus = run_engine(make_engine, model, sents, te, td, roots, sweeps)
    tb = orc.Tables(model, te, td)
    off = corpus.var_off
    B, T1, LP, G, RK = (x.cpu().numpy() for x in (r.beliefs, r.top1, r.logp, r.grad, r.rank))
    worst = 0.0
    for i, s in enumerate(sents):
        o = orc.run_fast(tb, s, roots[i], sweeps)
        b = B[off[i]:off[i + 1], :model['V']]
        worst = max(worst, float(np.abs(b - o['marginals']).max()))
        assert np.abs(b - o['marginals']).max() < belief_atol, i
        np.testing.assert_array_equal(T1[off[i]:off[i + 1]], o['top1'])
        lp_rtol = 1e-5 if model['V'] >= 8192 else 2e-6 * belief_atol / BELIEF_ATOL
        np.testing.assert_allclose(LP[i], o['logp'], rtol=lp_rtol)
        np.testing.assert_allclose(G[i][:3], o['g_ee_unreg'][0], rtol=GRAD_RTOL, atol=GRAD_ATOL)
        np.testing.assert_allclose(G[i][3:], o['g_ed_unreg'][0], rtol=GRAD_RTOL, atol=GRAD_ATOL)
        rk, ork = RK[off[i]:off[i + 1]], o['label_rank']      # oracle: V when outside the top-50 list
        assert ((rk == ork) | ((ork >= 50) & (rk >= 50))).all()
    return worst
