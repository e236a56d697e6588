"""Host logic end to end on the CPU: engine.py + the C++ schedule compiler (csrc/plan.cpp) driving a NumPy
emulation of the device kernels (tests/fake_kernels.py), compared with the oracle and the reference fixtures.
The GPU tier (test_gpu_*.py) runs the same comparisons through the real kernels."""
import glob
import json
import os

import numpy as np
import pytest

from fake_kernels import FakeKernels
from macaronicusermodeling_b200 import build, synth
from macaronicusermodeling_b200.engine import Corpus, Engine
from oracle import lbp_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'graph_*.npz')))


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build()


def run_engine(model, sents, theta_ee, theta_ed, roots_pos, sweeps, beliefs=True):
    eng = Engine(model, kernels=FakeKernels())
    eng.set_theta(theta_ee, theta_ed)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(roots_pos)
    return eng.run(corpus, roots, sweeps, want_grad=True, want_marg=True, want_beliefs=beliefs), corpus


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_engine_matches_reference_fixture(path):
    z = np.load(path, allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'], 'ed': z['ed'],
             'ped': z['ped']}
    spec = json.loads(str(z['spec']))
    sent = synth.sentence_to_arrays(str(z['sentence']))
    r, corpus = run_engine(model, [sent], z['theta_ee'], z['theta_ed'], [list(z['roots'])], spec['sweeps'])
    V = model['V']
    b = r.beliefs.numpy()[:, :V]
    assert np.abs(b - z['marginals']).max() < 1e-6
    if os.path.basename(path) != 'graph_zeros.npz':          # theta = 0: all beliefs tie
        np.testing.assert_array_equal(r.top1.numpy(), z['top1'])
    np.testing.assert_allclose(r.logp.numpy()[0], float(z['logp']), rtol=2e-6)
    g = r.grad.numpy()[0]
    np.testing.assert_allclose(g[:3], z['g_ee_unreg'][0], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(g[3:], z['g_ed_unreg'][0], rtol=1e-4, atol=2e-6)


def test_batch_of_mixed_sentences_matches_oracle():
    model = synth.make_model(96, 24, seed=3)
    layouts = ['pppp', 'gpgpp', 'ppgpgp', 'pp', 'pgppg', 'gpg', 'ppppppp', 'prpgp', 'ppp', 'gppg']
    sents = [synth.sentence_to_arrays(synth.make_sentence(model, l, seed=50 + i, n_history=3)) for i, l in enumerate(layouts)]
    roots = synth.draw_roots(sents, 3, seed=9)
    te, td = [0.6, -0.5, 0.1], [0.8, -0.3, 0.6, 0.2, 0.5, -0.2]
    r, corpus = run_engine(model, sents, te, td, roots, 3)
    tb = orc.Tables(model, te, td)
    off = corpus.var_off
    for i, s in enumerate(sents):
        o = orc.run_fast(tb, s, roots[i], 3)
        b = r.beliefs.numpy()[off[i]:off[i + 1], :model['V']]
        assert np.abs(b - o['marginals']).max() < 1e-6, layouts[i]
        np.testing.assert_array_equal(r.top1.numpy()[off[i]:off[i + 1]], o['top1'])
        np.testing.assert_allclose(r.logp.numpy()[i], o['logp'], rtol=2e-6)
        g = r.grad.numpy()[i]
        np.testing.assert_allclose(g[:3], o['g_ee_unreg'][0], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(g[3:], o['g_ed_unreg'][0], rtol=1e-4, atol=2e-6)
        rk, ork = r.rank.numpy()[off[i]:off[i + 1]], o['label_rank']          # oracle: V when outside the top-50 list
        assert ((rk == ork) | ((ork >= 50) & (rk >= 50))).all()


def test_microbatching_is_transparent():
    model = synth.make_model(64, 16, seed=4)
    sents = synth.make_corpus(model, 12, k=4, g=1, seed=2)
    roots_pos = synth.draw_roots(sents, 3, seed=1)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(roots_pos)
    te, td = [0.4, 0.3, 0.0], [0.5, 0.2, 0.1, 0.1, 0.1, 0.0]
    eng = Engine(model, kernels=FakeKernels())
    eng.set_theta(te, td)
    whole = eng.run(corpus, roots, 3)
    eng2 = Engine(model, kernels=FakeKernels(), workspace_bytes=1)   # forces tiny micro-batches
    eng2.set_theta(te, td)
    eng2.rows_budget = lambda: 120
    assert len(eng2.microbatches(corpus, 3, True)) > 2
    g, lp, t1, rk = eng2.run_many(corpus, roots, 3)
    np.testing.assert_allclose(g.numpy(), whole.grad.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(lp.numpy(), whole.logp.numpy(), rtol=1e-12)
    np.testing.assert_array_equal(t1.numpy(), whole.top1.numpy())


def test_inference_only_drops_dead_updates():
    """without the gradient stage the last sweep's variable->factor messages feed nothing (dead code)"""
    model = synth.make_model(64, 16, seed=5)
    sents = synth.make_corpus(model, 3, k=5, g=0, seed=3)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(synth.draw_roots(sents, 3, seed=2))
    eng = Engine(model, kernels=FakeKernels())
    eng.set_theta([0.3, 0.2, 0.1], [0.5, 0.2, 0.1, 0.1, 0.1, 0.0])
    full = eng.run(corpus, roots, 3, want_grad=True, want_marg=True, want_beliefs=True)
    inf = eng.run(corpus, roots, 3, want_grad=False, want_marg=True, want_beliefs=True)
    assert inf.stats['dead'] > full.stats['dead']
    np.testing.assert_allclose(inf.beliefs.numpy(), full.beliefs.numpy(), rtol=1e-6)
    np.testing.assert_allclose(inf.logp.numpy(), full.logp.numpy(), rtol=1e-9)
