"""Host logic end to end on the CPU: engine.py + the C++ schedule compiler (csrc/plan.cpp) driving a NumPy
emulation of the device kernels (tests/fake_kernels.py), compared with the oracle and the reference fixtures.
The GPU tier (test_gpu_*.py) runs the same comparisons through the real kernels."""
import glob
import json
import os

import numpy as np
import pytest

import common_checks
from fake_kernels import FakeKernels
from macaronicusermodeling_b200 import build, synth
from macaronicusermodeling_b200.engine import Corpus, Engine
from oracle import lbp_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'graph_*.npz')))


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build()


def make_engine(model):
    return Engine(model, kernels=FakeKernels())


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_engine_matches_reference_fixture(path):
    common_checks.check_fixture(make_engine, path)


def test_batch_of_mixed_sentences_matches_oracle():
    model = synth.make_model(96, 24, seed=3)
    layouts = ['pppp', 'gpgpp', 'ppgpgp', 'pp', 'pgppg', 'gpg', 'ppppppp', 'prpgp', 'ppp', 'gppg']
    sents = [synth.sentence_to_arrays(synth.make_sentence(model, l, seed=50 + i, n_history=3)) for i, l in enumerate(layouts)]
    roots = synth.draw_roots(sents, 3, seed=9)
    common_checks.check_against_oracle(make_engine, model, sents, roots, [0.6, -0.5, 0.1], [0.8, -0.3, 0.6, 0.2, 0.5, -0.2])


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


def test_sliced_level_gemms_are_transparent():
    """Engine.gemm_slice_rows: a level's message GEMM issued as several launches over row slices (the GPU engine does it to
    re-align the CTA pairs of the tcgen05 kernel) addresses the same A / D rows as one launch"""
    model = synth.make_model(64, 16, seed=6)
    sents = synth.make_corpus(model, 30, k=6, g=1, seed=5)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(synth.draw_roots(sents, 3, seed=2))
    te, td = [0.4, 0.3, 0.0], [0.5, 0.2, 0.1, 0.1, 0.1, 0.0]
    res = []
    for pairs in (None, 1):
        eng = Engine(model, kernels=FakeKernels(), gemm_slice_pairs=pairs)
        assert eng.gemm_slice_rows == (256 if pairs else 0)
        eng.set_theta(te, td)
        r = eng.run(corpus, roots, 3, want_beliefs=True)
        res.append((r.grad.numpy(), r.logp.numpy(), r.top1.numpy(), r.beliefs.numpy(), eng.gemm_launches))
    assert res[1][4] > res[0][4]
    for x, y in zip(res[0][:4], res[1][:4]):
        np.testing.assert_array_equal(x, y)


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


def test_empty_and_degenerate_inputs():
    """empty batch -> empty result; a sentence without predicted tokens has no factor graph (LBP.py:193 asserts)"""
    model = synth.make_model(64, 16, seed=5)
    eng = make_engine(model)
    eng.set_theta([0.1, 0.2, 0.3], [0.1] * 6)
    r = eng.run(Corpus([]), np.zeros((0, 4), dtype=np.int32), 3)
    assert r.grad.shape == (0, 9) and r.logp.shape == (0,) and r.top1.shape == (0,)
    g, lp, t1, rk = eng.run_many(Corpus([]), np.zeros((0, 4), dtype=np.int32), 3)
    assert g.shape == (0, 9)
    only_given = synth.sentence_to_arrays(synth.make_sentence(model, 'ggg', seed=1))
    with pytest.raises(ValueError):
        Corpus([only_given])
