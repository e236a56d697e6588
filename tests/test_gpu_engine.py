"""GPU tier, end to end: the CUDA path through the C ABI against the reference fixtures and the oracle."""
import glob
import os

import numpy as np
import pytest

import common_checks
from macaronicusermodeling_b200 import build, synth
from macaronicusermodeling_b200.engine import Corpus, Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'graph_*.npz')))


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build()


def make_engine(model):
    return Engine(model)                     # product path: tcgen05 GEMM


def make_engine_simt(model):
    return Engine(model, gemm_impl=1)        # CUDA-core cross-check GEMM


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_reference_fixture_simt(path):
    common_checks.check_fixture(make_engine_simt, path)


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_reference_fixture(path):
    common_checks.check_fixture(make_engine, path)


def test_mixed_batch_vs_oracle():
    model = synth.make_model(96, 24, seed=3)
    layouts = ['pppp', 'gpgpp', 'ppgpgp', 'pp', 'pgppg', 'gpg', 'ppppppp', 'prpgp', 'ppp', 'gppg']
    sents = [synth.sentence_to_arrays(synth.make_sentence(model, l, seed=50 + i, n_history=3)) for i, l in enumerate(layouts)]
    roots = synth.draw_roots(sents, 3, seed=9)
    common_checks.check_against_oracle(make_engine, model, sents, roots, [0.6, -0.5, 0.1], [0.8, -0.3, 0.6, 0.2, 0.5, -0.2])


def test_c2_single_user_vs_oracle():
    """BASELINE config C2 shape: V = 1000, 20 predicted tokens (190 pairwise factors), 3 sweeps; plus the mixed
    k = 10 / g = 10 variant."""
    model = synth.make_model(1000, 200, seed=11)
    sents = synth.make_corpus(model, 2, k=20, g=0, seed=5) + synth.make_corpus(model, 2, k=10, g=10, seed=6)
    roots = synth.draw_roots(sents, 3, seed=4)
    worst = common_checks.check_against_oracle(make_engine, model, sents, roots, [0.9, 0.5, -0.2],
                                               [1.1, -0.7, 0.5, 0.3, 0.4, -0.1])
    print('worst belief abs err', worst)


def test_microbatched_equals_whole():
    model = synth.make_model(512, 64, seed=12)
    sents = synth.make_corpus(model, 24, k=6, g=2, seed=7)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(synth.draw_roots(sents, 3, seed=3))
    eng = Engine(model)
    eng.set_theta([0.4, 0.3, 0.0], [0.5, 0.2, 0.1, 0.1, 0.1, 0.0])
    whole = eng.run(corpus, roots, 3)
    eng.rows_budget = lambda: 600
    g, lp, t1, rk = eng.run_many(corpus, roots, 3)
    np.testing.assert_allclose(g.cpu().numpy(), whole.grad.cpu().numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(t1.cpu().numpy(), whole.top1.cpu().numpy())
