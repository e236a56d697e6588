"""Pin the CPU oracle (oracle/lbp_oracle.py) against outputs of the reference itself.

The fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs the
(py3-patched) reference LBP.py / train.py / c_array_utils.pyx in the build container."""
import glob
import json
import os

import numpy as np
import pytest

from macaronicusermodeling_b200 import synth
from oracle import lbp_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'graph_*.npz')))


def load_case(path):
    z = np.load(path, allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'],
             'ed': z['ed'], 'ped': z['ped']}
    spec = json.loads(str(z['spec']))
    sent = synth.sentence_to_arrays(str(z['sentence']))
    return z, model, spec, sent


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_literal_oracle_matches_reference(path):
    z, model, spec, sent = load_case(path)
    N, lr = spec.get('N', 10), spec.get('lr', 0.1)
    out = orc.run_literal(model, sent, z['theta_ee'], z['theta_ed'], list(z['roots']), spec['sweeps'],
                          reg=0.2 / N, lr=lr, approx_inference=spec.get('approx_inference', False),
                          approx_beliefs=spec.get('approx_beliefs', False))
    assert int(out['is_loopy']) == int(z['is_loopy'])
    assert list(out['var_ids']) == list(z['var_ids'])
    # graph construction: factor ids / types / variables / gaps / observed dims (train.py:255-297)
    g = orc.Graph(sent)
    desc = np.array([[f.id, 0 if f.ftype == orc.T_EN_DE else 1, f.arity, f.vars[0], f.vars[1] if f.arity > 1 else -1,
                      f.gap, -1 if f.obs is None else f.obs] for f in g.factors])
    np.testing.assert_array_equal(desc, z['factor_desc'])
    # every message the reference holds after the sweeps
    keys = [k for k in z.files if k.startswith('msg|')]
    assert len(keys) == len(out['messages'])
    for k in keys:
        _, a, b = k.split('|')
        np.testing.assert_allclose(out['messages'][a, b], z[k], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(out['marginals'], z['marginals'], rtol=1e-11)
    np.testing.assert_array_equal(out['top1'], z['top1'])
    np.testing.assert_allclose(out['logp'], float(z['logp']), rtol=1e-12)
    np.testing.assert_allclose(out['g_ee_unreg'], z['g_ee_unreg'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out['g_ed_unreg'], z['g_ed_unreg'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out['g_ee_ret'], z['g_ee_ret'], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(out['g_ed_ret'], z['g_ed_ret'], rtol=1e-9, atol=1e-13)
    np.testing.assert_array_equal(out['precision_counts'], z['precision_counts'])


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_fast_oracle_matches_reference(path):
    """closed-form gradient + hoisted potentials + level-batched GEMMs == the reference."""
    z, model, spec, sent = load_case(path)
    N, lr = spec.get('N', 10), spec.get('lr', 0.1)
    tb = orc.Tables(model, z['theta_ee'], z['theta_ed'])
    out = orc.run_fast(tb, sent, list(z['roots']), spec['sweeps'], reg=0.2 / N, lr=lr,
                       approx_inference=spec.get('approx_inference', False), approx_beliefs=spec.get('approx_beliefs', False))
    np.testing.assert_allclose(out['marginals'], z['marginals'], rtol=1e-10)
    np.testing.assert_array_equal(out['top1'], z['top1'])
    np.testing.assert_allclose(out['logp'], float(z['logp']), rtol=1e-11)
    np.testing.assert_allclose(out['g_ee_unreg'], z['g_ee_unreg'], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(out['g_ed_unreg'], z['g_ed_unreg'], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(out['g_ee_ret'], z['g_ee_ret'], rtol=1e-8, atol=1e-12)
    np.testing.assert_array_equal(out['precision_counts'], z['precision_counts'])


def test_max_vocab_order_matches_reference():
    """LBP.py:402-411: top-50 labels in descending belief order with '%0.4f' log-probabilities."""
    z, model, spec, sent = load_case(os.path.join(GOLDEN, 'graph_toy5.npz'))
    out = orc.run_literal(model, sent, z['theta_ee'], z['theta_ed'], list(z['roots']), spec['sweeps'])
    m = out['marginals'][0]
    top = np.argsort(-m, kind='stable')[:50]
    assert [synth.en_word(int(i)) for i in top] == [str(w) for w in z['maxvocab0_words']]
    np.testing.assert_allclose(np.round(np.log(m[top]), 4), z['maxvocab0_logp'], atol=1.01e-4)


@pytest.mark.parametrize('fast', [False, True])
def test_sgd_trajectory_matches_reference(fast):
    """train.py:617-638 per-sentence SGD (2 epochs, lr schedule, reg_param/N, in-place theta)."""
    z = np.load(os.path.join(GOLDEN, 'sgd_trajectory.npz'), allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'],
             'ed': z['ed'], 'ped': z['ped']}
    sents = [synth.sentence_to_arrays(str(s)) for s in z['sentences']]
    roots = z['roots'].tolist()
    traj, logps = orc.sgd_trajectory(model, sents, roots, epochs=2, fast=fast)
    np.testing.assert_allclose(traj, z['traj'], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(logps, z['logps'], rtol=1e-10)


def test_message_invariants():
    """The reference's own `if __debug__` assertions (LBP.py:639-640, 658-659) as properties."""
    z, model, spec, sent = load_case(os.path.join(GOLDEN, 'graph_k8.npz'))
    out = orc.run_literal(model, sent, z['theta_ee'], z['theta_ed'], list(z['roots']), spec['sweeps'])
    for m in out['messages'].values():
        assert (m >= 0).all()
        assert abs(m.sum() - 1.0) < 1e-10
    # bias feature gradient is identically zero up to rounding
    assert abs(out['g_ee_unreg'][0, 2]) < 1e-9 and abs(out['g_ed_unreg'][0, 5]) < 1e-9


def test_user_adapt_trajectory_matches_reference():
    """--user_adapt semantics (train.py:160-173, :224-245, :379-390, :402-409)"""
    z = np.load(os.path.join(GOLDEN, 'sgd_trajectory_user_adapt.npz'), allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'],
             'ed': z['ed'], 'ped': z['ped']}
    raw = [json.loads(str(s)) for s in z['sentences']]
    sents = [synth.sentence_to_arrays(r) for r in raw]
    users = [str(u) for u in z['users']]
    traj = orc.sgd_trajectory_user_adapt(model, sents, [r['user_id'] for r in raw], z['roots'].tolist(), users,
                                         epochs=2, ua_scale=0.5)
    np.testing.assert_allclose(traj, z['traj'], rtol=1e-8, atol=1e-12)


def test_chunked_evaluator_matches_fast():
    """run_chunked (tables built on the fly in row blocks: the checker of BASELINE config C5 at V = 50 000) gives run_fast's
    numbers, incl. given tokens (unary en_en columns), a tree, a single variable and 10 sweeps"""
    model = synth.make_model(300, 40, seed=3, dtype=np.float32)
    lay = ['ppppp', 'pgppgp', 'pp', 'p', 'pprpp']
    sents = [synth.sentence_to_arrays(synth.make_sentence(model, l, seed=5 + i, n_history=3)) for i, l in enumerate(lay)]
    roots = synth.draw_roots(sents, 10, seed=2)
    te, td = [0.8, 0.5, -0.3], [1.0, -0.6, 0.5, 0.3, 0.4, -0.2]
    m64 = {k: (np.asarray(v, dtype=np.float64) if hasattr(v, 'dtype') else v) for k, v in model.items()}
    tb = orc.Tables(m64, te, td)
    res = orc.run_chunked(model, sents, te, td, roots, 10, block=64, workers=3)
    for s, r, o in zip(sents, roots, res):
        f = orc.run_fast(tb, s, r, 10, want_grad=False)
        assert np.abs(f['marginals'] - o['marginals']).max() < 1e-15
        np.testing.assert_array_equal(f['top1'], o['top1'])
        np.testing.assert_array_equal(f['label_rank'], o['label_rank'])
        assert abs(f['logp'] - o['logp']) < 1e-12
