"""CPU tier: the oracle against BRUTE-FORCE enumeration, independent of belief propagation.  A sentence with two predicted
tokens has one pairwise factor and is a tree (FactorGraph.has_loops false, LBP.py:174-190, :219), so one sweep of the
reference's sum-product gives the exact marginals of  p(x1, x2) ~ u1(x1) u2(x2) T[x1, x2]  (SURVEY.md section 8(c)(iv)),
and the pairwise part of the gradient is  phi[l1, l2] - E_p[phi]  with the exact joint (LBP.py:544-569, :592-619)."""
import numpy as np
import pytest

from macaronicusermodeling_b200 import synth
from oracle import lbp_oracle as orc


def brute_force(model, sent, theta_ee, theta_ed):
    g = orc.Graph(sent)
    V = model['pmi'].shape[0]
    phi_ee, phi_ee_w1 = orc.dense_phi_en_en(model)
    phi_ed = orc.dense_phi_en_de(model, sent)
    te, td = np.asarray(theta_ee, dtype=np.float64), np.asarray(theta_ed, dtype=np.float64)
    unary = {v: np.ones(V) for v in g.var_ids}
    pair = None
    for f in g.factors:
        if f.arity == 1:
            phi = phi_ed if f.ftype == orc.T_EN_DE else (phi_ee if f.gap > 1 else phi_ee_w1)
            th = td if f.ftype == orc.T_EN_DE else te
            unary[f.vars[0]] = unary[f.vars[0]] * np.exp(phi[:, f.obs, :].dot(th))
        else:
            assert pair is None
            pair = f
    phi_p = phi_ee if pair.gap > 1 else phi_ee_w1
    a, b = pair.vars
    joint = unary[a][:, None] * unary[b][None, :] * np.exp(phi_p.dot(te))
    joint /= joint.sum()
    marg = {a: joint.sum(axis=1), b: joint.sum(axis=0)}
    expect = np.tensordot(joint, phi_p, axes=([0, 1], [0, 1]))
    g_pair = phi_p[g.label[a], g.label[b]] - expect
    return g, np.stack([marg[v] for v in g.var_ids]), g_pair


@pytest.mark.parametrize('layout', ['pp', 'pgp', 'gpgpg', 'prp'])
def test_tree_marginals_and_pairwise_gradient_by_enumeration(layout):
    model = synth.make_model(60, 12, seed=4)
    sent = synth.sentence_to_arrays(synth.make_sentence(model, layout, seed=11, n_history=2))
    te, td = [0.7, -0.4, 0.2], [1.1, -0.6, 0.4, 0.3, -0.5, 0.1]
    g, marg, g_pair = brute_force(model, sent, te, td)
    assert len(g.var_ids) == 2 and not g.has_loops(g.var_ids[0])
    roots = [g.var_ids[0], g.var_ids[1], g.var_ids[0], g.var_ids[1]]
    lit = orc.run_literal(model, sent, te, td, roots, sweeps=3)
    fast = orc.run_fast(orc.Tables(model, te, td), sent, roots, sweeps=3)
    for out in (lit, fast):
        np.testing.assert_allclose(out['marginals'], marg, rtol=1e-10, atol=1e-15)
    # en_en gradient = the pairwise factor's part + the unary en_en factors of given tokens (belief = table / sum, messages
    # ignored: LBP.py:540); without given tokens it is the pairwise part alone
    if 'g' not in layout and 'r' not in layout:
        np.testing.assert_allclose(lit['g_ee_unreg'][0], g_pair, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(fast['g_ee_unreg'][0], g_pair, rtol=1e-9, atol=1e-12)
