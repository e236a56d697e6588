"""Checks of the drop-in LBP / train_compat modules against the reference fixtures, shared by the CPU tier
(emulated kernels) and the GPU tier."""
import io
import json

import numpy as np

from macaronicusermodeling_b200 import train_compat as tc


def graph_from_fixture(z):
    V, Vd = z['pmi'].shape[0], z['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    en2id = dict((e, i) for i, e in enumerate(en_domain))
    de2id = dict((d, i) for i, d in enumerate(de_domain))
    pw = tc.make_phi_wrapper(z['pmi'], z['pmi_w1'], z['ed'], z['ped'])
    spec = json.loads(str(z['spec']))
    t_ee = np.array(z['theta_ee'], dtype=np.float64).reshape(1, 3)
    t_ed = np.array(z['theta_ed'], dtype=np.float64).reshape(1, 6)
    opts = tc.default_options(session_history=True, use_approx_inference=spec.get('approx_inference', False),
                              use_approx_beliefs=spec.get('approx_beliefs', False))
    fg = tc.create_factor_graph(str(z['sentence']), spec.get('lr', 0.1), tc.F_EN_EN_NAMES, tc.F_EN_DE_NAMES, t_ee, t_ed, pw,
                                en_domain, de2id, en2id, {}, options=opts, N=spec.get('N', 10), de_domain=de_domain)
    return fg, spec


def check_dynamic_features_are_fixed_at_build_time(path):
    """train.py:176-215 writes a sentence's history features into the SHARED phi_en_de planes and train.py:239-250 turns them
    into potentials right away; a second graph built from the same PhiWrapper before the first one is evaluated must not
    change the first one's results (the lazy shim snapshots the features in PotentialTable.slice_potentials)"""
    z = np.load(path, allow_pickle=False)
    V, Vd = z['pmi'].shape[0], z['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    en2id = dict((e, i) for i, e in enumerate(en_domain))
    de2id = dict((d, i) for i, d in enumerate(de_domain))
    pw = tc.make_phi_wrapper(z['pmi'], z['pmi_w1'], z['ed'], z['ped'])
    spec = json.loads(str(z['spec']))
    t_ee = np.array(z['theta_ee'], dtype=np.float64).reshape(1, 3)
    t_ed = np.array(z['theta_ed'], dtype=np.float64).reshape(1, 6)
    opts = tc.default_options(session_history=True)
    mk = lambda sent: tc.create_factor_graph(sent, spec.get('lr', 0.1), tc.F_EN_EN_NAMES, tc.F_EN_DE_NAMES, t_ee, t_ed, pw,
                                             en_domain, de2id, en2id, {}, options=opts, N=spec.get('N', 10), de_domain=de_domain)
    first = mk(str(z['sentence']))
    other = json.loads(str(z['sentence']))
    assert len(other['past_correct_guesses']) + len(other['past_guesses_for_current_sent']) > 0
    other['past_correct_guesses'], other['past_guesses_for_current_sent'] = [], []
    second = mk(json.dumps(other))                               # rewrites the dynamic planes of the shared wrapper
    roots = [int(r) for r in z['roots']]
    marg = []
    for fg in (first, second):
        fg.initialize(roots[0])
        fg.treelike_inference(spec['sweeps'], roots[1:])
        marg.append(np.stack([fg.variables[v].get_marginal().m[:, 0] for v in sorted(fg.variables.keys())]))
    assert np.abs(marg[0] - z['marginals']).max() < 1e-6        # the first graph still sees ITS history features
    assert np.abs(marg[1] - z['marginals']).max() > 1e-4        # ... and the second one its own (none)


def check_lbp_api_fixture(path, check_messages=True):
    z = np.load(path, allow_pickle=False)
    fg, spec = graph_from_fixture(z)
    roots = [int(r) for r in z['roots']]
    fg.initialize(roots[0])
    assert bool(fg.isLoopy) == bool(int(z['is_loopy']))
    fg.treelike_inference(spec['sweeps'], roots[1:])
    # graph wiring (train.py:255-297)
    desc = np.array([[f.id, 0 if f.factor_type == 'en_de' else 1, len(f.varset), f.varset[0].id,
                      f.varset[1].id if len(f.varset) > 1 else -1, f.gap,
                      -1 if f.potential_table.observed_dim is None else f.potential_table.observed_dim] for f in fg.factors])
    np.testing.assert_array_equal(desc, z['factor_desc'])
    assert sorted(fg.variables.keys()) == list(z['var_ids'])
    marg = np.stack([fg.variables[v].get_marginal().m[:, 0] for v in sorted(fg.variables.keys())])
    assert np.abs(marg - z['marginals']).max() < 1e-6
    np.testing.assert_allclose(fg.get_posterior_probs(), float(z['logp']), rtol=2e-6)
    g_ee, g_ed = fg.get_unregularized_gradeint()
    assert g_ee.shape == (1, 3) and g_ed.shape == (1, 6)
    np.testing.assert_allclose(g_ee, z['g_ee_unreg'], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(g_ed, z['g_ed_unreg'], rtol=1e-4, atol=2e-6)
    r_ee, r_ed = fg.return_gradient()
    np.testing.assert_allclose(r_ee, z['g_ee_ret'], rtol=1e-4, atol=2e-7)
    np.testing.assert_allclose(r_ed, z['g_ed_ret'], rtol=1e-4, atol=2e-7)
    if 'zeros' not in path:
        np.testing.assert_array_equal(np.array(fg.get_precision_counts()), z['precision_counts'])
        got, want = fg.to_string(), [str(s) for s in z['to_string']]
        assert len(got) == len(want)
        for a, b in zip(got, want):            # same words in the same order; '%0.4f' log-probs may differ in the last digit
            ta, tb = a.split(' '), b.split(' ')
            assert len(ta) == len(tb)
            for x, y in zip(ta, tb):
                try:
                    assert abs(float(x) - float(y)) <= 1.5e-4, (x, y)
                except ValueError:
                    assert x == y, (x, y)
    if check_messages:
        keys = [k for k in z.files if k.startswith('msg|')]
        assert len(keys) == len(fg.messages)
        for k in keys:
            _, a, b = k.split('|')
            m = fg.messages[a, b].m
            assert m.shape == (len(z[k]), 1)
            assert np.abs(m[:, 0] - z[k]).max() < 1e-6, k


def check_params_roundtrip(tmp_path):
    ee = np.array([[0.123456789, -1.5, 0.0]])
    ed = np.array([[1.0, 2.0, -3.25, 0.5, 0.25, 0.125]])
    p = str(tmp_path / 'params')
    tc.save_params(io.open(p, 'w', encoding='utf8'), ee, ed, list(tc.F_EN_EN_NAMES), list(tc.F_EN_DE_NAMES), {})
    een, eet, edn, edt, d2t = tc.read_params(p)
    assert een == tc.F_EN_EN_NAMES and edn == tc.F_EN_DE_NAMES and d2t == {}
    np.testing.assert_allclose(eet, ee, atol=5e-7)
    np.testing.assert_allclose(edt, ed, atol=5e-7)


def _adapt_fixture():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'sgd_trajectory_user_adapt.npz'), allow_pickle=False)
    V, Vd = z['pmi'].shape[0], z['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    return z, en_domain, de_domain, dict((e, i) for i, e in enumerate(en_domain)), dict((d, i) for i, d in enumerate(de_domain))


def check_user_adapt_drop_in():
    """train.py's loop with --user_adapt written against the drop-in API (batch_sgd + batch_sgd_accumulate)"""
    z, en_domain, de_domain, en2id, de2id = _adapt_fixture()
    pw = tc.make_phi_wrapper(z['pmi'], z['pmi_w1'], z['ed'], z['ped'])
    users = [str(u) for u in z['users']]
    te, td = np.zeros((1, 3)), np.zeros((1, 6))
    d2t = {}
    for u in users:
        d2t['en_en', u], d2t['en_de', u] = np.zeros((1, 3)), np.zeros((1, 6))
    opts = tc.default_options(session_history=True, user_adapt=True, reg_param_ua_scale='0.5')
    sents = [str(s) for s in z['sentences']]
    roots = z['roots'].tolist()
    traj = []
    for epoch in range(2):
        lr = 0.1 / float(1.0 + epoch * 0.3)
        for si, s in enumerate(sents):
            res = tc.batch_sgd(s, tc.F_EN_EN_NAMES, tc.F_EN_DE_NAMES, te, td, pw, lr, en_domain, de2id, en2id, d2t,
                               options=opts, N=len(sents), de_domain=de_domain, roots=roots[epoch][si])
            tc.batch_sgd_accumulate(res, te, td, d2t)
            row = [te[0], td[0]]
            for u in users:
                row += [d2t['en_en', u][0], d2t['en_de', u][0]]
            traj.append(np.concatenate(row))
    np.testing.assert_allclose(np.array(traj), z['traj'], rtol=1e-4, atol=3e-7)


def check_adapt_trainer(make_engine):
    """trainer.AdaptTrainer with one-sentence domain batches == the reference's --user_adapt trajectory"""
    from macaronicusermodeling_b200 import synth
    from macaronicusermodeling_b200.engine import Corpus
    from macaronicusermodeling_b200.trainer import AdaptTrainer
    z, *_ = _adapt_fixture()
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'], 'ed': z['ed'], 'ped': z['ped']}
    raw = [json.loads(str(s)) for s in z['sentences']]
    sents = [synth.sentence_to_arrays(r) for r in raw]
    users = [str(u) for u in z['users']]
    roots = z['roots'].tolist()
    tr = AdaptTrainer(make_engine(model), users, reg_param=0.2, ua_scale=0.5, N=len(sents))
    traj = []
    for epoch in range(2):
        lr = tr.lr(epoch)
        for si, s in enumerate(sents):
            c = Corpus([s])
            red = tr.step_domains([(raw[si]['user_id'], c, c.roots_from_positions([roots[epoch][si]]))], lr)
            tr.apply(red, lr)
            row = [tr.theta_ee, tr.theta_ed]
            for u in users:
                row += list(tr.domain2theta[u])
            traj.append(np.concatenate(row))
    np.testing.assert_allclose(np.array(traj), z['traj'], rtol=1e-4, atol=3e-7)


# ----------------------------------------------------------------------------- eager (one message at a time) mode
def explicit_graph_from_fixture(z):
    """the explicit PotentialTable(table=...) graph of tests/golden/make_golden.py::explicit_table_case, through the drop-in"""
    from macaronicusermodeling_b200 import LBP
    V = z['unary'].shape[1]
    dom = ['w%d' % i for i in range(V)]
    fg = LBP.FactorGraph(theta_en_en_names=['a'], theta_en_de_names=['b'], theta_en_en=np.zeros((1, 1)),
                         theta_en_de=np.zeros((1, 1)), phi_en_en_w1=None, phi_en_en=None, phi_en_de=None)
    vs = [LBP.VariableNode(id=i, var_type=LBP.VAR_TYPE_PREDICTED, domain_type='en', domain=dom,
                           supervised_label=dom[(7 * i + 3) % V]) for i in range(4)]
    fid = 0
    for i, v in enumerate(vs):
        f = LBP.FactorNode(id=fid, factor_type='en_de', observed_domain_size=5)
        f.add_varset_with_potentials(varset=[v], ptable=LBP.PotentialTable(v_id2dim={v.id: 0}, table=z['unary'][i],
                                                                          observed_dim=i % 5))
        fg.add_factor(f)
        fid += 1
    for i, j in ((0, 1), (1, 2), (0, 2), (2, 3)):
        f = LBP.FactorNode(id=fid, factor_type='en_en')
        f.add_varset_with_potentials(varset=[vs[i], vs[j]],
                                     ptable=LBP.PotentialTable(v_id2dim={vs[i].id: 0, vs[j].id: 1}, table=z['pair_%d_%d' % (i, j)]))
        fg.add_factor(f)
        fid += 1
    return fg, vs


def check_explicit_graph_structure(z):
    """no arithmetic: initialize() of an explicit-table graph creates the reference's uniform messages (LBP.py:200-216)"""
    fg, vs = explicit_graph_from_fixture(z)
    fg.initialize(int(z['roots'][0]))
    assert bool(fg.isLoopy) == bool(int(z['is_loopy']))
    keys = sorted(k[4:] for k in z.files if k.startswith('msg|'))
    assert sorted('%s|%s' % k for k in fg.messages.keys()) == keys
    for m in fg.messages.values():
        np.testing.assert_array_equal(m.m, np.full((60, 1), 1.0 / 60))


def check_explicit_graph(z):
    fg, vs = explicit_graph_from_fixture(z)
    roots = [int(r) for r in z['roots']]
    fg.initialize(roots[0])
    fg.treelike_inference(3, roots[1:])
    for k in z.files:
        if k.startswith('msg|'):
            _, a, b = k.split('|')
            np.testing.assert_allclose(fg.messages[a, b].m[:, 0], z[k], rtol=1e-11, atol=1e-300)
    marg = np.stack([v.get_marginal().m[:, 0] for v in vs])
    np.testing.assert_allclose(marg, z['marginals'], rtol=1e-11)
    np.testing.assert_allclose(fg.get_posterior_probs(), float(z['logp']), rtol=1e-11)


def check_per_node_updates(path):
    """The reference's treelike_inference loop (LBP.py:225-243) written by the CALLER with the per-node API:
    get_message_schedule + VariableNode / FactorNode.update_message_to, then marginals and the per-factor gradient."""
    from macaronicusermodeling_b200 import LBP
    z = np.load(path, allow_pickle=False)
    fg, spec = graph_from_fixture(z)
    roots = [int(r) for r in z['roots']]
    fg.initialize(roots[0])
    n_sweeps = spec['sweeps'] if fg.isLoopy else 1
    for it in range(n_sweeps):
        schedule = fg.get_message_schedule(fg.variables[roots[1 + it]])
        for frm, to in reversed(schedule):
            if not (isinstance(to, LBP.FactorNode) and len(to.varset) < 2):
                frm.update_message_to(to)
        for to, frm in schedule:
            if not (isinstance(to, LBP.FactorNode) and len(to.varset) < 2):
                frm.update_message_to(to)
    keys = [k for k in z.files if k.startswith('msg|')]
    assert len(keys) == len(fg.messages)
    for k in keys:
        _, a, b = k.split('|')
        assert np.abs(fg.messages[a, b].m[:, 0] - z[k]).max() < 1e-6, k
    marg = np.stack([fg.variables[v].get_marginal().m[:, 0] for v in sorted(fg.variables.keys())])
    assert np.abs(marg - z['marginals']).max() < 1e-6
    np.testing.assert_allclose(fg.get_posterior_probs(), float(z['logp']), rtol=2e-6)
    g_ee, g_ed = fg.get_unregularized_gradeint()
    np.testing.assert_allclose(g_ee, z['g_ee_unreg'], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(g_ed, z['g_ed_unreg'], rtol=1e-4, atol=2e-6)


def check_au_remaining(au):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'au_cases_more.npz'), allow_pickle=False)
    a, b, d1, d2, W, M, phi = (z[k] for k in ('a', 'b', 'd1', 'd2', 'W', 'M', 'phi'))
    tol = dict(rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(au.induce_s_pointwise_multiply_clip(d1, d2), z['induce_s_pointwise_multiply_clip'], **tol)
    np.testing.assert_array_equal(au.induce_s(a.copy()), z['induce_s'])
    np.testing.assert_array_equal(au.induce_s(a[:50].copy()), z['induce_s_small'])
    np.testing.assert_allclose(au.induce_s_mutliply_clip(a - 0.5, W), z['induce_s_mutliply_clip'], rtol=1e-11, atol=1e-13)
    d = au.make_sparse_and_dot(a, np.ascontiguousarray(b.T))
    keys = sorted(d.keys())
    np.testing.assert_array_equal(np.array(keys, dtype=np.int64), z['msd_keys'])
    np.testing.assert_allclose(np.array([d[k] for k in keys]), z['msd_vals'], **tol)
    mz, md = au.sparse_multiply_and_normalize(d, M)
    np.testing.assert_allclose(mz, z['smn_dense'], rtol=1e-11)
    np.testing.assert_allclose(np.array([md[k] for k in keys]), z['smn_vals'], rtol=1e-11)
    np.testing.assert_allclose(au.sd_matrix_multiply(d1, d2.T.copy()), z['sd_matrix_multiply'], rtol=1e-12)
    np.testing.assert_allclose(au.ss_matix_multiply(d1.T.copy(), d2), z['ss_matix_multiply'], rtol=1e-12)
    ap = au.make_adapt_phi(phi, 4)
    np.testing.assert_array_equal(ap, z['make_adapt_phi'])
    np.testing.assert_array_equal(au.set_adaptation(3, ap.copy(), [1, 3]), z['set_adaptation'])
    np.testing.assert_array_equal(au.set_adaptation_off(3, au.set_adaptation(3, ap.copy(), [1, 3]), [3]), z['set_adaptation_off'])
    np.testing.assert_array_equal(au.set_original(phi * 2.0, ap.copy()), z['set_original'])
    # the first fixture's top-K functions (they are what LBP.py's approximate paths call)
    y = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'au_cases.npz'), allow_pickle=False)
    a, b, T = y['a'], y['b'], y['T']
    np.testing.assert_allclose(au.sparse_vec_mat_dot(a, T), y['sparse_vec_mat_dot_col'], rtol=1e-12)
    np.testing.assert_allclose(au.sparse_vec_mat_dot(np.ascontiguousarray(a.T), T), y['sparse_vec_mat_dot_row'], rtol=1e-12)
    sd, ci, ri = au.sparse_dot(a, np.ascontiguousarray(b.T))
    np.testing.assert_allclose(sd, y['sparse_dot_out'], rtol=1e-12)
    np.testing.assert_array_equal(np.sort(ci), y['sparse_dot_cidx'])
    np.testing.assert_array_equal(np.sort(ri), y['sparse_dot_ridx'])
    spm = au.sparse_pointwise_multiply(sd, ci.astype(np.int64), ri.astype(np.int64), T)
    np.testing.assert_allclose(spm, y['sparse_pointwise_multiply'], rtol=1e-12)
    np.testing.assert_allclose(au.sparse_normalize(spm.copy(), ci, ri), y['sparse_normalize'], rtol=1e-11)
    np.testing.assert_array_equal(au.clip(y['clip_in'].copy()), y['clip'])
