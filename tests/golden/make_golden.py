#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

The reference (/root/reference, Python 2) publishes no golden vectors, so the
oracle in oracle/lbp_oracle.py is pinned against outputs of the reference's own code executed
here.  This script

  1. copies LBP.py / train.py / training_classes.py / utils / array_utils into a scratch
     directory OUTSIDE the repo and applies the mechanical, arithmetic-free py2->py3 patch of
     SURVEY.md §8(c) (print statements, iteritems, dict.keys() indexing, np.int -> np.int64 ...),
  2. builds array_utils/c_array_utils.pyx with Cython exactly like array_utils/setup.py does,
  3. drives the reference through its own call sites (train.create_factor_graph, train.batch_sgd,
     FactorGraph.initialize / treelike_inference / return_gradient / get_posterior_probs /
     get_precision_counts, VariableNode.get_marginal / get_max_vocab, the `au` helpers) on
     seeded synthetic macaronic sentences,
  4. writes the inputs and the reference's outputs as .npz fixtures next to this file.

Nothing from the reference is copied into the repository; only numbers it computed are.
The fixtures travel to the GPU box; /root/reference does not.  Re-run with:

    python tests/golden/make_golden.py            # needs /root/reference, Cython, gcc

The BFS root of every sweep (LBP.py:176 and :223 draw it from the global `random`) is injected
through a stub `random.sample`, because Python 2 and Python 3 streams differ for the same seed
(SURVEY.md §8(c)); the roots are stored in the fixtures and fed to the oracle and the GPU path.
"""
import argparse
import importlib
import io
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from macaronicusermodeling_b200 import synth  # noqa: E402  (sentence/model generator shared with tests)


# ----------------------------------------------------------------------------- patching
def _patch_py(src):
    out = []
    for line in src.split('\n'):
        m = re.match(r'^(\s*)print (.*)$', line)
        if m and not line.lstrip().startswith('#'):
            line = '%sprint(%s)' % (m.group(1), m.group(2))
        elif re.match(r'^\s*print\s*$', line):
            line = line.replace('print', 'print()')
        out.append(line)
    s = '\n'.join(out)
    s = s.replace('.iteritems()', '.items()').replace('xrange(', 'range(')
    s = s.replace('v_id2dim[v_id2dim.keys()[0]]', 'v_id2dim[list(v_id2dim.keys())[0]]')
    s = s.replace('random.sample(self.variables.keys(), 1)', 'random.sample(sorted(self.variables.keys()), 1)')
    s = re.sub(r'except (\w+), (\w+):', r'except \1 as \2:', s)
    # LBP.py:110,127 sort (position, factor) tuples; ties need a key on py3
    s = s.replace('sorted([(f.position, f) for f in self.factors if f.position is not None])',
                  'sorted([(f.position, f) for f in self.factors if f.position is not None], key=lambda t: t[0])')
    return s


def _patch_pyx(src):
    s = src.replace('np.int_t', 'np.int64_t')
    s = s.replace('dtype=np.int)', 'dtype=np.int64)').replace('dtype=np.float)', 'dtype=np.float64)')
    return s


def build_patched_reference(ref_root, work):
    for name in ['LBP.py', 'train.py', 'training_classes.py']:
        with io.open(os.path.join(ref_root, name), encoding='utf8') as f:
            s = _patch_py(f.read())
        with io.open(os.path.join(work, name), 'w', encoding='utf8') as f:
            f.write(s)
    os.makedirs(os.path.join(work, 'utils'))
    for name in os.listdir(os.path.join(ref_root, 'utils')):
        if name.endswith('.py'):
            with io.open(os.path.join(ref_root, 'utils', name), encoding='utf8') as f:
                s = _patch_py(f.read())
            with io.open(os.path.join(work, 'utils', name), 'w', encoding='utf8') as f:
                f.write(s)
    au = os.path.join(work, 'array_utils')
    os.makedirs(au)
    open(os.path.join(au, '__init__.py'), 'w').close()
    with io.open(os.path.join(ref_root, 'array_utils', 'c_array_utils.pyx'), encoding='utf8') as f:
        s = _patch_pyx(f.read())
    with io.open(os.path.join(au, 'c_array_utils.pyx'), 'w', encoding='utf8') as f:
        f.write(s)
    with open(os.path.join(work, 'setup_au.py'), 'w') as f:
        f.write("from setuptools import setup\nfrom Cython.Build import cythonize\nimport numpy\n"
                "setup(name='c_array_utils', ext_modules=cythonize('array_utils/c_array_utils.pyx', "
                "compiler_directives={'language_level': 2}), include_dirs=[numpy.get_include()])\n")
    subprocess.check_call([sys.executable, 'setup_au.py', 'build_ext', '--inplace'], cwd=work,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


class _RootFeeder(object):
    """Stands in for the `random` module inside the patched LBP: sample() pops injected roots."""

    def __init__(self):
        self.queue = []
        self.log = []

    def sample(self, population, k):
        assert k == 1
        r = self.queue.pop(0)
        assert r in list(population), (r, list(population))
        self.log.append(r)
        return [r]


def import_reference(work):
    sys.path.insert(0, work)
    sys.modules['enchant'] = types.ModuleType('enchant')
    LBP = importlib.import_module('LBP')
    train = importlib.import_module('train')
    feeder = _RootFeeder()
    LBP.random = feeder
    return LBP, train, feeder


# ----------------------------------------------------------------------------- driving the reference
def ref_options(session_history=True, history=True, approx_inference=False, approx_beliefs=False, user_adapt=False,
                ua_scale='1.0'):
    return argparse.Namespace(use_approx_beliefs=approx_beliefs, use_approx_inference=approx_inference, report_times=False,
                              reg_param='0.2', reg_param_ua_scale=ua_scale, user_adapt=user_adapt,
                              experience_adapt=False, use_correct_feat=True, history=history,
                              session_history=session_history)


def make_phi(model):
    ones = np.ones_like(model['pmi'])
    phi_ee_w1 = np.stack([model['pmi'], model['pmi_w1'], ones], axis=2)             # train.py:594
    phi_ee = np.stack([model['pmi'], np.zeros_like(model['pmi']), ones], axis=2)    # train.py:595
    z = np.zeros_like(model['ed'])
    phi_ed = np.stack([model['ed'], model['ped'], z, z.copy(), z.copy(), np.ones_like(model['ed'])], axis=2)  # :609
    return phi_ee, phi_ee_w1, phi_ed


def run_graph_case(LBP, train, feeder, model, sent, theta_ee, theta_ed, roots, n_sweeps, N, lr, opts):
    """create_factor_graph -> initialize -> treelike_inference -> everything the callers read."""
    V, Vd = model['pmi'].shape[0], model['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    en2id = dict((e, i) for i, e in enumerate(en_domain))
    de2id = dict((d, i) for i, d in enumerate(de_domain))
    phi_ee, phi_ee_w1, phi_ed = make_phi(model)
    pw = LBP.PhiWrapper(phi_ee, phi_ee_w1, phi_ed)
    train.options = opts
    train.N = N
    train.de_domain = de_domain
    train.domain2theta = {}
    ti = train.TrainingInstance.from_dict(json.loads(synth.sentence_to_json(sent)))
    t_ee = np.array(theta_ee, dtype=np.float64).reshape(1, -1)
    t_ed = np.array(theta_ed, dtype=np.float64).reshape(1, -1)
    fg = train.create_factor_graph(ti=ti, learning_rate=lr,
                                   theta_en_en_names=['pmi', 'pmi_w1', 'bias'],
                                   theta_en_de_names=['ed', 'ped', 'correct', 'full_history', 'hit_history', 'bias'],
                                   theta_en_en=t_ee, theta_en_de=t_ed, phi_wrapper=pw, en_domain=en_domain,
                                   de2id=de2id, en2id=en2id, d2t={})
    feeder.queue = [roots[0]]            # has_loops draw (LBP.py:176)
    fg.initialize()
    n_eff = n_sweeps if fg.isLoopy else 1
    feeder.queue = list(roots[1:1 + n_eff])
    fg.treelike_inference(n_sweeps)
    assert not feeder.queue
    out = {}
    out['is_loopy'] = np.array(int(bool(fg.isLoopy)))
    out['var_ids'] = np.array(sorted(fg.variables.keys()), dtype=np.int64)
    marg = np.stack([fg.variables[v].get_marginal().m[:, 0] for v in sorted(fg.variables.keys())])
    out['marginals'] = marg
    out['top1'] = marg.argmax(axis=1)
    out['logp'] = np.array(fg.get_posterior_probs())
    g_ee, g_ed = fg.get_unregularized_gradeint()
    out['g_ee_unreg'] = g_ee.copy()
    out['g_ed_unreg'] = g_ed.copy()
    r_ee, r_ed = fg.return_gradient()
    out['g_ee_ret'] = r_ee.copy()
    out['g_ed_ret'] = r_ed.copy()
    out['precision_counts'] = np.array(fg.get_precision_counts(), dtype=np.int64)
    # every message, keyed 'msg|F_3|X_1'
    for (a, b), m in fg.messages.items():
        out['msg|%s|%s' % (a, b)] = m.m[:, 0].copy()
    # factor table layout so that the oracle test can check its own graph construction
    out['factor_desc'] = np.array([[f.id, 0 if f.factor_type == 'en_de' else 1, len(f.varset), f.varset[0].id,
                                    f.varset[1].id if len(f.varset) > 1 else -1, f.gap,
                                    -1 if f.potential_table.observed_dim is None else f.potential_table.observed_dim]
                                   for f in fg.factors], dtype=np.int64)
    sl, slp, top = fg.variables[sorted(fg.variables.keys())[0]].get_max_vocab(min(50, V - 1))
    out['maxvocab0_words'] = np.array([w for w, _ in top])
    out['maxvocab0_logp'] = np.array([float(p) for _, p in top])
    out['to_string'] = np.array(fg.to_string())
    return out


def run_batch_sgd(LBP, train, feeder, model, sents, theta_ee, theta_ed, roots_per_sent, lr, N, opts):
    V, Vd = model['pmi'].shape[0], model['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    en2id = dict((e, i) for i, e in enumerate(en_domain))
    de2id = dict((d, i) for i, d in enumerate(de_domain))
    phi_ee, phi_ee_w1, phi_ed = make_phi(model)
    pw = LBP.PhiWrapper(phi_ee, phi_ee_w1, phi_ed)
    train.options = opts
    train.N = N
    train.de_domain = de_domain
    train.domain2theta = {}
    res = []
    for s, roots in zip(sents, roots_per_sent):
        feeder.queue = list(roots)
        r = train.batch_sgd(synth.sentence_to_json(s), ['pmi', 'pmi_w1', 'bias'],
                            ['ed', 'ped', 'correct', 'full_history', 'hit_history', 'bias'],
                            theta_ee, theta_ed, pw, lr, en_domain, de2id, en2id, {})
        feeder.queue = []
        res.append(r)
    return res


def run_sgd_trajectory_user_adapt(LBP, train, feeder, model, sents, roots, epochs, opts, users):
    """train.py:617-638 with --user_adapt: per-user theta REPLACES the base theta in the potentials (:224-229, :242-245);
    both get the gradient, with their own regularisation (:379-390, :402-409)."""
    V, Vd = model['pmi'].shape[0], model['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    en2id = dict((e, i) for i, e in enumerate(en_domain))
    de2id = dict((d, i) for i, d in enumerate(de_domain))
    phi_ee, phi_ee_w1, phi_ed = make_phi(model)
    pw = LBP.PhiWrapper(phi_ee, phi_ee_w1, phi_ed)
    train.options = opts
    train.N = len(sents)
    train.de_domain = de_domain
    train.f_en_en_theta = np.zeros((1, 3))
    train.f_en_de_theta = np.zeros((1, 6))
    train.train_prediction_probs = 0.0
    d2t = {}
    for u in users:
        d2t['en_en', u] = np.zeros((1, 3))
        d2t['en_de', u] = np.zeros((1, 6))
    train.domain2theta = d2t
    traj = []
    for epoch in range(epochs):
        lr = 0.1 / float(1.0 + epoch * 0.3)
        for si, s in enumerate(sents):
            feeder.queue = list(roots[epoch][si])
            r = train.batch_sgd(synth.sentence_to_json(s), ['pmi', 'pmi_w1', 'bias'],
                                ['ed', 'ped', 'correct', 'full_history', 'hit_history', 'bias'],
                                train.f_en_en_theta, train.f_en_de_theta, pw, lr, en_domain, de2id, en2id, d2t)
            feeder.queue = []
            train.batch_sgd_accumulate(r)
            row = [train.f_en_en_theta[0], train.f_en_de_theta[0]]
            for u in users:
                row += [d2t['en_en', u][0], d2t['en_de', u][0]]
            traj.append(np.concatenate(row))
    return np.array(traj)


def run_sgd_trajectory(LBP, train, feeder, model, sents, roots, epochs, opts):
    """train.py:617-638 re-enacted with the reference's batch_sgd / batch_sgd_accumulate (no shuffle:
    the order is an input)."""
    train.f_en_en_theta = np.zeros((1, 3))
    train.f_en_de_theta = np.zeros((1, 6))
    train.train_prediction_probs = 0.0
    N = len(sents)
    traj = []
    logps = []
    for epoch in range(epochs):
        lr = 0.1 / float(1.0 + epoch * 0.3)                                        # train.py:621
        for si, s in enumerate(sents):
            r = run_batch_sgd(LBP, train, feeder, model, [s], train.f_en_en_theta, train.f_en_de_theta,
                              [roots[epoch][si]], lr, N, opts)[0]
            train.batch_sgd_accumulate(r)
            logps.append(r[1])
            traj.append(np.concatenate([train.f_en_en_theta[0], train.f_en_de_theta[0]]))
    return np.array(traj), np.array(logps)


def au_cases(au):
    rng = np.random.default_rng(7)
    out = {}
    a = rng.random((150, 1))
    b = rng.random((150, 1))
    T = rng.random((150, 150))
    out['a'], out['b'], out['T'] = a, b, T
    out['pointwise_multiply'] = au.pointwise_multiply(a, b)
    out['normalize'] = au.normalize(a.copy())
    z = np.zeros((5, 1))
    out['normalize_zero'] = au.normalize(z)
    out['dense_dot_Tv'] = au.dense_dot(T, a)
    out['dense_dot_vT'] = au.dense_dot(np.ascontiguousarray(a.T), T)
    out['dense_dot_outer'] = au.dense_dot(a, np.ascontiguousarray(b.T))
    out['dense_pointwise_multiply'] = au.dense_pointwise_multiply(T, T.T.copy())
    out['sparse_vec_mat_dot_col'] = au.sparse_vec_mat_dot(a, T)
    out['sparse_vec_mat_dot_row'] = au.sparse_vec_mat_dot(np.ascontiguousarray(a.T), T)
    sd, ci, ri = au.sparse_dot(a, np.ascontiguousarray(b.T))
    out['sparse_dot_out'], out['sparse_dot_cidx'], out['sparse_dot_ridx'] = sd, np.sort(ci), np.sort(ri)
    spm = au.sparse_pointwise_multiply(sd, ci.astype(np.int64), ri.astype(np.int64), T)
    out['sparse_pointwise_multiply'] = spm
    out['sparse_normalize'] = au.sparse_normalize(spm.copy(), ci, ri)
    c = rng.random((6, 1)) * 1e-100
    out['clip_in'] = c.copy()
    out['clip'] = au.clip(c.copy())
    return out


def au_cases_more(au):
    """the remaining c_array_utils functions (none of them is called by LBP.py; kept for API completeness)"""
    rng = np.random.default_rng(11)
    out = {}
    a = rng.random((160, 1))
    b = rng.random((160, 1))
    d1 = rng.random((40, 30))
    d2 = rng.random((40, 30))
    W = rng.random((12, 160))
    out['a'], out['b'], out['d1'], out['d2'], out['W'] = a, b, d1, d2, W
    out['induce_s_pointwise_multiply_clip'] = au.induce_s_pointwise_multiply_clip(d1, d2)
    out['induce_s'] = au.induce_s(a.copy())
    out['induce_s_small'] = au.induce_s(a[:50].copy())
    out['induce_s_mutliply_clip'] = au.induce_s_mutliply_clip(a - 0.5, W)
    d = au.make_sparse_and_dot(a, np.ascontiguousarray(b.T))
    keys = sorted(d.keys())
    out['msd_keys'] = np.array(keys, dtype=np.int64)
    out['msd_vals'] = np.array([d[k] for k in keys])
    M = rng.random((160, 160))
    out['M'] = M
    mz, md = au.sparse_multiply_and_normalize(d, M)
    out['smn_dense'] = mz
    out['smn_vals'] = np.array([md[k] for k in keys])
    out['sd_matrix_multiply'] = au.sd_matrix_multiply(d1, d2.T.copy())
    out['ss_matix_multiply'] = au.ss_matix_multiply(d1.T.copy(), d2)
    phi = rng.random((7, 3))
    out['phi'] = phi
    ap = au.make_adapt_phi(phi, 4)
    out['make_adapt_phi'] = ap.copy()
    out['set_adaptation'] = au.set_adaptation(3, ap.copy(), [1, 3]).copy()
    out['set_adaptation_off'] = au.set_adaptation_off(3, au.set_adaptation(3, ap.copy(), [1, 3]), [3]).copy()
    out['set_original'] = au.set_original(phi * 2.0, ap.copy()).copy()
    return out


def explicit_table_case(LBP):
    """a graph built from explicit PotentialTable(table=...) arrays (the run.py / toy style): messages and marginals"""
    rng = np.random.default_rng(3)
    V = 60
    dom = ['w%d' % i for i in range(V)]
    fg = LBP.FactorGraph(theta_en_en_names=['a'], theta_en_de_names=['b'], theta_en_en=np.zeros((1, 1)),
                         theta_en_de=np.zeros((1, 1)), phi_en_en_w1=None, phi_en_en=None, phi_en_de=None)
    vs = [LBP.VariableNode(id=i, var_type=LBP.VAR_TYPE_PREDICTED, domain_type='en', domain=dom, supervised_label=dom[(7 * i + 3) % V])
          for i in range(4)]
    unary = rng.random((4, V, 5)) + 0.05
    pair = {}
    fid = 0
    for i, v in enumerate(vs):
        f = LBP.FactorNode(id=fid, factor_type='en_de', observed_domain_size=5)
        f.add_varset_with_potentials(varset=[v], ptable=LBP.PotentialTable(v_id2dim={v.id: 0}, table=unary[i], observed_dim=i % 5))
        fg.add_factor(f)
        fid += 1
    for i, j in ((0, 1), (1, 2), (0, 2), (2, 3)):
        t = rng.random((V, V)) + 0.01
        pair['%d_%d' % (i, j)] = t
        f = LBP.FactorNode(id=fid, factor_type='en_en')
        f.add_varset_with_potentials(varset=[vs[i], vs[j]], ptable=LBP.PotentialTable(v_id2dim={vs[i].id: 0, vs[j].id: 1}, table=t))
        fg.add_factor(f)
        fid += 1
    return fg, vs, unary, pair


def run_explicit_case(LBP, feeder):
    fg, vs, unary, pair = explicit_table_case(LBP)
    roots = [1, 2, 0, 3]
    feeder.queue = [roots[0]]
    fg.initialize()
    feeder.queue = list(roots[1:])
    fg.treelike_inference(3)
    out = {'roots': np.array(roots), 'unary': unary, 'is_loopy': np.array(int(fg.isLoopy))}
    for k, t in pair.items():
        out['pair_' + k] = t
    for (a, b), m in fg.messages.items():
        out['msg|%s|%s' % (a, b)] = m.m[:, 0]
    out['marginals'] = np.stack([v.get_marginal().m[:, 0] for v in vs])
    out['logp'] = np.array(fg.get_posterior_probs())
    return out


# ----------------------------------------------------------------------------- cases
def run_formats_case(LBP, train, feeder, work):
    """On-disk formats written BY THE REFERENCE: a parameter checkpoint from train.save_params
    (train.py:77-99) with adapted domains, what train.read_params (:46-74) parses back from it, and the prediction / .dist
    lines of FactorGraph.to_string / to_dist (LBP.py:109-143) for one graph."""
    import codecs
    rng = np.random.default_rng(77)
    ee = rng.normal(size=(1, 3)) * 0.7
    ed = rng.normal(size=(1, 6)) * 0.7
    d2t = {}
    for d in ('user_a', 'u2', 'a_rather_long_user_name'):
        d2t['en_en', d] = rng.normal(size=(1, 3))
        d2t['en_de', d] = rng.normal(size=(1, 6)) * 1e-3        # small values: the 6-decimal rounding is visible
    ee_names, ed_names = ['pmi', 'pmi_w1', 'bias'], ['ed', 'ped', 'correct', 'full_history', 'hit_history', 'bias']
    path = os.path.join(work, 'golden.params')
    train.save_params(codecs.open(path, 'w', 'utf8'), ee, ed, ee_names, ed_names, d2t)
    text = open(path, 'rb').read().decode('utf8')
    een, eet, edn, edt, back = train.read_params(path)
    out = {'params_text': np.array(text), 'ee': ee, 'ed': ed, 'domains': np.array([d for ft, d in d2t if ft == 'en_en']),
           'd_ee': np.stack([d2t['en_en', d][0] for ft, d in d2t if ft == 'en_en']),
           'd_ed': np.stack([d2t['en_de', d][0] for ft, d in d2t if ft == 'en_en']),
           'read_een': np.array(een), 'read_edn': np.array(edn), 'read_ee': eet, 'read_ed': edt,
           'read_d_ee': np.stack([back['en_en', d][0] for ft, d in d2t if ft == 'en_en']),
           'read_d_ed': np.stack([back['en_de', d][0] for ft, d in d2t if ft == 'en_en'])}
    # prediction + .dist lines of one graph (the toy5 case: given + predicted tokens, history features)
    spec = synth.golden_case_specs()['toy5']
    model = synth.make_model(spec['V'], spec['Vd'], seed=spec['model_seed'])
    sent = synth.make_sentence(model, spec['layout'], seed=spec['sent_seed'], n_history=spec.get('n_history', 2))
    r = np.random.default_rng(spec['sent_seed'] + 99)
    pred = [p for p, k in enumerate(spec['layout']) if k == 'p']
    roots = [int(r.choice(pred)) for _ in range(1 + spec['sweeps'])]
    V, Vd = spec['V'], spec['Vd']
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    phi_ee, phi_ee_w1, phi_ed = make_phi(model)
    pw = LBP.PhiWrapper(phi_ee, phi_ee_w1, phi_ed)
    train.options = ref_options()
    train.N = 10
    train.de_domain = de_domain
    train.domain2theta = {}
    ti = train.TrainingInstance.from_dict(json.loads(synth.sentence_to_json(sent)))
    fg = train.create_factor_graph(ti=ti, learning_rate=0.1, theta_en_en_names=ee_names, theta_en_de_names=ed_names,
                                   theta_en_en=np.array(spec['theta_ee'], dtype=np.float64).reshape(1, -1),
                                   theta_en_de=np.array(spec['theta_ed'], dtype=np.float64).reshape(1, -1), phi_wrapper=pw,
                                   en_domain=en_domain, de2id=dict((d, i) for i, d in enumerate(de_domain)),
                                   en2id=dict((e, i) for i, e in enumerate(en_domain)), d2t={})
    feeder.queue = [roots[0]]
    fg.initialize()
    feeder.queue = list(roots[1:1 + (spec['sweeps'] if fg.isLoopy else 1)])
    fg.treelike_inference(spec['sweeps'])
    out['case'] = np.array('toy5')
    out['roots'] = np.array(roots)
    out['to_string'] = np.array(fg.to_string())
    out['to_dist'] = np.array(fg.to_dist())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=HERE)
    ap.add_argument('--only-new', action='store_true', help='write only au_cases_more.npz and graphx_explicit.npz')
    ap.add_argument('--only-formats', action='store_true', help='write only formats.npz (checkpoint / prediction / .dist text)')
    args = ap.parse_args()
    work = tempfile.mkdtemp(prefix='mlbp_ref_')
    try:
        build_patched_reference(args.ref, work)
        LBP, train, feeder = import_reference(work)
        au = importlib.import_module('array_utils.c_array_utils')
        np.savez_compressed(os.path.join(args.out, 'formats.npz'), **run_formats_case(LBP, train, feeder, work))
        if args.only_formats:
            return
        np.savez_compressed(os.path.join(args.out, 'au_cases_more.npz'), **au_cases_more(au))
        np.savez_compressed(os.path.join(args.out, 'graphx_explicit.npz'), **run_explicit_case(LBP, feeder))
        if args.only_new:
            return
        np.savez_compressed(os.path.join(args.out, 'au_cases.npz'), **au_cases(au))

        # ---- graph cases (BASELINE config C1-style toy graphs + structural edge cases)
        cases = synth.golden_case_specs()
        for name, spec in cases.items():
            model = synth.make_model(spec['V'], spec['Vd'], seed=spec['model_seed'], w1_density=spec.get('w1_density', 1.0))
            sent = synth.make_sentence(model, spec['layout'], seed=spec['sent_seed'], n_history=spec.get('n_history', 2))
            rng = np.random.default_rng(spec['sent_seed'] + 99)
            pred = [p for p, k in enumerate(spec['layout']) if k == 'p']
            roots = [int(rng.choice(pred)) for _ in range(1 + spec['sweeps'])]
            theta_ee = spec['theta_ee']
            theta_ed = spec['theta_ed']
            out = run_graph_case(LBP, train, feeder, model, sent, theta_ee, theta_ed, roots, spec['sweeps'],
                                 N=spec.get('N', 10), lr=spec.get('lr', 0.1),
                                 opts=ref_options(approx_inference=spec.get('approx_inference', False),
                                                  approx_beliefs=spec.get('approx_beliefs', False)))
            np.savez_compressed(os.path.join(args.out, 'graph_%s.npz' % name),
                                spec=json.dumps(spec), sentence=synth.sentence_to_json(sent), roots=np.array(roots),
                                pmi=model['pmi'], pmi_w1=model['pmi_w1'], ed=model['ed'], ped=model['ped'],
                                theta_ee=np.array(theta_ee), theta_ed=np.array(theta_ed), **out)
            print('graph case', name, 'logp', float(out['logp']), 'loopy', int(out['is_loopy']))

        # ---- trainer-level: 2 epochs of train.py's per-sentence SGD on 5 sentences (BASELINE config C2 semantics)
        model = synth.make_model(80, 30, seed=21)
        layouts = ['pppp', 'gpgpp', 'ppgpgp', 'pp', 'pgppg']
        sents = [synth.make_sentence(model, l, seed=300 + i, n_history=3) for i, l in enumerate(layouts)]
        rng = np.random.default_rng(5)
        roots = [[[int(rng.choice([p for p, k in enumerate(l) if k == 'p'])) for _ in range(4)] for l in layouts]
                 for _ in range(2)]
        traj, logps = run_sgd_trajectory(LBP, train, feeder, model, sents, roots, 2, ref_options())
        np.savez_compressed(os.path.join(args.out, 'sgd_trajectory.npz'),
                            sentences=np.array([synth.sentence_to_json(s) for s in sents]),
                            roots=np.array(roots), traj=traj, logps=logps,
                            pmi=model['pmi'], pmi_w1=model['pmi_w1'], ed=model['ed'], ped=model['ped'])
        print('sgd trajectory final theta', traj[-1])
        users = ['ua', 'ub']
        sents_u = [synth.make_sentence(model, l, seed=400 + i, n_history=3, user_id=users[i % 2]) for i, l in enumerate(layouts)]
        traj_u = run_sgd_trajectory_user_adapt(LBP, train, feeder, model, sents_u, roots, 2,
                                               ref_options(user_adapt=True, ua_scale='0.5'), users)
        np.savez_compressed(os.path.join(args.out, 'sgd_trajectory_user_adapt.npz'),
                            sentences=np.array([synth.sentence_to_json(s) for s in sents_u]), roots=np.array(roots),
                            traj=traj_u, users=np.array(users), pmi=model['pmi'], pmi_w1=model['pmi_w1'], ed=model['ed'],
                            ped=model['ped'])
        print('user-adapt trajectory final', traj_u[-1][:9])
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    main()
