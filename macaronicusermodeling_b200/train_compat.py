"""The reference trainer's entry points on top of the drop-in LBP module: ``create_factor_graph`` (train.py:133-305),
``batch_sgd`` (:357-397), ``batch_predictions`` (:308-338), ``batch_sgd_accumulate`` (:400-416), ``save_params`` /
``read_params`` (:46-99).

Differences from the reference, all at the edges:
* the module globals the reference reads (``options``, ``N``, ``de_domain``, ``domain2theta``; train.py:14-19,
  :145, :155-158) are explicit keyword arguments here;
* ``create_factor_graph`` does NOT materialise exp(phi . theta) over the (V,V,3) / (V,Vd,6) tensors on the CPU
  (train.py:218-253, 7 s per sentence at V = 10k): ``fg.pot_*`` stay None and the GPU engine rebuilds the
  potentials from theta once per step;
* sentences come as the JSON objects of training_classes.TrainingInstance (dicts), not pickled class instances.
For throughput use trainer.Trainer / trainer.batch_sgd_many, which batch many sentences per launch; this module is
the one-sentence-at-a-time API the reference's callers know.
"""
import argparse
import json
import sys

import numpy as np
from numpy import float64 as DTYPE

from . import synth
from .LBP import (FactorGraph, FactorNode, PhiWrapper, PotentialTable, VariableNode, VAR_TYPE_GIVEN,  # noqa: F401
                  VAR_TYPE_PREDICTED)

PRED2GIVEN = 'pred2given'
PRED2PRED = 'pred2pred'
F_EN_EN_NAMES = ['pmi', 'pmi_w1', 'bias']                                          # train.py:510
F_EN_DE_NAMES = ['ed', 'ped', 'correct', 'full_history', 'hit_history', 'bias']    # train.py:513


def default_options(**kw):
    """the flags create_factor_graph / batch_sgd read, with train.py's defaults (train.py:447-461)"""
    o = argparse.Namespace(use_approx_beliefs=False, use_approx_inference=False, report_times=False, reg_param='0.2',
                           reg_param_ua_scale='1.0', user_adapt=False, experience_adapt=False, use_correct_feat=True,
                           history=True, session_history=False)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def make_phi_wrapper(pmi, pmi_w1, ed, ped):
    """train.py:592-612"""
    ones = np.ones_like(pmi)
    phi_w1 = np.stack([pmi, pmi_w1, ones], axis=2).astype(DTYPE)
    phi = np.stack([pmi, np.zeros_like(pmi), ones], axis=2).astype(DTYPE)
    z = np.zeros_like(ed)
    phi_ed = np.stack([ed, ped, z, z.copy(), z.copy(), np.ones_like(ed)], axis=2).astype(DTYPE)
    return PhiWrapper(phi, phi_w1, phi_ed)


def apply_regularization(reg, grad, lr, theta):
    """train.py:31-35"""
    grad -= reg * theta
    grad *= lr
    return grad


def _guess_word(g):
    return synth._norm_guess(g['guess'])


def create_factor_graph(ti, learning_rate, theta_en_en_names, theta_en_de_names, theta_en_en, theta_en_de, phi_wrapper,
                        en_domain, de2id, en2id, d2t, options=None, N=1, de_domain=None):
    """train.py:133-305 with the drop-in LBP classes."""
    options = options if options is not None else default_options()
    if isinstance(ti, str):
        ti = json.loads(ti)
    nodes = sorted(ti['current_sent'], key=lambda n: int(n['position']))
    cg = {}
    for g in reversed(ti['current_guesses']):
        cg[tuple(g['id'])] = g
    rg = {}
    for g in reversed(ti['current_revealed_guesses']):
        rg[tuple(g['id'])] = g
    var_node_pairs = []
    for idx, n in enumerate(nodes):                                              # get_var_node_pair, train.py:102-130
        if n['lang'] == 'en':
            v = VariableNode(id=idx, var_type=VAR_TYPE_GIVEN, domain_type='en', domain=en_domain,
                             supervised_label=n['l2_word'].lower().replace("'", ""))
        else:
            nid = tuple(n['id'])
            guess, var_type = (cg[nid], VAR_TYPE_PREDICTED) if nid in cg else (rg[nid], VAR_TYPE_GIVEN)
            v = VariableNode(id=idx, var_type=var_type, domain_type='en', domain=en_domain, supervised_label=_guess_word(guess))
            if var_type == VAR_TYPE_PREDICTED:
                v.set_truth_label(n['l1_parent'].lower().replace("'", ""))
        var_node_pairs.append((v, n))
    len_en_domain = len(en_domain)
    len_de_domain = len(de_domain) if de_domain is not None else phi_wrapper.phi_en_de.shape[1]
    fg = FactorGraph(theta_en_en_names=theta_en_en_names, theta_en_de_names=theta_en_de_names, theta_en_en=theta_en_en,
                     theta_en_de=theta_en_de, phi_en_en=phi_wrapper.phi_en_en, phi_en_en_w1=phi_wrapper.phi_en_en_w1,
                     phi_en_de=phi_wrapper.phi_en_de)
    fg.learning_rate = learning_rate
    fg.use_approx_beliefs = options.use_approx_beliefs
    fg.use_approx_inference = options.use_approx_inference
    fg.report_times = options.report_times
    fg.regularization_param = float(options.reg_param) / float(N)               # train.py:158
    if options.user_adapt and options.experience_adapt:
        raise BaseException("2 domains not supported simultaniously")
    if options.user_adapt or options.experience_adapt:                            # train.py:160-173
        d = ti['user_id'] if options.user_adapt else len(ti['past_sentences_seen'])
        fg.active_domains['en_en', d] = 1
        fg.active_domains['en_de', d] = 1
        sys.stderr.write('+' if options.user_adapt else '=')
        fg.pot_theta_en_en = d2t['en_en', d]                                      # train.py:224-229: REPLACES the base theta
        fg.pot_theta_en_de = d2t['en_de', d]                                      # train.py:242-245
    # per-sentence dynamic features written into the SHARED phi_en_de planes (train.py:176-215)
    phi_ed = fg.phi_en_de
    if options.use_correct_feat:
        plane = np.zeros((len_en_domain, len_de_domain), dtype=DTYPE)
        for g in ti['current_guesses']:
            if g.get('reference') is not None and _guess_word(g) == g['reference']:
                plane[en2id[_guess_word(g)], de2id[g['l2_word']]] += 1.00
        phi_ed[:, :, theta_en_de_names.index('correct')] = plane
    if options.history:
        plane = np.zeros((len_en_domain, len_de_domain), dtype=DTYPE)
        for g in ti['past_correct_guesses']:
            plane[en2id[_guess_word(g)], de2id[g['l2_word']]] += 1.00
        phi_ed[:, :, theta_en_de_names.index('full_history')] = plane
    if options.session_history:
        plane = np.zeros((len_en_domain, len_de_domain), dtype=DTYPE)
        for g in ti['past_guesses_for_current_sent']:
            if not g['revealed']:
                plane[en2id[_guess_word(g)], de2id[g['l2_word']]] -= 1.00
        phi_ed[:, :, theta_en_de_names.index('hit_history')] = plane
    factors = []
    for v, n in var_node_pairs:                                                   # train.py:255-264
        if v.var_type == VAR_TYPE_PREDICTED:
            f = FactorNode(id=len(factors), factor_type='en_de', observed_domain_size=len_de_domain)
            p = PotentialTable(v_id2dim={v.id: 0}, table=None, observed_dim=de2id[n['l2_word']])
            f.add_varset_with_potentials(varset=[v], ptable=p)
            f.position = v.id
            f.gap = 0
            f.word_label = n['l2_word']
            factors.append(f)
    for idx_1, (v1, n1) in enumerate(var_node_pairs):                              # train.py:270-297
        for v2, n2 in var_node_pairs[idx_1 + 1:]:
            if v1.var_type == VAR_TYPE_PREDICTED and v2.var_type == VAR_TYPE_PREDICTED:
                f = FactorNode(id=len(factors), factor_type='en_en')
                p = PotentialTable(v_id2dim={v1.id: 0, v2.id: 1}, table=None, observed_dim=None)
                f.add_varset_with_potentials(varset=[v1, v2], ptable=p)
                f.gap = abs(v1.id - v2.id)
                f.connect_type = PRED2PRED
                factors.append(f)
            elif v1.var_type == VAR_TYPE_GIVEN and v2.var_type == VAR_TYPE_GIVEN:
                pass
            else:
                v_given = v1 if v1.var_type == VAR_TYPE_GIVEN else v2
                v_pred = v1 if v1.var_type == VAR_TYPE_PREDICTED else v2
                f = FactorNode(id=len(factors), factor_type='en_en', observed_domain_type='en',
                               observed_domain_size=len_en_domain)
                p = PotentialTable(v_id2dim={v_pred.id: 0}, table=None, observed_dim=en2id[v_given.supervised_label])
                f.add_varset_with_potentials(varset=[v_pred], ptable=p)
                f.position = v_given.id
                f.gap = abs(v_given.id - v_pred.id)
                f.connect_type = PRED2GIVEN
                f.word_label = v_given.supervised_label
                factors.append(f)
    for f in factors:
        fg.add_factor(f)
    for f in fg.factors:
        f.potential_table.slice_potentials()
    sys.stderr.write('.')
    return fg


def batch_sgd(training_instance, theta_en_en_names, theta_en_de_names, theta_en_en, theta_en_de, phi_wrapper, lr, en_domain,
              de2id, en2id, d2t, options=None, N=1, de_domain=None, roots=None):
    """train.py:357-397 -> [sent_id, logp, g_en_en (1,3), g_en_de (1,6), sample_ag].  ``roots`` (optional) pins the BFS
    roots [has_loops draw, sweep 1, sweep 2, sweep 3] instead of drawing them from ``random``."""
    ti = json.loads(training_instance) if isinstance(training_instance, str) else training_instance
    sent_id = ti['current_sent'][0]['sent_id']
    fg = create_factor_graph(ti, lr, theta_en_en_names, theta_en_de_names, theta_en_en, theta_en_de, phi_wrapper, en_domain,
                             de2id, en2id, d2t, options, N, de_domain)
    fg.initialize(None if roots is None else roots[0])
    fg.treelike_inference(3, None if roots is None else roots[1:])
    if options is not None and (options.user_adapt or options.experience_adapt):  # train.py:379-390
        g_en_en, g_en_de = fg.get_unregularized_gradeint()
        sample_ag = {}
        r, l = fg.regularization_param, fg.learning_rate
        scale_reg = float(options.reg_param_ua_scale)
        for f_type, d in fg.active_domains:
            g = g_en_en.copy() if f_type == 'en_en' else g_en_de.copy()
            sample_ag[f_type, d] = apply_regularization(r * scale_reg, g, l, d2t[f_type, d])
        g_en_en = apply_regularization(r, g_en_en, l, fg.theta_en_en)
        g_en_de = apply_regularization(r, g_en_de, l, fg.theta_en_de)
    else:
        sample_ag = None
        g_en_en, g_en_de = fg.return_gradient()
    fg.display_timing_info()
    p = fg.get_posterior_probs()
    return [sent_id, p, g_en_en, g_en_de, sample_ag]


def batch_predictions(training_instance, theta_en_en_names, theta_en_de_names, theta_en_en, theta_en_de, phi_wrapper, lr,
                      en_domain, de2id, en2id, d2t, qp=False, options=None, N=1, de_domain=None, roots=None):
    """train.py:308-338 -> [logp, prediction string, dist string, (p@0, p@25, p@50, total)]"""
    ti = json.loads(training_instance) if isinstance(training_instance, str) else training_instance
    sent_id = ti['current_sent'][0]['sent_id']
    fg = create_factor_graph(ti, lr, theta_en_en_names, theta_en_de_names, theta_en_en, theta_en_de, phi_wrapper, en_domain,
                             de2id, en2id, d2t, options, N, de_domain)
    fg.initialize(None if roots is None else roots[0])
    fg.treelike_inference(3, None if roots is None else roots[1:])
    p = fg.get_posterior_probs()
    if qp:
        factor_dist = fgs = None
    else:
        fgs = '\n'.join(['*SENT_ID:' + str(sent_id)] + fg.to_string())
        factor_dist = fg.to_dist()
    return [p, fgs, factor_dist, fg.get_precision_counts()]


def batch_sgd_accumulate(result, f_en_en_theta, f_en_de_theta, domain2theta=None):
    """train.py:400-416: in-place theta update (and, when adapting, of the active domains' thetas, :406-409); returns
    the log-probability to add to train_prediction_probs"""
    f_en_en_theta += result[2]
    f_en_de_theta += result[3]
    if result[4] is not None:
        for key, ag in result[4].items():
            domain2theta[key] += ag
    sys.stderr.write('*')
    return result[1]


def save_params(w, ee_theta, ed_theta, ee_names, ed_names, d2t):
    """train.py:77-99 (same text format, 6 decimals)"""
    w.write('\t'.join(['EE_F:'] + ee_names) + '\n')
    fl = [item for sublist in ee_theta.tolist() for item in sublist]
    w.write('\t'.join(['Original'.ljust(15)] + ['%0.6f' % i for i in fl]) + '\n')
    for ft, d in d2t:
        if ft == 'en_en':
            fl = [item for sublist in d2t[ft, d].tolist() for item in sublist]
            w.write('\t'.join([d.ljust(15)] + ['%0.6f' % i for i in fl]) + '\n')
    w.write('\t'.join(['ED_F:'] + ed_names) + '\n')
    fl = [item for sublist in ed_theta.tolist() for item in sublist]
    w.write('\t'.join(['Original'.ljust(15)] + ['%0.6f' % i for i in fl]) + '\n')
    for ft, d in d2t:
        if ft == 'en_de':
            fl = [item for sublist in d2t[ft, d].tolist() for item in sublist]
            w.write('\t'.join([d.ljust(15)] + ['%0.6f' % i for i in fl]) + '\n')
    w.flush()
    w.close()


def read_params(params_file):
    """train.py:46-74"""
    import codecs
    d2t = {}
    p1, p2 = codecs.open(params_file, 'r', 'utf8').read().split('ED_F:')
    _, p1 = p1.strip().split('EE_F:')
    p1_lines = p1.split('\n')
    een = p1_lines[0].split()
    eet = np.array([float(i) for i in p1_lines[1].split()[1:]]).reshape(1, -1)
    for line in p1_lines[2:]:
        items = line.split()
        d2t['en_en', items[0].strip()] = np.array([float(i) for i in items[1:]]).reshape(1, -1)
    p2_lines = p2.strip().split('\n')
    edn = p2_lines[0].split()
    edt = np.array([float(i) for i in p2_lines[1].split()[1:]]).reshape(1, -1)
    for line in p2_lines[2:]:
        items = line.split()
        d2t['en_de', items[0]] = np.array([float(i) for i in items[1:]]).reshape(1, -1)
    return een, eet, edn, edt, d2t
