"""GPU tier: the reference-flag command line (real-data front end) on files written in the reference's formats."""
import codecs
import json
import os
import random

import numpy as np
import pytest

from macaronicusermodeling_b200 import build, synth, train_cli
from macaronicusermodeling_b200.engine import Corpus
from macaronicusermodeling_b200.train_compat import read_params
from oracle import lbp_oracle as orc

pytestmark = pytest.mark.gpu


def write_inputs(tmp, model, sents):
    V, Vd = model['V'], model['Vd']
    p = {}
    for name, words in (('en', [synth.en_word(i) for i in range(V)]), ('de', [synth.de_word(i) for i in range(Vd)])):
        p[name] = str(tmp / (name + '.vocab'))
        with codecs.open(p[name], 'w', 'utf8') as f:
            f.write('\n'.join(words) + '\n')
    for k in ('pmi', 'pmi_w1', 'ed', 'ped'):
        p[k] = str(tmp / (k + '.mat'))
        np.savetxt(p[k], model[k], fmt='%.17g')
    p['ti'] = str(tmp / 'train.json')
    with codecs.open(p['ti'], 'w', 'utf8') as f:
        for s in sents:
            f.write(json.dumps(s) + '\n')
    return p


def test_cli_train_then_predict(tmp_path):
    build.build()
    model = synth.make_model(120, 24, seed=41)
    layouts = ['pppp', 'gpgpp', 'ppgp', 'pp', 'pgppg', 'ppp']
    raw = [synth.make_sentence(model, l, seed=500 + i, n_history=3, sent_id=i) for i, l in enumerate(layouts)]
    p = write_inputs(tmp_path, model, raw)
    params = str(tmp_path / 'model.params')
    base = ['--ti', p['ti'], '--end', p['en'], '--ded', p['de'], '--phi_pmi', p['pmi'], '--phi_pmi_w1', p['pmi_w1'],
            '--phi_ed', p['ed'], '--phi_ped', p['ped'], '--history', '--session_history']
    assert train_cli.main(base + ['--save_params', params, '--minibatch', '1', '--epochs', '2', '--seed', '7']) == 0
    een, eet, edn, edt, d2t = read_params(params)
    assert os.path.exists(params + '.iter0') and os.path.exists(params + '.iter1')
    # the same run on the oracle: same shuffle, same roots (train.py's per-sentence SGD, minibatch 1)
    sents = [synth.sentence_to_arrays(r) for r in raw]
    random.seed(7)
    rng = random.Random(8)
    order = list(range(len(sents)))
    te, td = np.zeros((1, 3)), np.zeros((1, 6))
    reg = 0.2 / len(sents)
    for epoch in range(2):
        lr = 0.1 / (1.0 + 0.3 * epoch)
        random.shuffle(order)
        for i in order:
            c = Corpus([sents[i]])
            roots_local = train_cli.draw_roots(c, 3, rng)[0]
            roots_pos = [int(c.var_pos[r]) for r in roots_local]
            r = orc.run_fast(orc.Tables(model, te, td), sents[i], roots_pos, 3, reg, lr)
            te, td = te + r['g_ee_ret'], td + r['g_ed_ret']
    np.testing.assert_allclose(eet, te, atol=2e-6)           # checkpoints carry 6 decimals (train.py:80)
    np.testing.assert_allclose(edt, td, atol=2e-6)
    # prediction mode writes the files eval.py / get_corr.py parse
    pred = str(tmp_path / 'out.pred')
    assert train_cli.main(base + ['--load_params', params, '--save_predictions', pred]) == 0
    text = codecs.open(pred, 'r', 'utf8').read().split('\n')
    assert sum(1 for l in text if l.startswith('*SENT_ID:')) == len(raw)
    first = [l for l in text if l and not l.startswith('*SENT_ID:') and not l.startswith(' ')][0].split(' ')
    assert first[0].startswith('d') and first[1].startswith('e') and len(first) == 3 + 2 * 50
    dist = codecs.open(pred + '.dist', 'r', 'utf8').read().strip().split('\n')
    assert len(dist) == sum(l.count('p') for l in layouts)
    assert len(dist[0].split(' ||| ')[2].split(' ')) == 120


def test_cli_user_adapt_train_then_predict(tmp_path):
    """--user_adapt through the command line: per-user theta lines in <params>.user_adapt (train.py:46-99 format),
    minibatch 1 == train.py's per-sentence trajectory (checked on the oracle, itself pinned on a reference fixture)"""
    build.build()
    model = synth.make_model(120, 24, seed=43)
    layouts = ['pppp', 'gpgpp', 'ppgp', 'pp', 'pgppg', 'ppp']
    users = ['ua', 'ub']
    raw = [synth.make_sentence(model, l, seed=600 + i, n_history=3, sent_id=i, user_id=users[i % 2]) for i, l in enumerate(layouts)]
    p = write_inputs(tmp_path, model, raw)
    with codecs.open(p['ti'] + '.users', 'w', 'utf8') as f:
        f.write('\n'.join(users) + '\n')
    params = str(tmp_path / 'model.params')
    base = ['--ti', p['ti'], '--end', p['en'], '--ded', p['de'], '--phi_pmi', p['pmi'], '--phi_pmi_w1', p['pmi_w1'],
            '--phi_ed', p['ed'], '--phi_ped', p['ped'], '--history', '--session_history', '--user_adapt',
            '--reg_param_ua_scale', '0.5']
    assert train_cli.main(base + ['--save_params', params, '--minibatch', '1', '--epochs', '2', '--seed', '7']) == 0
    assert os.path.exists(params + '.user_adapt.iter0') and os.path.exists(params + '.user_adapt.iter1')
    een, eet, edn, edt, d2t = read_params(params + '.user_adapt')
    assert sorted(d2t.keys()) == [('en_de', 'ua'), ('en_de', 'ub'), ('en_en', 'ua'), ('en_en', 'ub')]
    sents = [synth.sentence_to_arrays(r) for r in raw]
    random.seed(7)
    rng = random.Random(8)
    order = list(range(len(sents)))
    te, td = np.zeros((1, 3)), np.zeros((1, 6))
    ut = {u: (np.zeros((1, 3)), np.zeros((1, 6))) for u in users}
    reg = 0.2 / len(sents)
    for epoch in range(2):
        lr = 0.1 / (1.0 + 0.3 * epoch)
        random.shuffle(order)
        for i in order:
            c = Corpus([sents[i]])
            roots_pos = [int(c.var_pos[r]) for r in train_cli.draw_roots(c, 3, rng)[0]]
            ue, ud = ut[raw[i]['user_id']]
            r = orc.run_fast(orc.Tables(model, ue, ud), sents[i], roots_pos, 3)
            g_ee, g_ed = r['g_ee_unreg'], r['g_ed_unreg']
            ut[raw[i]['user_id']] = (ue + lr * (g_ee - reg * 0.5 * ue), ud + lr * (g_ed - reg * 0.5 * ud))
            te, td = te + lr * (g_ee - reg * te), td + lr * (g_ed - reg * td)
    np.testing.assert_allclose(eet, te, atol=2e-6)
    np.testing.assert_allclose(edt, td, atol=2e-6)
    for u in users:
        np.testing.assert_allclose(d2t['en_en', u], ut[u][0], atol=2e-6)
        np.testing.assert_allclose(d2t['en_de', u], ut[u][1], atol=2e-6)
    pred = str(tmp_path / 'out.pred')
    assert train_cli.main(base + ['--load_params', params, '--save_predictions', pred]) == 0
    text = codecs.open(pred, 'r', 'utf8').read().split('\n')
    assert [l for l in text if l.startswith('*SENT_ID:')] == ['*SENT_ID:%d' % i for i in range(len(raw))]   # file order kept
