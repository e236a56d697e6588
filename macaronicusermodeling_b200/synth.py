"""Synthetic macaronic sentences and feature planes (SURVEY.md §8(d) generator).

The reference ships no data (README:3 points at an external download), so every test, fixture and
benchmark in this repo uses seeded synthetic inputs of the reference's own shapes:

* feature planes  PMI, PMI_w1 in [0,1]^{V x V} and ED, PED in [0,1]^{V x Vd}  (the real files are
  ``*.scaled.01`` matrices read by np.loadtxt, train.py:589-603),
* sentences in the reference's JSON wire format (training_classes.py:30-39, :141-147, :176-183):
  one ``TrainingInstance`` dict per sentence.

``sentence_to_arrays`` is the host-side front end of the hot path: it turns one such JSON sentence
into the integer arrays the batched engine consumes, restating train.py:102-130 (which positions are
GIVEN / PREDICTED) and train.py:176-215 (the per-sentence sparse ``correct`` / ``full_history`` /
``hit_history`` features) without ever materialising a (V, Vd) plane.
"""
import json

import numpy as np

EN_DE_NAMES = ['ed', 'ped', 'correct', 'full_history', 'hit_history', 'bias']   # train.py:513
EN_EN_NAMES = ['pmi', 'pmi_w1', 'bias']                                         # train.py:510
F_CORRECT, F_FULL_HISTORY, F_HIT_HISTORY = 2, 3, 4

KIND_GIVEN, KIND_PREDICTED = 0, 1


def make_model(V, Vd, seed=1234, w1_density=1.0, dtype=np.float64, pmi_density=1.0):
    """iid U[0,1] feature planes; PMI and PMI_w1 optionally sparsified (real PMI matrices are mostly zeros: only word
    pairs that co-occur carry a value; the ``*.scaled.01`` files of run-training.sh:10 keep the zeros)."""
    rng = np.random.default_rng(seed)
    pmi = rng.random((V, V)).astype(dtype)
    pmi_w1 = rng.random((V, V)).astype(dtype)
    if w1_density < 1.0:
        pmi_w1 *= (rng.random((V, V)) < w1_density)
    if pmi_density < 1.0:
        pmi *= (np.random.default_rng(seed + 7).random((V, V)) < pmi_density)
    ed = rng.random((V, Vd)).astype(dtype)
    ped = rng.random((V, Vd)).astype(dtype)
    return {'V': V, 'Vd': Vd, 'pmi': pmi, 'pmi_w1': pmi_w1, 'ed': ed, 'ped': ped}


def make_model_large(V, Vd, seed=1234, block=4096):
    """float32 feature planes generated in row blocks (BASELINE config C5: V = 50 000 -> 10 GB per plane; a float64
    intermediate of the whole plane would not be affordable)"""
    rng = np.random.default_rng(seed)

    def plane(rows, cols):
        out = np.empty((rows, cols), dtype=np.float32)
        for r0 in range(0, rows, block):
            out[r0:r0 + block] = rng.random((min(block, rows - r0), cols), dtype=np.float32)
        return out
    return {'V': V, 'Vd': Vd, 'pmi': plane(V, V), 'pmi_w1': plane(V, V), 'ed': plane(V, Vd), 'ped': plane(V, Vd)}


def en_word(i):
    return 'e%d' % i


def de_word(i):
    return 'd%d' % i


def make_sentence(model, layout, seed=0, n_history=2, sent_id=0, user_id='u0', p_correct=0.3):
    """layout: string over {'p' predicted German token, 'g' given English token, 'r' revealed German token}."""
    rng = np.random.default_rng(seed)
    V, Vd = model['V'], model['Vd']
    cur, guesses, revealed = [], [], []
    de_here = []
    for pos, k in enumerate(layout):
        nid = [sent_id, pos]
        if k == 'g':
            cur.append({'sent_id': sent_id, 'id': nid, 'l2_word': en_word(int(rng.integers(V))), 'l1_parent': '',
                        'position': pos, 'lang': 'en'})
        else:
            d = int(rng.integers(Vd))
            truth = int(rng.integers(V))
            de_here.append(d)
            cur.append({'sent_id': sent_id, 'id': nid, 'l2_word': de_word(d), 'l1_parent': en_word(truth),
                        'position': pos, 'lang': 'de'})
            if k == 'p':
                g = truth if rng.random() < p_correct else int(rng.integers(V))
                guesses.append({'id': nid, 'guess': en_word(g), 'revealed': False, 'l2_word': de_word(d),
                                'reference': en_word(truth)})
            elif k == 'r':
                revealed.append({'id': nid, 'guess': en_word(truth), 'revealed': True, 'l2_word': de_word(d),
                                 'reference': en_word(truth)})
            else:
                raise ValueError(k)
    past_correct, past_here = [], []
    for _ in range(n_history):
        d = int(rng.choice(de_here)) if (de_here and rng.random() < 0.7) else int(rng.integers(Vd))
        past_correct.append({'id': [sent_id + 1000, 0], 'guess': en_word(int(rng.integers(V))), 'revealed': False,
                             'l2_word': de_word(d)})
        d = int(rng.choice(de_here)) if de_here else int(rng.integers(Vd))
        past_here.append({'id': [sent_id, 0], 'guess': en_word(int(rng.integers(V))),
                          'revealed': bool(rng.random() < 0.25), 'l2_word': de_word(d)})
    return {'user_id': user_id, 'past_correct_guesses': past_correct, 'past_sentences_seen': [],
            'past_guesses_for_current_sent': past_here, 'current_sent': cur,
            'current_revealed_guesses': revealed, 'current_guesses': guesses}


def sentence_to_json(sent):
    return json.dumps(sent)


def _norm_guess(g):
    """training_classes.py:94-110 (Guess.__init__) guess normalisation."""
    s = g.strip()
    if s == '':
        return '__blank__'
    if s.lower() in ('__blank__', '__unk__', '__copy__'):
        return g
    g = sorted([(len(t), t) for t in g.split()])[-1][1]
    g = g[:-1] if g[-1] == '*' and len(g) > 1 else g
    return g.lower().replace("'", "")


class SentenceArrays(object):
    """One sentence lowered to integers.

    kind[p]   KIND_GIVEN / KIND_PREDICTED per position (train.py:107-117)
    label[p]  en index of the supervised label (given word, revealed word or the user's guess)
    de[p]     de index of the observed German word for PREDICTED positions, -1 otherwise
    sparse    (n,4) float64 rows (en index, de index, feature index in EN_DE_NAMES, value) that
              train.py:176-215 would write into the shared dense phi_en_de planes
    """
    __slots__ = ('kind', 'label', 'de', 'sparse', 'sent_id', 'user_id', 'words')

    def __init__(self, kind, label, de, sparse, sent_id=0, user_id=None, words=None):
        self.kind = np.asarray(kind, dtype=np.int32)
        self.label = np.asarray(label, dtype=np.int32)
        self.de = np.asarray(de, dtype=np.int32)
        self.sparse = np.asarray(sparse, dtype=np.float64).reshape(-1, 4)
        self.sent_id = sent_id
        self.user_id = user_id
        self.words = words

    @property
    def predicted(self):
        return np.nonzero(self.kind == KIND_PREDICTED)[0]

    @property
    def given(self):
        return np.nonzero(self.kind == KIND_GIVEN)[0]


def sentence_to_arrays(sent, en2id=None, de2id=None, history=True, session_history=True, use_correct_feat=True):
    """JSON sentence -> SentenceArrays (train.py:102-130 + :176-215).  ``en2id``/``de2id`` default to the
    synthetic 'e<i>' / 'd<i>' vocabularies."""
    if isinstance(sent, str):
        sent = json.loads(sent)
    e2i = (lambda w: int(w[1:])) if en2id is None else (lambda w: en2id[w])
    d2i = (lambda w: int(w[1:])) if de2id is None else (lambda w: de2id[w])
    nodes = sorted(sent['current_sent'], key=lambda n: int(n['position']))      # train.py:139-140
    kind, label, de = [], [], []
    cg = {tuple(g['id']): g for g in reversed(sent['current_guesses'])}          # find_guess: first match wins
    rg = {tuple(g['id']): g for g in reversed(sent['current_revealed_guesses'])}
    for n in nodes:
        nid = tuple(n['id'])
        if n['lang'] == 'en':
            kind.append(KIND_GIVEN)
            label.append(e2i(n['l2_word'].lower().replace("'", "")))            # training_classes.py:161-162
            de.append(-1)
        elif nid in cg:
            kind.append(KIND_PREDICTED)
            label.append(e2i(_norm_guess(cg[nid]['guess'])))
            de.append(d2i(n['l2_word']))
        else:
            kind.append(KIND_GIVEN)                                               # revealed: train.py:113-115
            label.append(e2i(_norm_guess(rg[nid]['guess'])))
            de.append(-1)
    sparse = []
    if use_correct_feat:                                                          # train.py:176-186
        for g in sent['current_guesses']:
            ref = g.get('reference', None)
            if ref is not None and _norm_guess(g['guess']) == ref:
                sparse.append((e2i(_norm_guess(g['guess'])), d2i(g['l2_word']), F_CORRECT, 1.0))
    if history:                                                                   # train.py:190-200
        for g in sent['past_correct_guesses']:
            sparse.append((e2i(_norm_guess(g['guess'])), d2i(g['l2_word']), F_FULL_HISTORY, 1.0))
    if session_history:                                                           # train.py:204-215
        for g in sent['past_guesses_for_current_sent']:
            if not g['revealed']:
                sparse.append((e2i(_norm_guess(g['guess'])), d2i(g['l2_word']), F_HIT_HISTORY, -1.0))
    sid = nodes[0]['sent_id'] if nodes else 0
    return SentenceArrays(kind, label, de, sparse, sent_id=sid, user_id=sent.get('user_id'))


def golden_case_specs():
    """The graph cases tests/golden/make_golden.py runs through the reference (BASELINE config C1 and the
    structural edge cases: tree, single variable, adjacent-only, revealed tokens, peaked potentials, theta=0)."""
    t_ee = [0.7, -0.4, 0.2]
    t_ed = [0.9, -0.6, 0.5, 0.3, 0.4, -0.1]
    return {
        'toy3': dict(V=64, Vd=20, layout='gpppg', sweeps=3, model_seed=1, sent_seed=11, theta_ee=t_ee, theta_ed=t_ed),
        'toy5': dict(V=200, Vd=50, layout='pgpgppgp', sweeps=3, model_seed=2, sent_seed=12, theta_ee=t_ee, theta_ed=t_ed,
                     n_history=4),
        'tree2': dict(V=64, Vd=20, layout='pgp', sweeps=3, model_seed=3, sent_seed=13, theta_ee=t_ee, theta_ed=t_ed),
        'single': dict(V=64, Vd=20, layout='gpg', sweeps=3, model_seed=4, sent_seed=14, theta_ee=t_ee, theta_ed=t_ed),
        'adjacent4': dict(V=96, Vd=24, layout='pppp', sweeps=3, model_seed=5, sent_seed=15, theta_ee=[0.5, 0.8, -0.3],
                          theta_ed=t_ed, w1_density=0.3),
        'revealed': dict(V=72, Vd=20, layout='prpgp', sweeps=3, model_seed=6, sent_seed=16, theta_ee=t_ee, theta_ed=t_ed),
        'peaked': dict(V=80, Vd=20, layout='gpppp', sweeps=3, model_seed=7, sent_seed=17, theta_ee=[4.0, 2.5, -1.0],
                       theta_ed=[3.0, 2.0, 1.5, 1.0, 1.0, 0.5], n_history=5),
        'zeros': dict(V=64, Vd=20, layout='ppgp', sweeps=3, model_seed=8, sent_seed=18, theta_ee=[0.0, 0.0, 0.0],
                      theta_ed=[0.0] * 6),
        'k8': dict(V=128, Vd=32, layout='pppppppp', sweeps=3, model_seed=9, sent_seed=19, theta_ee=t_ee, theta_ed=t_ed),
        'approx_inf': dict(V=160, Vd=24, layout='pgppp', sweeps=3, model_seed=11, sent_seed=21, theta_ee=[2.0, 1.0, -0.5],
                           theta_ed=[2.0, -1.0, 0.5, 0.3, 0.4, -0.1], approx_inference=True),
        'approx_both': dict(V=160, Vd=24, layout='ppgpp', sweeps=3, model_seed=12, sent_seed=22, theta_ee=[2.0, 1.0, -0.5],
                            theta_ed=[2.0, -1.0, 0.5, 0.3, 0.4, -0.1], approx_inference=True, approx_beliefs=True),
        'approx_bel': dict(V=160, Vd=24, layout='pppp', sweeps=3, model_seed=13, sent_seed=23, theta_ee=[2.0, 1.0, -0.5],
                           theta_ed=[2.0, -1.0, 0.5, 0.3, 0.4, -0.1], approx_beliefs=True),
        'sweeps5': dict(V=64, Vd=20, layout='pgppp', sweeps=5, model_seed=10, sent_seed=20, theta_ee=t_ee, theta_ed=t_ed),
    }


def make_corpus(model, n_sentences, k=20, g=0, seed=1234, n_history=2, layouts=None, users=None):
    """n synthetic sentences (list of SentenceArrays).  Default layout: k predicted German tokens (adjacent)
    followed by g given English tokens interleaved from the right (BASELINE configs C2/C3)."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(n_sentences):
        if layouts is not None:
            layout = layouts[s % len(layouts)]
        else:
            lay = ['p'] * k + ['g'] * g
            if g:
                rng.shuffle(lay)
            layout = ''.join(lay)
        uid = None if users is None else users[s % len(users)]
        sent = make_sentence(model, layout, seed=int(rng.integers(1 << 31)), n_history=n_history, sent_id=s,
                             user_id=uid if uid is not None else 'u0')
        out.append(sentence_to_arrays(sent))
    return out


def draw_roots(sentences, sweeps, seed=0):
    """BFS roots: one draw for has_loops (LBP.py:176) + one per sweep (LBP.py:223), each uniform over the
    sentence's variables (= its predicted positions).  Returned as a list of int lists of POSITION ids."""
    rng = np.random.default_rng(seed)
    roots = []
    for s in sentences:
        pred = s.predicted
        roots.append([int(pred[int(rng.integers(len(pred)))]) for _ in range(1 + sweeps)])
    return roots
