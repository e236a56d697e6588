"""Command-line trainer / predictor with the reference's flags (train.py:437-462, train_mp.py:457,463) on the batched
B200 engine -- the real-data front end: TrainingInstance JSON lines, vocabulary files, np.loadtxt feature matrices
(train.py:551-610), text parameter checkpoints (train.py:46-99) and the prediction / .dist files eval.py and get_corr.py
read (LBP.py:109-143).

    python -m macaronicusermodeling_b200.train_cli --ti train.json --end en.vocab --ded de.vocab \
        --phi_pmi pmi.mat --phi_pmi_w1 pmi_w1.mat --phi_ed ed.mat --phi_ped ped.mat --save_params model.params
    python -m macaronicusermodeling_b200.train_cli --ti test.json ... --load_params model.params --save_predictions out.pred

Where train_mp.py hands sentences to `--cpu N` pool workers that each see a theta about N sentences old
(train_mp.py:634-649), this trainer takes synchronous minibatches of `--minibatch` sentences (default: the value of
--cpu, 4) that all see the same theta; `--minibatch 1` is train.py's exact per-sentence SGD.  Under torchrun the
minibatch is sharded over the ranks and the gradient all-reduced (trainer.Trainer).
"""
import argparse
import codecs
import json
import random
import sys

import numpy as np

from . import synth
from .train_compat import F_EN_DE_NAMES, F_EN_EN_NAMES, read_params, save_params


def parse(argv=None):
    opt = argparse.ArgumentParser(description='batched B200 trainer for the macaronic user model')
    opt.add_argument('--ti', dest='training_instances', default='')
    opt.add_argument('--tune', dest='tuning_instances', default='')
    opt.add_argument('--end', dest='en_domain', default='')
    opt.add_argument('--ded', dest='de_domain', default='')
    opt.add_argument('--phi_pmi', dest='phi_pmi', default='')
    opt.add_argument('--phi_pmi_w1', dest='phi_pmi_w1', default='')
    opt.add_argument('--phi_ed', dest='phi_ed', default='')
    opt.add_argument('--phi_ped', dest='phi_ped', default='')
    opt.add_argument('--cpu', dest='cpus', default='')
    opt.add_argument('--minibatch', dest='minibatch', type=int, default=0)
    opt.add_argument('--epochs', dest='epochs', type=int, default=3)                  # train_mp.py:629
    opt.add_argument('--reg_param', dest='reg_param', default='0.2')
    opt.add_argument('--reg_param_ua_scale', dest='reg_param_ua_scale', default='1.0')
    opt.add_argument('--save_params', dest='save_params_file', default='')
    opt.add_argument('--load_params', dest='load_params_file', default='')
    opt.add_argument('--save_predictions', dest='save_predictions_file', default='')
    opt.add_argument('--history', dest='history', default=False, action='store_true')  # train_mp.py:463
    opt.add_argument('--session_history', dest='session_history', default=False, action='store_true')
    opt.add_argument('--user_adapt', dest='user_adapt', default=False, action='store_true')
    opt.add_argument('--experience_adapt', dest='experience_adapt', default=False, action='store_true')
    opt.add_argument('--quick_predict', dest='quick_predict', default=False, action='store_true')
    opt.add_argument('--use_approx_inference', dest='use_approx_inference', default=False, action='store_true')
    opt.add_argument('--use_approx_beliefs', dest='use_approx_beliefs', default=False, action='store_true')
    opt.add_argument('--report_times', dest='report_times', default=False, action='store_true')
    opt.add_argument('--use_correct_feat', dest='use_correct_feat', default=True, action='store_true')
    opt.add_argument('--seed', dest='seed', type=int, default=1234)                   # train.py:436
    return opt.parse_args(argv)


def gather_domain_thetas(tr, owner, rank, world):
    """Domain thetas never leave the rank that owns the domain during training; before a checkpoint is written rank 0
    collects every owner's current values (no-op on one rank)."""
    if world <= 1:
        return
    import torch.distributed as dist
    mine = dict((d, t) for d, t in tr.domain2theta.items() if owner.get(d, 0) == rank)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    for part in parts:
        tr.domain2theta.update(part)


def error_msg():
    sys.stderr.write('Usage: python -m macaronicusermodeling_b200.train_cli\n  --ti [training instance file]\n  --tune [dev set]\n'
                     '  --end [en domain file]\n  --ded [de domain file]\n  --phi_pmi [pmi file]\n  --phi_pmi_w1 [pmi w1 file]\n'
                     '  --phi_ed [ed file]\n  --phi_ped [ped file]\n  --save_params [file] or --load_params [file] --save_predictions [file]\n')


def load_inputs(options):
    de_domain = [i.strip() for i in codecs.open(options.de_domain, 'r', 'utf8').readlines()]       # train.py:582-585
    en_domain = [i.strip() for i in codecs.open(options.en_domain, 'r', 'utf8').readlines()]
    en2id = dict((e, idx) for idx, e in enumerate(en_domain))
    de2id = dict((d, idx) for idx, d in enumerate(de_domain))
    model = {'pmi': np.loadtxt(options.phi_pmi, dtype=np.float64, ndmin=2),                        # train.py:589-603
             'pmi_w1': np.loadtxt(options.phi_pmi_w1, dtype=np.float64, ndmin=2),
             'ed': np.loadtxt(options.phi_ed, dtype=np.float64, ndmin=2),
             'ped': np.loadtxt(options.phi_ped, dtype=np.float64, ndmin=2)}
    model['V'], model['Vd'] = len(en_domain), len(de_domain)
    return en_domain, de_domain, en2id, de2id, model


def lower(lines, options, en2id, de2id):
    return [synth.sentence_to_arrays(json.loads(l), en2id, de2id, history=options.history,
                                     session_history=options.session_history, use_correct_feat=options.use_correct_feat)
            for l in lines if l.strip()]


def domain_of(raw, options):
    """adaptation domain of a training instance: the user (train.py:160-164) or the experience level (:165-169)"""
    if options.user_adapt:
        return str(raw['user_id'])
    return str(len(raw.get('past_sentences_seen', [])))


def read_domains(options):
    """train.py:488-505: <ti>.users / <ti>.experience list the adaptation domains"""
    ext = '.users' if options.user_adapt else '.experience'
    try:
        return [d.strip() for d in codecs.open(options.training_instances + ext).readlines() if d.strip()]
    except IOError:
        sys.stderr.write('%s option can not find %s file. (looked for: %s)\n'
                         % ('user_adapt' if options.user_adapt else 'experience_adapt', ext, options.training_instances + ext))
        raise SystemExit(1)


def adapt_ext(options):
    return '.user_adapt' if options.user_adapt else ('.exp_adapt' if options.experience_adapt else '')


def read_params_adapt(path, options):
    """train.py:524-531: try <file><ext>, then <file>"""
    for cand in (path + adapt_ext(options), path):
        try:
            print('trying to read:', cand)
            return read_params(cand)
        except IOError:
            continue
    raise IOError('no parameter file: ' + path)


def d2t_arrays(tr):
    """AdaptTrainer.domain2theta -> the reference's {('en_en', d): (1,3), ('en_de', d): (1,6)} dict (save_params)"""
    d2t = {}
    for d, (te, td) in tr.domain2theta.items():
        d2t['en_en', d] = np.asarray(te, dtype=np.float64).reshape(1, -1)
        d2t['en_de', d] = np.asarray(td, dtype=np.float64).reshape(1, -1)
    return d2t


def draw_roots(corpus, sweeps, rng):
    k = np.diff(corpus.var_off)
    return np.array([[rng.randrange(int(kk)) for _ in range(1 + sweeps)] for kk in k], dtype=np.int32)


def predict(engine, sents, raw, en_domain, de_domain, options, rng, qp, batch=64):
    """train.py:308-338 batched: returns (sum log-posterior, prediction strings, dist strings, (p0, p25, p50, total))"""
    from .engine import Corpus
    logp_sum, preds, dists = 0.0, [], []
    p0 = p25 = p50 = tot = 0
    V = len(en_domain)
    for lo in range(0, len(sents), batch):
        chunk = sents[lo:lo + batch]
        corpus = Corpus(chunk)
        r = engine.run(corpus, draw_roots(corpus, 3, rng), 3, want_grad=False, want_marg=True, want_beliefs=not qp,
                       approx_inference=options.use_approx_inference, want_topk=0 if qp else min(50, V - 1))
        logp_sum += float(r.logp.sum().item())
        rank = r.rank.cpu().numpy()
        p0 += int((rank == 0).sum()); p25 += int((rank < 26).sum()); p50 += int((rank < 50).sum()); tot += len(rank)
        if qp:
            continue
        B = r.beliefs.cpu().numpy()[:, :V].astype(np.float64)
        TI, _, TT = (x.cpu().numpy() for x in r.topk)        # the 50 best words per variable, listed on the device (mlbp_topk_rows)
        for si, s in enumerate(chunk):
            ti = raw[lo + si]
            nodes = sorted(ti['current_sent'], key=lambda n: int(n['position']))
            lines, dl = ['*SENT_ID:' + str(nodes[0]['sent_id'])], []
            vi = int(corpus.var_off[si])
            for p, n in enumerate(nodes):                                            # FactorGraph.to_string, LBP.py:109-123
                if s.kind[p] == 1:
                    b = B[vi]
                    vi += 1
                    top = min(50, V - 1)
                    if TT[vi - 1] == 0:
                        idx = TI[vi - 1]
                    else:                                    # exact ties: the reference's order is NumPy's (LBP.py:405-406)
                        idx = np.argpartition(b, -top)[-top:]
                        idx = idx[np.argsort(b[idx])][::-1]
                    with np.errstate(divide='ignore'):
                        lb = np.log(b)
                    sl = en_domain[int(s.label[p])]
                    pred = ' '.join(en_domain[i] + ' ' + '%0.4f' % lb[i] for i in idx)
                    lines.append(' '.join([n['l2_word'], sl, '%0.4f' % lb[int(s.label[p])], pred]))
                    truth = n['l1_parent'].lower().replace("'", "") if n.get('l1_parent') else 'None'
                    dl.append(' ||| '.join([truth, sl, ' '.join('%0.6f' % x for x in lb)]))   # to_dist, LBP.py:125-143
                else:
                    lines.append(' '.join(['', en_domain[int(s.label[p])], '']))
            preds.append('\n'.join(lines))
            dists.append('\n'.join(dl))
    return logp_sum, preds, dists, (p0, p25, p50, tot)


def main(argv=None):
    options = parse(argv)
    need = [options.training_instances, options.en_domain, options.de_domain, options.phi_pmi_w1, options.phi_pmi,
            options.phi_ed, options.phi_ped]
    if '' in need or (options.save_params_file == '' and options.load_params_file == '' and options.save_predictions_file == ''):
        error_msg()
        return 1
    if options.user_adapt and options.experience_adapt:
        sys.stderr.write('Currently only supports 1 type of adaptation.')                  # train.py:489-491
        return 1
    adapt = options.user_adapt or options.experience_adapt
    import os
    import torch
    from .engine import Corpus, Engine
    from .trainer import AdaptTrainer, Trainer, dist_info
    # one process per GPU under torchrun: rank / world from the environment, NCCL over NVLink for the 16 x f64 all-reduce
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
    # the reference trains AND predicts with these flags (train.py:155-156 -> LBP.py:506-516, :554-563)
    approx_kw = dict(approx_inference=options.use_approx_inference, approx_beliefs=options.use_approx_beliefs)
    random.seed(options.seed)
    rng = random.Random(options.seed + 1)
    en_domain, de_domain, en2id, de2id, model = load_inputs(options)
    print(len(en_domain), len(de_domain))
    train_lines = [l for l in codecs.open(options.training_instances, 'r', 'utf8').readlines() if l.strip()]
    tune_lines = ([l for l in codecs.open(options.tuning_instances, 'r', 'utf8').readlines() if l.strip()]
                  if options.tuning_instances else None)
    engine = Engine(model)
    mode = 'training' if options.save_predictions_file == '' else 'predicting'
    f_en_en_names, f_en_de_names = list(F_EN_EN_NAMES), list(F_EN_DE_NAMES)
    theta_ee, theta_ed = np.zeros((1, 3)), np.zeros((1, 6))
    d2t_loaded = {}
    if options.load_params_file:
        f_en_en_names, theta_ee, f_en_de_names, theta_ed, d2t_loaded = read_params_adapt(options.load_params_file, options)
    rank, world = dist_info()
    if mode == 'training' and adapt:
        # --user_adapt / --experience_adapt (train.py:160-173, :224-245, :379-390, :402-409): the sentences of a minibatch are
        # grouped by domain; each domain's theta builds its own potentials and stays on the rank that owns the domain
        domains = read_domains(options)
        raws = [json.loads(l) for l in train_lines]
        sents = lower(train_lines, options, en2id, de2id)
        doms = [domain_of(r, options) for r in raws]
        for d in doms:
            if d not in domains:
                domains.append(d)
        owner = dict((d, i % world) for i, d in enumerate(domains))
        mb = options.minibatch or (4 if options.cpus.strip() == '' else int(options.cpus))
        tr = AdaptTrainer(engine, domains, reg_param=float(options.reg_param), ua_scale=float(options.reg_param_ua_scale),
                          N=len(train_lines))
        tr.theta_ee, tr.theta_ed = theta_ee.reshape(-1).copy(), theta_ed.reshape(-1).copy()
        for (ft, d), t in d2t_loaded.items():
            te, td = tr.domain2theta.get(d, (np.zeros(3), np.zeros(6)))
            tr.domain2theta[d] = (t.reshape(-1).copy(), td) if ft == 'en_en' else (te, t.reshape(-1).copy())
        order = list(range(len(sents)))
        ext = adapt_ext(options)
        for epoch in range(options.epochs):
            lr = tr.lr(epoch)
            random.shuffle(order)
            logp = 0.0
            for lo in range(0, len(order), mb):
                groups = {}
                for i in order[lo:lo + mb]:
                    if owner[doms[i]] == rank:
                        groups.setdefault(doms[i], []).append(sents[i])
                batches = []
                for d in sorted(groups):
                    c = Corpus(groups[d])
                    batches.append((d, c, draw_roots(c, 3, rng)))
                red = tr.step_domains(batches, lr, **approx_kw)
                logp += tr.apply(red, lr)[9]
            print('\nepoch:', epoch)
            print(f_en_en_names, tr.theta_ee)
            print(f_en_de_names, tr.theta_ed)
            print('\ntrain prediction probs:', logp / float(len(sents)))
            if options.save_params_file:
                gather_domain_thetas(tr, owner, rank, world)              # a domain's theta lives on the rank that owns it
            if rank == 0 and options.save_params_file:
                save_params(codecs.open(options.save_params_file + ext + '.iter' + str(epoch), 'w', 'utf8'),
                            tr.theta_ee.reshape(1, -1), tr.theta_ed.reshape(1, -1), f_en_en_names, f_en_de_names, d2t_arrays(tr))
                print('saved params')
        print('\ntheta final:', tr.theta_ee, tr.theta_ed)
        if options.save_params_file:
            gather_domain_thetas(tr, owner, rank, world)
        if rank == 0 and options.save_params_file:
            save_params(codecs.open(options.save_params_file + ext, 'w', 'utf8'), tr.theta_ee.reshape(1, -1),
                        tr.theta_ed.reshape(1, -1), f_en_en_names, f_en_de_names, d2t_arrays(tr))
        return 0
    if mode == 'training':
        mb = options.minibatch or (4 if options.cpus.strip() == '' else int(options.cpus))     # train_mp.py:493
        tr = Trainer(engine, reg_param=float(options.reg_param), N=len(train_lines))
        tr.theta_ee, tr.theta_ed = theta_ee.reshape(-1).copy(), theta_ed.reshape(-1).copy()
        sents = lower(train_lines, options, en2id, de2id)
        order = list(range(len(sents)))
        for epoch in range(options.epochs):
            lr = tr.lr(epoch)
            random.shuffle(order)                                                          # train.py:624
            logp = 0.0
            for lo in range(0, len(order), mb):
                batch = [sents[i] for i in order[lo:lo + mb]][rank::world]
                corpus = Corpus(batch)                  # an empty shard contributes a zero vector to the all-reduce
                red = tr.step(corpus, draw_roots(corpus, 3, rng), lr, **approx_kw)
                logp += tr.apply(red, lr)[9]
            print('\nepoch:', epoch)
            print(f_en_en_names, tr.theta_ee)
            print(f_en_de_names, tr.theta_ed)
            print('\ntrain prediction probs:', logp / float(len(sents)))
            if rank == 0 and options.save_params_file:
                save_params(codecs.open(options.save_params_file + '.iter' + str(epoch), 'w', 'utf8'), tr.theta_ee.reshape(1, -1),
                            tr.theta_ed.reshape(1, -1), f_en_en_names, f_en_de_names, {})
                print('saved params')
            if tune_lines is not None:
                engine.set_theta(tr.theta_ee, tr.theta_ed, with_grad=True)
                lp, _, _, (p0, p25, p50, tot) = predict(engine, lower(tune_lines, options, en2id, de2id),
                                                        [json.loads(l) for l in tune_lines], en_domain, de_domain, options, rng, True)
                print('\ntune prediction probs:', lp / float(len(tune_lines)))
                for name, p in (('Prec at 0:', p0), ('prec at 25:', p25), ('prec at 50:', p50)):
                    print(name, '%0.2f' % (float(100 * p) / float(max(tot, 1))), 'total:', tot)
        print('\ntheta final:', tr.theta_ee, tr.theta_ed)
        if rank == 0 and options.save_params_file:
            save_params(codecs.open(options.save_params_file, 'w', 'utf8'), tr.theta_ee.reshape(1, -1), tr.theta_ed.reshape(1, -1),
                        f_en_en_names, f_en_de_names, {})
        return 0
    # ---- predicting (train.py:678-757)
    raw = [json.loads(l) for l in train_lines]
    all_sents = lower(train_lines, options, en2id, de2id)
    if adapt:
        # every sentence is scored with its domain's theta (train.py:224-245 via create_factor_graph); results keep file order
        doms = [domain_of(r, options) for r in raw]
        lp, p0, p25, p50, tot = 0.0, 0, 0, 0, 0
        preds, dists = [None] * len(raw), [None] * len(raw)
        for d in sorted(set(doms)):
            idx = [i for i, x in enumerate(doms) if x == d]
            engine.set_theta(d2t_loaded['en_en', d].reshape(-1), d2t_loaded['en_de', d].reshape(-1), with_grad=True)
            l_, pr, di, (a0, a25, a50, at) = predict(engine, [all_sents[i] for i in idx], [raw[i] for i in idx], en_domain,
                                                     de_domain, options, rng, options.quick_predict)
            lp += l_; p0 += a0; p25 += a25; p50 += a50; tot += at
            for j, i in enumerate(idx):
                if pr:
                    preds[i], dists[i] = pr[j], di[j]
    else:
        engine.set_theta(theta_ee.reshape(-1), theta_ed.reshape(-1), with_grad=True)
        lp, preds, dists, (p0, p25, p50, tot) = predict(engine, all_sents, raw, en_domain, de_domain, options, rng,
                                                         options.quick_predict)
    if not options.quick_predict:
        with codecs.open(options.save_predictions_file, 'w', 'utf8') as w, \
                codecs.open(options.save_predictions_file + '.dist', 'w', 'utf8') as wd:
            for p, d in zip(preds, dists):
                w.write(p + '\n')
                wd.write(d + '\n')
    print('\nprediction probs:', lp / float(len(train_lines)))
    for name, p in (('Prec at 0:', p0), ('prec at 25:', p25), ('prec at 50:', p50)):
        print(name, '%0.2f' % (float(100 * p) / float(max(tot, 1))), 'total:', tot)
    torch.cuda.synchronize()
    return 0


if __name__ == '__main__':
    sys.exit(main())
