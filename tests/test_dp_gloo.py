"""N > 1 host logic on the CPU: two gloo ranks shard the sentences (like train_mp.py hands them to pool workers,
train_mp.py:634-649), all-reduce the 16-float64 gradient vector and must end at the single-process theta."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _setup():
    from fake_kernels import FakeKernels
    from macaronicusermodeling_b200 import synth
    from macaronicusermodeling_b200.engine import Corpus, Engine
    from macaronicusermodeling_b200.trainer import Trainer
    model = synth.make_model(64, 16, seed=2)
    sents = synth.make_corpus(model, 8, k=4, g=1, seed=4)
    roots = Corpus(sents).roots_from_positions(synth.draw_roots(sents, 3, seed=6))
    return FakeKernels, Corpus, Engine, Trainer, model, sents, roots


def _one_step(Trainer, Engine, Corpus, FakeKernels, model, sents, roots, N):
    tr = Trainer(Engine(model, kernels=FakeKernels()), reg_param=0.2, N=N)
    tr.theta_ee = np.array([0.3, 0.2, -0.1]); tr.theta_ed = np.array([0.5, -0.2, 0.3, 0.1, 0.2, 0.0])
    red = tr.step(Corpus(sents), roots, 0.01)
    h = tr.apply(red, 0.01)
    return tr, h


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    FakeKernels, Corpus, Engine, Trainer, model, sents, roots = _setup()
    lo, hi = rank * len(sents) // world, (rank + 1) * len(sents) // world
    tr, h = _one_step(Trainer, Engine, Corpus, FakeKernels, model, sents[lo:hi], roots[lo:hi], len(sents))
    out[rank] = np.concatenate([tr.theta_ee, tr.theta_ed, h]).tolist()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process():
    from macaronicusermodeling_b200 import build
    build.build()
    FakeKernels, Corpus, Engine, Trainer, model, sents, roots = _setup()
    tr, h = _one_step(Trainer, Engine, Corpus, FakeKernels, model, sents, roots, len(sents))
    want = np.concatenate([tr.theta_ee, tr.theta_ed, h])
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for r in (0, 1):
        np.testing.assert_allclose(np.array(out[r]), want, rtol=1e-10, atol=1e-12)
    assert want[9 + 14] == len(sents)


# ---------------------------------------------------------------------------------------------- per-user theta (C4)
def _adapt_setup():
    from fake_kernels import FakeKernels
    from macaronicusermodeling_b200 import synth
    from macaronicusermodeling_b200.engine import Corpus, Engine
    from macaronicusermodeling_b200.trainer import AdaptTrainer
    model = synth.make_model(64, 16, seed=3)
    users = ['ua', 'ub', 'uc', 'ud']
    per_user = {u: synth.make_corpus(model, 3, k=4, g=1, seed=20 + i, users=[u]) for i, u in enumerate(users)}
    return FakeKernels, Corpus, Engine, AdaptTrainer, model, users, per_user


def _adapt_step(AdaptTrainer, Engine, Corpus, FakeKernels, model, users_all, mine, per_user):
    from macaronicusermodeling_b200 import synth
    tr = AdaptTrainer(Engine(model, kernels=FakeKernels()), mine, reg_param=0.2, ua_scale=0.5, N=3 * len(users_all))
    tr.theta_ee = np.array([0.3, 0.2, -0.1]); tr.theta_ed = np.array([0.5, -0.2, 0.3, 0.1, 0.2, 0.0])
    for i, u in enumerate(mine):
        tr.domain2theta[u] = (tr.theta_ee * (1.0 + 0.1 * users_all.index(u)), tr.theta_ed.copy())
    batches = []
    for u in mine:
        c = Corpus(per_user[u])
        batches.append((u, c, c.roots_from_positions(synth.draw_roots(per_user[u], 3, seed=40 + users_all.index(u)))))
    red = tr.step_domains(batches, 0.01)
    tr.apply(red, 0.01)
    return tr


def _adapt_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    FakeKernels, Corpus, Engine, AdaptTrainer, model, users, per_user = _adapt_setup()
    mine = users[rank::world]                               # users are the unit of sharding; their theta never leaves the rank
    tr = _adapt_step(AdaptTrainer, Engine, Corpus, FakeKernels, model, users, mine, per_user)
    out[rank] = {'base': np.concatenate([tr.theta_ee, tr.theta_ed]).tolist(),
                 'users': {u: np.concatenate(tr.domain2theta[u]).tolist() for u in mine}}
    dist.destroy_process_group()


def test_two_rank_user_adapt_equals_single_process():
    """C4 with --user_adapt on two ranks: the base theta (all-reduced gradient) and every user's theta (rank-local) end where
    a single process that owns all users ends"""
    from macaronicusermodeling_b200 import build
    build.build()
    FakeKernels, Corpus, Engine, AdaptTrainer, model, users, per_user = _adapt_setup()
    tr = _adapt_step(AdaptTrainer, Engine, Corpus, FakeKernels, model, users, users, per_user)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_adapt_worker, args=(2, port, out), nprocs=2, join=True)
    want_base = np.concatenate([tr.theta_ee, tr.theta_ed])
    for r in (0, 1):
        np.testing.assert_allclose(np.array(out[r]['base']), want_base, rtol=1e-10, atol=1e-12)
        for u, th in out[r]['users'].items():
            np.testing.assert_allclose(np.array(th), np.concatenate(tr.domain2theta[u]), rtol=1e-10, atol=1e-12)
    assert sorted(list(out[0]['users']) + list(out[1]['users'])) == users
