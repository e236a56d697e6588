#!/usr/bin/env python
"""BASELINE config C4: train_mp.py-style multi-user training -- 512 synthetic users x 100 sentences (V = 10 000, k = 20,
3 sweeps), users sharded over the GPUs, one 16 x f64 all-reduce of the base-theta gradient per step.

    python scripts/c4_multi_user.py                                   # one GPU: this rank's share of the users
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/c4_multi_user.py --gpus 8

Two modes: shared theta (train_mp.py: one model, sentences of all users in one minibatch per rank) and --user_adapt
(train.py:160-173, :224-245: every user owns a theta pair that builds its own potential tables; the per-user theta never
leaves its rank).  A step = one pass over every user's 100 sentences (an epoch of the config).  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macaronicusermodeling_b200 import synth  # noqa: E402
from macaronicusermodeling_b200.engine import Corpus, Engine  # noqa: E402
from macaronicusermodeling_b200.trainer import AdaptTrainer, Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--users', type=int, default=512)
    ap.add_argument('--sentences-per-user', type=int, default=100)
    ap.add_argument('--V', type=int, default=10000)
    ap.add_argument('--Vd', type=int, default=2000)
    ap.add_argument('--k', type=int, default=20)
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--user_adapt', action='store_true')
    ap.add_argument('--workspace-gb', type=float, default=96.0)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        os.environ.setdefault('MLBP_PLAN_THREADS', str(max(2, min(16, (os.cpu_count() or 8) // world))))
    users = ['user%03d' % u for u in range(a.users)]
    mine = users[rank::world]                                      # users are the unit of sharding (train_mp.py hands out sentences)
    model = synth.make_model(a.V, a.Vd, seed=1234, dtype=np.float32)
    per_user = {u: synth.make_corpus(model, a.sentences_per_user, k=a.k, g=0, seed=1000 + 17 * users.index(u), users=[u]) for u in mine}
    eng = Engine(model, workspace_bytes=int(a.workspace_gb * (1 << 30)))
    N = a.users * a.sentences_per_user
    rng = np.random.default_rng(7 + rank)
    roots_of = lambda c: (rng.random((c.n_sent, 4)) * np.diff(c.var_off)[:, None]).astype(np.int32)
    te0, td0 = np.array([0.8, 0.5, -0.3]), np.array([1.0, -0.6, 0.5, 0.3, 0.4, -0.2])
    if a.user_adapt:
        tr = AdaptTrainer(eng, mine, reg_param=0.2, ua_scale=0.5, N=N)
        for u in mine:
            tr.domain2theta[u] = (te0.copy(), td0.copy())
        batches = [(u, Corpus(per_user[u]), None) for u in mine]
    else:
        tr = Trainer(eng, reg_param=0.2, N=N)
        corpus = Corpus([s for u in mine for s in per_user[u]])
        parts = eng.prepare(corpus, 3, True)
    tr.theta_ee, tr.theta_ed = te0.copy(), td0.copy()
    lr = 0.01 / N                                                   # small steps: theta stays in the benchmark's regime

    def epoch():
        if a.user_adapt:
            red = tr.step_domains([(u, c, roots_of(c)) for u, c, _ in batches], lr)
        else:
            red = tr.step(parts, roots_of(corpus), lr)
        return tr.apply(red, lr)

    epoch()                                                        # warm-up (allocations, tensor maps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.epochs):
        h = epoch()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) / a.epochs
        print(json.dumps({'config': 'C4 multi-user training: %d users x %d sentences, V=%d, k=%d, 3 sweeps, %s' % (
            a.users, a.sentences_per_user, a.V, a.k, 'per-user theta (--user_adapt)' if a.user_adapt else 'shared theta'),
            'n_gpus': world, 'users_on_rank0': len(mine), 'sentences_per_s': N * (len(mine) * world / a.users) / (ms / 1e3)
            if len(mine) * world != a.users else N / (ms / 1e3), 'ms_per_epoch': ms, 'mean_logp_per_sentence': float(h[9] / h[14])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
