"""GPU tier: trainer semantics (train.py:357-416, :617-638) and the bench entry points."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from macaronicusermodeling_b200 import build, synth
from macaronicusermodeling_b200.engine import Corpus, Engine
from macaronicusermodeling_b200.trainer import Trainer, batch_sgd_many

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build()


def test_sgd_trajectory_matches_reference():
    """minibatch = 1 sentence on one GPU reproduces train.py's per-sentence SGD trajectory (2 epochs)"""
    z = np.load(os.path.join(GOLDEN, 'sgd_trajectory.npz'), allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'], 'ed': z['ed'],
             'ped': z['ped']}
    sents = [synth.sentence_to_arrays(str(s)) for s in z['sentences']]
    roots = z['roots'].tolist()
    eng = Engine(model)
    tr = Trainer(eng, reg_param=0.2, N=len(sents), sweeps=3)
    traj, logps = [], []
    for epoch in range(2):
        lr = tr.lr(epoch)
        for si, s in enumerate(sents):
            c = Corpus([s])
            red = tr.step(c, c.roots_from_positions([roots[epoch][si]]), lr)
            h = tr.apply(red, lr)
            logps.append(h[9])
            traj.append(np.concatenate([tr.theta_ee, tr.theta_ed]))
    np.testing.assert_allclose(np.array(traj), z['traj'], rtol=1e-4, atol=2e-7)
    np.testing.assert_allclose(np.array(logps), z['logps'], rtol=1e-5)


def test_batch_sgd_many_returns_reference_result_lists():
    z = np.load(os.path.join(GOLDEN, 'sgd_trajectory.npz'), allow_pickle=False)
    model = {'V': z['pmi'].shape[0], 'Vd': z['ed'].shape[1], 'pmi': z['pmi'], 'pmi_w1': z['pmi_w1'], 'ed': z['ed'],
             'ped': z['ped']}
    sents = [synth.sentence_to_arrays(str(s)) for s in z['sentences']]
    roots = z['roots'].tolist()[0]
    eng = Engine(model)
    res = batch_sgd_many(eng, sents, np.zeros(3), np.zeros(6), 0.1, roots, reg_param=0.2, N=len(sents))
    # sentence 0 at theta = 0 is the first step of the reference trajectory
    np.testing.assert_allclose(np.concatenate([res[0][2][0], res[0][3][0]]), z['traj'][0], rtol=1e-4, atol=2e-7)
    np.testing.assert_allclose(res[0][1], z['logps'][0], rtol=1e-5)
    assert res[0][2].shape == (1, 3) and res[0][3].shape == (1, 6)


def test_bench_line_small():
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '1', '--warmup', '1', '--sentences',
                          '16', '--V', '1024', '--Vd', '128', '--k', '6', '--cpu-sample', '2'], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'ms_per_step', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
        assert key in line
    assert line['gpu_launches'] > 0 and line['value'] > 0 and line['e2e']['value'] > 0
