"""bench.py's GPU arm driven on the CPU: the real Workload / Trainer / Engine / schedule compiler with the NumPy kernel
emulation (tests/fake_kernels.py) and stubbed CUDA events, at toy sizes -- catches host-side bugs of the bench (JSON keys,
config plumbing, the separate profiling pass) before a GPU minute is spent.  The numbers it prints mean nothing."""
import io
import json
import sys
import time
from contextlib import redirect_stdout

import pytest
import torch

import bench
from fake_kernels import FakeKernels
from macaronicusermodeling_b200 import engine


class _Event(object):
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


@pytest.mark.parametrize('extra', [[], ['--config', 'c4', '--users', '3', '--sentences-per-user', '4'],
                                   ['--config', 'c4', '--users', '3', '--sentences-per-user', '4', '--user-adapt'],
                                   ['--config', 'c5', '--V', '96', '--sweeps', '4'], ['--scaling', 'strong', '--msg-passes', '3']],
                         ids=['c3', 'c4', 'c4_adapt', 'c5', 'c3_strong'])
def test_bench_gpu_arm_dry_run(monkeypatch, extra):
    monkeypatch.setattr(engine, 'Kernels', FakeKernels)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda *_: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *_: None)
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, 'tensor', lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != 'device'}))
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--steps', '2', '--warmup', '1', '--sentences', '6', '--V', '96', '--Vd', '12', '--k', '4',
                                      '--no-cpu-baseline'] + extra)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.ours(bench.parse())
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'hbm_kernels', 'cpu_baseline', 'host',
                'message_rows'):
        assert key in line, key
    assert line['gpu_launches'] > 0 and line['value'] > 0 and line['e2e']['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] == 128
    assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'by_passes'}
    assert line['scaling'] == ('strong' if ('c4' in extra or 'strong' in extra) else 'weak')
    assert 'workload' in line['config']
