"""CPU tier: `bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) at a toy size: one JSON line,
both CPU variants timed (one process with threaded BLAS; train_mp.py-style pool of forked single-threaded workers,
train_mp.py:634-649), the faster one reported; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ['--impl', 'reference', '--steps', '2', '--warmup', '1', '--sentences', '16', '--V', '128', '--Vd', '32', '--k', '4',
        '--ref-sample', '3']


def run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py')] + ARGS, capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_line():
    out = run({})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['steps'] == 2 and line['gpu_launches'] == 0
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['variant'] in ('process_pool', 'blas_threads') and cb['cores'] >= 1
    assert cb['value'] == line['value'] == max(cb['variants'][k]['value'] for k in ('process_pool', 'blas_threads'))
    lit = cb['variants']['literal']                          # the op-for-op restatement, timed on two short sentences, never the arm's value
    assert lit['extrapolated'] and lit['value'] > 0 and set(lit['measured_s']) == {'k=3 (3 pairwise factors)', 'k=4 (6 pairwise factors)'}
    assert abs(lit['seconds_per_workload_sentence'] - (lit['per_sentence_constant_s'] + 6 * lit['per_pairwise_factor_s'])) < 1e-9   # k = 4
    assert cb['variants']['process_pool']['sentences_per_step'] == 3
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_are_silent():
    out = run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.returncode == 0 and out.stdout.strip() == ''
