"""CPU: the bench.py contract pieces that do not need a GPU — the reference arm's JSON line, and that the product arm
refuses to run without a CUDA device (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '1'])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['metric'] == 'images/sec' and line['unit'] == 'images/s'
    assert line['higher_is_better'] is True and line['n_gpus'] == 1 and line['steps'] == 1
    assert line['value'] > 0 and line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and 'oracle' in cb['sample']
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


@pytest.mark.skipif(torch.cuda.is_available(), reason='only meaningful on a box without a GPU')
def test_product_arm_fails_loudly_without_cuda():
    r = _run(['--steps', '1', '--warmup', '1'], timeout=300)
    assert r.returncode != 0 and 'no CUDA device' in (r.stderr + r.stdout)
