"""bench.py's contract on a box without a GPU: the reference arm (the oracle port on the host cores - the one leg of bench.py
that may execute oracle/) prints the JSON line the driver parses, and the product arm fails loudly instead of falling back
to a CPU path."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), *args], cwd=REPO, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--height', '240', '--width', '320')
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['steps'] == 1 and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['vs_baseline'] is None and line['data'] == 'synthetic' and 'workload' in line['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1'],
                       cwd=REPO, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a box without a GPU')
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run('--steps', '1', '--warmup', '1', '--no-cpu-baseline', '--no-extras', timeout=120)
    assert r.returncode != 0
    assert r.stdout.strip() == ''                      # no JSON line from a CPU fallback
