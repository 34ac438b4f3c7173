"""bench.py contract checks that need no GPU: the reference arm (lime's CPU algorithm, oracle port) prints one JSON
line with the keys the driver reads, and rank != 0 under torchrun does no work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'] + extra, capture_output=True, text=True, env=e, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_json_line():
    line = _run(['--workload', 'redfield_batch', '--batch', '64'])
    d = json.loads(line.splitlines()[-1])
    assert d['impl'] == 'reference' and d['metric'] == 'redfield_rho_steps_per_s' and d['unit'] == 'rho-steps/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(['--workload', 'redfield_batch'], env={'RANK': '1', 'WORLD_SIZE': '2'}) == ''


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a GPU is visible')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and 'no CPU fallback' in (out.stderr + out.stdout)
    assert out.stdout.strip() == ''          # no JSON line, no number
