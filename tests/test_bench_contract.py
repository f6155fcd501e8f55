"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on the host cores) prints exactly ONE JSON
line on stdout with the keys of the bench contract, and non-zero ranks of a torchrun launch print nothing."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_json_line():
    lines = [l for l in _run({}).splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/sec' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='images/sec', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_are_silent():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}).strip() == ''
