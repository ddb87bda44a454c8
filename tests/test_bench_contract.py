"""bench.py contract on CPU: the reference arm (the only arm that runs without a GPU) prints exactly one JSON line with the
keys the driver reads; the default arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--cpu-agents', '1500'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'agent-steps/sec' and d['unit'] == 'agent-steps/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert 'workload' in d['config'] and d['vs_baseline'] is None


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_default_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1', '--agents', '1000'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and 'no CPU fallback' in (out.stderr + out.stdout)
