"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the CPU
restatement timed on the host cores) prints ONE JSON line with the keys the driver reads, on the same `config` object as the
GPU arm; the GPU arm itself refuses to run without a CUDA device (no CPU fallback behind the headline number)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, 'bench.py')


def _run(*args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, BENCH, *args], env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run('--impl', 'reference', '--steps', '1', '--warmup', '1')
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['steps'] == 1 and d['warmup'] == 1 and d['value'] > 0 and d['vs_baseline'] is None
    assert d['config']['workload'].startswith('cfg2') and 'model' not in d['config']
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    out = _run('--steps', '1', '--warmup', '1', '--workloads', 'none', timeout=300)
    assert out.returncode != 0
    assert not any(ln.startswith('{') for ln in out.stdout.splitlines())
