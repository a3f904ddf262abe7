"""bench.py's JSON contract, exercised on the CPU through the reference arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), cwd=ROOT, env=e,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = [l for l in _run('--impl', 'reference', '--workload', 'tiny', '--steps', '2', '--warmup', '1').splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'images/sec' and d['unit'] == 'images/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 2 and d['value'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and 1 <= cb['cores'] <= (os.cpu_count() or 1)
    assert 'workload' in d['config']


def test_reference_arm_non_zero_ranks_exit_silently():
    out = _run('--impl', 'reference', '--workload', 'tiny', '--gpus', '2', '--steps', '1', '--warmup', '0',
               env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert out.strip() == ''


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', 'tiny', '--steps', '1'], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and p.stdout.strip() == ''
