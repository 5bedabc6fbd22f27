"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints exactly ONE JSON line on stdout with the
keys the driver reads, and the non-zero ranks of a torchrun launch of that arm exit 0 without output."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '1', '--steps', '1',
                           '--warmup', '1', '--cpu-sample', '1'], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'images/s'
    assert d['value'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
    for k in ('metric', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ''
