"""bench.py's reference arm (the oracle port on host cores) needs no GPU: run it once and check
the JSON line against the driver's contract.  The arm under test IS the checker timed as the CPU
baseline -- the one place besides tests/ where oracle/ may execute."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--steps', '1', '--warmup', '0'], capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line['impl'] == 'reference'
    assert line['metric'] == 'caption samples/sec (train fwd+bwd)' and line['unit'] == 'samples/s'
    assert line['higher_is_better'] is True and line['vs_baseline'] is None
    assert line['value'] > 0 and line['ms_per_step'] > 0 and line['steps'] == 1
    assert 'workload' in line['config'] and 'model' not in line['config']
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and cb['sample']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['unit'] == line['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
