"""The CPU arm of bench.py (`--impl reference`: the CPU oracle on a bounded sample of the
workload) runs without a GPU and prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                        '--steps', '1', '--cpu-atoms', '3'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    out = json.loads(line)
    assert out['impl'] == 'reference' and out['unit'] == 's' and out['higher_is_better'] is False
    for key in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'config',
                'cpu_baseline', 'e2e'):
        assert key in out, key
    assert out['cpu_baseline']['kind'] == 'port' and out['cpu_baseline']['cores'] >= 1
    assert out['e2e']['h2d_bytes_per_step'] == 0 and out['e2e']['d2h_bytes_per_step'] == 0
    assert out['value'] > 0
