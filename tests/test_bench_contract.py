"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the keys the bench
contract names -- the unmodified reference package from baseline/_ref (or /root/reference) timed on the
host cores on a bounded sample of the job; and bench.py refuses to run its own arm without a GPU
(the product has no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, REPO)
    from oracle import refarm
    if refarm.reference_root() is None:
        pytest.skip('the reference package is not installed (run __graft_entry__.build() where /root/reference exists)')
    env = dict(os.environ, BSI_BENCH_REF_BUDGET='0.5')
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--n-free', '14'], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['unit'] == 'trajectory-steps/s' and line['higher_is_better'] is True
    assert line['steps'] == 1 and line['n_gpus'] == 1 and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] == line['e2e']['value']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and 'attract on random N=64 K=3 net' in line['metric']


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and 'no CPU fallback' in (out.stderr + out.stdout)
