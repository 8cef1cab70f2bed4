"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/bsi.h declares,
specialises kernels without a GPU, and refuses to run without one (no CPU fallback)."""
import ctypes
import os
import re
from math import inf

import pytest

from boolsi_b200 import _lib, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.lib()


def test_exports_match_header(built):
    header = open(os.path.join(REPO, 'include', 'bsi.h')).read()
    declared = set(re.findall(r'^(?:int|void|const char \*)\s*\*?(bsi_[a-z0-9_]+)\(', header, re.M))
    assert declared == set(_lib.EXPORTS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), name
    assert built.bsi_abi_version() == 2


def _pk(n, k, seed, n_free):
    cfg = synth.random_attract_problem(n, k, seed, n_free=n_free)
    return _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                              cfg['origin simulation problem'], cfg['simulation problem variations'])


def test_offline_specialisation_and_op_counts(built):
    pk = _pk(64, 3, 0, 26)
    with _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True) as plan:
        assert plan.n_problems == 2 ** 26
        ops = plan.op_counts()
        assert ops['compare_lop3'] == 64 and 0 < ops['update_lop3'] <= 64
        assert 'lop3.b32' in plan.source()
        assert len(plan.cubin()) > 10000
        with pytest.raises(_lib.BsiError) as e:
            plan.attract_run(0, 32, 10, inf)
        assert e.value.status == -2          # BSI_E_CUDA: no device, no fallback


def test_wide_rules_specialise(built):
    pk = _pk(24, 6, 1, 8)
    with _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True) as plan:
        assert plan.op_counts()['update_lop3'] > 24


def test_bad_arguments(built):
    pk = _pk(8, 2, 0, 4)
    with pytest.raises(_lib.BsiError) as e:
        _lib.Plan(pk, 7, offline=True)
    assert e.value.status == -1
    with pytest.raises(_lib.BsiError):
        _lib.Plan(pk, _lib.MODE_TARGET, offline=True)      # target mode without a target


def test_no_gpu_is_an_error_not_a_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    pk = _pk(8, 2, 0, 4)
    with pytest.raises(_lib.BsiError) as e:
        _lib.Plan(pk, _lib.MODE_ATTRACT, device=0)
    assert e.value.status == -2


def test_disk_cache_hands_the_next_process_identical_cubins(tmp_path):
    """The disk level of the kernel cache (BSI_CACHE_DIR; by default the user's cache directory, "off"
    disables it): the first process compiles and stores one file per unit,
    the second one loads them instead of running NVRTC and gets byte-identical kernels; a
    damaged file is ignored (recompiled), never trusted."""
    import hashlib
    import os
    import subprocess
    import sys
    import time
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, time, hashlib; sys.path.insert(0, %r)\n"
        "from boolsi_b200 import _lib, synth\n"
        "cfg = synth.random_attract_problem(24, 3, 5, n_free=22)\n"
        "pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],\n"
        "                        cfg['origin simulation problem'], cfg['simulation problem variations'])\n"
        "t = time.perf_counter()\n"
        "plan = _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True)\n"
        "print(time.perf_counter() - t, hashlib.sha1(plan.cubin()).hexdigest(), plan.cubin_bytes())\n" % repo)

    def run(cache_dir):
        env = dict(os.environ)
        env['BSI_CACHE_DIR'] = 'off' if cache_dir is None else str(cache_dir)
        out = subprocess.run([sys.executable, '-c', script], env=env, capture_output=True, text=True, check=True)
        seconds, digest, size = out.stdout.split()
        return float(seconds), digest, int(size)

    cold = run(tmp_path)
    files = sorted(os.listdir(tmp_path))
    # (2^22 initial states: every kernel unit; a small problem space would leave the queue kernel and
    #  the throughput-mode resolution kernels to the first run that needs them)
    assert len(files) == 6 and all(f.startswith('bsi-') and f.endswith('.cubin') for f in files), files
    warm = run(tmp_path)
    assert warm[1:] == cold[1:]
    assert warm[0] < 0.5 * cold[0], (cold, warm)
    # a truncated file and a file with foreign content fall back to compiling
    victim = os.path.join(tmp_path, files[0])
    data = open(victim, 'rb').read()
    open(victim, 'wb').write(data[:len(data) // 2])
    open(os.path.join(tmp_path, files[1]), 'wb').write(b'not a cache file')
    again = run(tmp_path)
    assert again[1:] == cold[1:]
    assert open(victim, 'rb').read() == data                      # rewritten after the recompile
    assert run(None)[1:] == cold[1:]                              # and with the disk level off: plain NVRTC


def test_small_problem_spaces_leave_three_kernel_units_for_later(tmp_path):
    """A problem space whose queues can never be long (here 2^10 initial states) is specialised for the
    dense trajectory kernel, the latency-oriented resolution kernels and the rest only: half the cold
    start of a job that consists of little else.  The other units are compiled by the first run that
    needs them (GPU tests: retry rounds and forced throughput mode on small spaces)."""
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys; sys.path.insert(0, %r)\n"
        "from boolsi_b200 import _lib, synth\n"
        "cfg = synth.random_attract_problem(24, 3, 5, n_free=10)\n"
        "pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],\n"
        "                        cfg['origin simulation problem'], cfg['simulation problem variations'])\n"
        "plan = _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True)\n" % repo)
    env = dict(os.environ)
    env['BSI_CACHE_DIR'] = str(tmp_path)
    subprocess.run([sys.executable, '-c', script], env=env, check=True)
    files = sorted(os.listdir(tmp_path))
    assert sorted(f[-8:-6] for f in files) == ['u1', 'u4', 'u5'], files
