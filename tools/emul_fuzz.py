"""Longer differential run of the host-emulated kernels against the C oracle (the per-seed body of
tests/test_kernel_emul.py::test_emulated_kernels_on_random_networks over a seed range).
usage: python tools/emul_fuzz.py FIRST_SEED LAST_SEED"""
import os
import pathlib
import sys
import tempfile
import traceback

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tests')]
import test_kernel_emul as T  # noqa: E402

bad = []
for seed in range(int(sys.argv[1]), int(sys.argv[2]) + 1):
    with tempfile.TemporaryDirectory() as tmp:
        try:
            T.test_emulated_kernels_on_random_networks(seed, pathlib.Path(tmp))
            print('seed', seed, 'ok', flush=True)
        except Exception:
            bad.append(seed)
            print('seed', seed, 'FAILED', flush=True)
            traceback.print_exc()
print('failed seeds:', bad)
sys.exit(1 if bad else 0)
