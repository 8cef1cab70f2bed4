"""CPU: the C ABI is usable from plain C (no Python, no torch types): examples/attract_example.c
compiles against include/bsi.h with gcc -std=c99, links libbsi.so, specialises the kernels
offline, and an offline plan refuses to run (there is no CPU fallback)."""
import os
import shutil
import subprocess

import pytest

from boolsi_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('gcc') is None, reason='needs gcc')
def test_plain_c_program_builds_against_the_header_and_runs_offline(tmp_path):
    _lib.lib()                                              # make sure libbsi.so is built
    libdir = os.path.join(REPO, 'boolsi_b200')
    exe = str(tmp_path / 'attract_example')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(REPO, 'include'),
                    os.path.join(REPO, 'examples', 'attract_example.c'), '-L', libdir, '-lbsi',
                    '-Wl,-rpath,' + libdir, '-o', exe], check=True)
    out = subprocess.run([exe, '--offline'], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert 'problems 8,' in out.stdout
    assert 'run refused with status -2' in out.stdout and 'no CPU fallback' in out.stdout
