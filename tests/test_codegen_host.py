"""CPU: the generated per-network macro header, EXECUTED on the host.

The header libbsi prepends to the kernel template is plain preprocessor text over 32-bit words
(32 trajectories per word), so tests/host_emul/harness.cpp can run its macros with a software
`lop3`: the synchronous update (LOP3 synthesis incl. Shannon trees for wide rules, fixed nodes as
constants), the perturbation switch, the index -> initial-state generators, the bit-matrix
pack / unpack and the filter compares are checked here against the truth tables and the problem
decoder -- the same arithmetic the GPU threads execute, without a GPU."""
import itertools
import os
import shutil
import subprocess

import numpy as np
import pytest

from boolsi_b200 import _lib
from boolsi_b200.problems import decode_simulation_problem

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')


def _random_network(seed):
    rng = np.random.default_rng([seed, 77])
    n = int(rng.integers(2, 41)) if seed % 4 else int(rng.integers(60, 100))
    preds, tts = [], []
    for _ in range(n):
        k = min(int(rng.choice([0, 1, 1, 2, 2, 3, 3, 3, 4, 5, 6])), n)
        p = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
        bits = rng.integers(0, 2, size=2 ** k)
        preds.append(p)
        tts.append({key: bool(bits[j]) for j, key in enumerate(itertools.product((False, True), repeat=k))})
    initial = [bool(b) for b in rng.integers(0, 2, size=n)]
    nodes = [int(x) for x in rng.permutation(n)]
    n_free = int(rng.integers(0, min(n, 9) + 1))
    free = sorted(nodes[:n_free])
    fixed = {node: bool(rng.integers(0, 2)) for node in nodes[n_free:n_free + int(rng.integers(0, 3))]}
    pert = {}
    for _ in range(int(rng.integers(0, 4))):
        pert.setdefault(int(rng.integers(1, 9)), {})[int(rng.integers(0, n))] = bool(rng.integers(0, 2))
    return n, preds, tts, (initial, fixed, pert), (free, [], []), rng


def _harness(tmp_path, packed):
    with _lib.Plan(packed, _lib.MODE_ATTRACT, offline=True, source_only=1) as plan:
        header = plan.source().split('#line 1 "bsi_template.cuh"')[0]
    path = tmp_path / 'header.h'
    path.write_text(header)
    exe = str(tmp_path / 'harness')
    subprocess.run(['g++', '-O1', '-std=c++17', '-w', '-DBSI_HEADER="%s"' % path,
                    os.path.join(HERE, 'host_emul', 'harness.cpp'), '-o', exe], check=True)

    def ask(lines):
        out = subprocess.run([exe], input='\n'.join(lines) + '\n', capture_output=True, text=True, check=True)
        return [[int(x) for x in row.split()] for row in out.stdout.splitlines()]
    return ask


def _lanes(words, n):
    """node words -> 32 lane states (lists of bool)"""
    return [[bool((words[i] >> lane) & 1) for i in range(n)] for lane in range(32)]


def _words(states, n):
    return [sum(int(states[lane][i]) << lane for lane in range(32)) for i in range(n)]


@pytest.mark.parametrize('seed', range(12))
def test_generated_macros_executed_on_the_host(seed, tmp_path):
    n, preds, tts, origin, variations, rng = _random_network(seed)
    initial, fixed, pert = origin
    pk = _lib.PackedProblem(preds, tts, origin, variations)
    ask = _harness(tmp_path, pk)
    kw = pk.kw

    # ---- BSI_UPDATE: one synchronous step of 32 random states (model.py:16-49) ------------------
    cmds, want = [], []
    for _ in range(6):
        a = [int(x) for x in rng.integers(0, 2 ** 32, size=n)]
        nxt = []
        for state in _lanes(a, n):
            new = [tts[i][tuple(state[p] for p in preds[i])] for i in range(n)]
            for node, value in fixed.items():            # fixed nodes are constants of the rules
                new[node] = value
            nxt.append(new)
        cmds.append('U ' + ' '.join(map(str, a)))
        want.append(_words(nxt, n))
    # ---- BSI_UPDATE_QUIET: the same step from states that hold their fixed values (any time >= 1 that is
    #      not a perturbation time): fixed nodes enter their readers as constants ------------------
    for _ in range(6):
        a = [int(x) for x in rng.integers(0, 2 ** 32, size=n)]
        for node, value in fixed.items():
            a[node] = 0xFFFFFFFF if value else 0
        nxt = []
        for state in _lanes(a, n):
            new = [tts[i][tuple(state[p] for p in preds[i])] for i in range(n)]
            for node, value in fixed.items():
                new[node] = value
            nxt.append(new)
        cmds.append('Q ' + ' '.join(map(str, a)))
        want.append(_words(nxt, n))
    # ---- BSI_PERTURB: overrides at their times (model.py:70-71) -----------------------------------
    for t in range(0, 10):
        a = [int(x) for x in rng.integers(0, 2 ** 32, size=n)]
        b = list(a)
        for node, value in pert.get(t, {}).items():
            b[node] = 0xFFFFFFFF if value else 0
        cmds.append('P %d ' % t + ' '.join(map(str, a)))
        want.append(b)
    # ---- BSI_INIT_DENSE / BSI_LANE_SETUP: problem index -> initial state (batching.py) ------------
    n_words = (pk.n_problems + 31) // 32
    for W in sorted({0, n_words - 1, int(rng.integers(0, n_words))}):
        states = []
        for lane in range(32):
            index = min(32 * W + lane, pk.n_problems - 1) if 32 * W + lane >= pk.n_problems else 32 * W + lane
            states.append(decode_simulation_problem(index, origin, variations)[0])
        cmds.append('D %d' % W)
        want.append(('dense', W, _words(states, n)))
    for index in sorted({0, pk.n_problems - 1, int(rng.integers(0, pk.n_problems))}):
        state = decode_simulation_problem(index, origin, variations)[0]
        code = sum(int(v) << i for i, v in enumerate(state))
        cmds.append('L %d' % index)
        want.append([(code >> (64 * j)) & (2 ** 64 - 1) for j in range(kw)])
    # ---- BSI_UNPACK / BSI_PACK -------------------------------------------------------------------
    a = [int(x) for x in rng.integers(0, 2 ** 32, size=n)]
    codes = [sum(int(v) << i for i, v in enumerate(state)) for state in _lanes(a, n)]
    cmds.append('T ' + ' '.join(map(str, a)))
    want.append([(c >> (64 * j)) & (2 ** 64 - 1) for c in codes for j in range(kw)] + a)
    # ---- filter compares -------------------------------------------------------------------------
    for trial in range(6):
        a = [int(x) for x in rng.integers(0, 2 ** 32, size=n)]
        k = list(a)
        for node in rng.choice(n, size=int(rng.integers(0, 3)), replace=False):
            k[int(node)] ^= int(rng.integers(0, 2 ** 32))                  # some lanes differ somewhere
        alive = int(rng.integers(0, 2 ** 32))
        cmds.append('F %d ' % alive + ' '.join(map(str, a + k)))
        differ = 0
        for x, y in zip(a, k):
            differ |= x ^ y
        want.append(('filter', (~alive & 0xFFFFFFFF) | differ))

    got = ask(cmds)
    assert len(got) == len(want)
    for cmd, g, w in zip(cmds, got, want):
        if isinstance(w, tuple) and w[0] == 'dense':
            # lanes beyond the last problem are don't-care (masked by the valid mask on the device)
            valid = 0xFFFFFFFF if 32 * w[1] + 32 <= pk.n_problems else (1 << (pk.n_problems - 32 * w[1])) - 1
            assert [x & valid for x in g] == [x & valid for x in w[2]], cmd
        elif isinstance(w, tuple) and w[0] == 'filter':
            p, ref_filter, full, ref_full, p_hot = g
            assert p == ref_filter == p_hot, cmd            # both macro forms == term-by-term filter
            assert full == ref_full == w[1], cmd            # filter + rest == full compare
            assert p & ~full == 0                           # a filter pass is necessary for a match
        else:
            assert g == w, cmd
