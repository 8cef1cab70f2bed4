"""CPU: the attract kernels themselves -- the text of csrc/bsi_template.cuh, not a restatement --
executed by 32 host threads per warp (tests/host_emul: a small CUDA shim + the launch sequence of
csrc/bsi.cpp) and compared with the C oracle: attractor table, trajectory-step count, conservation.

What this covers without a GPU: the checkpoint schedule and the three-level match filter, early /
deferred match records, the early-word queue, retry rounds, key / mu resolution in both the
latency and the throughput form, warp-aggregated table updates, and the plan options that move
checkpoints between shared memory and scratch.  What it cannot cover: anything that depends on
the real memory system or scheduler (that is what the -m gpu tests are for)."""
import itertools
import os
import shutil
import subprocess
from math import inf

import numpy as np
import pytest

from boolsi_b200 import _lib, synth
from oracle import coracle

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
pytestmark = pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')


def _build(tmp_path, packed, block_threads=32, **opts):
    with _lib.Plan(packed, _lib.MODE_ATTRACT, offline=True, block_threads=block_threads, source_only=1, **opts) as plan:
        header = plan.source().split('#line 1 "bsi_template.cuh"')[0]
    (tmp_path / 'header.h').write_text(header)
    exe = str(tmp_path / 'kernel_emul')
    libdir = os.path.join(REPO, 'boolsi_b200')
    subprocess.run(['g++', '-O1', '-std=c++17', '-w', '-pthread',
                    '-DBSI_HEADER="%s"' % (tmp_path / 'header.h'),
                    '-DBSI_TEMPLATE="%s"' % os.path.join(libdir, 'csrc', 'bsi_template.cuh'),
                    '-I', os.path.join(HERE, 'host_emul'), os.path.join(HERE, 'host_emul', 'kernel_emul.cpp'),
                    '-L', libdir, '-lbsi', '-Wl,-rpath,' + libdir, '-o', exe], check=True)
    return exe


def _run(exe, lo, hi, max_t, max_a, first_horizon=0, few_words=0, no_early=0, table_log2=12, no_dict=0,
         dict_log2=12):
    cap = lambda x: 0 if x == inf else int(x)            # noqa: E731  (0 = no cap on the command line)
    out = subprocess.run([exe] + [str(int(v)) for v in (lo, hi, cap(max_t), cap(max_a), first_horizon, few_words,
                                                         no_early, table_log2, no_dict, dict_log2)],
                         capture_output=True, text=True, timeout=600, check=True).stdout.splitlines()
    head = out[0].split()
    counters = dict(steps=int(head[1]), accepted=int(head[2]), without=int(head[3]), table_full=int(head[4]),
                    internal_error=int(head[5]))
    rows = []
    for line in out[1:]:
        key_part, stat_part = line[4:].split('|')
        key = sum(int(w) << (64 * j) for j, w in enumerate(key_part.split()))
        n, s1, s2, lam = (int(x) for x in stat_part.split())
        rows.append((key, n, s1, s2, lam))
    return counters, sorted(rows)


def _oracle(space, max_t, max_a, lo, hi):
    rec = coracle.attract_records(space, max_t, max_a, lo, hi)
    rows = sorted((k, v[0], v[1], v[2], v[3]) for k, v in coracle.attract_table(rec, space.kw).items())
    return rows, rec['steps'], int((rec['found'] == 0).sum())


def _check(exe, space, lo, hi, max_t, max_a, **run_opts):
    want_rows, want_steps, want_without = _oracle(space, max_t, max_a, lo, hi)
    counters, rows = _run(exe, lo, hi, max_t, max_a, **run_opts)
    assert counters['table_full'] == 0 and counters['internal_error'] == 0
    assert rows == want_rows, run_opts
    assert counters['steps'] == want_steps, run_opts
    assert counters['without'] == want_without
    assert counters['accepted'] + counters['without'] == hi - lo


def _random_case(seed):
    rng = np.random.default_rng([seed, 31])
    n = int(rng.integers(2, 41)) if seed % 3 else int(rng.integers(60, 75))
    preds, tts = [], []
    for _ in range(n):
        k = min(int(rng.choice([0, 1, 1, 2, 2, 3, 3, 3, 4, 5, 6])), n)
        p = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
        bits = rng.integers(0, 2, size=2 ** k)
        preds.append(p)
        tts.append({key: bool(bits[j]) for j, key in enumerate(itertools.product((False, True), repeat=k))})
    initial = [bool(b) for b in rng.integers(0, 2, size=n)]
    nodes = [int(x) for x in rng.permutation(n)]
    n_free = int(rng.integers(1, min(n, 9) + 1))
    free = sorted(nodes[:n_free])
    fixed = {node: bool(rng.integers(0, 2)) for node in nodes[n_free:n_free + int(rng.integers(0, 3))]}
    pert = {}
    for _ in range(int(rng.integers(0, 4))):
        pert.setdefault(int(rng.integers(1, 12)), {})[int(rng.integers(0, n))] = bool(rng.integers(0, 2))
    return (preds, tts, (initial, fixed, pert), (free, [], [])), rng


@pytest.mark.parametrize('seed', range(7))
def test_emulated_kernels_on_random_networks(seed, tmp_path):
    """Mixed in-degrees 0..6, fixed nodes, perturbations, caps, sub-ranges; every run form: default,
    tiny first horizon (retry rounds through the queue kernel), throughput-mode resolution,
    per-lane records only."""
    args, rng = _random_case(seed)
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk, filter_nodes=[0, 3][seed % 2])
    max_t = [inf, int(rng.integers(1, 40)), int(rng.integers(40, 400))][seed % 3]
    max_a = [inf, inf, int(rng.integers(1, 6))][(seed // 3) % 3]
    n = pk.n_problems
    lo = int(rng.integers(0, n))
    hi = int(rng.integers(lo, n + 1))
    _check(exe, space, 0, n, max_t, max_a)
    _check(exe, space, lo, hi, max_t, max_a, first_horizon=[1, 2, 7][seed % 3])
    _check(exe, space, 0, n, max_t, max_a, few_words=1, first_horizon=5)
    _check(exe, space, lo, hi, max_t, max_a, few_words=1, no_early=1)


@pytest.mark.parametrize('opts', [dict(smem_slots=2), dict(smem_slots=1), dict(smem_slots=2, min_blocks=6)],
                         ids=['slots2', 'slots1', 'slots2-cap168'])
def test_emulated_kernels_with_checkpoints_in_scratch(opts, tmp_path):
    cfg = synth.random_attract_problem(64, 3, 1, n_free=9)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk, **opts)
    _check(exe, space, 0, pk.n_problems, 1000, inf)
    _check(exe, space, 3, 300, 120, inf, first_horizon=16, few_words=1)


@pytest.mark.parametrize('n_nodes,k,seed,tune', [(64, 3, 2, 32), (64, 3, 2, 64), (72, 3, 0, 32)],
                         ids=['64-by-transpose', '64-bit-by-bit', '72-by-transpose'])
def test_emulated_extraction_of_final_checkpoint_states(n_nodes, k, seed, tune, tmp_path):
    """Lanes that match a FINAL checkpoint ship y (that checkpoint) and z (the one before it): both
    extraction forms -- 32x32 transposes of the checkpoints for all lanes of the word, single-bit moves
    per lane -- forced on the same inputs: chaotic networks whose closing lanes match the final checkpoints
    (4 to 32 such lanes per word; 72 nodes = two key words and a partial third 32-node block)."""
    cfg = synth.random_attract_problem(n_nodes, k, seed, n_free=9)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk, tune=tune)
    for first_horizon, few_words in ((16, 1), (0, 0)):
        _check(exe, space, 0, pk.n_problems, 1000, inf, first_horizon=first_horizon, few_words=few_words, no_early=1)
    _check(exe, space, 5, 300, 90, 40, first_horizon=4)


# (160 and 200 nodes run on the device: tests/test_gpu_large.py)
@pytest.mark.parametrize('n_nodes,k,seed,max_t', [(256, 2, 1, 400)])
def test_emulated_kernels_at_the_maximum_network_size(n_nodes, k, seed, max_t, tmp_path):
    """Three and four key words per state: beyond the golden fixtures (128 nodes)."""
    cfg = synth.random_attract_problem(n_nodes, k, seed, n_free=7)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    assert pk.kw == (n_nodes + 63) // 64
    exe = _build(tmp_path, pk)
    _check(exe, space, 0, pk.n_problems, max_t, inf)
    _check(exe, space, 5, 100, max_t, inf, few_words=1, first_horizon=8)


@pytest.mark.parametrize('k,seed', [(2, 0), (2, 1), (1, 0)])
def test_emulated_kernels_in_the_ordered_regime(k, seed, tmp_path):
    """Every lane falls into one of one or two fixed points within a few steps: whole words hold
    one key (the uniform early-word path), in both resolution forms and on ragged ranges."""
    cfg = synth.random_attract_problem(64, k, seed, n_free=11)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk)
    for max_t, max_a in ((inf, inf), (1000, inf), (6, inf), (inf, 1)):
        _check(exe, space, 0, pk.n_problems, max_t, max_a, few_words=1)
        _check(exe, space, 37, 1501, max_t, max_a)


# ---- simulate: bsi_trajectories with per-lane fixed-node / perturbation variations -------------------

def _build_simulate(tmp_path, packed):
    with _lib.Plan(packed, _lib.MODE_SIMULATE, offline=True, block_threads=32, source_only=1) as plan:
        header = plan.source().split('#line 1 "bsi_template.cuh"')[0]
    (tmp_path / 'header_sim.h').write_text(header)
    exe = str(tmp_path / 'simulate_emul')
    libdir = os.path.join(REPO, 'boolsi_b200')
    subprocess.run(['g++', '-O1', '-std=c++17', '-w', '-pthread',
                    '-DBSI_HEADER="%s"' % (tmp_path / 'header_sim.h'),
                    '-DBSI_TEMPLATE="%s"' % os.path.join(libdir, 'csrc', 'bsi_template.cuh'),
                    '-I', os.path.join(HERE, 'host_emul'), os.path.join(HERE, 'host_emul', 'simulate_emul.cpp'),
                    '-o', exe], check=True)
    return exe


@pytest.mark.parametrize('seed', range(6))
def test_emulated_trajectories_with_every_kind_of_variation(seed, tmp_path):
    """simulate mode: initial-state, fixed-node and perturbation variations of every NodeStateRange
    (radix-2 and radix-3 digits, per-lane masks, per-lane last perturbation time) against the C
    oracle's plain iteration; contiguous ranges and explicit index lists."""
    from boolsi_b200.constants import NodeStateRange as R
    ranges = [R.TRUE_OR_FALSE, R.MAYBE_FALSE, R.MAYBE_TRUE, R.MAYBE_TRUE_OR_FALSE]
    args, rng = _random_case(seed + 50)
    preds, tts, origin, variations = args
    n_nodes = len(preds)
    free = variations[0][:5]
    others = [i for i in range(n_nodes) if i not in free and i not in origin[1]]
    var_fixed = [(int(node), ranges[int(rng.integers(0, 4))]) for node in others[:int(rng.integers(0, 3))]]
    var_pert = [(int(rng.integers(1, 10)), int(rng.integers(0, n_nodes)), ranges[int(rng.integers(0, 4))])
                for _ in range(int(rng.integers(0, 3)))]
    args = (preds, tts, origin, (free, var_fixed, var_pert))
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    assert pk.n_problems == space.n_problems
    exe = _build_simulate(tmp_path, pk)
    max_t = 14 + seed
    n = pk.n_problems
    lo = int(rng.integers(0, n))
    hi = int(rng.integers(lo, n + 1))
    for b, e in ((0, n), (lo, hi)):
        out = tmp_path / 'states.bin'
        subprocess.run([exe, str(b), str(e), str(max_t), str(out)], check=True, timeout=600)
        got = np.fromfile(out, dtype='<u8').reshape(e - b, max_t + 1, pk.kw)
        want, _, _ = coracle.simulate_states(space, max_t, b, e)
        assert np.array_equal(got, np.asarray(want, np.uint64)), (b, e)
    pick = sorted(int(i) for i in rng.choice(n, size=min(n, 9), replace=False))
    out = tmp_path / 'list.bin'
    subprocess.run([exe, '--list', str(max_t), str(out)] + [str(i) for i in pick], check=True, timeout=600)
    got = np.fromfile(out, dtype='<u8').reshape(len(pick), max_t + 1, pk.kw)
    for row, index in zip(got, pick):
        want, _, _ = coracle.simulate_states(space, max_t, index, index + 1)
        assert np.array_equal(row, np.asarray(want, np.uint64)[0]), index


def test_emulated_trajectories_of_a_256_node_network(tmp_path):
    cfg = synth.random_attract_problem(256, 3, 2, n_free=6)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    assert pk.kw == 4
    exe = _build_simulate(tmp_path, pk)
    out = tmp_path / 'states.bin'
    subprocess.run([exe, '0', str(pk.n_problems), '40', str(out)], check=True, timeout=600)
    got = np.fromfile(out, dtype='<u8').reshape(pk.n_problems, 41, 4)
    want, _, _ = coracle.simulate_states(space, 40, 0, pk.n_problems)
    assert np.array_equal(got, np.asarray(want, np.uint64))


# ---- target: bsi_target_main with horizon rounds and the closure test --------------------------------

@pytest.mark.parametrize('seed', range(6))
def test_emulated_target_search(seed, tmp_path):
    """target mode with variations of every kind: first hit time per problem from its own last
    perturbation on, capped and uncapped (closure rounds), whole space and a ragged sub-range."""
    from boolsi_b200.constants import NodeStateRange as R
    ranges = [R.TRUE_OR_FALSE, R.MAYBE_FALSE, R.MAYBE_TRUE, R.MAYBE_TRUE_OR_FALSE]
    args, rng = _random_case(seed + 80)
    preds, tts, origin, variations = args
    n_nodes = len(preds)
    free = variations[0][:6]
    others = [i for i in range(n_nodes) if i not in free and i not in origin[1]]
    var_fixed = [(int(node), ranges[int(rng.integers(0, 4))]) for node in others[:int(rng.integers(0, 3))]]
    var_pert = [(int(rng.integers(1, 10)), int(rng.integers(0, n_nodes)), ranges[int(rng.integers(0, 4))])
                for _ in range(int(rng.integers(0, 3)))]
    args = (preds, tts, origin, (free, var_fixed, var_pert))
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    target_nodes = sorted(int(x) for x in rng.choice(n_nodes, size=int(rng.integers(1, min(n_nodes, 4) + 1)),
                                                     replace=False))
    target_code = sum((1 << node) for node in target_nodes if rng.integers(0, 2))
    tgt = _lib.make_target(set(target_nodes), target_code, pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt, offline=True, block_threads=32, source_only=1) as plan:
        header = plan.source().split('#line 1 "bsi_template.cuh"')[0]
    (tmp_path / 'header_tgt.h').write_text(header)
    exe = str(tmp_path / 'target_emul')
    libdir = os.path.join(REPO, 'boolsi_b200')
    subprocess.run(['g++', '-O1', '-std=c++17', '-w', '-pthread',
                    '-DBSI_HEADER="%s"' % (tmp_path / 'header_tgt.h'),
                    '-DBSI_TEMPLATE="%s"' % os.path.join(libdir, 'csrc', 'bsi_template.cuh'),
                    '-I', os.path.join(HERE, 'host_emul'), os.path.join(HERE, 'host_emul', 'target_emul.cpp'),
                    '-o', exe], check=True)
    n = pk.n_problems
    lo = int(rng.integers(0, n))
    hi = int(rng.integers(lo, n + 1))
    lp_max = max([t for t in origin[2]] + [t for t, _, _ in var_pert] + [0])
    for max_t in (inf, lp_max + int(rng.integers(0, 40))):
        for b, e in ((0, n), (lo, hi)):
            rec = coracle.target_records(space, set(target_nodes), target_code, max_t, b, e)
            want = [(b + int(i), int(rec['stop_t'][i])) for i in np.nonzero(rec['reached'])[0]]
            out = subprocess.run([exe, str(b), str(e), '0' if max_t == inf else str(max_t)], capture_output=True,
                                 text=True, check=True, timeout=600).stdout.splitlines()
            got = [tuple(int(x) for x in line.split()[1:]) for line in out if line.startswith('hit')]
            assert got == want, (max_t, b, e)
            assert int(out[-1].split()[1]) == len(want)


@pytest.mark.parametrize('run_opts', [dict(few_words=1), dict(few_words=1, no_early=1), {}],
                         ids=['throughput', 'throughput-records', 'auto'])
def test_emulated_kernels_with_thousands_of_attractors(run_opts, tmp_path):
    """Every node keeps its state except a few on short cycles: 2560 different attractors, every
    thread of a warp holding many keys at once (the warp-level aggregation falls back to per-thread
    table updates; the table fills to a fifth)."""
    n, n_free = 20, 10
    preds = [[i] for i in range(n)]
    tts = [{(False,): False, (True,): True} for _ in range(n)]
    preds[17], preds[18], preds[19] = [18], [19], [17]                   # 17 <- 18 <- 19 <- not 17
    tts[19] = {(False,): True, (True,): False}
    preds[16], tts[16] = [3, 16], {(False, False): False, (False, True): True, (True, False): True,
                                   (True, True): False}
    args = (preds, tts, ([False] * n, {}, {}), (list(range(n_free)) + [17, 19], [], []))
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk)
    want_rows, _, _ = _oracle(space, 50, inf, 0, pk.n_problems)
    assert len(want_rows) == 2560
    _check(exe, space, 0, pk.n_problems, 50, inf, table_log2=14, **run_opts)


@pytest.mark.parametrize('block_threads', [32, 128])
def test_emulated_cycle_dictionary(block_threads, tmp_path):
    """Seed-1 network of the bench family (every trajectory ends in one of seven cycles of length 14,
    29 or 58): throughput-mode resolution with the cycle dictionary -- learnt by the first words,
    looked up by the rest; several independent warps, so that lookups race with a learner -- on,
    off, and with orbit storage for two of the seven attractors only.  The harness also checks
    that every stored orbit starts at its minimum state."""
    cfg = synth.random_attract_problem(64, 3, 1, n_free=12)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk, block_threads=block_threads)
    for opts in (dict(few_words=1), dict(few_words=1, no_early=1), dict(few_words=1, no_dict=1),
                 dict(few_words=1, dict_log2=6), dict(few_words=1, first_horizon=8), {}):
        _check(exe, space, 0, pk.n_problems, 1000, inf, **opts)
    _check(exe, space, 100, 3001, 90, 30, few_words=1)


def test_emulated_early_words_of_the_cambium_network(tmp_path):
    """cambium2 (30 nodes, every state an initial state): two thirds of the lanes end in fixed points,
    the rest on cycles of length 2 to 18 -- words that mix fixed points, walked cycles and, with the
    18-cycle, the dictionary; the event / group resolution of the early-word queue in the throughput
    and the latency form, with and without caps, on ragged ranges spread over the 2^30 states."""
    import sys
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import codec
    from boolsi_b200.constants import NodeStateRange
    pub = codec.load('published.json')
    _, preds, tts = codec.decode_network(codec.load('networks.json')['networks']['cambium'])
    origin, variations = codec.decode_problem(pub['cambium2_problem'], NodeStateRange)
    args = (preds, tts, origin, variations)
    space = coracle.Space(*args)
    pk = _lib.PackedProblem(*args)
    exe = _build(tmp_path, pk)
    for lo, n, max_t, max_a, opts in ((0, 4096, inf, inf, dict(few_words=1)),
                                      (123456789, 3001, inf, inf, dict(few_words=1)),
                                      (700000003, 2500, 12, 6, dict(few_words=1)),
                                      (1000000000, 2048, inf, inf, {}),
                                      (1073741824 - 1500, 1500, inf, inf, dict(few_words=1, first_horizon=6)),
                                      (5, 2000, inf, inf, dict(few_words=1, no_dict=1))):
        _check(exe, space, lo, lo + n, max_t, max_a, table_log2=10, **opts)
