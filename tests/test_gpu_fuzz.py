"""GPU: randomized differential test of all three modes against the C oracle -- random small
networks with mixed in-degrees (constants, copies, up to 6 inputs), random fixed nodes,
perturbations and variations of every NodeStateRange, random caps and index sub-ranges."""
import itertools
from math import inf

import numpy as np
import pytest

from boolsi_b200 import _lib
from boolsi_b200.constants import NodeStateRange as R
from oracle import coracle

pytestmark = pytest.mark.gpu
RANGES = [R.TRUE_OR_FALSE, R.MAYBE_FALSE, R.MAYBE_TRUE, R.MAYBE_TRUE_OR_FALSE]


def _random_case(seed, mode):
    rng = np.random.default_rng([seed, {'attract': 1, 'simulate': 2, 'target': 3}[mode]])
    n = int(rng.integers(2, 41)) if seed % 5 else int(rng.integers(60, 75))
    preds, tts = [], []
    for _ in range(n):
        k = int(rng.choice([0, 1, 1, 2, 2, 3, 3, 3, 4, 5, 6]))
        k = min(k, n)
        p = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
        bits = rng.integers(0, 2, size=2 ** k)
        preds.append(p)
        tts.append({key: bool(bits[j]) for j, key in enumerate(itertools.product((False, True), repeat=k))})
    initial = [bool(b) for b in rng.integers(0, 2, size=n)]
    nodes = [int(x) for x in rng.permutation(n)]
    n_free = int(rng.integers(0, min(n, 9) + 1))
    free = sorted(nodes[:n_free])
    rest = nodes[n_free:]
    fixed = {node: bool(rng.integers(0, 2)) for node in rest[:int(rng.integers(0, 3))]}
    pert = {}
    for _ in range(int(rng.integers(0, 4))):
        t = int(rng.integers(1, 12))
        pert.setdefault(t, {})[int(rng.integers(0, n))] = bool(rng.integers(0, 2))
    var_fixed, var_pert = [], []
    if mode != 'attract':                      # attract forbids these (input.py:392-397, 537-543)
        for node in rest[3:3 + int(rng.integers(0, 3))]:
            var_fixed.append((node, RANGES[int(rng.integers(0, 4))]))
        for _ in range(int(rng.integers(0, 3))):
            var_pert.append((int(rng.integers(1, 10)), int(rng.integers(0, n)), RANGES[int(rng.integers(0, 4))]))
    return n, preds, tts, (initial, fixed, pert), (free, var_fixed, var_pert), rng


@pytest.mark.parametrize('seed', range(24))
def test_attract_fuzz(seed):
    n, preds, tts, origin, variations, rng = _random_case(seed, 'attract')
    sp = coracle.Space(preds, tts, origin, variations)
    pk = _lib.PackedProblem(preds, tts, origin, variations)
    assert pk.n_problems == sp.n_problems
    max_t = [inf, int(rng.integers(1, 40)), int(rng.integers(40, 400))][seed % 3]
    max_a = [inf, inf, int(rng.integers(1, 6))][(seed // 3) % 3]
    lo = int(rng.integers(0, sp.n_problems))
    hi = int(rng.integers(lo, sp.n_problems + 1))
    # resolve_few_words = 1 sends even these small cases through the throughput-mode resolution
    # kernels (warp-aggregated table updates); no_early_words = 1 through per-lane records only
    with _lib.Plan(pk, _lib.MODE_ATTRACT, first_horizon=[0, 1, 7, 64][seed % 4],
                   resolve_few_words=[0, 1][(seed // 2) % 2], no_early_words=[0, 0, 1][seed % 3]) as plan:
        for b, e in ((0, sp.n_problems), (lo, hi)):
            plan.table_clear()
            st = plan.attract_run(b, e, max_t, max_a)
            rec = coracle.attract_records(sp, max_t, max_a, b, e)
            want = sorted((k, v[0], v[1], v[2], v[3]) for k, v in coracle.attract_table(rec, sp.kw).items())
            got = sorted((_lib.code_from_words(r['key'][:pk.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']),
                          int(r['lam'])) for r in plan.table_read())
            assert got == want
            assert st['steps'] == rec['steps']
            assert st['n_without'] == int((rec['found'] == 0).sum())


@pytest.mark.parametrize('opts', [dict(resolve_few_words=1), dict(resolve_few_words=1, no_early_words=1), {}],
                         ids=['throughput', 'throughput-records', 'auto'])
@pytest.mark.parametrize('n_free', [5, 12, 15])
def test_attract_many_attractors_per_warp(n_free, opts):
    """Every node keeps its state except a few that fall into short cycles: up to 2^n_free
    different attractors, so that the threads of a warp hold many different keys at once (the
    warp-level aggregation must fall back to per-thread table updates) -- against the C oracle."""
    n = 20
    preds = [[i] for i in range(n)]
    tts = [{(False,): False, (True,): True} for _ in range(n)]
    preds[17], tts[17] = [18], {(False,): False, (True,): True}          # 17 <- 18 <- 19 <- not 17
    preds[18], tts[18] = [19], {(False,): False, (True,): True}
    preds[19], tts[19] = [17], {(False,): True, (True,): False}
    preds[16], tts[16] = [3, 16], {(False, False): False, (False, True): True, (True, False): True, (True, True): False}
    origin = ([False] * n, {}, {})
    variations = (list(range(n_free)) + [17, 19], [], [])
    sp = coracle.Space(preds, tts, origin, variations)
    pk = _lib.PackedProblem(preds, tts, origin, variations)
    rec = coracle.attract_records(sp, 50, inf)
    want = sorted((k, v[0], v[1], v[2], v[3]) for k, v in coracle.attract_table(rec, sp.kw).items())
    assert len(want) >= 1 << (n_free - 1)
    with _lib.Plan(pk, _lib.MODE_ATTRACT, table_log2_slots=18, **opts) as plan:
        st = plan.attract_run(0, pk.n_problems, 50, inf)
        got = sorted((_lib.code_from_words(r['key'][:pk.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']),
                      int(r['lam'])) for r in plan.table_read(1 << 17))
        assert got == want
        assert st['steps'] == rec['steps']


@pytest.mark.parametrize('seed', range(16))
def test_simulate_fuzz(seed):
    n, preds, tts, origin, variations, rng = _random_case(seed, 'simulate')
    sp = coracle.Space(preds, tts, origin, variations)
    pk = _lib.PackedProblem(preds, tts, origin, variations)
    lp = max([t for t in origin[2]] + [t for t, _, _ in variations[2]] + [0])
    max_t = lp + int(rng.integers(0, 30))
    n_run = min(sp.n_problems, 700)
    lo = int(rng.integers(0, sp.n_problems - n_run + 1))
    want, _, _ = coracle.simulate_states(sp, max_t, lo, lo + n_run)
    with _lib.Plan(pk, _lib.MODE_SIMULATE) as plan:
        got, _ = plan.simulate_run(lo, lo + n_run, max_t)
        assert np.array_equal(got, want)
        pick = sorted(int(x) for x in rng.choice(sp.n_problems, size=min(sp.n_problems, 40), replace=False))
        sub, _ = plan.simulate_list(pick, max_t)
        for row, index in zip(sub, pick):
            one, _, _ = coracle.simulate_states(sp, max_t, index, index + 1)
            assert np.array_equal(row, one[0])


@pytest.mark.parametrize('seed', range(16))
def test_target_fuzz(seed):
    n, preds, tts, origin, variations, rng = _random_case(seed, 'target')
    sp = coracle.Space(preds, tts, origin, variations)
    pk = _lib.PackedProblem(preds, tts, origin, variations)
    target_nodes = sorted(int(x) for x in rng.choice(n, size=int(rng.integers(1, min(n, 4) + 1)), replace=False))
    target_code = sum((1 << node) for node in target_nodes if rng.integers(0, 2))
    max_t = [inf, int(rng.integers(12, 60))][seed % 2]
    rec = coracle.target_records(sp, set(target_nodes), target_code, max_t)
    want = np.nonzero(rec['reached'])[0]
    tgt = _lib.make_target(set(target_nodes), target_code, pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
        hit_index, hit_t, st = plan.target_run(0, sp.n_problems, max_t)
        assert [int(i) for i in hit_index] == [int(i) for i in want]
        assert [int(t) for t in hit_t] == [int(rec['stop_t'][i]) for i in want]
        assert st['n_found'] == len(want)
        few, _, _ = plan.target_run(0, sp.n_problems, max_t, 3)
        assert [int(i) for i in few] == [int(i) for i in want[:3]]
