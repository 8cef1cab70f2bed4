"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box):
the oracle restatements against the UNMODIFIED reference run live on fresh random inputs that are
not among the committed fixtures -- a second pin of the oracle besides tests/golden/."""
import itertools
from math import inf

import numpy as np
import pytest

import _refshim
from oracle import coracle, oracle

pytestmark = pytest.mark.skipif(not _refshim.reference_available(),
                                reason='reference tree not present (GPU box)')


@pytest.fixture(scope='module')
def ref():
    _refshim.install()
    import boolsi.attract as ra
    import boolsi.simulate as rs
    import boolsi.target as rt
    from boolsi.constants import NodeStateRange
    return ra, rs, rt, NodeStateRange


def _net(rng, n):
    preds, tts = [], []
    for _ in range(n):
        k = min(n, int(rng.choice([0, 1, 2, 2, 3, 3, 4])))
        p = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
        bits = rng.integers(0, 2, size=2 ** k)
        preds.append(p)
        tts.append({key: bool(bits[j]) for j, key in enumerate(itertools.product((False, True), repeat=k))})
    return preds, tts


def _code(state):
    return sum(1 << i for i, b in enumerate(state) if b)


@pytest.mark.parametrize('seed', range(100, 108))
def test_attract_against_the_live_reference(ref, seed):
    ra, _, _, _ = ref
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 18))
    preds, tts = _net(rng, n)
    free = sorted(int(x) for x in rng.choice(n, size=int(rng.integers(1, min(n, 7) + 1)), replace=False))
    origin = ([bool(b) for b in rng.integers(0, 2, size=n)],
              {int(rng.integers(0, n)): bool(rng.integers(0, 2))} if seed % 2 else {},
              {int(rng.integers(1, 6)): {int(rng.integers(0, n)): bool(rng.integers(0, 2))}} if seed % 3 else {})
    variations = (free, [], [])
    max_t = [inf, int(rng.integers(6, 30))][seed % 2]
    max_a = [inf, inf, 2][seed % 3]
    n_problems = 2 ** len(free)
    db = _refshim.new_db_conn()
    ra.init_attractor_db_structure(db)
    ra.attract_master(_refshim.NoMpiComm(), 3, origin, variations, preds, tts, max_t, max_a, n_problems,
                      True, db, False, None)
    want = {key: (obj.frequency, [_code(s) for s in obj.states])
            for (_, key), obj in db.root.aggregated_attractors.items()}
    res = oracle.attract(preds, tts, origin, variations, max_t, max_a)
    assert {k: (v[0], v[3]) for k, v in res['table'].items()} == want
    for (_, key), obj in db.root.aggregated_attractors.items():
        mean, m2 = oracle.mean_and_variation_sum(*res['table'][key][:3])
        assert abs(mean - obj.trajectory_l_mean) <= 1e-9 * max(1.0, abs(mean))
        assert abs(m2 - obj.trajectory_l_variation_sum) <= 1e-6 * max(1.0, abs(m2))
    sp = coracle.Space(preds, tts, origin, variations)
    rec = coracle.attract_records(sp, max_t, max_a)
    table = coracle.attract_table(rec, sp.kw)
    assert {k: v[0] for k, v in table.items()} == {k: v[0] for k, v in want.items()}
    assert rec['steps'] == res['steps']


@pytest.mark.parametrize('seed', range(200, 206))
def test_simulate_and_target_against_the_live_reference(ref, seed):
    _, rs, rt, R = ref
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 14))
    preds, tts = _net(rng, n)
    nodes = [int(x) for x in rng.permutation(n)]
    origin = ([bool(b) for b in rng.integers(0, 2, size=n)], {nodes[0]: bool(rng.integers(0, 2))},
              {2: {nodes[1]: True}})
    ranges = [R.TRUE_OR_FALSE, R.MAYBE_FALSE, R.MAYBE_TRUE, R.MAYBE_TRUE_OR_FALSE]
    variations = (sorted(nodes[2:4]), [(nodes[4 % n], ranges[seed % 4])],
                  [(int(rng.integers(1, 7)), nodes[5 % n], ranges[(seed + 1) % 4])])
    n_problems = oracle.count_problems(variations)
    max_t = 12
    db = _refshim.new_db_conn()
    rs.init_simulation_db_structure(db)
    rs.simulate_master(_refshim.NoMpiComm(), 3, origin, variations, preds, tts, max_t, n_problems, db, None)
    want = [[_code(s) for s in sim.states] for sim in db.root.simulations.values()]
    got = oracle.simulate(preds, tts, origin, variations, max_t)
    assert [s['states'] for s in got['simulations']] == want
    sp = coracle.Space(preds, tts, origin, variations)
    states, _, _ = coracle.simulate_states(sp, max_t)
    assert [coracle.codes_from_words(st) for st in states] == want
    # target: a substate of two nodes
    tnodes = {nodes[0], nodes[2]}
    tcode = sum(1 << x for x in tnodes if rng.integers(0, 2))
    db = _refshim.new_db_conn()
    rs.init_simulation_db_structure(db)
    rt.target_master(_refshim.NoMpiComm(), 3, origin, variations, tcode, tnodes, preds, tts, inf, 40,
                     n_problems, db, None)
    want_t = [[_code(s) for s in sim.states] for sim in db.root.simulations.values()]
    got_t = oracle.target(preds, tts, origin, variations, tnodes, tcode, 40)
    assert [s['states'] for s in got_t['simulations']] == want_t
    rec = coracle.target_records(sp, tnodes, tcode, 40)
    assert int(rec['reached'].sum()) == len(want_t)
