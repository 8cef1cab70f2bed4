"""The C oracle (oracle/bsi_oracle.c) against the reference-generated golden fixtures."""
from math import inf

import numpy as np
import pytest

import goldenlib
from oracle import coracle, oracle


def _space(a):
    return coracle.Space(a['preds'], a['tts'], a['origin'], a['variations'])


@pytest.mark.parametrize('case', goldenlib.cases('attract'), ids=lambda c: c['name'])
def test_attract(case):
    a = goldenlib.case_args(case)
    sp = _space(a)
    assert sp.n_problems == case['n_problems']
    rec = coracle.attract_records(sp, a['max_t'], a['max_attractor_l'])
    exp = case['expected']
    assert rec['steps'] == exp['steps']
    assert [int(t) for t in rec['stop_t']] == exp['stop_t']
    keys = coracle.codes_from_words(rec['key'])
    got = [None if not f else [format(k, 'x'), int(l), int(m)]
           for f, k, l, m in zip(rec['found'], keys, rec['lam'], rec['traj_l'])]
    assert got == exp['records']
    table = coracle.attract_table(rec, sp.kw)
    rows = sorted(table.items(), key=lambda kv: (-kv[1][0], kv[0]))
    want = goldenlib.expected_attractor_rows(case)
    assert [(k, r[0], r[1], r[2], r[3]) for k, r in rows] == \
        [(k, n, s1, s2, len(cyc)) for k, n, s1, s2, cyc in want]


@pytest.mark.parametrize('case', goldenlib.cases('simulate'), ids=lambda c: c['name'])
def test_simulate(case):
    a = goldenlib.case_args(case)
    sp = _space(a)
    states, stop_t, steps = coracle.simulate_states(sp, a['max_t'])
    assert steps == case['expected']['steps']
    want = goldenlib.expected_simulations(case)
    assert len(want) == sp.n_problems
    for index, codes, _, _ in want:
        assert coracle.codes_from_words(states[index]) == codes


@pytest.mark.parametrize('case', goldenlib.cases('target'), ids=lambda c: c['name'])
def test_target(case):
    a = goldenlib.case_args(case)
    sp = _space(a)
    rec = coracle.target_records(sp, a['target_nodes'], a['target_code'], a['max_t'])
    assert rec['steps'] == case['expected']['steps']
    assert [int(t) for t in rec['stop_t']] == case['expected']['stop_t']
    hits = [int(i) for i in np.nonzero(rec['reached'])[0]]
    assert len(hits) == case['expected']['n_hits_total']
    want = goldenlib.expected_simulations(case)
    if a['n_to_find'] is not inf:
        hits = hits[:a['n_to_find']]
    assert hits == [w[0] for w in want]
    # the trajectory of a hit is plain iteration up to the hit time
    for index, codes, _, _ in want[:20]:
        t_hit = int(rec['stop_t'][index])
        assert len(codes) == t_hit + 1
        if t_hit > 0:
            states, _, _ = coracle.simulate_states(sp, t_hit, index, index + 1)
            assert coracle.codes_from_words(states[0]) == codes


def test_python_and_c_oracles_agree_on_a_fresh_network():
    from boolsi_b200 import synth
    cfg = synth.random_attract_problem(24, 2, 5, n_free=8)
    preds, tts = cfg['incoming node lists'], cfg['truth tables']
    origin, variations = cfg['origin simulation problem'], cfg['simulation problem variations']
    res = oracle.attract(preds, tts, origin, variations, 64, inf)
    sp = coracle.Space(preds, tts, origin, variations)
    rec = coracle.attract_records(sp, 64, inf)
    table = coracle.attract_table(rec, sp.kw)
    assert {k: v[:3] for k, v in res['table'].items()} == {k: v[:3] for k, v in table.items()}
    assert res['steps'] == rec['steps']


@pytest.mark.parametrize('n_nodes,k,seed,max_t', [(200, 2, 0, 300), (256, 2, 1, 400)])
def test_python_and_c_oracles_agree_at_the_maximum_network_size(n_nodes, k, seed, max_t):
    """Four key words per state (129..256 nodes): what tools/next_round/test_gpu_large.py
    will check the GPU path against (fixtures stop at 128 nodes)."""
    from boolsi_b200 import synth
    cfg = synth.random_attract_problem(n_nodes, k, seed, n_free=5)
    preds, tts = cfg['incoming node lists'], cfg['truth tables']
    origin, variations = cfg['origin simulation problem'], cfg['simulation problem variations']
    res = oracle.attract(preds, tts, origin, variations, max_t, inf)
    sp = coracle.Space(preds, tts, origin, variations)
    assert sp.kw == 4
    rec = coracle.attract_records(sp, max_t, inf)
    table = coracle.attract_table(rec, sp.kw)
    assert {key: v[:3] for key, v in res['table'].items()} == {key: v[:3] for key, v in table.items()}
    assert res['steps'] == rec['steps']
