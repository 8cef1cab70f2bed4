"""The Python oracle (oracle/oracle.py) against the reference-generated golden fixtures."""
import math

import pytest

import goldenlib
from oracle import oracle


@pytest.mark.parametrize('case', goldenlib.cases('attract'), ids=lambda c: c['name'])
def test_attract(case):
    if case['expected']['steps'] > 150000:
        pytest.skip('covered by the C oracle (too slow for the pure-Python restatement)')
    a = goldenlib.case_args(case)
    res = oracle.attract(a['preds'], a['tts'], a['origin'], a['variations'], a['max_t'],
                         a['max_attractor_l'])
    exp = case['expected']
    got = [(k, r[0], r[1], r[2], r[3]) for k, r in oracle.sorted_attractors(res['table'])]
    assert got == goldenlib.expected_attractor_rows(case)
    assert res['steps'] == exp['steps']
    assert sum(r[1] for r in got) == exp['total_frequency']
    assert res['n_without'] == case['n_problems'] - exp['total_frequency']
    recs = [None if r is None else [format(r[0], 'x'), r[1], r[2]] for r in res['records']]
    assert recs == exp['records']
    # the reference's floats agree with the exact integer triple to rounding noise
    for (k, n, s1, s2, _), att in zip(got, exp['attractors']):
        mean, m2 = oracle.mean_and_variation_sum(n, s1, s2)
        assert math.isclose(mean, att['mean'], rel_tol=1e-12, abs_tol=1e-12)
        assert math.isclose(m2, att['variation_sum'], rel_tol=1e-9, abs_tol=1e-9)


@pytest.mark.parametrize('case', goldenlib.cases('simulate') + goldenlib.cases('target'),
                         ids=lambda c: c['name'])
def test_trajectories(case):
    a = goldenlib.case_args(case)
    if case['mode'] == 'simulate':
        res = oracle.simulate(a['preds'], a['tts'], a['origin'], a['variations'], a['max_t'])
    else:
        res = oracle.target(a['preds'], a['tts'], a['origin'], a['variations'],
                            a['target_nodes'], a['target_code'], a['max_t'], a['n_to_find'])
    got = [(s['index'], s['states'], s['fixed_nodes'], s['perturbations'])
           for s in res['simulations']]
    assert got == goldenlib.expected_simulations(case)
    assert res['steps'] == case['expected']['steps']
    assert res['n_hits_total'] == case['expected']['n_hits_total']


def test_decode_matches_enumeration_order():
    """Digit order: initial-state nodes first (LSD), then fixed, then perturbations."""
    from boolsi_b200.constants import NodeStateRange as R
    origin = ([False, True, False], {2: True}, {3: {0: True}})
    variations = ([1], [(0, R.MAYBE_TRUE_OR_FALSE)], [(5, 2, R.MAYBE_FALSE)])
    assert oracle.count_problems(variations) == 12
    code, fixed, pert = oracle.decode_problem(1 + 2 * 2 + 6 * 1, origin, variations)
    assert code == 0b010 and fixed == {2: True, 0: True} and pert == {3: {0: True}, 5: {2: False}}
    code, fixed, pert = oracle.decode_problem(0, origin, variations)
    assert code == 0 and fixed == {2: True} and pert == {3: {0: True}}
