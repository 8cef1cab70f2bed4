"""`-r` (reference-point detection, reference attract.py:305-371): the reference-generated fixtures of
tests/golden/rmode.json (unmodified reference, `storing_all_states=False`) against the default-mode
results this repo computes for both flags.

The modes agree whenever no time cap cuts a detection short; rmode.json records, per attract case,
what `-r` reports and whether that equals the default mode.  boolsi_b200 always follows the default
mode and warns when `-r` is combined with a cap."""
import logging
import math
from math import inf

import pytest

import codec
import goldenlib
from oracle import oracle

RMODE = {c['name']: c for c in codec.load('rmode.json')['cases']}


def test_every_uncapped_case_agrees_between_the_modes():
    uncapped = [c for c in goldenlib.cases('attract') if c['max_t'] is None]
    assert len(uncapped) >= 15
    assert all(RMODE[c['name']]['equal_to_default_mode'] for c in uncapped)
    # and the cap artefacts are real: some capped cases differ (the ones DESIGN.md documents)
    assert any(not c['equal_to_default_mode'] for c in RMODE.values())


@pytest.mark.parametrize('case', [c for c in goldenlib.cases('attract')
                                  if RMODE[c['name']]['equal_to_default_mode'] and c['expected']['steps'] <= 150000],
                         ids=lambda c: c['name'])
def test_oracle_default_mode_equals_reference_r_mode_where_they_agree(case):
    a = goldenlib.case_args(case)
    res = oracle.attract(a['preds'], a['tts'], a['origin'], a['variations'], a['max_t'], a['max_attractor_l'])
    want = RMODE[case['name']]['attractors']
    got = oracle.sorted_attractors(res['table'])
    assert [(format(k, 'x'), r[0], [format(c, 'x') for c in r[3]]) for k, r in got] == \
        [(w['key'], w['frequency'], w['states']) for w in want]
    for (k, r), w in zip(got, want):
        mean, m2 = oracle.mean_and_variation_sum(r[0], r[1], r[2])
        assert math.isclose(mean, w['mean'], rel_tol=1e-12, abs_tol=1e-12)
        assert math.isclose(m2, w['variation_sum'], rel_tol=1e-9, abs_tol=1e-9)


class _Comm:
    n_workers = 0
    rank = 0


@pytest.mark.gpu
@pytest.mark.parametrize('case', [c for c in goldenlib.cases('attract') if RMODE[c['name']]['equal_to_default_mode']],
                         ids=lambda c: c['name'])
def test_attract_master_r_flag(case, caplog):
    import boolsi_b200
    from boolsi_b200.problems import state_from_code
    a = goldenlib.case_args(case)
    db = boolsi_b200.ResultStore()
    with caplog.at_level(logging.INFO):
        n = boolsi_b200.attract_master(_Comm(), 100, a['origin'], a['variations'], a['preds'], a['tts'],
                                       a['max_t'], a['max_attractor_l'], a['n_problems'], False, db, False, None)
    want = RMODE[case['name']]['attractors']
    assert n == len(want)
    assert db.root.total_frequency() == RMODE[case['name']]['total_frequency']
    for (sort_key, att), w in zip(db.root.aggregated_attractors.items(), want):
        assert sort_key == (-w['frequency'], int(w['key'], 16))
        assert att.states == [state_from_code(int(c, 16), a['n_nodes']) for c in w['states']]
        assert math.isclose(att.trajectory_l_mean, w['mean'], rel_tol=1e-12, abs_tol=1e-12)
        assert math.isclose(att.trajectory_l_variation_sum, w['variation_sum'], rel_tol=1e-9, abs_tol=1e-9)
    capped = a['max_t'] != inf
    warned = [r for r in caplog.records if r.levelno >= logging.WARNING and '-r' in r.getMessage()]
    assert bool(warned) == capped        # a cap + -r is announced; without a cap the modes are the same thing
