"""GPU: the drop-in `*_master` functions against the reference-generated golden fixtures --
the same checks the reference's own attract/simulate/target tests make on `db_conn.root`."""
from math import inf

import pytest

import goldenlib
import boolsi_b200
from boolsi_b200.problems import state_from_code
from oracle import oracle

pytestmark = pytest.mark.gpu


class _Comm:
    n_workers = 0
    rank = 0


@pytest.mark.parametrize('case', goldenlib.cases('attract'), ids=lambda c: c['name'])
def test_attract_master(case):
    a = goldenlib.case_args(case)
    db = boolsi_b200.ResultStore()
    n = boolsi_b200.attract_master(_Comm(), 100, a['origin'], a['variations'], a['preds'], a['tts'],
                                   a['max_t'], a['max_attractor_l'], a['n_problems'], True, db, False, None)
    want = goldenlib.expected_attractor_rows(case)
    assert n == len(want) == db.root.n_aggregated_attractors()
    assert db.root.total_frequency() == sum(r[1] for r in want)
    got = list(db.root.aggregated_attractors.items())
    assert [k for k, _ in got] == [(-f, key) for key, f, _, _, _ in want]
    for (_, att), (key, f, s1, s2, cyc) in zip(got, want):
        assert att.frequency == f
        assert att.states == [state_from_code(c, a['n_nodes']) for c in cyc]
        mean, m2 = oracle.mean_and_variation_sum(f, s1, s2)
        assert att.trajectory_l_mean == mean and att.trajectory_l_variation_sum == m2
        # node activity counted on the device == what attractor_analysis.py:49-51 sums over the states
        assert [int(c) for c in att.node_on_counts] == [sum(st[i] for st in att.states) for i in range(a['n_nodes'])]
    if want:
        assert sum(len(v) for v in db.root.aggregated_attractor_keys_by_batch_index.values()) == len(want)


@pytest.mark.parametrize('case', goldenlib.cases('simulate'), ids=lambda c: c['name'])
def test_simulate_master(case):
    a = goldenlib.case_args(case)
    db = boolsi_b200.ResultStore()
    boolsi_b200.simulate_master(_Comm(), 100, a['origin'], a['variations'], a['preds'], a['tts'],
                                a['max_t'], a['n_problems'], db, None)
    want = goldenlib.expected_simulations(case)
    got = list(db.root.simulations.values())
    assert db.root.n_simulations() == len(want) == len(got)
    for sim, (index, codes, fixed, pert) in zip(got, want):
        assert sim.states == [state_from_code(c, a['n_nodes']) for c in codes]
        assert sim.fixed_nodes == fixed and sim.perturbed_nodes_by_t == pert
        assert sim.n_perturbations == sum(len(d) for d in pert.values())


@pytest.mark.parametrize('case', goldenlib.cases('target'), ids=lambda c: c['name'])
def test_target_master(case):
    a = goldenlib.case_args(case)
    db = boolsi_b200.ResultStore()
    n = boolsi_b200.target_master(_Comm(), 100, a['origin'], a['variations'], a['target_code'],
                                  a['target_nodes'], a['preds'], a['tts'], a['n_to_find'], a['max_t'],
                                  a['n_problems'], db, None)
    want = goldenlib.expected_simulations(case)
    got = list(db.root.simulations.values())
    assert n == len(want) == len(got)
    for sim, (index, codes, fixed, pert) in zip(got, want):
        assert sim.states == [state_from_code(c, a['n_nodes']) for c in codes]
        assert sim.fixed_nodes == fixed and sim.perturbed_nodes_by_t == pert
