"""GPU: the library's groups (one plan per device, ONE host process -- the replacement of the MPI task
farm, reference mpi.py:193-340) against single-plan runs and the C oracle; and the drop-in masters
using every visible GPU.  Multi-device cases skip on a one-GPU box."""
from math import inf

import numpy as np
import pytest

import boolsi_b200
from boolsi_b200 import _lib, synth
from oracle import coracle

pytestmark = pytest.mark.gpu


def _n_gpus():
    return _lib.device_count()


def _rows(plan):
    return sorted((_lib.code_from_words(r['key'][:plan.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']),
                   int(r['lam'])) for r in plan.table_read())


def _attract_problem(n_nodes, k, seed, n_free):
    cfg = synth.random_attract_problem(n_nodes, k, seed, n_free=n_free)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    return args, _lib.PackedProblem(*args)


@pytest.mark.parametrize('k,seed,max_t', [(3, 1, 1000), (3, 2, 1000), (2, 0, inf)])
def test_group_of_one_equals_the_plan_and_the_oracle(k, seed, max_t):
    args, pk = _attract_problem(64, k, seed, 13)
    sp = coracle.Space(*args)
    rec = coracle.attract_records(sp, max_t, inf)
    want = sorted((key, v[0], v[1], v[2], v[3]) for key, v in coracle.attract_table(rec, sp.kw).items())
    with _lib.Group(pk, _lib.MODE_ATTRACT, devices=[0]) as group:
        assert group.n_devices == 1
        for lo, hi in ((0, pk.n_problems), (37, 5001)):
            total, per = group.attract_run(lo, hi, max_t, inf)      # tables start from empty on every call
            if (lo, hi) == (0, pk.n_problems):
                assert _rows(group.plan(0)) == want
                assert total['steps'] == rec['steps'] and total['n_found'] + total['n_without'] == pk.n_problems
            else:
                sub = coracle.attract_records(sp, max_t, inf, lo, hi)
                assert _rows(group.plan(0)) == sorted((key, v[0], v[1], v[2], v[3])
                                                      for key, v in coracle.attract_table(sub, sp.kw).items())
                assert total['steps'] == sub['steps']


def test_group_grows_the_attractor_table_instead_of_failing():
    """More attractors than table slots (the reference's store just grows, attract.py:405-455): the group
    run enlarges the tables and repeats the job; a bare plan still reports BSI_E_CAPACITY."""
    n = 14
    preds = [[i] for i in range(n)]                         # identity network: every state is a fixed point
    tts = [{(False,): False, (True,): True} for _ in range(n)]
    pk = _lib.PackedProblem(preds, tts, ([False] * n, {}, {}), (list(range(n)), [], []))
    with _lib.Group(pk, _lib.MODE_ATTRACT, devices=[0], table_log2_slots=10) as group:
        total, _ = group.attract_run(0, 1 << n, inf, inf)
        assert total['n_found'] == 1 << n and total['steps'] == 1 << n
        assert group.plan(0).table_count() == 1 << n
    db = boolsi_b200.ResultStore()

    class Comm:
        n_workers = 0
    got = boolsi_b200.attract_master(Comm(), 100, ([False] * n, {}, {}), (list(range(n)), [], []), preds, tts, inf,
                                     inf, 1 << n, True, db, False, None)
    assert got == 1 << n == db.root.n_aggregated_attractors()


@pytest.mark.parametrize('k,seed,n_free,max_t', [(3, 0, 20, 1000), (3, 1, 18, 1000), (3, 2, 18, 3000), (2, 1, 18, inf)])
def test_all_devices_give_the_one_device_table(k, seed, n_free, max_t):
    if _n_gpus() < 2:
        pytest.skip('needs at least two GPUs')
    args, pk = _attract_problem(64, k, seed, n_free)
    with _lib.Group(pk, _lib.MODE_ATTRACT, devices=[0]) as one:
        t1, _ = one.attract_run(0, pk.n_problems, max_t, inf)
        want = _rows(one.plan(0))
    with _lib.Group(pk, _lib.MODE_ATTRACT) as group:
        assert group.n_devices == _n_gpus()
        for lo, hi in ((0, pk.n_problems), (0, pk.n_problems), (5, 77)):     # twice: buffers are reused
            total, per = group.attract_run(lo, hi, max_t, inf)
            if hi == pk.n_problems:
                assert _rows(group.plan(0)) == want
                assert (total['steps'], total['n_found'], total['n_without']) == (t1['steps'], t1['n_found'], t1['n_without'])
                assert sum(p['n_problems'] for p in per) == pk.n_problems and all(p['n_problems'] for p in per)
        states = group.plan(0).attractor_states_batch([(key, lam) for key, _, _, _, lam in want])
        for (key, _, _, _, lam), orbit in zip(want, states):
            assert orbit[0] == key == min(orbit) and len(set(orbit)) == lam


def test_all_devices_many_attractors_overflow_the_first_payload():
    """More records per device than the first gather buffer holds (1024): the exchange is repeated with
    room, nothing is counted twice."""
    if _n_gpus() < 2:
        pytest.skip('needs at least two GPUs')
    n = 13
    preds = [[i] for i in range(n)]
    tts = [{(False,): False, (True,): True} for _ in range(n)]
    pk = _lib.PackedProblem(preds, tts, ([False] * n, {}, {}), (list(range(n)), [], []))
    with _lib.Group(pk, _lib.MODE_ATTRACT, devices=[0, 1]) as group:
        total, _ = group.attract_run(0, 1 << n, inf, inf)
        rows = _rows(group.plan(0))
    assert total['n_found'] == 1 << n and len(rows) == 1 << n and all(r[1] == 1 for r in rows)


def test_target_and_simulate_groups_equal_the_plans():
    cfg = synth.random_target_problem(64, 3, 2, n_free=14, n_fixed=2, n_perturbed=2, n_perturbation_times=3,
                                      max_perturbation_t=20, n_target_nodes=5)          # 12320 of 16384 problems hit
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    pk = _lib.PackedProblem(*args)
    tgt = _lib.make_target(cfg['target node set'], cfg['target substate code'], pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
        want_i, want_t, want_st = plan.target_run(0, pk.n_problems, 200)
    assert len(want_i) > 50
    for devices in ([0], None):
        with _lib.Group(pk, _lib.MODE_TARGET, devices=devices, target=tgt) as group:
            for n_to_find in (inf, 7):
                got_i, got_t, st = group.target_run(0, pk.n_problems, 200, n_to_find)
                k = len(want_i) if n_to_find == inf else n_to_find
                assert np.array_equal(got_i, want_i[:k]) and np.array_equal(got_t, want_t[:k])
                assert st['n_found'] == want_st['n_found'] and st['steps'] == want_st['steps']
    with _lib.Plan(pk, _lib.MODE_SIMULATE) as plan:
        want, _ = plan.simulate_run(100, 1200, 30)
    with _lib.Group(pk, _lib.MODE_SIMULATE) as group:
        got, _ = group.simulate_run(100, 1200, 30)
    assert np.array_equal(got, want)


def test_masters_use_every_visible_gpu(caplog):
    """...for jobs that can use them: a job of fewer words than one GPU holds in flight stays on one device
    (more devices could not shorten it and each costs a CUDA context)."""
    import logging
    from boolsi_b200.attract import SMALL_JOB_PROBLEMS
    args, pk = _attract_problem(64, 3, 1, 21)
    assert pk.n_problems > SMALL_JOB_PROBLEMS

    class Comm:
        n_workers = 0
    db = boolsi_b200.ResultStore()
    with caplog.at_level(logging.INFO):
        n = boolsi_b200.attract_master(Comm(), 100, args[2], args[3], args[0], args[1], 1000, inf, pk.n_problems,
                                       True, db, False, None)
    assert any('Used %d GPU(s)' % _n_gpus() in r.getMessage() for r in caplog.records)
    small_args, small_pk = _attract_problem(64, 3, 1, 12)
    caplog.clear()
    with caplog.at_level(logging.INFO):
        boolsi_b200.attract_master(Comm(), 100, small_args[2], small_args[3], small_args[0], small_args[1], 1000, inf,
                                   small_pk.n_problems, True, boolsi_b200.ResultStore(), False, None)
    assert any('Used 1 GPU(s)' in r.getMessage() for r in caplog.records)
    db1 = boolsi_b200.ResultStore()
    n1 = boolsi_b200.attract_master(Comm(), 100, args[2], args[3], args[0], args[1], 1000, inf, pk.n_problems,
                                    True, db1, False, None, devices=[0])
    assert n == n1
    a, b = list(db.root.aggregated_attractors.items()), list(db1.root.aggregated_attractors.items())
    assert [k for k, _ in a] == [k for k, _ in b]
    for (_, x), (_, y) in zip(a, b):
        assert (x.states, x.frequency, x.sum_l, x.sum_l2) == (y.states, y.frequency, y.sum_l, y.sum_l2)


@pytest.mark.parametrize('opts', [{}, {'no_dict': 1}], ids=['stored-orbits', 'regenerated'])
def test_node_activity_counts_on_the_device(opts):
    """bsi_attractor_activity from the orbits the run stored (cycle dictionary) and from regenerated
    ones (dictionary off; several length classes) against a count over the explicit states; the
    observation matrix of the node-correlation analysis follows as one array division."""
    from boolsi_b200 import analysis
    from boolsi_b200.attract import read_attractors
    args, pk = _attract_problem(64, 3, 1, 14)
    with _lib.Plan(pk, _lib.MODE_ATTRACT, **opts) as plan:
        plan.attract_run(0, pk.n_problems, 1000, inf)
        attractors = read_attractors(plan, pk.n_nodes)
    assert len(attractors) >= 3 and len({len(a.states) for a in attractors}) >= 2
    for a in attractors:
        assert [int(c) for c in a.node_on_counts] == np.asarray(a.states).sum(axis=0).tolist()
    obs, freq = analysis.node_activity(attractors)
    for a in attractors:
        a.node_on_counts = None
    obs2, freq2 = analysis.node_activity(attractors)          # the reference's way: mean over the states
    assert np.array_equal(obs, obs2) and np.array_equal(freq, freq2)


def test_progress_callback_reports_finished_waves():
    """bsi_plan_set_progress: one call per finished wave of problems (chunk_log2), monotone, ending at
    the total -- the feed of the reference's progress bar (utils.py:19-54)."""
    args, pk = _attract_problem(64, 2, 1, 14)
    seen = []
    with _lib.Group(pk, _lib.MODE_ATTRACT, devices=[0], chunk_log2=11) as group:
        group.set_progress(lambda member, done, total: seen.append((member, done, total)))
        group.attract_run(0, pk.n_problems, inf, inf)
        group.set_progress(None)
        n_calls = len(seen)
        group.attract_run(0, pk.n_problems, inf, inf)
    assert n_calls == len(seen) == 8                      # 2^14 problems in waves of 2^11; none after switching off
    assert [d for _, d, _ in seen] == [2048 * (i + 1) for i in range(8)]
    assert all(m == 0 and t == pk.n_problems for m, _, t in seen)
