"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the reference-generated
golden fixtures and against the C oracle on seeded inputs.  Bit-exact: integers only."""
from math import inf

import numpy as np
import pytest

import goldenlib
from boolsi_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _packed(a):
    return _lib.PackedProblem(a['preds'], a['tts'], a['origin'], a['variations'])


def _table_rows(plan):
    recs = plan.table_read()
    rows = [(_lib.code_from_words(r['key'][:plan.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']),
             int(r['lam'])) for r in recs]
    return sorted(rows, key=lambda r: (-r[1], r[0]))     # attract.py:170-186 order


@pytest.mark.parametrize('case', goldenlib.cases('attract'), ids=lambda c: c['name'])
def test_attract_golden(case):
    a = goldenlib.case_args(case)
    pk = _packed(a)
    assert pk.n_problems == case['n_problems']
    with _lib.Plan(pk, _lib.MODE_ATTRACT) as plan:
        st = plan.attract_run(0, pk.n_problems, a['max_t'], a['max_attractor_l'])
        want = goldenlib.expected_attractor_rows(case)
        got = _table_rows(plan)
        assert [(k, n, s1, s2, lam) for k, n, s1, s2, lam in got] == \
            [(k, n, s1, s2, len(cyc)) for k, n, s1, s2, cyc in want]
        exp = case['expected']
        assert st['steps'] == exp['steps']
        assert st['n_found'] == sum(r[1] for r in want)
        assert st['n_without'] == sum(1 for r in exp['records'] if r is None)
        for k, n, s1, s2, cyc in want[:8]:
            assert plan.attractor_states(k, len(cyc)) == cyc


@pytest.mark.parametrize('first_horizon', [1, 2, 5, 64])
@pytest.mark.parametrize('name', ['attract_rnd64k2s0_f10_t5000_ainf', 'attract_rnd20k2s3_f12_tinf_ainf',
                                  'attract_cambium2_rand11', 'attract_rulesA_fixed_perturbed',
                                  'attract_rnd33k3s2_f10_t400_ainf'])
def test_attract_horizon_rounds(name, first_horizon):
    """The result may not depend on the detection-round schedule (retry queue path)."""
    case = [c for c in goldenlib.cases('attract') if c['name'] == name][0]
    a = goldenlib.case_args(case)
    pk = _packed(a)
    with _lib.Plan(pk, _lib.MODE_ATTRACT, first_horizon=first_horizon, filter_nodes=3) as plan:
        st = plan.attract_run(0, pk.n_problems, a['max_t'], a['max_attractor_l'])
        want = goldenlib.expected_attractor_rows(case)
        assert _table_rows(plan) == [(k, n, s1, s2, len(cyc)) for k, n, s1, s2, cyc in want]
        assert st['steps'] == case['expected']['steps']


@pytest.mark.parametrize('opts', [dict(smem_slots=3, min_blocks=1), dict(smem_slots=2, min_blocks=6),
                                  dict(smem_slots=1, min_blocks=1), dict(smem_slots=2, block_threads=32)],
                         ids=lambda o: '-'.join('%s%d' % (k[:4], v) for k, v in o.items()))
@pytest.mark.parametrize('name', ['attract_rnd64k2s0_f10_t5000_ainf', 'attract_rnd20k2s3_f12_tinf_ainf',
                                  'attract_rulesA_fixed_perturbed', 'attract_rnd33k3s2_f10_t400_ainf'])
def test_attract_checkpoint_placement(name, opts):
    """Where the checkpoint slots live (shared memory or the L2-resident scratch) and the register
    cap are tuning choices: every combination gives the reference's result."""
    case = [c for c in goldenlib.cases('attract') if c['name'] == name][0]
    a = goldenlib.case_args(case)
    pk = _packed(a)
    with _lib.Plan(pk, _lib.MODE_ATTRACT, **opts) as plan:
        st = plan.attract_run(0, pk.n_problems, a['max_t'], a['max_attractor_l'])
        want = goldenlib.expected_attractor_rows(case)
        assert _table_rows(plan) == [(k, n, s1, s2, len(cyc)) for k, n, s1, s2, cyc in want]
        assert st['steps'] == case['expected']['steps']
        for k, n, s1, s2, cyc in want[:4]:
            assert plan.attractor_states(k, len(cyc)) == cyc


def test_attract_ranges_accumulate_and_clear():
    case = [c for c in goldenlib.cases('attract') if c['name'] == 'attract_cambium2_first12'][0]
    a = goldenlib.case_args(case)
    pk = _packed(a)
    want = [(k, n, s1, s2, len(cyc)) for k, n, s1, s2, cyc in goldenlib.expected_attractor_rows(case)]
    with _lib.Plan(pk, _lib.MODE_ATTRACT, chunk_log2=9) as plan:
        steps = 0
        for b, e in [(0, 7), (7, 1000), (1000, 1000), (1000, 4001), (4001, 4096)]:   # ragged, empty
            steps += plan.attract_run(b, e, a['max_t'], a['max_attractor_l'])['steps']
        assert _table_rows(plan) == want
        assert steps == case['expected']['steps']
        plan.table_clear()
        assert plan.table_count() == 0
        plan.attract_run(0, 4096, a['max_t'], a['max_attractor_l'])
        assert _table_rows(plan) == want


@pytest.mark.parametrize('case', goldenlib.cases('simulate'), ids=lambda c: c['name'])
def test_simulate_golden(case):
    a = goldenlib.case_args(case)
    pk = _packed(a)
    want = goldenlib.expected_simulations(case)
    with _lib.Plan(pk, _lib.MODE_SIMULATE) as plan:
        states, st = plan.simulate_run(0, pk.n_problems, a['max_t'])
        assert len(want) == pk.n_problems
        for index, codes, _, _ in want:
            assert [_lib.code_from_words(w) for w in states[index]] == codes
        # explicit index lists give the same trajectories
        pick = [w[0] for w in want[::max(1, len(want) // 7)]]
        sub, _ = plan.simulate_list(pick, a['max_t'])
        for row, index in zip(sub, pick):
            assert [_lib.code_from_words(w) for w in row] == want[index][1]


@pytest.mark.parametrize('case', goldenlib.cases('target'), ids=lambda c: c['name'])
def test_target_golden(case):
    a = goldenlib.case_args(case)
    pk = _packed(a)
    want = goldenlib.expected_simulations(case)
    tgt = _lib.make_target(a['target_nodes'], a['target_code'], pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
        hit_index, hit_t, st = plan.target_run(0, pk.n_problems, a['max_t'], a['n_to_find'])
        assert st['n_found'] == case['expected']['n_hits_total']
        assert [int(i) for i in hit_index] == [w[0] for w in want]
        assert [int(t) for t in hit_t] == [len(w[1]) - 1 for w in want]
        # the reported trajectory of a hit = plain iteration up to the hit time (target.py:129-131)
        for (index, codes, _, _), t in list(zip(want, hit_t))[:16]:
            sub, _ = plan.simulate_list([index], int(t))
            assert [_lib.code_from_words(w) for w in sub[0]] == codes


# ---- against the C oracle on seeded inputs larger than the fixtures -------------------------------

def _oracle_table(cfg, max_t, max_a, idx_begin, idx_end):
    from oracle import coracle
    sp = coracle.Space(cfg['incoming node lists'], cfg['truth tables'],
                       cfg['origin simulation problem'], cfg['simulation problem variations'])
    rec = coracle.attract_records(sp, max_t, max_a, idx_begin, idx_end)
    table = coracle.attract_table(rec, sp.kw)
    rows = sorted(((k, v[0], v[1], v[2], v[3]) for k, v in table.items()), key=lambda r: (-r[1], r[0]))
    return rows, rec['steps'], int((rec['found'] == 0).sum())


@pytest.mark.parametrize('n_nodes,k,seed,n_free,max_t,max_a,lo,hi', [
    (64, 3, 0, 26, 1000, inf, 0, 1 << 13),              # the headline network, first 8192 problems
    (64, 3, 1, 26, 1000, inf, (1 << 25) + 5, (1 << 25) + 5 + 3000),   # unaligned range high up
    (64, 3, 2, 26, 300, 50, 12345, 12345 + 4096),
    (64, 2, 0, 20, inf, inf, 0, 1 << 14),               # ordered regime, no caps
    (32, 2, 5, 16, inf, inf, 0, 1 << 16),
    (48, 3, 7, 14, 4000, inf, 0, 1 << 12),
    (17, 4, 3, 17, inf, inf, 0, 1 << 15),
])
def test_attract_vs_c_oracle(n_nodes, k, seed, n_free, max_t, max_a, lo, hi):
    cfg = synth.random_attract_problem(n_nodes, k, seed, n_free=n_free)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                            cfg['origin simulation problem'], cfg['simulation problem variations'])
    want_rows, want_steps, want_without = _oracle_table(cfg, max_t, max_a, lo, hi)
    for opts in ({}, dict(smem_slots=3, min_blocks=1), dict(smem_slots=2, min_blocks=6)):
        with _lib.Plan(pk, _lib.MODE_ATTRACT, **opts) as plan:
            st = plan.attract_run(lo, hi, max_t, max_a)
            assert _table_rows(plan) == want_rows
            assert st['steps'] == want_steps
            assert st['n_without'] == want_without


def test_attract_random_free_subset_vs_c_oracle():
    rng = np.random.default_rng(11)
    free = sorted(int(x) for x in rng.choice(64, size=13, replace=False))
    cfg = synth.random_attract_problem(64, 3, 4, n_free=13, free_nodes=free)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                            cfg['origin simulation problem'], cfg['simulation problem variations'])
    want_rows, want_steps, want_without = _oracle_table(cfg, 700, inf, 0, 1 << 13)
    with _lib.Plan(pk, _lib.MODE_ATTRACT) as plan:
        st = plan.attract_run(0, 1 << 13, 700, inf)
        assert _table_rows(plan) == want_rows
        assert (st['steps'], st['n_without']) == (want_steps, want_without)


def test_target_vs_c_oracle_config5_shape():
    from oracle import coracle
    cfg = synth.random_target_problem(128, 4, 0, n_free=14, n_target_nodes=5)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    sp = coracle.Space(*args)
    rec = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 500)
    want = np.nonzero(rec['reached'])[0]
    pk = _lib.PackedProblem(*args)
    tgt = _lib.make_target(cfg['target node set'], cfg['target substate code'], pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
        hit_index, hit_t, st = plan.target_run(0, pk.n_problems, 500)
        assert [int(i) for i in hit_index] == [int(i) for i in want]
        assert [int(t) for t in hit_t] == [int(rec['stop_t'][i]) for i in want]


@pytest.mark.parametrize('opts', [dict(target_hit_list=1), dict(target_hit_list=7), dict(chunk_log2=9)],
                         ids=['chunk-by-chunk', 'hit-list-overflow', 'one-pass-many-chunks'])
def test_target_run_forms_agree(opts):
    """A finite max_t runs all chunks back to back into one hit list; chunk by chunk (the form of
    max_t = inf, and the fallback when the list overflows) must give the same hits, times and totals --
    and both the C oracle's."""
    from oracle import coracle
    cfg = synth.random_target_problem(128, 4, 0, n_free=13, n_target_nodes=4)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    sp = coracle.Space(*args)
    rec = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 300, 100, 8000)
    want = np.nonzero(rec['reached'])[0] + 100
    assert len(want) > 7
    pk = _lib.PackedProblem(*args)
    tgt = _lib.make_target(cfg['target node set'], cfg['target substate code'], pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
        _, _, base = plan.target_run(100, 8000, 300)                  # one chunk, one pass
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt, **dict(dict(chunk_log2=10), **opts)) as plan:
        hit_index, hit_t, st = plan.target_run(100, 8000, 300)
        assert [int(i) for i in hit_index] == [int(i) for i in want]
        assert [int(t) for t in hit_t] == [int(rec['stop_t'][i - 100]) for i in want]
        assert st['n_found'] == len(want)
        assert (st['steps'], st['n_without'], st['n_found']) == (base['steps'], base['n_without'], base['n_found'])
        few, _, st3 = plan.target_run(100, 8000, 300, 3)
        assert [int(i) for i in few] == [int(i) for i in want[:3]] and st3['n_found'] == len(want)


def test_no_device_plan_cannot_run():
    cfg = synth.random_attract_problem(8, 2, 0, n_free=4)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                            cfg['origin simulation problem'], cfg['simulation problem variations'])
    with _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True) as plan:
        with pytest.raises(_lib.BsiError):
            plan.attract_run(0, 16, 10, inf)


def test_two_shards_export_payload_and_merge_equal_the_whole():
    """The N>1 data path on one GPU: two plans take the two halves of the index space, export
    header + records to device buffers (no host round trip) and fold each other's records in."""
    import torch
    from boolsi_b200.problems import shard_range
    case = [c for c in goldenlib.cases('attract') if c['name'] == 'attract_rnd20k2s3_f12_tinf_ainf'][0]
    a = goldenlib.case_args(case)
    pk = _packed(a)
    want = [(k, n, s1, s2, len(cyc)) for k, n, s1, s2, cyc in goldenlib.expected_attractor_rows(case)]
    plans = [_lib.Plan(pk, _lib.MODE_ATTRACT) for _ in range(2)]
    bufs = [torch.zeros((257, 8), dtype=torch.int64, device='cuda') for _ in range(2)]
    stats = []
    for r, (plan, buf) in enumerate(zip(plans, bufs)):
        lo, hi = shard_range(pk.n_problems, r, 2)
        if r == 0:
            stats.append(plan.attract_run(lo, hi, a['max_t'], a['max_attractor_l']))
            plan.table_export_payload(buf.data_ptr(), buf.shape[0])
        else:
            # the asynchronous form: run and export are enqueued, statistics collected afterwards
            plan.attract_run_async(lo, hi, a['max_t'], a['max_attractor_l'])
            plan.table_export_payload(buf.data_ptr(), buf.shape[0])
            with pytest.raises(_lib.BsiError):
                plan.attract_run_async(lo, hi, a['max_t'], a['max_attractor_l'])    # one pending run per plan
            stats.append(plan.attract_run_finish())
    torch.cuda.synchronize()
    headers = [b[0].cpu().tolist() for b in bufs]
    for h, st, plan in zip(headers, stats, plans):
        assert h[:6] == [plan.table_count(), st['steps'], st['n_found'], st['n_without'], 0, 0]
    assert sum(h[1] for h in headers) == case['expected']['steps']
    plans[0].table_merge_device(bufs[1].data_ptr(), bufs[1].shape[0])     # header row is skipped (n = 0)
    plans[1].table_merge_device(bufs[0].data_ptr(), bufs[0].shape[0])
    for plan in plans:
        assert _table_rows(plan) == want
        plan.close()


def test_full_size_headline_job_properties():
    """BASELINE.json configs[3] at full size (2^26 initial states): size-independent properties.
    (i) conservation, (ii) the table of 8 shards run one after the other equals the table of one
    run, (iii) the table does not depend on which index bit drives which free node (lanes vs
    word bits), (iv) an oracle-checked slice is consistent with it."""
    import random
    cfg = synth.random_attract_problem(64, 3, 0, n_free=26)
    args = [cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations']]
    n = 1 << 26
    with _lib.Plan(_lib.PackedProblem(*args), _lib.MODE_ATTRACT) as plan:
        st = plan.attract_run(0, n, 1000, inf)
        whole = _table_rows(plan)
        assert st['n_found'] + st['n_without'] == n
        assert sum(r[1] for r in whole) == st['n_found']
        assert all(r[2] <= 1000 * r[1] for r in whole)           # every trajectory length <= max_t
        plan.table_clear()
        steps = 0
        from boolsi_b200.problems import shard_range
        for r in range(8):
            lo, hi = shard_range(n, r, 8)
            steps += plan.attract_run(lo, hi, 1000, inf)['steps']
        assert _table_rows(plan) == whole and steps == st['steps']
    shuffled = list(args[3][0])
    random.Random(5).shuffle(shuffled)
    args2 = args[:3] + [(shuffled, [], [])]
    with _lib.Plan(_lib.PackedProblem(*args2), _lib.MODE_ATTRACT) as plan:
        st2 = plan.attract_run(0, n, 1000, inf)
        assert _table_rows(plan) == whole and st2['steps'] == st['steps']


def test_full_size_target_job_properties():
    """BASELINE.json configs[4] shape at reduced count (2^24 of the 2^28 sampled initial states):
    N=128, K=4, fixed nodes + per-step perturbations, max_t=500.  Hits come back in index order,
    a sub-range run gives exactly the hits of that sub-range, and an oracle slice agrees."""
    from oracle import coracle
    cfg = synth.random_target_problem(128, 4, 0, n_free=24, n_target_nodes=6)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    pk = _lib.PackedProblem(*args)
    tgt = _lib.make_target(cfg['target node set'], cfg['target substate code'], pk.kw)
    with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
        hit_index, hit_t, st = plan.target_run(0, pk.n_problems, 500)
        assert st['n_found'] == len(hit_index)
        assert np.all(np.diff(hit_index.astype(np.int64)) > 0)                 # enumeration order
        assert np.all(hit_t <= 500)
        lo, hi = (1 << 23) + 77, (1 << 23) + 77 + 40000
        sub_index, sub_t, _ = plan.target_run(lo, hi, 500)
        sel = (hit_index >= lo) & (hit_index < hi)
        assert np.array_equal(sub_index, hit_index[sel]) and np.array_equal(sub_t, hit_t[sel])
        first5, _, st5 = plan.target_run(0, pk.n_problems, 500, 5)
        assert np.array_equal(first5, hit_index[:5]) and st5['n_found'] == st['n_found']
    sp = coracle.Space(*args)
    rec = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 500, lo, lo + 4096)
    want = np.nonzero(rec['reached'])[0] + lo
    got = hit_index[(hit_index >= lo) & (hit_index < lo + 4096)]
    assert np.array_equal(got.astype(np.int64), want.astype(np.int64))


def test_attractor_table_overflow_is_reported():
    """More attractors than table slots: BSI_E_CAPACITY, not a hang or a silent drop."""
    n = 14
    preds = [[i] for i in range(n)]                         # identity network: every state is a fixed point
    tts = [{(False,): False, (True,): True} for _ in range(n)]
    pk = _lib.PackedProblem(preds, tts, ([False] * n, {}, {}), (list(range(n)), [], []))
    with _lib.Plan(pk, _lib.MODE_ATTRACT, table_log2_slots=10) as plan:
        with pytest.raises(_lib.BsiError) as e:
            plan.attract_run(0, 1 << n, inf, inf)
        assert e.value.status == -4
    with _lib.Plan(pk, _lib.MODE_ATTRACT, table_log2_slots=16) as plan:
        st = plan.attract_run(0, 1 << n, inf, inf)
        assert plan.table_count() == 1 << n and st['n_found'] == 1 << n and st['steps'] == 1 << n


def test_engine_grows_its_table_instead_of_failing():
    """The per-rank engine (what attract_sharded and bench.py run) repeats the job with a larger table when
    the attractors do not fit, as the reference's store simply grows (attract.py:405-455)."""
    from boolsi_b200.attract import AttractEngine
    n = 14
    preds = [[i] for i in range(n)]
    tts = [{(False,): False, (True,): True} for _ in range(n)]
    eng = AttractEngine(preds, tts, ([False] * n, {}, {}), (list(range(n)), [], []), table_log2_slots=10)
    try:
        st = eng.run(inf, inf)
        assert eng.plan_opts['table_log2_slots'] >= 16
        assert eng.plan.table_count() == 1 << n and st['n_found'] == 1 << n and st['steps'] == 1 << n
    finally:
        eng.close()


def test_attract_long_cap_goes_through_several_rounds():
    """max_t beyond the first horizon in the chaotic regime: rounds 1024 -> cap (most lanes stay
    open), results equal to the oracle's."""
    cfg = synth.random_attract_problem(64, 3, 2, n_free=12)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                            cfg['origin simulation problem'], cfg['simulation problem variations'])
    want_rows, want_steps, want_without = _oracle_table(cfg, 3000, inf, 0, 1 << 12)
    with _lib.Plan(pk, _lib.MODE_ATTRACT) as plan:
        st = plan.attract_run(0, 1 << 12, 3000, inf)
        assert st['rounds'] >= 2
        assert _table_rows(plan) == want_rows
        assert (st['steps'], st['n_without']) == (want_steps, want_without)
