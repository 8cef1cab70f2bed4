"""CPU tests of the host-side logic: problem decoding, sharding, the result store, and the
world_size-2 table exchange over gloo (the N>1 path without GPUs)."""
import os
import sys
from math import inf

import numpy as np
import pytest

import goldenlib
from boolsi_b200 import problems, synth
from boolsi_b200.db import ResultStore, init_attractor_db_structure, init_simulation_db_structure
from oracle import coracle, oracle


@pytest.mark.parametrize('case', [c for c in goldenlib.cases() if c['n_problems'] <= 1152],
                         ids=lambda c: c['name'])
def test_decode_matches_oracle(case):
    a = goldenlib.case_args(case)
    n = problems.count_simulation_problems(a['variations'])
    assert n == case['n_problems']
    for index in sorted({0, 1, n // 3, n // 2, n - 1}):
        state, fixed, pert = problems.decode_simulation_problem(index, a['origin'], a['variations'])
        code, ofixed, opert = oracle.decode_problem(index, a['origin'], a['variations'])
        assert sum(1 << i for i, b in enumerate(state) if b) == code
        assert fixed == ofixed and pert == opert
        assert problems.state_from_code(code, a['n_nodes']) == state


@pytest.mark.parametrize('n,world', [(8, 1), (8, 2), (2 ** 26, 8), (1000, 3), (31, 4), (33, 2), (0, 2)])
def test_shard_range_partitions(n, world):
    spans = [problems.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(lo % 32 == 0 for lo, _ in spans)


def test_result_store_surface():
    db = ResultStore()
    init_attractor_db_structure(db)
    init_simulation_db_structure(db)
    db.root.aggregated_attractors[(-3, 9)] = 'b'
    db.root.aggregated_attractors[(-7, 2)] = 'a'
    assert list(db.root.aggregated_attractors.values()) == ['a', 'b']     # BTree order
    db.root.total_frequency.change(10)
    db.root.n_aggregated_attractors.set(2)
    assert db.root.total_frequency() == 10 and db.root.n_aggregated_attractors() == 2
    db.cacheMinimize()


def _records_from_table(table, kw):
    rec = np.zeros((len(table), 8), np.int64)
    for i, (key, (n, s1, s2, lam)) in enumerate(sorted(table.items())):
        for j in range(kw):
            rec[i, j] = np.array((key >> (64 * j)) & (2 ** 64 - 1), np.uint64).astype(np.int64)
        rec[i, 4:] = [n, s1, s2, lam]
    return rec


def _gloo_worker(rank, world, port, n_free, capacity, out_dir):
    import torch
    import torch.distributed as dist
    from boolsi_b200.sharding import exchange_tables
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg = synth.random_attract_problem(20, 2, 3, n_free=n_free)
    sp = coracle.Space(cfg['incoming node lists'], cfg['truth tables'],
                       cfg['origin simulation problem'], cfg['simulation problem variations'])
    lo, hi = problems.shard_range(sp.n_problems, rank, world)
    rec = coracle.attract_records(sp, inf, inf, lo, hi, n_threads=1)
    mine = _records_from_table(coracle.attract_table(rec, sp.kw), sp.kw)

    def fill(buf):
        m = min(len(mine), buf.shape[0] - 1)
        buf.zero_()
        buf[0, :4] = torch.tensor([len(mine), rec['steps'], int(rec['found'].sum()), 0])
        buf[1:1 + m].copy_(torch.from_numpy(mine[:m]))

    everyone, headers = exchange_tables(fill, world, 'cpu', capacity=capacity)
    merged = {}
    for r in range(world):
        for row in everyone[r, 1:1 + int(headers[r, 0])].numpy():
            key = int(np.array(row[0], np.int64).astype(np.uint64))
            cur = merged.setdefault(key, [0, 0, 0, int(row[7])])
            cur[0] += int(row[4]); cur[1] += int(row[5]); cur[2] += int(row[6])
    np.save(os.path.join(out_dir, 'r%d.npy' % rank),
            np.array([[k] + v for k, v in sorted(merged.items())] + [[-1, int(headers[:, 1].sum()), 0, 0, 0]],
                     dtype=object), allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.parametrize('capacity', [4096, 1])     # 1 forces the overflow-and-retry protocol
def test_two_rank_exchange_over_gloo(tmp_path, capacity):
    import torch.multiprocessing as mp
    n_free, world = 10, 2
    port = 29500 + (os.getpid() + capacity) % 2000
    mp.spawn(_gloo_worker, args=(world, port, n_free, capacity, str(tmp_path)), nprocs=world, join=True)
    cfg = synth.random_attract_problem(20, 2, 3, n_free=n_free)
    sp = coracle.Space(cfg['incoming node lists'], cfg['truth tables'],
                       cfg['origin simulation problem'], cfg['simulation problem variations'])
    rec = coracle.attract_records(sp, inf, inf)
    want = [[k] + v for k, v in sorted(coracle.attract_table(rec, sp.kw).items())]
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), 'r%d.npy' % r), allow_pickle=True).tolist()
        assert got[:-1] == want                       # every rank holds the complete, identical table
        assert got[-1][1] == rec['steps']


@pytest.mark.parametrize('n_to_find', [inf, 1, 7, 10 ** 9])
def test_sharded_target_hits_merge_in_enumeration_order(n_to_find):
    """The N>1 `target` path without GPUs: per-shard hit lists (from the C oracle) merged by
    `merge_sharded_hits` equal the single-range answer, for every rank count."""
    from boolsi_b200.target import merge_sharded_hits
    cfg = synth.random_target_problem(24, 3, 1, n_free=11, n_fixed=2, n_perturbed=2, n_perturbation_times=3,
                                      max_perturbation_t=9, n_target_nodes=3)
    sp = coracle.Space(cfg['incoming node lists'], cfg['truth tables'],
                       cfg['origin simulation problem'], cfg['simulation problem variations'])
    rec = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 60)
    want = np.nonzero(rec['reached'])[0]
    assert len(want) > 10
    cap = len(want) if n_to_find == inf else min(len(want), n_to_find)
    for world in (1, 2, 3, 8):
        per_rank = []
        for r in range(world):
            lo, hi = problems.shard_range(sp.n_problems, r, world)
            sub = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 60, lo, hi)
            hits = np.nonzero(sub['reached'])[0]
            keep = hits if n_to_find == inf else hits[:n_to_find]
            per_rank.append((keep + lo, sub['stop_t'][keep], len(hits)))
        index, t, total = merge_sharded_hits(per_rank, n_to_find)
        assert total == len(want)
        assert [int(i) for i in index] == [int(i) for i in want[:cap]]
        assert [int(x) for x in t] == [int(rec['stop_t'][i]) for i in want[:cap]]


def test_problem_decoder_masks_equal_the_full_decode():
    """ProblemDecoder (what the masters use for the Simulation records) against
    decode_simulation_problem on every problem of spaces with all NodeStateRange kinds."""
    from boolsi_b200.constants import NodeStateRange as R
    from boolsi_b200.problems import ProblemDecoder, count_simulation_problems, decode_simulation_problem
    origin = ([False, True, False, True, False, False], {4: True}, {3: {1: False}, 7: {0: True}})
    spaces = [
        ([0, 2, 5], [], []),
        ([], [(3, R.MAYBE_TRUE)], [(3, 2, R.MAYBE_TRUE_OR_FALSE)]),
        ([1, 2], [(3, R.TRUE_OR_FALSE), (5, R.MAYBE_FALSE)], [(3, 1, R.MAYBE_TRUE_OR_FALSE), (9, 0, R.TRUE_OR_FALSE),
                                                            (7, 0, R.MAYBE_FALSE)]),
        ([], [], []),
    ]
    for variations in spaces:
        decoder = ProblemDecoder(origin, variations)
        for index in range(count_simulation_problems(variations)):
            _, fixed, perturbed = decode_simulation_problem(index, origin, variations)
            assert decoder.masks(index) == (fixed, perturbed), (variations, index)


def _gloo_hits_worker(rank, world, port, n_to_find, out_dir):
    import torch.distributed as dist
    from boolsi_b200.target import exchange_hits, merge_sharded_hits
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg = synth.random_target_problem(24, 3, 1, n_free=11, n_fixed=2, n_perturbed=2, n_perturbation_times=3,
                                      max_perturbation_t=9, n_target_nodes=3)
    sp = coracle.Space(cfg['incoming node lists'], cfg['truth tables'],
                       cfg['origin simulation problem'], cfg['simulation problem variations'])
    lo, hi = problems.shard_range(sp.n_problems, rank, world)
    if rank == 1:
        hi = lo                     # an empty shard: a rank with no hits still takes part
    sub = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 60, lo, hi)
    hits = np.nonzero(sub['reached'])[0]
    keep = hits if n_to_find == inf else hits[:n_to_find]
    everyone = exchange_hits((keep + lo).astype(np.uint64), sub['stop_t'][keep], len(hits), sub['steps'], world, 'cpu')
    index, t, total = merge_sharded_hits(everyone, n_to_find)
    np.save(os.path.join(out_dir, 'h%d.npy' % rank), np.array([index.tolist(), t.tolist(), [total]], dtype=object),
            allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.parametrize('n_to_find', [inf, 5])
def test_three_rank_hit_exchange_over_gloo(tmp_path, n_to_find):
    """`target_sharded`'s exchange: fixed-size integer records through all_gather_into_tensor (no pickled
    objects), ragged per-rank counts and an empty rank included; every rank ends with the same list."""
    import torch.multiprocessing as mp
    world = 3
    port = 31500 + (os.getpid() + (0 if n_to_find == inf else 7)) % 2000
    mp.spawn(_gloo_hits_worker, args=(world, port, n_to_find, str(tmp_path)), nprocs=world, join=True)
    cfg = synth.random_target_problem(24, 3, 1, n_free=11, n_fixed=2, n_perturbed=2, n_perturbation_times=3,
                                      max_perturbation_t=9, n_target_nodes=3)
    sp = coracle.Space(cfg['incoming node lists'], cfg['truth tables'],
                       cfg['origin simulation problem'], cfg['simulation problem variations'])
    want_idx, want_t = [], []
    for r in (0, 2):                                   # rank 1's shard was left out on purpose
        lo, hi = problems.shard_range(sp.n_problems, r, world)
        sub = coracle.target_records(sp, cfg['target node set'], cfg['target substate code'], 60, lo, hi)
        hits = np.nonzero(sub['reached'])[0]
        want_idx += [int(i) + lo for i in hits]
        want_t += [int(sub['stop_t'][i]) for i in hits]
    total = len(want_idx)
    if n_to_find != inf:
        # each rank keeps its first n, the merge keeps the first n overall
        want_idx, want_t = want_idx[:n_to_find], want_t[:n_to_find]
    for r in range(world):
        index, t, tot = np.load(os.path.join(str(tmp_path), 'h%d.npy' % r), allow_pickle=True).tolist()
        assert [int(i) for i in index] == want_idx and [int(x) for x in t] == want_t and tot == [total]


def test_aggregated_attractor_keeps_packed_states_until_asked():
    """The device returns state codes; the record converts them to the reference's list of lists of
    bool (attract.py:18-45 consumers) only when `states` is read, and both forms agree."""
    from boolsi_b200 import _lib
    from boolsi_b200.attract import AggregatedAttractor
    n_nodes = 70
    codes = [0x3FFFFFFFFFFFFFFFFF, 5, (1 << 69) | 1]
    words = np.array([_lib.words_from_code(c, 2) for c in codes], np.uint64)
    packed = AggregatedAttractor(5, (words, n_nodes), 4, 10, 30)
    assert packed._states is None and packed.n_states == 3
    want = [[bool((c >> i) & 1) for i in range(n_nodes)] for c in codes]
    assert packed.states == want and packed.states is packed.states
    assert packed.states_array.shape == (3, n_nodes) and packed.states_array.tolist() == want
    plain = AggregatedAttractor(5, want, 4, 10, 30)
    assert plain.n_states == 3 and plain.states is want and plain.states_array.tolist() == want
    assert (packed.trajectory_l_mean, packed.trajectory_l_variation_sum) == (2.5, 5.0)


def test_small_jobs_stay_on_one_device(monkeypatch):
    """cached_group: no device list + a job of at most SMALL_JOB_PROBLEMS problems -> device 0 only (more
    devices cannot shorten a job of less than one wave and each costs a CUDA context); larger jobs and
    explicit lists are passed through."""
    from boolsi_b200 import _lib, attract
    seen = []

    class FakeGroup:
        def __init__(self, packed, mode, devices=None, target=None, **opts):
            seen.append(devices)

        def close(self):
            pass
    monkeypatch.setattr(_lib, 'Group', FakeGroup)
    attract.release_cached_plans()
    try:
        small = synth.random_attract_problem(64, 3, 0, n_free=10)
        big = synth.random_attract_problem(64, 3, 0, n_free=21)
        for cfg, devices in ((small, None), (big, None), (small, [0, 1])):
            pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                                    cfg['origin simulation problem'], cfg['simulation problem variations'])
            attract.cached_group(pk, _lib.MODE_ATTRACT, devices)
        assert seen == [[0], None, [0, 1]]
    finally:
        attract._GROUPS.clear()
