"""`target` mode on the GPU path: the drop-in for reference `target_master` (target.py:13-88)."""
import logging
from math import inf

from . import _lib
from .db import init_simulation_db_structure
from .problems import ProblemDecoder
from .simulate import Simulation, states_from_words, store_simulations


def merge_sharded_hits(per_rank, n_to_find=inf):
    """Hits of contiguous ascending index ranges, one entry per rank in rank order:
    [(hit_index array, hit_t array, total hits of the rank), ...] -> (hit_index, hit_t, total).
    Ranges are ascending and each rank's hits are in index order, so concatenation in rank order IS
    the enumeration order the reference reports in (simulate.py:149-178); `n_to_find` keeps the
    first n (simulate.py:163-164)."""
    import numpy as np
    index = np.concatenate([np.asarray(r[0], dtype=np.uint64) for r in per_rank]) if per_rank else np.zeros(0, np.uint64)
    t = np.concatenate([np.asarray(r[1], dtype=np.uint64) for r in per_rank]) if per_rank else np.zeros(0, np.uint64)
    total = sum(int(r[2]) for r in per_rank)
    if n_to_find is not None and n_to_find != inf:
        index, t = index[:int(n_to_find)], t[:int(n_to_find)]
    return index, t, total


def target_devices(predecessor_node_lists, truth_tables, origin_simulation_problem,
                   simulation_problem_variations, target_node_set, target_substate_code, max_t=inf,
                   n_to_find=inf, devices=None, **plan_opts):
    """The `target` job from ONE process: the library's group run splits the index space over
    `devices` (None = every visible GPU) and returns the hits in enumeration order.
    :return: dict(hit_index, hit_t, n_hits_total, steps_total, stats, group)"""
    from .attract import cached_group
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    tgt = _lib.make_target(target_node_set, target_substate_code, packed.kw)
    group = cached_group(packed, _lib.MODE_TARGET, devices, target=tgt, **plan_opts)
    hit_index, hit_t, st = group.target_run(0, packed.n_problems, max_t, n_to_find)
    return {'hit_index': hit_index, 'hit_t': hit_t, 'n_hits_total': st['n_found'], 'steps_total': st['steps'],
            'stats': st, 'group': group, 'packed': packed}


HIT_RECORD_WORDS = 2          # (problem index, hit time)


def target_sharded(predecessor_node_lists, truth_tables, origin_simulation_problem,
                   simulation_problem_variations, target_node_set, target_substate_code, max_t=inf,
                   n_to_find=inf, device=0, rank=0, world=1, **plan_opts):
    """One rank of a `torch.distributed` job: rank r searches the contiguous, word-aligned slice r of the
    index space on its GPU; the ranks exchange FIXED-SIZE integer records -- one all_gather of the
    per-rank totals, one of `(index, t_hit)` rows padded to the largest count -- and every rank merges
    them in index order (the replacement of the pickled result lists of mpi.py:296-332).
    :return: dict(hit_index, hit_t, n_hits_total, steps_total, stats)"""
    import numpy as np
    from .problems import shard_range
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    tgt = _lib.make_target(target_node_set, target_substate_code, packed.kw)
    lo, hi = shard_range(packed.n_problems, rank, world)
    with _lib.Plan(packed, _lib.MODE_TARGET, device=device, target=tgt, **plan_opts) as plan:
        hit_index, hit_t, st = plan.target_run(lo, hi, max_t, n_to_find)
    if world > 1:
        everyone = exchange_hits(hit_index, hit_t, st['n_found'], st['steps'], world,
                                 'cuda:%d' % device)
    else:
        everyone = [(hit_index, hit_t, st['n_found'], st['steps'])]
    index, t, total = merge_sharded_hits(everyone, n_to_find)
    return {'hit_index': index, 'hit_t': t, 'n_hits_total': total,
            'steps_total': sum(int(r[3]) for r in everyone), 'stats': st}


def exchange_hits(hit_index, hit_t, n_found, steps, world, device):
    """all_gather of one rank's hits as integer tensors: first the header (rows kept, total hits,
    trajectory-steps), then the rows padded to the largest count.  Works on NCCL (device tensors) and
    gloo (`device='cpu'`).  -> [(hit_index, hit_t, n_found, steps)] in rank order."""
    import numpy as np
    import torch
    import torch.distributed as dist
    head = torch.tensor([len(hit_index), int(n_found), int(steps)], dtype=torch.int64, device=device)
    heads = torch.zeros((world, 3), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(heads.view(world * 3), head)
    heads = heads.cpu().numpy()
    width = int(heads[:, 0].max())
    out = []
    rows = None
    if width:
        mine = torch.zeros((width, HIT_RECORD_WORDS), dtype=torch.int64, device=device)
        if len(hit_index):
            pair = np.stack([np.asarray(hit_index, np.uint64), np.asarray(hit_t, np.uint64)], axis=1).view(np.int64)
            mine[:len(hit_index)] = torch.from_numpy(pair).to(device)
        rows = torch.zeros((world, width, HIT_RECORD_WORDS), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(rows.view(world * width, HIT_RECORD_WORDS), mine)
        rows = rows.cpu().numpy().view(np.uint64)
    for r in range(world):
        k = int(heads[r, 0])
        idx = rows[r, :k, 0].copy() if k else np.zeros(0, np.uint64)
        tt = rows[r, :k, 1].copy() if k else np.zeros(0, np.uint64)
        out.append((idx, tt, int(heads[r, 1]), int(heads[r, 2])))
    return out


def target_master(mpi_comm, n_simulation_problem_batches_per_process, origin_simulation_problem,
                  simulation_problem_variations, target_substate_code, target_node_set,
                  predecessor_node_lists, truth_tables, n_simulations_to_reach_target_substate,
                  max_t, n_simulation_problems, db_conn, output_dirpath, devices=None, device=None):
    """Drop-in for reference `target_master`: keeps, in enumeration order, the first
    `n_simulations_to_reach_target_substate` simulations that reach the target substate after
    their last perturbation; each as the trajectory up to the first hit (target.py:129-131).
    The search runs on every visible GPU from this one process (`devices` / `device` restrict it)."""
    log = logging.getLogger()
    if n_simulations_to_reach_target_substate is not inf and \
            n_simulations_to_reach_target_substate > n_simulation_problems:
        log.warning('Requested {} simulations that reach target state, but only {} simulation '
                    'problems provided. Will look for all simulations that reach target state.'.format(
                        n_simulations_to_reach_target_substate, n_simulation_problems))
        n_simulations_to_reach_target_substate = inf
    if not hasattr(db_conn.root, 'simulations'):
        init_simulation_db_structure(db_conn)
    if device is not None:
        devices = [device]
    res = target_devices(predecessor_node_lists, truth_tables, origin_simulation_problem,
                         simulation_problem_variations, target_node_set, target_substate_code, max_t,
                         n_simulations_to_reach_target_substate, devices=devices)
    packed, hit_index, hit_t = res['packed'], res['hit_index'], res['hit_t']
    plan = res['group'].plan(0)
    simulations = []
    # trajectories of the reported hits: plain iteration up to the hit time (target.py:129-131),
    # regenerated in batches -- one launch per power-of-two class of hit times
    states_by_pos = [None] * len(hit_index)
    groups = {}
    for pos, t in enumerate(hit_t):
        groups.setdefault(int(t).bit_length(), []).append(pos)
    for _, members in sorted(groups.items()):
        t_max = max(int(hit_t[pos]) for pos in members)
        rows, _ = plan.simulate_list([int(hit_index[pos]) for pos in members], t_max)
        for row, pos in zip(rows, members):
            states_by_pos[pos] = states_from_words(row[:int(hit_t[pos]) + 1], packed.n_nodes)
    decoder = ProblemDecoder(origin_simulation_problem, simulation_problem_variations)
    for pos, index in enumerate(hit_index):
        fixed, perturbed = decoder.masks(int(index))
        simulations.append(Simulation(states_by_pos[pos], fixed, perturbed))
    store_simulations(db_conn, simulations)
    log.info('{} simulations reach target state.'.format(db_conn.root.n_simulations()))
    return db_conn.root.n_simulations()
