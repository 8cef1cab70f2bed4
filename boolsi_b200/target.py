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


def target_sharded(predecessor_node_lists, truth_tables, origin_simulation_problem,
                   simulation_problem_variations, target_node_set, target_substate_code, max_t=inf,
                   n_to_find=inf, device=0, rank=0, world=1, **plan_opts):
    """The `target` job sharded like `attract`: rank r searches the contiguous, word-aligned slice
    r of the index space on its GPU; the per-rank hit lists (at most `n_to_find` each) are
    exchanged with one all_gather and merged in index order on every rank.
    :return: dict(hit_index, hit_t, n_hits_total, steps_total, stats)"""
    from .problems import shard_range
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    tgt = _lib.make_target(target_node_set, target_substate_code, packed.kw)
    lo, hi = shard_range(packed.n_problems, rank, world)
    with _lib.Plan(packed, _lib.MODE_TARGET, device=device, target=tgt, **plan_opts) as plan:
        hit_index, hit_t, st = plan.target_run(lo, hi, max_t, n_to_find)
    mine = (hit_index, hit_t, st['n_found'], st['steps'])
    if world > 1:
        import torch.distributed as dist
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    else:
        everyone = [mine]
    index, t, total = merge_sharded_hits(everyone, n_to_find)
    return {'hit_index': index, 'hit_t': t, 'n_hits_total': total,
            'steps_total': sum(int(r[3]) for r in everyone), 'stats': st}


def target_master(mpi_comm, n_simulation_problem_batches_per_process, origin_simulation_problem,
                  simulation_problem_variations, target_substate_code, target_node_set,
                  predecessor_node_lists, truth_tables, n_simulations_to_reach_target_substate,
                  max_t, n_simulation_problems, db_conn, output_dirpath, device=0):
    """Drop-in for reference `target_master`: keeps, in enumeration order, the first
    `n_simulations_to_reach_target_substate` simulations that reach the target substate after
    their last perturbation; each as the trajectory up to the first hit (target.py:129-131)."""
    log = logging.getLogger()
    if n_simulations_to_reach_target_substate is not inf and \
            n_simulations_to_reach_target_substate > n_simulation_problems:
        log.warning('Requested {} simulations that reach target state, but only {} simulation '
                    'problems provided. Will look for all simulations that reach target state.'.format(
                        n_simulations_to_reach_target_substate, n_simulation_problems))
        n_simulations_to_reach_target_substate = inf
    if not hasattr(db_conn.root, 'simulations'):
        init_simulation_db_structure(db_conn)
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    tgt = _lib.make_target(target_node_set, target_substate_code, packed.kw)
    simulations = []
    with _lib.Plan(packed, _lib.MODE_TARGET, device=device, target=tgt) as plan:
        hit_index, hit_t, _ = plan.target_run(0, packed.n_problems, max_t,
                                              n_simulations_to_reach_target_substate)
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
