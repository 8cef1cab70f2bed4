"""`target` mode on the GPU path: the drop-in for reference `target_master` (target.py:13-88)."""
import logging
from math import inf

from . import _lib
from .db import init_simulation_db_structure
from .problems import decode_simulation_problem
from .simulate import Simulation, states_from_words, store_simulations


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
        for pos, index in enumerate(hit_index):
            _, fixed, perturbed = decode_simulation_problem(int(index), origin_simulation_problem,
                                                            simulation_problem_variations)
            simulations.append(Simulation(states_by_pos[pos], fixed, perturbed))
    store_simulations(db_conn, simulations)
    log.info('{} simulations reach target state.'.format(db_conn.root.n_simulations()))
    return db_conn.root.n_simulations()
