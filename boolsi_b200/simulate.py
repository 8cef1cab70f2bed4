"""`simulate` mode on the GPU path: the drop-in for reference `simulate_master`
(simulate.py:46-78) and the `Simulation` result record (simulate.py:16-28)."""
import logging

from . import _lib
from .db import init_simulation_db_structure
from .problems import ProblemDecoder


def count_perturbations(perturbed_nodes_by_t):
    """reference model.py:239-247"""
    return sum(len(nodes) for nodes in perturbed_nodes_by_t.values())


class Simulation:
    """reference simulate.py:16-28.  `states` may be given as a bool array [n_states, n_nodes]
    (what the device results unpack to): it is kept as `states_array` for the vectorised CSV
    writer, and `states` -- the reference's list of lists of bool -- is made on first use."""

    def __init__(self, states, fixed_nodes, perturbed_nodes_by_t):
        if hasattr(states, 'tolist'):
            self.states_array = states
            self._states = None
        else:
            self.states_array = None
            self._states = states
        self.fixed_nodes = fixed_nodes
        self.perturbed_nodes_by_t = perturbed_nodes_by_t
        self.n_perturbations = count_perturbations(perturbed_nodes_by_t)

    @property
    def states(self):
        if self._states is None:
            self._states = self.states_array.tolist()
        return self._states


def states_from_words(rows, n_nodes):
    """-> bool array [n_states, n_nodes]"""
    return _lib.state_array_from_word_rows(rows, n_nodes)


def store_simulations(db_conn, simulations):
    """Post-conditions of write_simulations_to_db (simulate.py:149-178) for one batch."""
    root = db_conn.root
    for i, simulation in enumerate(simulations):
        root.simulations[(0, i)] = simulation
    if simulations:
        root.n_simulations.change(len(simulations))
        root.n_simulation_batches.change(1)
        root.simulation_batch_sizes_lcm.set(len(simulations))


def simulate_master(mpi_comm, n_simulation_problem_batches_per_process, origin_simulation_problem,
                    simulation_problem_variations, predecessor_node_lists, truth_tables, max_t,
                    n_simulation_problems, db_conn, output_directory, device=0):
    """Drop-in for reference `simulate_master`: every problem simulated for `max_t` steps, results
    in `db_conn.root.simulations` in enumeration order."""
    logging.getLogger().info('GPU path will be used to perform {} simulations...'.format(
        n_simulation_problems))
    if not hasattr(db_conn.root, 'simulations'):
        init_simulation_db_structure(db_conn)
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    n_nodes = packed.n_nodes
    with _lib.Plan(packed, _lib.MODE_SIMULATE, device=device) as plan:
        states, _ = plan.simulate_run(0, packed.n_problems, max_t)
    simulations = []
    chunk = 1024                       # unpack many trajectories per numpy call, bounded memory
    decoder = ProblemDecoder(origin_simulation_problem, simulation_problem_variations)
    for begin in range(0, packed.n_problems, chunk):
        words = states[begin:begin + chunk]
        bools = states_from_words(words.reshape(-1, words.shape[-1]), n_nodes).reshape(
            words.shape[0], words.shape[1], n_nodes)
        for k in range(words.shape[0]):
            fixed, perturbed = decoder.masks(begin + k)
            simulations.append(Simulation(bools[k], fixed, perturbed))
    store_simulations(db_conn, simulations)
