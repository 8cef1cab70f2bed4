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
    """reference simulate.py:16-28.  `states` may be given as the reference's list of lists of bool, as
    a bool array [n_states, n_nodes], or PACKED -- `(uint64 words [n_states, key_words], n_nodes)`,
    what the device returns.  Packed states stay packed (8 bytes per 64 nodes and time step) until a
    consumer asks: `states_array` (bool array, what the vectorised CSV writer reads) and `states`
    (list of lists, what the reference's writers read) are made on first use."""

    def __init__(self, states, fixed_nodes, perturbed_nodes_by_t):
        self._words = self._array = self._states = None
        if isinstance(states, tuple):
            self._words, self._n_nodes = states
        elif hasattr(states, 'tolist'):
            self._array = states
        else:
            self._states = states
        self.fixed_nodes = fixed_nodes
        self.perturbed_nodes_by_t = perturbed_nodes_by_t
        self.n_perturbations = count_perturbations(perturbed_nodes_by_t)

    @property
    def states_array(self):
        if self._array is None and self._words is not None:
            self._array = _lib.state_array_from_word_rows(self._words, self._n_nodes)
        return self._array

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


#: upper bound on the packed states one device call of `simulate_master` brings back (bytes)
SIMULATE_CHUNK_BYTES = 256 << 20


def simulate_master(mpi_comm, n_simulation_problem_batches_per_process, origin_simulation_problem,
                    simulation_problem_variations, predecessor_node_lists, truth_tables, max_t,
                    n_simulation_problems, db_conn, output_directory, devices=None, device=None):
    """Drop-in for reference `simulate_master`: every problem simulated for `max_t` steps, results
    in `db_conn.root.simulations` in enumeration order.  The job goes to the device(s) in chunks of
    bounded size (the reference's batches, mpi.py:131-190), every visible GPU taking a slice of each;
    trajectories stay packed in the result records until a writer reads them."""
    from .attract import cached_group
    logging.getLogger().info('GPU path will be used to perform {} simulations...'.format(
        n_simulation_problems))
    if not hasattr(db_conn.root, 'simulations'):
        init_simulation_db_structure(db_conn)
    if device is not None:
        devices = [device]
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    n_nodes = packed.n_nodes
    group = cached_group(packed, _lib.MODE_SIMULATE, devices)
    per_problem = (int(max_t) + 1) * packed.kw * 8
    chunk = max(32, (SIMULATE_CHUNK_BYTES // per_problem) // 32 * 32)
    decoder = ProblemDecoder(origin_simulation_problem, simulation_problem_variations)
    simulations = []
    for begin in range(0, packed.n_problems, chunk):
        end = min(packed.n_problems, begin + chunk)
        states, _ = group.simulate_run(begin, end, max_t)
        for k in range(end - begin):
            fixed, perturbed = decoder.masks(begin + k)
            simulations.append(Simulation((states[k], n_nodes), fixed, perturbed))
    store_simulations(db_conn, simulations)
