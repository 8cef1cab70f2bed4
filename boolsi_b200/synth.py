"""Synthetic networks for the benchmark configurations (BASELINE.json configs 4-5).

The reference ships no generator; these are classic Kauffman NK nets expressed in the exact
structures `parse_input` hands to the `*_master` functions (reference `input.py:124-157`):
`predecessor_node_lists[i]` sorted ascending (`input.py:796`) and `truth_tables[i]` a dict keyed
by `itertools.product((False, True), repeat=K)` tuples (`input.py:881`).
"""
import itertools

import numpy as np


def random_network(n_nodes, k, seed):
    """Random NK network: K distinct predecessors per node, fair-coin truth tables.

    :return: (predecessor_node_lists, truth_tables) in the reference's format
    """
    rng = np.random.default_rng(seed)
    predecessor_node_lists = []
    truth_tables = []
    for _ in range(n_nodes):
        preds = sorted(int(p) for p in rng.choice(n_nodes, size=min(k, n_nodes), replace=False))
        bits = rng.integers(0, 2, size=2 ** len(preds))
        predecessor_node_lists.append(preds)
        truth_tables.append({states: bool(bits[j]) for j, states in enumerate(
            itertools.product((False, True), repeat=len(preds)))})
    return predecessor_node_lists, truth_tables


def random_attract_problem(n_nodes, k, seed, n_free, free_nodes=None):
    """Config-4 style `attract` input: `n_free` nodes are `any`, the rest drawn once.

    :return: dict with the same keys `parse_input` produces (the ones the path consumes)
    """
    predecessor_node_lists, truth_tables = random_network(n_nodes, k, seed)
    rng = np.random.default_rng([seed, 0xB001])
    initial_state = [bool(b) for b in rng.integers(0, 2, size=n_nodes)]
    if free_nodes is None:
        free_nodes = list(range(n_free))
    else:
        free_nodes = sorted(int(f) for f in free_nodes)
    for node in free_nodes:
        initial_state[node] = False
    return {
        'node names': ['n%d' % i for i in range(n_nodes)],
        'incoming node lists': predecessor_node_lists,
        'truth tables': truth_tables,
        'origin simulation problem': (initial_state, dict(), dict()),
        'simulation problem variations': (free_nodes, [], []),
        'total combination count': 2 ** len(free_nodes),
        'target substate code': None,
        'target node set': None,
    }


def random_target_problem(n_nodes, k, seed, n_free, n_fixed=4, n_perturbed=4, n_perturbation_times=8,
                          max_perturbation_t=100, n_target_nodes=16):
    """Config-5 style `target` input: constant fixed nodes and perturbations, a partial target."""
    predecessor_node_lists, truth_tables = random_network(n_nodes, k, seed)
    rng = np.random.default_rng([seed, 0xB002])
    initial_state = [bool(b) for b in rng.integers(0, 2, size=n_nodes)]
    free_nodes = list(range(n_free))
    for node in free_nodes:
        initial_state[node] = False
    others = [int(x) for x in rng.permutation(np.arange(n_free, n_nodes))]
    fixed_nodes = {node: bool(rng.integers(0, 2)) for node in others[:n_fixed]}
    perturbed = others[n_fixed:n_fixed + n_perturbed]
    times = sorted(int(t) for t in rng.choice(np.arange(1, max_perturbation_t + 1),
                                             size=n_perturbation_times, replace=False))
    perturbed_nodes_by_t = {}
    for t in times:
        node = int(perturbed[int(rng.integers(0, len(perturbed)))])
        perturbed_nodes_by_t.setdefault(t, {})[node] = bool(rng.integers(0, 2))
    target_nodes = sorted(others[n_fixed + n_perturbed:n_fixed + n_perturbed + n_target_nodes])
    target_code = 0
    for node in target_nodes:
        if rng.integers(0, 2):
            target_code += 1 << node
    return {
        'node names': ['n%d' % i for i in range(n_nodes)],
        'incoming node lists': predecessor_node_lists,
        'truth tables': truth_tables,
        'origin simulation problem': (initial_state, fixed_nodes, perturbed_nodes_by_t),
        'simulation problem variations': (free_nodes, [], []),
        'total combination count': 2 ** len(free_nodes),
        'target substate code': target_code,
        'target node set': set(target_nodes),
    }
