"""Host-side view of the problem enumeration (reference batching.py:10-46, 160-229).

The device decodes problem indices itself; the host only needs the same decoding for the handful
of problems it reports (the `fixed_nodes` / `perturbed_nodes_by_t` of a `Simulation` record) and
for splitting the index space across GPUs.
"""
from .constants import NODE_STATE_BY_DIGIT, range_name, range_radix


def problem_space_radices(simulation_problem_variations):
    """One digit per variation, least significant first (batching.py:10-46)."""
    var_initial, var_fixed, var_perturbed = simulation_problem_variations
    return ([2] * len(var_initial) + [range_radix(r) for _, r in var_fixed] +
            [range_radix(r) for _, _, r in var_perturbed])


def count_simulation_problems(simulation_problem_variations):
    n = 1
    for radix in problem_space_radices(simulation_problem_variations):
        n *= radix
    return n


def decode_simulation_problem(index, origin_simulation_problem, simulation_problem_variations):
    """Problem `index` of the single-process enumeration order (batching.py:212-229 + 160-209).

    :return: (initial_state list[bool], fixed_nodes dict, perturbed_nodes_by_t dict)
    """
    initial_state, fixed_nodes, perturbed_nodes_by_t = origin_simulation_problem
    var_initial, var_fixed, var_perturbed = simulation_problem_variations
    digits = []
    for radix in problem_space_radices(simulation_problem_variations):
        index, digit = divmod(index, radix)
        digits.append(digit)
    state = list(initial_state)
    pos = 0
    for node in var_initial:
        state[node] = bool(digits[pos])
        pos += 1
    fixed = dict(fixed_nodes)
    for node, node_state_range in var_fixed:
        value = NODE_STATE_BY_DIGIT[range_name(node_state_range)][digits[pos]]
        if value is not None:
            fixed[node] = value
        pos += 1
    perturbed = {t: dict(nodes) for t, nodes in perturbed_nodes_by_t.items()}
    for t, node, node_state_range in var_perturbed:
        value = NODE_STATE_BY_DIGIT[range_name(node_state_range)][digits[pos]]
        if value is not None:
            perturbed.setdefault(t, {})[node] = value
        pos += 1
    return state, fixed, perturbed


class ProblemDecoder:
    """`fixed_nodes` / `perturbed_nodes_by_t` of many problems of one space, for the `Simulation`
    records of simulate / target (the initial state is not needed there: it is `states[0]`).
    Same result as `decode_simulation_problem(index, ...)[1:]`; the radices are worked out once,
    the initial-state digits are skipped with one division, and a space without fixed-node or
    perturbation variations hands out one shared pair of (read-only) dicts."""

    def __init__(self, origin_simulation_problem, simulation_problem_variations):
        _, self.fixed_nodes, self.perturbed_nodes_by_t = origin_simulation_problem
        var_initial, var_fixed, var_perturbed = simulation_problem_variations
        self.skip = 1 << len(var_initial)
        self.var_fixed = [(node, NODE_STATE_BY_DIGIT[range_name(r)], range_radix(r)) for node, r in var_fixed]
        self.var_perturbed = [(t, node, NODE_STATE_BY_DIGIT[range_name(r)], range_radix(r))
                              for t, node, r in var_perturbed]
        self.shared = None
        if not self.var_fixed and not self.var_perturbed:
            self.shared = (dict(self.fixed_nodes), {t: dict(nodes) for t, nodes in self.perturbed_nodes_by_t.items()})

    def masks(self, index):
        """:return: (fixed_nodes dict, perturbed_nodes_by_t dict) of problem `index`"""
        if self.shared is not None:
            return self.shared
        rest = index // self.skip
        fixed = dict(self.fixed_nodes)
        for node, values, radix in self.var_fixed:
            rest, digit = divmod(rest, radix)
            if values[digit] is not None:
                fixed[node] = values[digit]
        perturbed = {t: dict(nodes) for t, nodes in self.perturbed_nodes_by_t.items()}
        for t, node, values, radix in self.var_perturbed:
            rest, digit = divmod(rest, radix)
            if values[digit] is not None:
                perturbed.setdefault(t, {})[node] = values[digit]
        return fixed, perturbed


def shard_range(n_problems, rank, world):
    """Contiguous, word-aligned (multiple of 32) slice of [0, n_problems) for one GPU."""
    words = (n_problems + 31) // 32
    lo = (words * rank // world) * 32
    hi = (words * (rank + 1) // world) * 32
    return min(lo, n_problems), min(hi, n_problems)


def state_from_code(code, n_nodes):
    """Inverse of encode_state (model.py:131-149): bit i = node i."""
    return [bool((code >> i) & 1) for i in range(n_nodes)]
