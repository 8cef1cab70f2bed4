"""CPU restatement (pure Python, integer state codes) of the reference's hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as the
checker.  The product (`boolsi_b200/`) never imports anything under `oracle/`.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against the
fixtures in `tests/golden/` that `tests/golden/make_golden.py` produced by running the unmodified
reference (68 cases covering the reference's own known-answer tests, its published example CSVs
and random networks).

A network state is held as one Python int, bit i = node i, which is exactly the reference's
`state_code` (model.py:131-149).  Each function cites the reference lines it restates.  The
reference's `-r` (reference-point) detection mode (model.py:216-228, attract.py:305-371) is NOT
restated: it disagrees with the default mode at the time cap and the product
follows the default (storing-all-states) mode.
"""
from math import inf

RANGE_DIGITS = {  # batching.py:171-175; None = the variation is switched off for this digit
    'MAYBE_FALSE': (None, False),
    'MAYBE_TRUE': (None, True),
    'TRUE_OR_FALSE': (False, True),
    'MAYBE_TRUE_OR_FALSE': (None, False, True),
}


def _rname(r):
    return getattr(r, 'name', r)


class Network:
    """Predecessor lists + truth tables in index form.

    `tables[i][j]` is the output of node i when its predecessors, first one most significant,
    spell the integer j -- the `itertools.product((False, True), repeat=K)` order in which the
    reference builds (input.py:881) and looks up (model.py:26-27) its truth-table dicts.
    """

    def __init__(self, predecessor_node_lists, truth_tables):
        self.n = len(predecessor_node_lists)
        self.preds = [list(p) for p in predecessor_node_lists]
        self.tables = []
        for preds, table in zip(self.preds, truth_tables):
            k = len(preds)
            row = []
            for j in range(1 << k):
                key = tuple(bool((j >> (k - 1 - m)) & 1) for m in range(k))
                row.append(bool(table[key]))
            self.tables.append(row)

    def with_fixed_nodes(self, fixed_nodes):
        """model.py:31-49: a fixed node gets no predecessors and a constant table."""
        out = Network.__new__(Network)
        out.n = self.n
        out.preds = [list(p) for p in self.preds]
        out.tables = [list(t) for t in self.tables]
        for node, value in fixed_nodes.items():
            out.preds[node] = []
            out.tables[node] = [bool(value)]
        return out


def apply_update_rules(net, code):
    """model.py:16-28: synchronous truth-table update of every node from the old state."""
    new = 0
    for i in range(net.n):
        j = 0
        for p in net.preds[i]:
            j = (j << 1) | ((code >> p) & 1)
        if net.tables[i][j]:
            new |= 1 << i
    return new


def simulate_step(net, code, next_perturbed_nodes=None):
    """model.py:52-73: update, then perturbations scheduled for t+1 override the result."""
    new = apply_update_rules(net, code)
    if next_perturbed_nodes:
        for node, value in next_perturbed_nodes.items():
            if value:
                new |= 1 << node
            else:
                new &= ~(1 << node)
    return new


def problem_space_radices(variations):
    """batching.py:10-46: one digit per variation, least significant first."""
    var_initial, var_fixed, var_perturbed = variations
    radices = [2] * len(var_initial)
    radices += [3 if _rname(r) == 'MAYBE_TRUE_OR_FALSE' else 2 for _, r in var_fixed]
    radices += [3 if _rname(r) == 'MAYBE_TRUE_OR_FALSE' else 2 for _, _, r in var_perturbed]
    return radices


def count_problems(variations):
    """input.py:899-926."""
    n = 1
    for r in problem_space_radices(variations):
        n *= r
    return n


def decode_problem(index, origin, variations):
    """Problem `index` of the enumeration (single process: increment 1, batching.py:49-76).

    batching.py:212-229 (mixed-radix digits, LSD first) + batching.py:160-209 (digits -> problem).
    :return: (initial state code, fixed_nodes dict, perturbed_nodes_by_t dict)
    """
    initial_state, fixed_nodes, perturbed = origin
    var_initial, var_fixed, var_perturbed = variations
    digits = []
    for radix in problem_space_radices(variations):
        index, d = divmod(index, radix)
        digits.append(d)
    code = sum(1 << i for i, b in enumerate(initial_state) if b)
    pos = 0
    for node in var_initial:
        if digits[pos]:
            code |= 1 << node
        else:
            code &= ~(1 << node)
        pos += 1
    fixed = dict(fixed_nodes)
    for node, r in var_fixed:
        value = RANGE_DIGITS[_rname(r)][digits[pos]]
        if value is not None:
            fixed[node] = value
        pos += 1
    pert = {t: dict(d) for t, d in perturbed.items()}
    for t, node, r in var_perturbed:
        value = RANGE_DIGITS[_rname(r)][digits[pos]]
        if value is not None:
            pert.setdefault(t, {})[node] = value
        pos += 1
    return code, fixed, pert


def run_core(net, perturbed, initial_code, max_t, target_mask=0, target_code=None):
    """model.py:152-236 in storing-all-states mode.

    `net` must already be adjusted for the problem's fixed nodes (mpi.py:525-528).
    :return: (codes[0..t], last_perturbation_t, t, attractor_is_found, target_is_reached)
    """
    lp = max(perturbed) if perturbed else 0          # model.py:125
    codes = [initial_code]
    code = initial_code
    for t in range(lp):                               # model.py:93-103
        code = simulate_step(net, code, perturbed.get(t + 1))
        codes.append(code)
    seen = {code}
    t = lp
    found = False
    reached = (code & target_mask) == target_code     # model.py:200 (None never matches)
    while t < max_t and not found and not reached:    # model.py:201
        t += 1
        code = apply_update_rules(net, code)
        reached = (code & target_mask) == target_code
        found = code in seen
        codes.append(code)
        seen.add(code)
    return codes, lp, t, found, reached


def attract_problem(net, perturbed, initial_code, max_t, max_attractor_l):
    """attract.py:262-302.  :return: (None | (key, lambda, trajectory_l, cycle codes), stop t)"""
    codes, lp, t, found, _ = run_core(net, perturbed, initial_code, max_t)
    if not found:
        return None, t
    since = codes[lp:]
    mu = since.index(since[-1])                       # attract.py:290
    lam = len(since) - (mu + 1)
    if lam > max_attractor_l:                         # attract.py:294
        return None, t
    cycle = since[-lam:]
    return (min(cycle), lam, len(codes) - (lam + 1), cycle), t   # attract.py:295-300


def rotate_to_key(cycle):
    """attract.py:22-25: the min-code state first, then cycle order."""
    k = cycle.index(min(cycle))
    return cycle[k:] + cycle[:k]


def attract(predecessor_node_lists, truth_tables, origin, variations, max_t=inf,
            max_attractor_l=inf, idx_begin=0, idx_end=None):
    """mpi.py:498-544 + attract.py:374-402 with exact integer statistics.

    :return: dict(table={key: [n, sum_l, sum_l2, cycle codes (key first)]}, records=[...],
                  steps=sum of stop times, n_without=#problems with no attractor)
    """
    base = Network(predecessor_node_lists, truth_tables)
    if idx_end is None:
        idx_end = count_problems(variations)
    table, records, steps, adjusted_for, net = {}, [], 0, None, None
    for index in range(idx_begin, idx_end):
        code, fixed, pert = decode_problem(index, origin, variations)
        if fixed != adjusted_for:
            net, adjusted_for = base.with_fixed_nodes(fixed), fixed
        res, t = attract_problem(net, pert, code, max_t, max_attractor_l)
        steps += t
        if res is None:
            records.append(None)
            continue
        key, lam, traj_l, cycle = res
        records.append((key, lam, traj_l))
        row = table.get(key)
        if row is None:
            table[key] = [1, traj_l, traj_l * traj_l, rotate_to_key(cycle)]
        else:
            row[0] += 1
            row[1] += traj_l
            row[2] += traj_l * traj_l
    return {'table': table, 'records': records, 'steps': steps,
            'n_without': sum(1 for r in records if r is None)}


def sorted_attractors(table):
    """attract.py:170-186: output order = frequency descending, then key ascending."""
    return sorted(table.items(), key=lambda kv: (-kv[1][0], kv[0]))


def simulate_problem(net, perturbed, initial_code, max_t):
    """simulate.py:97-131: run until the cycle closes, then tile it out to max_t."""
    codes, lp, t, _, _ = run_core(net, perturbed, initial_code, max_t)
    if len(codes) - 1 < max_t:
        since = codes[lp:]
        lam = len(since) - since.index(since[-1]) - 1
        cycle = codes[-lam:]
        reps, rest = divmod(max_t - (len(codes) - 1), lam)
        codes = codes + cycle * reps + cycle[:rest]
    return codes, t


def target_problem(net, perturbed, initial_code, max_t, target_mask, target_code):
    """target.py:109-133: the trajectory up to the first hit at/after the last perturbation."""
    codes, _, t, _, reached = run_core(net, perturbed, initial_code, max_t, target_mask,
                                       target_code)
    return (codes if reached else None), t


def _run_trajectories(mode, predecessor_node_lists, truth_tables, origin, variations, max_t,
                      target_node_set, target_code, n_to_find, idx_begin, idx_end):
    base = Network(predecessor_node_lists, truth_tables)
    if idx_end is None:
        idx_end = count_problems(variations)
    target_mask = sum(1 << n for n in (target_node_set or ()))
    out, steps, adjusted_for, net, n_hits = [], 0, None, None, 0
    for index in range(idx_begin, idx_end):
        code, fixed, pert = decode_problem(index, origin, variations)
        if fixed != adjusted_for:
            net, adjusted_for = base.with_fixed_nodes(fixed), fixed
        if mode == 'simulate':
            codes, t = simulate_problem(net, pert, code, max_t)
        else:
            codes, t = target_problem(net, pert, code, max_t, target_mask, target_code)
        steps += t
        if codes is not None:
            n_hits += 1
            # simulate.py:163-164: no more than n_to_find are kept, first ones in index order
            if len(out) < n_to_find:
                out.append({'index': index, 'states': codes, 'fixed_nodes': fixed,
                            'perturbations': pert})
    return {'simulations': out, 'steps': steps, 'n_hits_total': n_hits}


def simulate(predecessor_node_lists, truth_tables, origin, variations, max_t, idx_begin=0,
             idx_end=None):
    return _run_trajectories('simulate', predecessor_node_lists, truth_tables, origin, variations,
                             max_t, None, None, inf, idx_begin, idx_end)


def target(predecessor_node_lists, truth_tables, origin, variations, target_node_set, target_code,
           max_t=inf, n_to_find=inf, idx_begin=0, idx_end=None):
    return _run_trajectories('target', predecessor_node_lists, truth_tables, origin, variations,
                             max_t, target_node_set, target_code, n_to_find, idx_begin, idx_end)


def mean_and_variation_sum(n, sum_l, sum_l2):
    """The floats the reference accumulates by Welford/Chan (attract.py:35-45), here derived
    from the exact integer triple: mean = S1/n, M2 = S2 - S1^2/n."""
    from fractions import Fraction
    mean = Fraction(sum_l, n)
    m2 = Fraction(sum_l2) - Fraction(sum_l * sum_l, n)
    return float(mean), float(m2)
