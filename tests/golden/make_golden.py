#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.json

The reference (pure Python, `/root/reference`) is imported through `_refshim` (storage stubs and
one py3.12 fix; no path arithmetic is replaced).  It cannot travel to the GPU box, so its outputs
are committed here as small JSON fixtures; this script is the provenance of every number in them.

What is recorded, per case, comes from two independent routes through the reference:
  (1) the mode driver end to end (`attract_master` / `simulate_master` / `target_master`,
      reference attract.py:67, simulate.py:46, target.py:13) -> the DB post-conditions;
  (2) the per-problem solvers driven by the reference's own enumeration
      (`generate_simulation_problems`, batching.py:258) -> exact per-problem records
      (key, lambda, trajectory length, stop time), from which the exact integer statistics
      (n, sum mu, sum mu^2) are accumulated.
Route (1) and (2) are cross-checked against each other before anything is written.
"""
import json
import zlib
import math
import os
import sys
import time
from functools import partial

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)

import _refshim  # noqa: E402

_refshim.install()

import numpy as np  # noqa: E402
from boolsi.constants import NodeStateRange, Mode, inf  # noqa: E402
from boolsi.input import parse_input, count_simulation_problems  # noqa: E402
from boolsi.model import (encode_state, adjust_update_rules_for_fixed_nodes,  # noqa: E402
                          simulate_until_attractor_or_target_substate_or_max_t)
from boolsi.batching import (generate_simulation_problem_batch_seeds,  # noqa: E402
                             generate_simulation_problems)
from boolsi.attract import (attract_master, init_attractor_db_structure,  # noqa: E402
                            simulate_until_attractor_or_max_t_storing_all_states)
from boolsi.simulate import simulate_master, init_simulation_db_structure, simulate_until_max_t  # noqa: E402
from boolsi.target import target_master, simulate_until_target_substate_or_max_t  # noqa: E402
from boolsi.testing_tools import (UPDATE_RULES_A, UPDATE_RULES_B,  # noqa: E402
                                  build_predecessor_nodes_lists_and_truth_tables)

import codec  # noqa: E402
from boolsi_b200 import synth  # noqa: E402

EXAMPLES = os.path.join(_refshim.REFERENCE_ROOT, 'examples')
R = NodeStateRange


def hexcode(code):
    return format(code, 'x')


def enumerate_problems(origin, variations, n_problems):
    """The reference's own enumeration in single-process order (chunks=1, batches=1)."""
    seeds = list(generate_simulation_problem_batch_seeds(variations, 1, n_problems, 1))
    assert len(seeds) == 1
    first, size, increment, radices = seeds[0]
    assert size == n_problems
    return generate_simulation_problems(first, size, increment, radices, origin, variations)


class Spy:
    """Wraps the reference core loop to capture the stop time t it returns (model.py:235)."""

    def __init__(self, fn):
        self.fn = fn
        self.t = None

    def __call__(self, *a):
        out = self.fn(*a)
        self.t = out[2]
        return out


def per_problem(mode, preds, tts, origin, variations, n_problems, max_t, max_attractor_l,
                target_nodes, target_code):
    encode = partial(encode_state, target_nodes if target_nodes is not None else set())
    core = Spy(partial(simulate_until_attractor_or_target_substate_or_max_t, True, max_t, encode,
                       target_code))
    records = []
    adjusted_for = None
    for index, (initial_state, fixed_nodes, perturbed) in enumerate(
            enumerate_problems(origin, variations, n_problems)):
        if fixed_nodes != adjusted_for:
            a_preds, a_tts = adjust_update_rules_for_fixed_nodes(preds, tts, fixed_nodes)
            adjusted_for = fixed_nodes
        if mode == 'attract':
            res = simulate_until_attractor_or_max_t_storing_all_states(
                max_attractor_l, core, initial_state, perturbed, a_preds, a_tts)
            if res is None:
                records.append({'t': core.t, 'res': None})
            else:
                key, codes, states, traj_l = res
                records.append({'t': core.t, 'res': (key, len(codes), traj_l, codes)})
        elif mode == 'simulate':
            states = simulate_until_max_t(max_t, core, initial_state, perturbed, a_preds, a_tts)
            records.append({'t': core.t, 'states': states, 'fixed': fixed_nodes,
                            'perturbed': perturbed})
        else:
            states = simulate_until_target_substate_or_max_t(
                core, initial_state, perturbed, a_preds, a_tts)
            records.append({'t': core.t, 'states': states, 'fixed': fixed_nodes,
                            'perturbed': perturbed})
    return records


def enc_sim(index, states, fixed, perturbed):
    return {'index': index,
            'states': [hexcode(codec.state_to_code(s)) for s in states],
            'fixed_nodes': {str(n): int(v) for n, v in sorted(fixed.items())},
            'perturbations': {str(t): {str(n): int(v) for n, v in sorted(d.items())}
                              for t, d in sorted(perturbed.items())}}


def run_case(name, mode, net_name, nets, origin, variations, max_t=inf, max_attractor_l=inf,
             n_to_find=inf, target_nodes=None, target_code=None, keep_records=True,
             batches_per_process=3):
    node_names, preds, tts = nets[net_name]
    n_problems = count_simulation_problems(*variations)
    t0 = time.time()
    case = {'name': name, 'mode': mode, 'network': net_name,
            'problem': codec.encode_problem(origin, variations), 'n_problems': n_problems,
            'max_t': None if max_t is inf else max_t,
            'max_attractor_l': None if max_attractor_l is inf else max_attractor_l,
            'n_to_find': None if n_to_find is inf else n_to_find,
            'target_nodes': None if target_nodes is None else sorted(target_nodes),
            'target_code': None if target_code is None else hexcode(target_code)}
    recs = per_problem(mode, preds, tts, origin, variations, n_problems, max_t, max_attractor_l,
                       target_nodes, target_code)
    exp = {'steps': sum(r['t'] for r in recs)}
    db = _refshim.new_db_conn()
    comm = _refshim.NoMpiComm()
    if mode == 'attract':
        init_attractor_db_structure(db)
        n_found = attract_master(comm, batches_per_process, origin, variations, preds, tts, max_t,
                                 max_attractor_l, n_problems, True, db, False, None)
        # exact integer statistics from route (2)
        agg = {}
        for r in recs:
            if r['res'] is not None:
                key, lam, traj_l, codes = r['res']
                a = agg.setdefault(key, {'n': 0, 's1': 0, 's2': 0, 'codes': None})
                a['n'] += 1
                a['s1'] += traj_l
                a['s2'] += traj_l * traj_l
        attractors = []
        for (neg_freq, key), obj in db.root.aggregated_attractors.items():
            a = agg[key]
            assert a['n'] == obj.frequency == -neg_freq
            codes = [codec.state_to_code(s) for s in obj.states]
            assert codes[0] == key == min(codes)
            mean = a['s1'] / a['n']
            m2 = a['s2'] - a['s1'] * a['s1'] / a['n']
            assert math.isclose(mean, obj.trajectory_l_mean, rel_tol=1e-9, abs_tol=1e-9)
            assert math.isclose(m2, obj.trajectory_l_variation_sum, rel_tol=1e-6, abs_tol=1e-6)
            attractors.append({'key': hexcode(key), 'states': [hexcode(c) for c in codes],
                               'frequency': obj.frequency, 'mean': obj.trajectory_l_mean,
                               'variation_sum': obj.trajectory_l_variation_sum,
                               'sum_l': a['s1'], 'sum_l2': a['s2']})
        assert n_found == len(attractors) == len(agg) == db.root.n_aggregated_attractors()
        exp['attractors'] = attractors
        exp['total_frequency'] = db.root.total_frequency()
        assert exp['total_frequency'] == sum(a['n'] for a in agg.values())
        if keep_records:
            exp['records'] = [None if r['res'] is None else
                              [hexcode(r['res'][0]), r['res'][1], r['res'][2]] for r in recs]
        exp['stop_t'] = [r['t'] for r in recs] if keep_records else None
    else:
        init_simulation_db_structure(db)
        if mode == 'simulate':
            simulate_master(comm, batches_per_process, origin, variations, preds, tts, max_t,
                            n_problems, db, None)
        else:
            target_master(comm, batches_per_process, origin, variations, target_code, target_nodes,
                          preds, tts, n_to_find, max_t, n_problems, db, None)
        mine = [enc_sim(i, r['states'], r['fixed'], r['perturbed'])
                for i, r in enumerate(recs) if r['states'] is not None]
        if n_to_find is not inf:
            mine = mine[:n_to_find]
        theirs = [enc_sim(-1, s.states, s.fixed_nodes, s.perturbed_nodes_by_t)
                  for s in db.root.simulations.values()]
        assert len(mine) == len(theirs) == db.root.n_simulations(), (len(mine), len(theirs))
        for a, b in zip(mine, theirs):
            b['index'] = a['index']
            assert a == b
        exp['simulations'] = mine
        exp['n_hits_total'] = sum(1 for r in recs if r['states'] is not None)
        exp['stop_t'] = [r['t'] for r in recs] if keep_records else None
    case['expected'] = exp
    print('  %-44s %-8s n=%-6d %.1fs' % (name, mode, n_problems, time.time() - t0), flush=True)
    return case


def parse_example(fname, mode, max_t):
    cfg = parse_input(os.path.join(EXAMPLES, fname), max_t, mode)
    return cfg


def sim_csv_states(path, n_nodes):
    """Parse the states (markers stripped) out of a published simulations.csv."""
    sims = {}
    with open(path) as f:
        rows = [line.rstrip('\n').split(',') for line in f][1:]
    for row in rows:
        sims.setdefault(row[0], []).append([int(c[0]) for c in row[2:2 + n_nodes]])
    return list(sims.values())


def main():
    nets = {}
    cases = []

    # ---- networks -------------------------------------------------------------------------
    for nm, rules in [('rulesA', UPDATE_RULES_A), ('rulesB', UPDATE_RULES_B)]:
        preds, tts = build_predecessor_nodes_lists_and_truth_tables(rules)
        nets[nm] = (sorted(rules), preds, tts)
    cfg_e2 = parse_example('example2.yaml', Mode.ATTRACT, inf)
    nets['example'] = (cfg_e2['node names'], cfg_e2['incoming node lists'], cfg_e2['truth tables'])
    cfg_gl = parse_example('glider.yaml', Mode.SIMULATE, 100)
    nets['glider'] = (cfg_gl['node names'], cfg_gl['incoming node lists'], cfg_gl['truth tables'])
    cfg_c2 = parse_example('cambium2.yaml', Mode.ATTRACT, inf)
    nets['cambium'] = (cfg_c2['node names'], cfg_c2['incoming node lists'], cfg_c2['truth tables'])
    for nm, (n, k, seed) in {'rnd64k3s0': (64, 3, 0), 'rnd64k3s1': (64, 3, 1),
                             'rnd64k2s0': (64, 2, 0), 'rnd20k2s3': (20, 2, 3),
                             'rnd40k4s0': (40, 4, 0), 'rnd70k3s0': (70, 3, 0),
                             'rnd100k5s0': (100, 5, 0), 'rnd128k4s0': (128, 4, 0),
                             'rnd33k3s2': (33, 3, 2), 'rnd12k6s0': (12, 6, 0),
                             'rnd128k2s1': (128, 2, 1)}.items():
        preds, tts = synth.random_network(n, k, seed)
        nets[nm] = (['n%d' % i for i in range(n)], preds, tts)
    # hand-made corner networks: constants, identity, self-negation, single node
    nets['single_not'] = (['A'], [[0]], [{(False,): True, (True,): False}])
    nets['consts'] = (['A', 'B', 'C', 'D'], [[], [], [0, 1], [3]],
                      [{(): True}, {(): False},
                       {(False, False): False, (False, True): True, (True, False): True,
                        (True, True): False},
                       {(False,): False, (True,): True}])

    # ---- attract --------------------------------------------------------------------------
    print('attract cases', flush=True)
    o, v = cfg_e2['origin simulation problem'], cfg_e2['simulation problem variations']
    cases.append(run_case('attract_example2', 'attract', 'example', nets, o, v))
    cases.append(run_case('attract_example2_t2', 'attract', 'example', nets, o, v, max_t=2))
    cases.append(run_case('attract_example2_t3', 'attract', 'example', nets, o, v, max_t=3))
    cases.append(run_case('attract_example2_t4', 'attract', 'example', nets, o, v, max_t=4))
    cases.append(run_case('attract_example2_a2', 'attract', 'example', nets, o, v,
                          max_attractor_l=2))
    cases.append(run_case('attract_example2_t1', 'attract', 'example', nets, o, v, max_t=1))
    # reference tests/attract_tests.py:242-288
    cases.append(run_case('attract_rulesA_fixed_perturbed', 'attract', 'rulesA', nets,
                          ([False] * 5, {2: True}, {5: {2: True}, 1000: {2: False}}),
                          ([0, 1, 2, 3, 4], [], [])))
    cases.append(run_case('attract_rulesA_fixed_perturbed_t1004', 'attract', 'rulesA', nets,
                          ([False] * 5, {2: True}, {5: {2: True}, 1000: {2: False}}),
                          ([0, 1, 2, 3, 4], [], []), max_t=1005))
    cases.append(run_case('attract_rulesA_all', 'attract', 'rulesA', nets,
                          ([False] * 5, {}, {}), ([0, 1, 2, 3, 4], [], [])))
    for mt in (inf, 3, 6, 7, 8):
        for ma in (inf, 3):
            cases.append(run_case('attract_rulesB_t%s_a%s' % (mt, ma), 'attract', 'rulesB', nets,
                                  ([False] * 6, {}, {}), ([0, 1, 2, 3, 4, 5], [], []),
                                  max_t=mt, max_attractor_l=ma))
    cases.append(run_case('attract_rulesB_perturbed', 'attract', 'rulesB', nets,
                          ([False] * 6, {}, {1: {1: True}, 3: {0: False, 4: True}}),
                          ([0, 1, 2, 3, 4, 5], [], [])))
    cases.append(run_case('attract_rulesB_perturbed_t5', 'attract', 'rulesB', nets,
                          ([False] * 6, {}, {1: {1: True}, 3: {0: False, 4: True}}),
                          ([0, 1, 2, 3, 4, 5], [], []), max_t=5))
    cases.append(run_case('attract_rulesB_perturbed_t3', 'attract', 'rulesB', nets,
                          ([False] * 6, {}, {1: {1: True}, 3: {0: False, 4: True}}),
                          ([0, 1, 2, 3, 4, 5], [], []), max_t=3))
    cases.append(run_case('attract_rulesB_partial_free', 'attract', 'rulesB', nets,
                          ([True, False, True, False, False, True], {4: False}, {}),
                          ([1, 3, 5], [], [])))
    cases.append(run_case('attract_rulesB_single_problem', 'attract', 'rulesB', nets,
                          ([True, False, True, False, False, True], {}, {}), ([], [], [])))
    cases.append(run_case('attract_single_not', 'attract', 'single_not', nets,
                          ([False], {}, {}), ([0], [], [])))
    cases.append(run_case('attract_consts', 'attract', 'consts', nets,
                          ([False] * 4, {}, {}), ([0, 1, 2, 3], [], [])))
    cases.append(run_case('attract_consts_fixed', 'attract', 'consts', nets,
                          ([False] * 4, {0: False, 3: True}, {2: {1: True}}),
                          ([0, 1, 2, 3], [], [])))
    # cambium2: published golden is the full 2^30; here slices of it
    c_init, c_fixed, c_pert = cfg_c2['origin simulation problem']
    c_free = cfg_c2['simulation problem variations'][0]
    cases.append(run_case('attract_cambium2_first12', 'attract', 'cambium', nets,
                          (list(c_init), c_fixed, c_pert), (c_free[:12], [], [])))
    rng = np.random.default_rng(7)
    init = [bool(b) for b in rng.integers(0, 2, size=30)]
    free = sorted(int(x) for x in rng.choice(c_free, size=11, replace=False))
    for f in free:
        init[f] = False
    cases.append(run_case('attract_cambium2_rand11', 'attract', 'cambium', nets,
                          (init, c_fixed, c_pert), (free, [], [])))
    cases.append(run_case('attract_cambium2_rand11_t9', 'attract', 'cambium', nets,
                          (init, c_fixed, c_pert), (free, [], []), max_t=9))
    cases.append(run_case('attract_cambium2_rand11_t12', 'attract', 'cambium', nets,
                          (init, c_fixed, c_pert), (free, [], []), max_t=12))
    # random nets
    for nm, n, nfree, mt, ma in [('rnd64k3s0', 64, 9, 1000, inf), ('rnd64k3s1', 64, 8, 1000, inf),
                                 ('rnd64k3s0', 64, 8, 200, 20),
                                 ('rnd64k2s0', 64, 10, 5000, inf), ('rnd20k2s3', 20, 12, inf, inf),
                                 ('rnd40k4s0', 40, 8, 300, inf), ('rnd70k3s0', 70, 8, 300, inf),
                                 ('rnd100k5s0', 100, 7, 150, inf), ('rnd128k4s0', 128, 7, 150, inf),
                                 ('rnd128k2s1', 128, 9, 2000, inf),
                                 ('rnd33k3s2', 33, 10, 400, inf), ('rnd12k6s0', 12, 12, inf, inf)]:
        r = np.random.default_rng(zlib.crc32(nm.encode()) % 1000 + n)
        init = [bool(b) for b in r.integers(0, 2, size=n)]
        free = sorted(int(x) for x in r.choice(n, size=nfree, replace=False))
        for f in free:
            init[f] = False
        cases.append(run_case('attract_%s_f%d_t%s_a%s' % (nm, nfree, mt, ma), 'attract', nm, nets,
                              (init, {}, {}), (free, [], []), max_t=mt, max_attractor_l=ma))
    # low free nodes first (index bits = node bits), with fixed nodes and perturbations
    init = [bool(b) for b in np.random.default_rng(5).integers(0, 2, size=64)]
    for f in range(9):
        init[f] = False
    cases.append(run_case('attract_rnd64k3s0_low9_fixed_perturbed', 'attract', 'rnd64k3s0', nets,
                          (init, {10: True, 40: False}, {3: {11: True}, 17: {0: False, 63: True}}),
                          (list(range(9)), [], []), max_t=600))

    # ---- simulate -------------------------------------------------------------------------
    print('simulate cases', flush=True)
    published = {}
    for fname, t, outdir in [('example1.yaml', 5, 'output2_example1'),
                             ('example1_fixed_nodes.yaml', 5, 'output5_example1_fixed_nodes'),
                             ('example1_perturbations.yaml', 5, 'output6_example1_perturbations'),
                             ('cambium1.yaml', 10, 'output7_cambium1')]:
        cfg = parse_example(fname, Mode.SIMULATE, t)
        net = 'cambium' if 'cambium' in fname else 'example'
        assert cfg['incoming node lists'] == nets[net][1]
        case = run_case('simulate_' + fname[:-5], 'simulate', net, nets,
                        cfg['origin simulation problem'], cfg['simulation problem variations'],
                        max_t=t)
        # the reference's published CSV must agree with what the reference computes here
        csv_path = os.path.join(EXAMPLES, outdir, 'simulations.csv')
        pub = sim_csv_states(csv_path, len(nets[net][0]))
        got = [[codec.code_to_state(int(c, 16), len(nets[net][0])) for c in s['states']]
               for s in case['expected']['simulations']]
        assert [[list(map(int, st)) for st in sim] for sim in got] == pub, fname
        with open(csv_path) as f:
            published[outdir + '/simulations.csv'] = f.read()
        with open(os.path.join(EXAMPLES, outdir, 'simulation_summaries.csv')) as f:
            published[outdir + '/simulation_summaries.csv'] = f.read()
        cases.append(case)
    for t in (4, 100):
        cases.append(run_case('simulate_glider_t%d' % t, 'simulate', 'glider', nets,
                              cfg_gl['origin simulation problem'],
                              cfg_gl['simulation problem variations'], max_t=t))
    # reference tests/simulate_tests.py:11-57 and :60-88
    cases.append(run_case('simulate_rulesA_maybe_perturbation', 'simulate', 'rulesA', nets,
                          ([False, True, True, False, False], {}, {}),
                          ([], [], [(40, 0, R.MAYBE_TRUE)]), max_t=101))
    cases.append(run_case('simulate_rulesA_tiling', 'simulate', 'rulesA', nets,
                          ([False, True, True, False, False], {}, {}), ([], [], []), max_t=23))
    # every NodeStateRange on fixed nodes and perturbations, heap-array order as input.py builds it
    cases.append(run_case('simulate_rulesB_all_ranges', 'simulate', 'rulesB', nets,
                          ([False, True, False, False, True, False], {5: True}, {2: {3: True}}),
                          ([0, 2], [(1, R.MAYBE_FALSE), (3, R.MAYBE_TRUE_OR_FALSE), (4, R.TRUE_OR_FALSE)],
                           [(2, 0, R.MAYBE_TRUE), (4, 1, R.MAYBE_TRUE_OR_FALSE), (9, 2, R.TRUE_OR_FALSE),
                            (6, 5, R.MAYBE_FALSE)]), max_t=14))
    r = np.random.default_rng(11)
    init = [bool(b) for b in r.integers(0, 2, size=64)]
    cases.append(run_case('simulate_rnd64k3s0', 'simulate', 'rnd64k3s0', nets,
                          (init, {7: True}, {5: {1: False, 2: True}, 30: {63: True}}),
                          ([3, 20, 41, 62, 63, 0], [(9, R.MAYBE_TRUE)], [(12, 33, R.MAYBE_TRUE_OR_FALSE)]),
                          max_t=48))
    init = [bool(b) for b in r.integers(0, 2, size=128)]
    cases.append(run_case('simulate_rnd128k4s0', 'simulate', 'rnd128k4s0', nets,
                          (init, {100: False, 127: True}, {2: {64: True}, 9: {0: True, 127: False}}),
                          ([1, 65, 126], [], [(4, 90, R.MAYBE_FALSE)]), max_t=40))
    init = [bool(b) for b in r.integers(0, 2, size=100)]
    cases.append(run_case('simulate_rnd100k5s0', 'simulate', 'rnd100k5s0', nets,
                          (init, {}, {}), ([0, 50, 99], [], []), max_t=33))

    # ---- target ---------------------------------------------------------------------------
    print('target cases', flush=True)
    cfg_e3 = parse_example('example3.yaml', Mode.TARGET, inf)
    case = run_case('target_example3', 'target', 'example', nets,
                    cfg_e3['origin simulation problem'], cfg_e3['simulation problem variations'],
                    target_nodes=cfg_e3['target node set'],
                    target_code=cfg_e3['target substate code'])
    pub = sim_csv_states(os.path.join(EXAMPLES, 'output4_example3', 'simulations.csv'), 3)
    got = [[list(map(int, codec.code_to_state(int(c, 16), 3))) for c in s['states']]
           for s in case['expected']['simulations']]
    assert got == pub
    with open(os.path.join(EXAMPLES, 'output4_example3', 'simulations.csv')) as f:
        published['output4_example3/simulations.csv'] = f.read()
    cases.append(case)
    # reference tests/target_tests.py:143-229
    tgt_nodes = {0, 1, 2, 3, 4, 5}
    tgt_code = codec.state_to_code([True, True, False, False, False, True])
    fixed_all = {0: True, 1: True, 2: False, 3: False, 4: False, 5: True}
    for nfind in (inf, 1, 5):
        cases.append(run_case('target_rulesB_all_fixed_n%s' % nfind, 'target', 'rulesB', nets,
                              ([False] * 6, fixed_all, {}), ([0, 1, 2, 3, 4, 5], [], []),
                              n_to_find=nfind, target_nodes=tgt_nodes, target_code=tgt_code))
    tgt_nodes = {1, 4}
    tgt_code = 1 << 4
    for mt in (inf, 2, 5):
        cases.append(run_case('target_rulesB_partial_t%s' % mt, 'target', 'rulesB', nets,
                              ([False] * 6, {}, {}), ([0, 1, 2, 3, 4, 5], [], []),
                              max_t=mt, target_nodes=tgt_nodes, target_code=tgt_code))
    # per-problem last perturbation time differs (any? perturbation late in time)
    cases.append(run_case('target_rulesB_var_perturbations', 'target', 'rulesB', nets,
                          ([False] * 6, {}, {2: {0: True}}),
                          ([0, 1, 2, 3], [(5, R.MAYBE_TRUE)],
                           [(7, 4, R.MAYBE_TRUE_OR_FALSE), (3, 1, R.MAYBE_FALSE)]),
                          max_t=12, target_nodes={1, 4}, target_code=1 << 4))
    cases.append(run_case('target_rulesA_cycle_closure', 'target', 'rulesA', nets,
                          ([False] * 5, {}, {}), ([0, 1, 2, 3, 4], [], []),
                          target_nodes={0, 1, 2, 3, 4}, target_code=0b00111))
    prob = synth.random_target_problem(128, 4, 0, n_free=9, n_target_nodes=5,
                                       max_perturbation_t=30)
    nets['rnd128k4s0_t'] = (prob['node names'], prob['incoming node lists'], prob['truth tables'])
    assert nets['rnd128k4s0_t'][1] == nets['rnd128k4s0'][1]
    del nets['rnd128k4s0_t']
    for nfind in (inf, 7):
        cases.append(run_case('target_rnd128k4s0_n%s' % nfind, 'target', 'rnd128k4s0', nets,
                              prob['origin simulation problem'],
                              prob['simulation problem variations'], max_t=120, n_to_find=nfind,
                              target_nodes=prob['target node set'],
                              target_code=prob['target substate code']))
    prob = synth.random_target_problem(64, 3, 1, n_free=10, n_target_nodes=6, max_perturbation_t=20)
    assert prob['incoming node lists'] == nets['rnd64k3s1'][1]
    cases.append(run_case('target_rnd64k3s1', 'target', 'rnd64k3s1', nets,
                          prob['origin simulation problem'], prob['simulation problem variations'],
                          max_t=200, target_nodes=prob['target node set'],
                          target_code=prob['target substate code']))

    # ---- published cambium2 golden (full 2^30 run; checked on the GPU) ----------------------
    with open(os.path.join(EXAMPLES, 'output8_cambium2', 'attractor_summaries.csv')) as f:
        published['output8_cambium2/attractor_summaries.csv'] = f.read()
    with open(os.path.join(EXAMPLES, 'output8_cambium2', 'attractors.csv')) as f:
        published['output8_cambium2/attractors.csv'] = f.read()
    with open(os.path.join(EXAMPLES, 'output3_example2', 'attractors.csv')) as f:
        published['output3_example2/attractors.csv'] = f.read()
    with open(os.path.join(EXAMPLES, 'output3_example2', 'attractor_summaries.csv')) as f:
        published['output3_example2/attractor_summaries.csv'] = f.read()
    cambium2_problem = codec.encode_problem(cfg_c2['origin simulation problem'],
                                            cfg_c2['simulation problem variations'])

    out_nets = {nm: codec.encode_network(*net) for nm, net in nets.items()}
    meta = {'generator': 'tests/golden/make_golden.py',
            'reference': 'boolsi/boolsi v1.0.5 at /root/reference, imported unmodified via _refshim',
            'reference_tests_under_shim': '134 passed (model, attract, target, simulate, batching, '
                                          'testing_tools, attractor_analysis, input)',
            'python': sys.version.split()[0], 'numpy': np.__version__}
    with open(os.path.join(HERE, 'networks.json'), 'w') as f:
        json.dump({'meta': meta, 'networks': out_nets}, f, separators=(',', ':'))
    with open(os.path.join(HERE, 'cases.json'), 'w') as f:
        json.dump({'meta': meta, 'cases': cases}, f, separators=(',', ':'))
    with open(os.path.join(HERE, 'published.json'), 'w') as f:
        json.dump({'meta': meta, 'files': published, 'cambium2_problem': cambium2_problem,
                   'cambium2_n_problems': cfg_c2['total combination count']}, f, indent=0)
    print('wrote %d networks, %d cases' % (len(out_nets), len(cases)))


if __name__ == '__main__':
    main()
