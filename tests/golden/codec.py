"""JSON <-> Python codecs for the golden fixtures (test infrastructure)."""
import itertools
import json
import os

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))


def tt_to_str(truth_table, k):
    return ''.join('1' if truth_table[s] else '0'
                   for s in itertools.product((False, True), repeat=k))


def str_to_tt(bits, k):
    return {s: bits[j] == '1'
            for j, s in enumerate(itertools.product((False, True), repeat=k))}


def encode_network(node_names, preds, tts):
    return {'node_names': list(node_names),
            'preds': [list(map(int, p)) for p in preds],
            'tt': [tt_to_str(t, len(p)) for p, t in zip(preds, tts)]}


def decode_network(net):
    preds = [list(p) for p in net['preds']]
    tts = [str_to_tt(b, len(p)) for p, b in zip(preds, net['tt'])]
    return net['node_names'], preds, tts


def encode_problem(origin, variations):
    initial_state, fixed_nodes, perturbed = origin
    var_initial, var_fixed, var_perturbed = variations
    return {
        'initial_state': [int(b) for b in initial_state],
        'fixed_nodes': {str(n): int(v) for n, v in fixed_nodes.items()},
        'perturbations': {str(t): {str(n): int(v) for n, v in d.items()}
                          for t, d in perturbed.items()},
        'var_initial': [int(n) for n in var_initial],
        'var_fixed': [[int(n), r.name] for n, r in var_fixed],
        'var_perturbed': [[int(t), int(n), r.name] for t, n, r in var_perturbed],
    }


def decode_problem(p, range_enum):
    origin = ([bool(b) for b in p['initial_state']],
              {int(n): bool(v) for n, v in p['fixed_nodes'].items()},
              {int(t): {int(n): bool(v) for n, v in d.items()}
               for t, d in p['perturbations'].items()})
    variations = (list(p['var_initial']),
                  [(n, range_enum[r]) for n, r in p['var_fixed']],
                  [(t, n, range_enum[r]) for t, n, r in p['var_perturbed']])
    return origin, variations


def state_to_code(state):
    return sum(1 << i for i, b in enumerate(state) if b)


def code_to_state(code, n):
    return [bool((code >> i) & 1) for i in range(n)]


def load(name):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)
