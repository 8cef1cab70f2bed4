"""Helpers shared by the tests: load golden cases and turn them into reference-style arguments."""
from math import inf

import codec
from boolsi_b200.constants import NodeStateRange

_cases = None
_nets = None


def networks():
    global _nets
    if _nets is None:
        _nets = {nm: codec.decode_network(net)
                 for nm, net in codec.load('networks.json')['networks'].items()}
    return _nets


def cases(mode=None):
    global _cases
    if _cases is None:
        _cases = codec.load('cases.json')['cases']
    return [c for c in _cases if mode is None or c['mode'] == mode]


def case_args(case):
    """-> dict(preds, tts, origin, variations, max_t, max_attractor_l, n_to_find, target_nodes,
    target_code) in the reference's Python types."""
    _, preds, tts = networks()[case['network']]
    origin, variations = codec.decode_problem(case['problem'], NodeStateRange)
    return {
        'n_nodes': len(preds), 'preds': preds, 'tts': tts, 'origin': origin,
        'variations': variations,
        'max_t': inf if case['max_t'] is None else case['max_t'],
        'max_attractor_l': inf if case['max_attractor_l'] is None else case['max_attractor_l'],
        'n_to_find': inf if case['n_to_find'] is None else case['n_to_find'],
        'target_nodes': None if case['target_nodes'] is None else set(case['target_nodes']),
        'target_code': None if case['target_code'] is None else int(case['target_code'], 16),
        'n_problems': case['n_problems'],
    }


def expected_attractor_rows(case):
    """[(key, n, sum_l, sum_l2, [cycle codes]) ...] in the reference's output order."""
    return [(int(a['key'], 16), a['frequency'], a['sum_l'], a['sum_l2'],
             [int(s, 16) for s in a['states']]) for a in case['expected']['attractors']]


def expected_simulations(case):
    return [(s['index'], [int(c, 16) for c in s['states']],
             {int(n): bool(v) for n, v in s['fixed_nodes'].items()},
             {int(t): {int(n): bool(v) for n, v in d.items()}
              for t, d in s['perturbations'].items()})
            for s in case['expected']['simulations']]
