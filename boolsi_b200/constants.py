"""Constants shared with the reference interface (mirror of reference `boolsi/constants.py:11-42`).

`NodeStateRange` carries the digit semantics of a simulation-problem variation
(reference `batching.py:171-175`); the member names and values are the reference's so that objects
produced by the reference's unchanged `input.py` compare equal by name/value.
"""
from enum import Enum
from math import inf  # noqa: F401  (re-exported, as in the reference)


class Mode(Enum):
    SIMULATE = 0
    ATTRACT = 1
    TARGET = 2


class NodeStateRange(Enum):
    MAYBE_FALSE = -0
    MAYBE_TRUE = -1
    MAYBE_TRUE_OR_FALSE = -10
    TRUE_OR_FALSE = 10


#: digit -> node state (None = variation switched off), reference `batching.py:171-175`
NODE_STATE_BY_DIGIT = {
    'MAYBE_FALSE': {0: None, 1: False},
    'MAYBE_TRUE': {0: None, 1: True},
    'TRUE_OR_FALSE': {0: False, 1: True},
    'MAYBE_TRUE_OR_FALSE': {0: None, 1: False, 2: True},
}


def range_name(node_state_range):
    """Name of a NodeStateRange given this enum, the reference's enum, or the name itself."""
    return getattr(node_state_range, 'name', node_state_range)


def range_radix(node_state_range):
    return 3 if range_name(node_state_range) == 'MAYBE_TRUE_OR_FALSE' else 2


simulation_name = 'simulation'
aggregated_attractor_name = 'attractor'
