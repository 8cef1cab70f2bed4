"""In-memory result store with the attribute surface the reference's writers/readers use on
`db_conn.root` (reference db.py:15-78, attract.py:48-64, simulate.py:31-43).

It replaces the ZODB temp database: results come off the device as a few KB of integer records,
so nothing needs spooling.  Only containers live here -- no path arithmetic.
"""
from sortedcontainers import SortedDict


class Length:
    """BTrees.Length.Length look-alike: a counter you call (attract.py:59-64)."""

    def __init__(self, value=0):
        self.value = value

    def __call__(self):
        return self.value

    def change(self, delta):
        self.value += delta

    def set(self, value):
        self.value = value


class BTree(SortedDict):
    """OOBTree.BTree look-alike: a mapping that iterates in key order."""


class _Root:
    pass


class ResultStore:
    """Stands in for a ZODB connection: `.root` attribute bag + the two methods consumers call."""

    def __init__(self):
        self.root = _Root()

    def cacheMinimize(self):
        pass

    def close(self):
        pass


def init_attractor_db_structure(db_conn):
    """reference attract.py:48-64"""
    db_conn.root.aggregated_attractors = BTree()
    db_conn.root.aggregated_attractor_keys_by_batch_index = BTree()
    db_conn.root.sorted_aggregated_attractors = BTree()
    db_conn.root.n_aggregated_attractors = Length()
    db_conn.root.total_frequency = Length()
    db_conn.root.n_aggregated_attractor_batches = Length()
    db_conn.root.aggregated_attractor_batch_sizes_lcm = Length(1)


def init_simulation_db_structure(db_conn):
    """reference simulate.py:31-43"""
    db_conn.root.simulations = BTree()
    db_conn.root.n_simulations = Length()
    db_conn.root.n_simulation_batches = Length()
    db_conn.root.simulation_batch_sizes_lcm = Length(1)
