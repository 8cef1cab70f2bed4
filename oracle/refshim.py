"""Import shim that makes the *unmodified* reference package importable on this image.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by `tests/golden/make_golden.py` (which writes the
committed fixtures), by `tests/test_reference_live.py`, and by `bench.py --impl reference`, which
times the reference's own `attract_master` on the host cores.  Nothing in the product imports this
file.  The package is looked for in `baseline/_ref/` (the unmodified `pip install --target` copy that
`__graft_entry__.build()` makes and that travels to the GPU box) and in `/root/reference`.

Why a shim: the image has Python 3.12.3 and no ZODB stack (SURVEY.md §7 risk 8):
  * `persistent`, `transaction`, `BTrees.{OOBTree,Length}`, `ZODB` are absent.  They are storage
    containers only (no path arithmetic), so they are replaced by plain in-memory containers.
  * `boolsi/input.py:884` uses `exec(..., locals=...)`, a 3.13+ signature.  The one function
    `build_truth_table_from_safe_update_rule` (input.py:867-896) is replaced by an `eval` with an
    explicit namespace; the reference's own tests (tests/input_tests.py:1451-1657) validate it.
  * `TickingProgressBar` (utils.py:19) is replaced by a no-op context manager so that no thread
    or progress file is created.
"""
import itertools
import os
import sys
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_ROOT = os.path.join(REPO, 'baseline', '_ref')


def _find_root():
    for root in (os.environ.get('BOOLSI_REFERENCE_ROOT'), '/root/reference', INSTALLED_ROOT):
        if root and os.path.isdir(os.path.join(root, 'boolsi')):
            return root
    return '/root/reference'


REFERENCE_ROOT = _find_root()


def reference_available(root=None):
    return os.path.isdir(os.path.join(root or REFERENCE_ROOT, 'boolsi'))


class _SortedDict(dict):
    """dict whose views iterate in sorted key order (stands in for OOBTree.BTree)."""

    def keys(self):
        return sorted(dict.keys(self))

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class _Length:
    """Stands in for BTrees.Length.Length."""

    def __init__(self, v=0):
        self.value = v

    def __call__(self):
        return self.value

    def change(self, d):
        self.value += d

    def set(self, v):
        self.value = v


class _Root:
    pass


class _Db:
    def pack(self):
        pass


class _Conn:
    def __init__(self):
        self.root = _Root()
        self._db = _Db()

    def cacheMinimize(self):
        pass

    def db(self):
        return self._db

    def close(self):
        pass


class _NoProgressBar:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a, **k):
        pass


class NoMpiComm:
    """What `MPICommWrapper()` yields without mpi4py (mpi.py:34-39): rank 0, no workers."""
    rank = 0
    n_workers = 0
    run_id = 'golden'


_installed = False


def install(root=None):
    """Register the stubs and patch the two py3.12-incompatible spots.  Idempotent.
    :param root: directory that holds the `boolsi` package (default: REFERENCE_ROOT)"""
    global _installed, REFERENCE_ROOT
    if _installed:
        return
    if root is not None:
        REFERENCE_ROOT = root
    if not reference_available():
        raise RuntimeError('reference package not found at %s' % REFERENCE_ROOT)

    persistent = types.ModuleType('persistent')
    persistent.Persistent = object
    transaction = types.ModuleType('transaction')
    transaction.commit = lambda: None
    transaction.abort = lambda: None
    btrees = types.ModuleType('BTrees')
    oobtree = types.ModuleType('BTrees.OOBTree')
    oobtree.BTree = _SortedDict
    length = types.ModuleType('BTrees.Length')
    length.Length = _Length
    btrees.OOBTree = oobtree
    btrees.Length = length
    zodb = types.ModuleType('ZODB')
    zodb.DB = object
    zodb.connection = lambda *_a, **_k: _Conn()
    zodb_fs = types.ModuleType('ZODB.FileStorage')
    zodb_fs.FileStorage = object
    zodb.FileStorage = zodb_fs
    for name, mod in [('persistent', persistent), ('transaction', transaction),
                      ('BTrees', btrees), ('BTrees.OOBTree', oobtree),
                      ('BTrees.Length', length), ('ZODB', zodb),
                      ('ZODB.FileStorage', zodb_fs)]:
        sys.modules.setdefault(name, mod)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    import boolsi.input as ref_input
    from boolsi.model import majority

    def build_truth_table_from_safe_update_rule(safe_update_rule, predecessor_nodes,
                                                safe_node_names, err_msg):
        truth_table = dict()
        for states in itertools.product((False, True), repeat=len(predecessor_nodes)):
            namespace = {safe_node_names[node]: states[i]
                         for i, node in enumerate(predecessor_nodes)}
            namespace['majority'] = majority
            try:
                value = eval(safe_update_rule, {'__builtins__': {}}, namespace)
            except BaseException:
                raise ValueError(err_msg)
            if value not in {False, True}:
                raise ValueError(err_msg)
            truth_table[states] = value
        return truth_table

    ref_input.build_truth_table_from_safe_update_rule = build_truth_table_from_safe_update_rule

    import boolsi.mpi as ref_mpi
    import boolsi.attract as ref_attract
    ref_mpi.TickingProgressBar = _NoProgressBar
    ref_attract.TickingProgressBar = _NoProgressBar
    _installed = True


def new_db_conn():
    return _Conn()
