"""ctypes binding of `libbsi.so` (C ABI in `include/bsi.h`).

The library is the product: there is no Python or CPU implementation of the path behind it, and a
missing library or missing CUDA device is an error, never a fallback.
"""
import ctypes as C
import functools
import itertools
import os
from math import inf

import numpy as np

from .constants import range_name

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libbsi.so')

BSI_INF = 2 ** 64 - 1
KEY_WORDS_MAX = 4
MODE_SIMULATE, MODE_ATTRACT, MODE_TARGET = 0, 1, 2
_RANGE_CODE = {'TRUE_OR_FALSE': 0, 'MAYBE_FALSE': 1, 'MAYBE_TRUE': 2, 'MAYBE_TRUE_OR_FALSE': 3}
STATUS_NAMES = {0: 'BSI_OK', -1: 'BSI_E_ARG', -2: 'BSI_E_CUDA', -3: 'BSI_E_NVRTC',
                -4: 'BSI_E_CAPACITY', -5: 'BSI_E_UNSUPPORTED', -6: 'BSI_E_NOMEM'}

#: every symbol include/bsi.h declares (checked by tests/test_abi.py)
EXPORTS = [
    'bsi_abi_version', 'bsi_last_error', 'bsi_device_count', 'bsi_plan_create',
    'bsi_plan_create_offline', 'bsi_plan_destroy', 'bsi_plan_n_problems', 'bsi_plan_key_words',
    'bsi_plan_set_stream', 'bsi_plan_source', 'bsi_plan_build_log', 'bsi_plan_cubin',
    'bsi_plan_launch_config', 'bsi_plan_op_counts', 'bsi_plan_update_quiet_lop3', 'bsi_compile_only', 'bsi_attract_run',
    'bsi_attract_table_clear', 'bsi_attract_table_count', 'bsi_attract_table_read',
    'bsi_attract_table_export_device', 'bsi_attract_table_merge_device', 'bsi_attractor_states',
    'bsi_simulate_run', 'bsi_simulate_list', 'bsi_target_run', 'bsi_measure_lop3_peak',
    'bsi_measure_update_rate', 'bsi_release_cached_memory', 'bsi_attract_table_export_payload',
    'bsi_attractor_states_batch', 'bsi_attract_schedule', 'bsi_attract_run_async',
    'bsi_attract_run_finish', 'bsi_group_create', 'bsi_group_destroy', 'bsi_group_size', 'bsi_group_plan',
    'bsi_group_attract_run', 'bsi_group_target_run', 'bsi_group_simulate_run', 'bsi_attractor_activity', 'bsi_attractor_orbits', 'bsi_plan_set_progress',
]


class BsiError(RuntimeError):
    def __init__(self, status, message):
        super().__init__('%s: %s' % (STATUS_NAMES.get(status, status), message))
        self.status = status


class NetworkDesc(C.Structure):
    _fields_ = [('n_nodes', C.c_uint32), ('pred_off', C.c_void_p), ('pred_idx', C.c_void_p),
                ('tt_off', C.c_void_p), ('tt', C.c_void_p)]


class SpaceDesc(C.Structure):
    _fields_ = [('initial_state', C.c_void_p),
                ('n_fixed', C.c_uint32), ('fixed_node', C.c_void_p), ('fixed_value', C.c_void_p),
                ('n_perturbations', C.c_uint32), ('pert_t', C.c_void_p), ('pert_node', C.c_void_p),
                ('pert_value', C.c_void_p),
                ('n_variations', C.c_uint32), ('var_kind', C.c_void_p), ('var_node', C.c_void_p),
                ('var_t', C.c_void_p), ('var_range', C.c_void_p)]


class TargetDesc(C.Structure):
    _fields_ = [('mask', C.c_uint64 * KEY_WORDS_MAX), ('code', C.c_uint64 * KEY_WORDS_MAX)]


class PlanOpts(C.Structure):
    _fields_ = [('first_horizon', C.c_uint32), ('filter_nodes', C.c_uint32),
                ('table_log2_slots', C.c_uint32), ('chunk_log2', C.c_uint32),
                ('no_dict', C.c_uint32), ('block_threads', C.c_uint32),
                ('min_blocks', C.c_uint32), ('debug_stats', C.c_uint32),
                ('resolve_few_words', C.c_uint32), ('no_early_words', C.c_uint32),
                ('smem_slots', C.c_uint32), ('source_only', C.c_uint32), ('keep_dict', C.c_uint32),
                ('tune', C.c_uint32), ('target_hit_list', C.c_uint32), ('reserved', C.c_uint32 * 1)]


class RunStats(C.Structure):
    _fields_ = [('n_problems', C.c_uint64), ('n_found', C.c_uint64), ('n_without', C.c_uint64),
                ('steps', C.c_uint64), ('device_steps', C.c_uint64), ('launches', C.c_uint64),
                ('rounds', C.c_uint64), ('ms_total', C.c_double), ('ms_main', C.c_double),
                ('ms_resolve', C.c_double), ('reserved', C.c_uint64 * 6)]

    def as_dict(self):
        d = {name: getattr(self, name) for name, _ in self._fields_ if name != 'reserved'}
        if self.reserved[0] or self.reserved[1]:
            d['debug'] = list(self.reserved[:4])
        return d


PROGRESS_FN = C.CFUNCTYPE(None, C.c_uint64, C.c_uint64, C.c_void_p)

ATTRACTOR_REC = np.dtype([('key', np.uint64, (KEY_WORDS_MAX,)), ('n', np.uint64),
                          ('sum_l', np.uint64), ('sum_l2', np.uint64), ('lam', np.uint64)])

_lib = None


def lib():
    """Load libbsi.so; raises if it has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsiError(-2, 'libbsi.so is not built (%s); run __graft_entry__.build(). There is '
                               'no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        p, u64, u32, i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.bsi_abi_version.restype = i
        L.bsi_last_error.restype = C.c_char_p
        L.bsi_device_count.argtypes = [C.POINTER(i)]
        L.bsi_plan_create.argtypes = [C.POINTER(NetworkDesc), C.POINTER(SpaceDesc), i,
                                      C.POINTER(TargetDesc), C.POINTER(PlanOpts), i, C.POINTER(p)]
        L.bsi_plan_create_offline.argtypes = [C.POINTER(NetworkDesc), C.POINTER(SpaceDesc), i,
                                              C.POINTER(TargetDesc), C.POINTER(PlanOpts),
                                              C.POINTER(p)]
        L.bsi_plan_destroy.argtypes = [p]
        L.bsi_plan_destroy.restype = None
        L.bsi_plan_n_problems.argtypes = [p, C.POINTER(u64)]
        L.bsi_plan_key_words.argtypes = [p, C.POINTER(u32)]
        L.bsi_plan_set_stream.argtypes = [p, p]
        L.bsi_plan_set_progress.argtypes = [p, PROGRESS_FN, p]
        L.bsi_plan_source.argtypes = [p]
        L.bsi_plan_source.restype = C.c_char_p
        L.bsi_plan_build_log.argtypes = [p]
        L.bsi_plan_build_log.restype = C.c_char_p
        L.bsi_plan_cubin.argtypes = [p, C.POINTER(p), C.POINTER(u64)]
        L.bsi_plan_launch_config.argtypes = [p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.bsi_plan_op_counts.argtypes = [p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.bsi_plan_update_quiet_lop3.argtypes = [p]
        L.bsi_compile_only.argtypes = [C.POINTER(NetworkDesc), C.POINTER(SpaceDesc), i,
                                       C.POINTER(TargetDesc), C.POINTER(PlanOpts),
                                       C.POINTER(u64), C.c_char_p, u64]
        L.bsi_attract_run.argtypes = [p, u64, u64, u64, u64, C.POINTER(RunStats)]
        L.bsi_attract_table_clear.argtypes = [p]
        L.bsi_attract_run_async.argtypes = [p, u64, u64, u64, u64]
        L.bsi_attract_run_finish.argtypes = [p, C.POINTER(RunStats)]
        L.bsi_attract_table_count.argtypes = [p, C.POINTER(u64)]
        L.bsi_attract_table_read.argtypes = [p, p, u64, C.POINTER(u64)]
        L.bsi_attract_table_export_device.argtypes = [p, p, u64, C.POINTER(u64)]
        L.bsi_attract_table_merge_device.argtypes = [p, p, u64]
        L.bsi_attract_table_export_payload.argtypes = [p, p, u64]
        L.bsi_attractor_states.argtypes = [p, p, u64, p]
        L.bsi_attractor_states_batch.argtypes = [p, p, u64, u64, p]
        L.bsi_attractor_activity.argtypes = [p, p, p, u64, p]
        L.bsi_attractor_orbits.argtypes = [p, p, p, u64, p, p]
        L.bsi_simulate_run.argtypes = [p, u64, u64, u64, p, C.POINTER(RunStats)]
        L.bsi_simulate_list.argtypes = [p, p, u64, u64, p, C.POINTER(RunStats)]
        L.bsi_target_run.argtypes = [p, u64, u64, u64, p, p, u64, C.POINTER(u64),
                                     C.POINTER(RunStats)]
        L.bsi_measure_lop3_peak.argtypes = [i, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(i)]
        L.bsi_measure_update_rate.argtypes = [p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.bsi_attract_schedule.argtypes = [u64, p, p, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.bsi_group_create.argtypes = [C.POINTER(NetworkDesc), C.POINTER(SpaceDesc), i, C.POINTER(TargetDesc),
                                       C.POINTER(PlanOpts), C.POINTER(i), i, C.POINTER(p)]
        L.bsi_group_destroy.argtypes = [p]
        L.bsi_group_destroy.restype = None
        L.bsi_group_size.argtypes = [p, C.POINTER(i)]
        L.bsi_group_plan.argtypes = [p, i, C.POINTER(p)]
        L.bsi_group_attract_run.argtypes = [p, u64, u64, u64, u64, C.POINTER(RunStats), C.POINTER(RunStats)]
        L.bsi_group_target_run.argtypes = [p, u64, u64, u64, p, p, u64, C.POINTER(u64), C.POINTER(RunStats)]
        L.bsi_group_simulate_run.argtypes = [p, u64, u64, u64, p, C.POINTER(RunStats)]
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise BsiError(status, lib().bsi_last_error().decode('utf-8', 'replace'))


def cap(value):
    """`math.inf` / None (reference cli.py:241-242) -> BSI_INF."""
    return BSI_INF if value is None or value == inf else int(value)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def words_from_code(code, kw):
    return np.array([(code >> (64 * j)) & (2 ** 64 - 1) for j in range(kw)], np.uint64)


def state_array_from_word_rows(rows, n_nodes):
    """uint64 [n_states, key_words] -> bool array [n_states, n_nodes] (bit i = node i)."""
    rows = np.ascontiguousarray(rows, dtype='<u8')
    bits = np.unpackbits(rows.view(np.uint8).reshape(rows.shape[0], -1), axis=1, bitorder='little')
    return bits[:, :n_nodes].astype(bool)


def states_from_word_rows(rows, n_nodes):
    """uint64 [n_states, key_words] -> list of n_states lists of n_nodes bools (bit i = node i)."""
    return state_array_from_word_rows(rows, n_nodes).tolist()


def code_from_words(words):
    v = 0
    for j, w in enumerate(words):
        v |= int(w) << (64 * j)
    return v


class PackedProblem:
    """The reference's Python structures packed into the C descriptors.

    Arguments are exactly what `parse_input` produces and the `*_master` functions take
    (reference input.py:124-157): predecessor lists, truth-table dicts, the origin simulation
    problem `(initial_state, fixed_nodes, perturbed_nodes_by_t)` and its variations
    `(initial nodes, [(node, range)], [(t, node, range)])`.  Keeps the numpy arrays alive.
    """

    def __init__(self, predecessor_node_lists, truth_tables, origin_simulation_problem,
                 simulation_problem_variations):
        n = len(predecessor_node_lists)
        self.n_nodes = n
        self.kw = max(1, (n + 63) // 64)
        pred_off = np.zeros(n + 1, np.uint32)
        tt_off = np.zeros(n + 1, np.uint64)
        pred_idx, tt = [], []
        for i, (preds, table) in enumerate(zip(predecessor_node_lists, truth_tables)):
            k = len(preds)
            pred_idx.extend(preds)
            # entry j: first predecessor = most significant bit of j, which is the order of
            # itertools.product((False, True), repeat=k) (input.py:881, model.py:26-27)
            tt.extend(map(table.__getitem__, _tt_keys(k)))
            pred_off[i + 1] = len(pred_idx)
            tt_off[i + 1] = len(tt)
        initial_state, fixed_nodes, perturbed_nodes_by_t = origin_simulation_problem
        var_initial, var_fixed, var_perturbed = simulation_problem_variations
        a = self.arrays = {
            'pred_off': pred_off, 'pred_idx': np.array(pred_idx + [0], np.uint32),
            'tt_off': tt_off, 'tt': np.array(tt + [0], np.uint8),
            'initial_state': np.array([1 if v else 0 for v in initial_state], np.uint8),
            'fixed_node': np.array(list(fixed_nodes.keys()) + [0], np.uint32),
            'fixed_value': np.array([1 if v else 0 for v in fixed_nodes.values()] + [0], np.uint8),
        }
        pt, pn, pv = [], [], []
        for t, nodes in perturbed_nodes_by_t.items():
            for node, value in nodes.items():
                pt.append(int(t)), pn.append(int(node)), pv.append(1 if value else 0)
        a['pert_t'] = np.array(pt + [0], np.uint64)
        a['pert_node'] = np.array(pn + [0], np.uint32)
        a['pert_value'] = np.array(pv + [0], np.uint8)
        vk, vn, vt, vr = [], [], [], []
        for node in var_initial:
            vk.append(0), vn.append(int(node)), vt.append(0), vr.append(0)
        for node, r in var_fixed:
            vk.append(1), vn.append(int(node)), vt.append(0), vr.append(_RANGE_CODE[range_name(r)])
        for t, node, r in var_perturbed:
            vk.append(2), vn.append(int(node)), vt.append(int(t)), vr.append(_RANGE_CODE[range_name(r)])
        a['var_kind'] = np.array(vk + [0], np.uint32)
        a['var_node'] = np.array(vn + [0], np.uint32)
        a['var_t'] = np.array(vt + [0], np.uint64)
        a['var_range'] = np.array(vr + [0], np.uint32)
        self.n_problems = 1
        for r in vr:
            self.n_problems *= 3 if r == 3 else 2
        net = self.network = NetworkDesc()
        net.n_nodes = n
        for name in ('pred_off', 'pred_idx', 'tt_off', 'tt'):
            setattr(net, name, _ptr(a[name]))
        sp = self.space = SpaceDesc()
        sp.n_fixed, sp.n_perturbations, sp.n_variations = len(fixed_nodes), len(pt), len(vk)
        for name in ('initial_state', 'fixed_node', 'fixed_value', 'pert_t', 'pert_node',
                     'pert_value', 'var_kind', 'var_node', 'var_t', 'var_range'):
            setattr(sp, name, _ptr(a[name]))


@functools.lru_cache(maxsize=None)
def _tt_keys(k):
    return tuple(itertools.product((False, True), repeat=k))


def make_target(target_node_set, target_substate_code, kw):
    t = TargetDesc()
    mask = sum(1 << int(i) for i in target_node_set)
    for j in range(KEY_WORDS_MAX):
        t.mask[j] = (mask >> (64 * j)) & (2 ** 64 - 1)
        t.code[j] = (int(target_substate_code) >> (64 * j)) & (2 ** 64 - 1)
    return t


def plan_opts(opts):
    o = PlanOpts()
    for k, v in opts.items():
        if not hasattr(o, k) or k == 'reserved':
            raise TypeError('unknown plan option %r' % k)
        setattr(o, k, int(v))
    return o


class Plan:
    """Owner of one `bsi_plan*` (network + problem space + mode specialised for one device)."""

    def __init__(self, packed, mode, device=0, target=None, offline=False, **opts):
        self.packed = packed
        self.mode = mode
        self.kw = packed.kw
        self._h = C.c_void_p()
        self._owned = True
        o = plan_opts(opts)
        tptr = C.byref(target) if target is not None else None
        if offline:
            check(lib().bsi_plan_create_offline(C.byref(packed.network), C.byref(packed.space),
                                                mode, tptr, C.byref(o), C.byref(self._h)))
        else:
            check(lib().bsi_plan_create(C.byref(packed.network), C.byref(packed.space), mode, tptr,
                                        C.byref(o), int(device), C.byref(self._h)))

    @classmethod
    def borrowed(cls, handle, packed, mode):
        """A view of a plan that something else owns (a group member): never destroyed from here."""
        self = cls.__new__(cls)
        self.packed, self.mode, self.kw = packed, mode, packed.kw
        self._h, self._owned = handle, False
        return self

    def close(self):
        if self._h and self._owned:
            lib().bsi_plan_destroy(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- introspection -------------------------------------------------------------------
    @property
    def n_problems(self):
        n = C.c_uint64()
        check(lib().bsi_plan_n_problems(self._h, C.byref(n)))
        return n.value

    def source(self):
        return lib().bsi_plan_source(self._h).decode()

    def build_log(self):
        return lib().bsi_plan_build_log(self._h).decode()

    def cubin(self):
        data, size = C.c_void_p(), C.c_uint64()
        check(lib().bsi_plan_cubin(self._h, C.byref(data), C.byref(size)))
        return C.string_at(data, size.value)

    def cubin_bytes(self):
        data, size = C.c_void_p(), C.c_uint64()
        check(lib().bsi_plan_cubin(self._h, C.byref(data), C.byref(size)))
        return int(size.value)

    def op_counts(self):
        u, c, f = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().bsi_plan_op_counts(self._h, C.byref(u), C.byref(c), C.byref(f)))
        return {'update_lop3': u.value, 'compare_lop3': c.value, 'filter_lop3': f.value,
                'update_quiet_lop3': self._quiet_lop3()}

    def _quiet_lop3(self):
        v = lib().bsi_plan_update_quiet_lop3(self._h)
        if v < 0:
            check(v)
        return v

    def launch_config(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().bsi_plan_launch_config(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {'n_sms': a.value, 'block_threads': b.value, 'blocks_per_sm': c.value}

    def measure_update_rate(self):
        w, l = C.c_double(), C.c_double()
        check(lib().bsi_measure_update_rate(self._h, C.byref(w), C.byref(l)))
        return {'word_steps_per_s': w.value, 'lop3_per_s': l.value}

    def set_progress(self, callback):
        """`callback(done, total)` after every finished wave of problems of a long run (None: off)."""
        self._progress = PROGRESS_FN((lambda done, total, _user: callback(done, total))) if callback else PROGRESS_FN()
        check(lib().bsi_plan_set_progress(self._h, self._progress, None))

    def set_stream(self, cuda_stream):
        check(lib().bsi_plan_set_stream(self._h, C.c_void_p(cuda_stream)))

    # ---- attract -------------------------------------------------------------------------
    def attract_run(self, idx_begin, idx_end, max_t=inf, max_attractor_l=inf):
        st = RunStats()
        check(lib().bsi_attract_run(self._h, int(idx_begin), int(idx_end), cap(max_t),
                                    cap(max_attractor_l), C.byref(st)))
        return st.as_dict()

    def attract_run_async(self, idx_begin, idx_end, max_t=inf, max_attractor_l=inf):
        """Enqueue the run on the plan's stream; collect with `attract_run_finish`."""
        check(lib().bsi_attract_run_async(self._h, int(idx_begin), int(idx_end), cap(max_t),
                                          cap(max_attractor_l)))

    def attract_run_finish(self):
        st = RunStats()
        check(lib().bsi_attract_run_finish(self._h, C.byref(st)))
        return st.as_dict()

    def table_clear(self):
        check(lib().bsi_attract_table_clear(self._h))

    def table_count(self):
        n = C.c_uint64()
        check(lib().bsi_attract_table_count(self._h, C.byref(n)))
        return n.value

    def table_read(self, first_capacity=1024):
        """-> structured array of ATTRACTOR_REC (unsorted).  One export when the table fits the
        first guess; otherwise the call reports the true count and is repeated."""
        capacity = first_capacity
        while True:
            out = np.zeros(capacity, ATTRACTOR_REC)
            got = C.c_uint64()
            status = lib().bsi_attract_table_read(self._h, _ptr(out), capacity, C.byref(got))
            if status == -4 and got.value > capacity:       # BSI_E_CAPACITY: got = records held
                capacity = int(got.value)
                continue
            check(status)
            return out[:got.value]

    def table_export_device(self, dev_ptr, capacity):
        n = C.c_uint64()
        check(lib().bsi_attract_table_export_device(self._h, C.c_void_p(dev_ptr), int(capacity),
                                                    C.byref(n)))
        return n.value

    def table_export_payload(self, dev_ptr, capacity):
        """Header record + records into a device buffer of `capacity` records, no host sync."""
        check(lib().bsi_attract_table_export_payload(self._h, C.c_void_p(dev_ptr), int(capacity)))

    def table_merge_device(self, dev_ptr, n):
        check(lib().bsi_attract_table_merge_device(self._h, C.c_void_p(dev_ptr), int(n)))

    def attractor_states(self, key, lam):
        """Orbit of the key state: `lam` state codes, key first (attract.py:22-25)."""
        kwords = words_from_code(int(key), self.kw)
        out = np.zeros((int(lam), self.kw), np.uint64)
        check(lib().bsi_attractor_states(self._h, _ptr(kwords), int(lam), _ptr(out)))
        return [code_from_words(row) for row in out]

    def attractor_states_batch(self, keys_and_lams, as_words=False):
        """[(key, lam), ...] -> [[state codes, key first], ...] (or, with `as_words`, uint64 arrays
        [lam, key_words]).  One launch when the padded result (every orbit as long as the longest)
        stays small; otherwise one launch per power-of-two length class, so that a long cycle does
        not inflate the buffers of many short ones."""
        out = [None] * len(keys_and_lams)
        if not keys_and_lams:
            return out
        groups = {}
        longest = max(int(lam) for _, lam in keys_and_lams)
        one_launch = len(keys_and_lams) * longest * self.kw * 8 <= (32 << 20)
        for pos, (key, lam) in enumerate(keys_and_lams):
            groups.setdefault(0 if one_launch else max(1, int(lam) - 1).bit_length(), []).append(pos)
        for _, members in sorted(groups.items()):
            n_states = max(int(keys_and_lams[pos][1]) for pos in members)
            keys = np.zeros((len(members), self.kw), np.uint64)
            for row, pos in enumerate(members):
                keys[row] = words_from_code(int(keys_and_lams[pos][0]), self.kw)
            buf = np.zeros((len(members), n_states, self.kw), np.uint64)
            check(lib().bsi_attractor_states_batch(self._h, _ptr(keys), len(members), n_states, _ptr(buf)))
            for row, pos in enumerate(members):
                lam = int(keys_and_lams[pos][1])
                out[pos] = buf[row, :lam].copy() if as_words else [code_from_words(w) for w in buf[row, :lam]]
        return out

    def attractor_orbits(self, keys_and_lams, with_activity=False):
        """[(key, lam), ...] -> list of uint64 arrays [lam, key_words] (key state first, attract.py:22-25),
        one library call: orbits the run stored are copied, the others are entered into the cycle
        dictionary first.  `with_activity`: also -> uint64 array [n, n_nodes], in how many states of each
        cycle every node is on (counted on the device in the same pass)."""
        n = len(keys_and_lams)
        counts = np.zeros((n, self.packed.n_nodes), np.uint64) if with_activity else None
        if not n:
            return ([], counts) if with_activity else []
        keys = np.zeros((n, self.kw), np.uint64)
        for row, (key, _) in enumerate(keys_and_lams):
            keys[row] = words_from_code(int(key), self.kw)
        lams = np.array([int(lam) for _, lam in keys_and_lams], np.uint64)
        ends = np.cumsum(lams.astype(np.int64))
        states = np.zeros((int(ends[-1]), self.kw), np.uint64)
        check(lib().bsi_attractor_orbits(self._h, _ptr(keys), _ptr(lams), n, _ptr(states),
                                         _ptr(counts) if with_activity else None))
        orbits = [states[int(e - l):int(e)] for e, l in zip(ends, lams.astype(np.int64))]
        return (orbits, counts) if with_activity else orbits

    def attractor_activity(self, keys_and_lams):
        """[(key, lam), ...] -> uint64 array [n, n_nodes]: in how many states of each cycle every node
        is on (attractor_analysis.py:49-51 divides by the cycle length), counted on the device."""
        n = len(keys_and_lams)
        counts = np.zeros((n, self.packed.n_nodes), np.uint64)
        if n:
            keys = np.zeros((n, self.kw), np.uint64)
            for row, (key, _) in enumerate(keys_and_lams):
                keys[row] = words_from_code(int(key), self.kw)
            lams = np.array([int(lam) for _, lam in keys_and_lams], np.uint64)
            check(lib().bsi_attractor_activity(self._h, _ptr(keys), _ptr(lams), n, _ptr(counts)))
        return counts

    # ---- simulate / target -----------------------------------------------------------------
    def simulate_run(self, idx_begin, idx_end, max_t):
        n = int(idx_end) - int(idx_begin)
        out = np.zeros((max(n, 1), int(max_t) + 1, self.kw), np.uint64)
        st = RunStats()
        check(lib().bsi_simulate_run(self._h, int(idx_begin), int(idx_end), int(max_t), _ptr(out),
                                     C.byref(st)))
        return out[:n], st.as_dict()

    def simulate_list(self, indices, max_t):
        idx = np.ascontiguousarray(indices, np.uint64)
        out = np.zeros((max(len(idx), 1), int(max_t) + 1, self.kw), np.uint64)
        st = RunStats()
        check(lib().bsi_simulate_list(self._h, _ptr(idx), len(idx), int(max_t), _ptr(out),
                                      C.byref(st)))
        return out[:len(idx)], st.as_dict()

    def target_run(self, idx_begin, idx_end, max_t=inf, n_to_find=inf):
        n = int(idx_end) - int(idx_begin)
        capacity = n if n_to_find is None or n_to_find == inf else min(n, int(n_to_find))
        hit_index = np.zeros(max(capacity, 1), np.uint64)
        hit_t = np.zeros(max(capacity, 1), np.uint64)
        got = C.c_uint64()
        st = RunStats()
        check(lib().bsi_target_run(self._h, int(idx_begin), int(idx_end), cap(max_t),
                                   _ptr(hit_index), _ptr(hit_t), capacity, C.byref(got),
                                   C.byref(st)))
        return hit_index[:got.value], hit_t[:got.value], st.as_dict()


class Group:
    """Owner of one `bsi_group*`: one plan per device in this process; the run functions split the
    index range over the devices (the replacement of the MPI task farm, reference mpi.py:193-340)."""

    def __init__(self, packed, mode, devices=None, target=None, **opts):
        self.packed, self.mode, self.kw = packed, mode, packed.kw
        self._h = C.c_void_p()
        o = plan_opts(opts)
        tptr = C.byref(target) if target is not None else None
        if devices is None:
            arr, n = None, 0                      # every visible device
        else:
            devices = [int(d) for d in devices]
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
        check(lib().bsi_group_create(C.byref(packed.network), C.byref(packed.space), mode, tptr, C.byref(o),
                                     arr, n, C.byref(self._h)))
        size = C.c_int()
        check(lib().bsi_group_size(self._h, C.byref(size)))
        self.n_devices = size.value

    def close(self):
        if self._h:
            lib().bsi_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def plan(self, i=0):
        h = C.c_void_p()
        check(lib().bsi_group_plan(self._h, int(i), C.byref(h)))
        return Plan.borrowed(h, self.packed, self.mode)

    def set_progress(self, callback):
        """`callback(member, done, total)` after every finished wave of problems on a member (called from
        that member's host thread; None: off)."""
        self._progress_plans = []
        for i in range(self.n_devices):
            member = self.plan(i)
            member.set_progress((lambda done, total, i=i: callback(i, done, total)) if callback else None)
            self._progress_plans.append(member)          # (keeps the ctypes callbacks alive)

    def attract_run(self, idx_begin, idx_end, max_t=inf, max_attractor_l=inf):
        """-> (totals, [per-device stats]); afterwards `plan(0)` holds the whole job's table."""
        total = RunStats()
        per = (RunStats * self.n_devices)()
        check(lib().bsi_group_attract_run(self._h, int(idx_begin), int(idx_end), cap(max_t), cap(max_attractor_l),
                                          C.byref(total), per))
        return total.as_dict(), [s.as_dict() for s in per]

    def target_run(self, idx_begin, idx_end, max_t=inf, n_to_find=inf):
        n = int(idx_end) - int(idx_begin)
        capacity = n if n_to_find is None or n_to_find == inf else min(n, int(n_to_find))
        hit_index = np.zeros(max(capacity, 1), np.uint64)
        hit_t = np.zeros(max(capacity, 1), np.uint64)
        got = C.c_uint64()
        st = RunStats()
        check(lib().bsi_group_target_run(self._h, int(idx_begin), int(idx_end), cap(max_t), _ptr(hit_index),
                                         _ptr(hit_t), capacity, C.byref(got), C.byref(st)))
        return hit_index[:got.value], hit_t[:got.value], st.as_dict()

    def simulate_run(self, idx_begin, idx_end, max_t):
        n = int(idx_end) - int(idx_begin)
        out = np.zeros((max(n, 1), int(max_t) + 1, self.kw), np.uint64)
        st = RunStats()
        check(lib().bsi_group_simulate_run(self._h, int(idx_begin), int(idx_end), int(max_t), _ptr(out),
                                           C.byref(st)))
        return out[:n], st.as_dict()


def attract_schedule(horizon):
    """-> (checkpoint times, slots, end, first_final) of a detection round with this horizon."""
    times, slots = np.zeros(96, np.uint32), np.zeros(96, np.uint32)
    n, end, ff = C.c_uint32(), C.c_uint32(), C.c_uint32()
    check(lib().bsi_attract_schedule(int(horizon), _ptr(times), _ptr(slots), 96, C.byref(n), C.byref(end),
                                     C.byref(ff)))
    return times[:n.value].tolist(), slots[:n.value].tolist(), end.value, ff.value


def release_cached_memory():
    check(lib().bsi_release_cached_memory())


def device_count():
    n = C.c_int()
    check(lib().bsi_device_count(C.byref(n)))
    return n.value


def measure_lop3_peak(device=0):
    ops, mhz, sms = C.c_double(), C.c_double(), C.c_int()
    check(lib().bsi_measure_lop3_peak(int(device), C.byref(ops), C.byref(mhz), C.byref(sms)))
    return {'lop3_per_s': ops.value, 'implied_sm_mhz': mhz.value, 'n_sms': sms.value}
