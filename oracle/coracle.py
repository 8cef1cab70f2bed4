"""ctypes front end of the C oracle (oracle/bsi_oracle.c).  TEST INFRASTRUCTURE -- see that file.

Takes the same Python structures the reference's `*_master` functions take and returns numpy
arrays of per-problem records, plus aggregation helpers that mirror oracle/oracle.py.
"""
import ctypes as C
import os
import subprocess
from math import inf

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'libbsi_oracle.so')
ORC_INF = 2 ** 64 - 1
_RANGE_CODE = {'TRUE_OR_FALSE': 0, 'MAYBE_FALSE': 1, 'MAYBE_TRUE': 2, 'MAYBE_TRUE_OR_FALSE': 3}


def build(force=False):
    src = os.path.join(HERE, 'bsi_oracle.c')
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', HERE])
    return LIB_PATH


class _Space(C.Structure):
    _fields_ = [('n_nodes', C.c_uint32), ('pred_off', C.c_void_p), ('pred_idx', C.c_void_p),
                ('tt_off', C.c_void_p), ('tt', C.c_void_p), ('initial_state', C.c_void_p),
                ('n_fixed', C.c_uint32), ('fixed_node', C.c_void_p), ('fixed_value', C.c_void_p),
                ('n_pert', C.c_uint32), ('pert_t', C.c_void_p), ('pert_node', C.c_void_p),
                ('pert_value', C.c_void_p),
                ('n_var', C.c_uint32), ('var_kind', C.c_void_p), ('var_node', C.c_void_p),
                ('var_t', C.c_void_p), ('var_range', C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        u64, p = C.c_uint64, C.c_void_p
        _lib.orc_attract.argtypes = [C.POINTER(_Space), u64, u64, u64, u64, p, p, p, p, p, p, C.c_int]
        _lib.orc_target.argtypes = [C.POINTER(_Space), u64, u64, u64, p, p, p, p, p, C.c_int]
        _lib.orc_simulate.argtypes = [C.POINTER(_Space), u64, u64, u64, p, p, p, C.c_int]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Space:
    """Packed problem space; keeps the numpy arrays alive for the lifetime of the struct."""

    def __init__(self, predecessor_node_lists, truth_tables, origin, variations):
        n = len(predecessor_node_lists)
        self.n = n
        self.kw = max(1, (n + 63) // 64)
        pred_off = np.zeros(n + 1, np.uint32)
        tt_off = np.zeros(n + 1, np.uint64)
        pred_idx, tt = [], []
        for i, (preds, table) in enumerate(zip(predecessor_node_lists, truth_tables)):
            k = len(preds)
            pred_idx += list(preds)
            for j in range(1 << k):
                tt.append(1 if table[tuple(bool((j >> (k - 1 - m)) & 1) for m in range(k))] else 0)
            pred_off[i + 1] = len(pred_idx)
            tt_off[i + 1] = len(tt)
        initial_state, fixed_nodes, perturbed = origin
        var_initial, var_fixed, var_perturbed = variations
        a = self.arrays = {
            'pred_off': pred_off, 'pred_idx': np.array(pred_idx + [0], np.uint32),
            'tt_off': tt_off, 'tt': np.array(tt, np.uint8),
            'initial_state': np.array([1 if b else 0 for b in initial_state], np.uint8),
            'fixed_node': np.array(list(fixed_nodes.keys()) + [0], np.uint32),
            'fixed_value': np.array([1 if v else 0 for v in fixed_nodes.values()] + [0], np.uint8),
        }
        pt, pn, pv = [], [], []
        for t, d in perturbed.items():
            for node, v in d.items():
                pt.append(t), pn.append(node), pv.append(1 if v else 0)
        a['pert_t'] = np.array(pt + [0], np.uint64)
        a['pert_node'] = np.array(pn + [0], np.uint32)
        a['pert_value'] = np.array(pv + [0], np.uint8)
        vk, vn, vt, vr = [], [], [], []
        for node in var_initial:
            vk.append(0), vn.append(node), vt.append(0), vr.append(0)
        for node, r in var_fixed:
            vk.append(1), vn.append(node), vt.append(0), vr.append(_RANGE_CODE[getattr(r, 'name', r)])
        for t, node, r in var_perturbed:
            vk.append(2), vn.append(node), vt.append(t), vr.append(_RANGE_CODE[getattr(r, 'name', r)])
        a['var_kind'] = np.array(vk + [0], np.uint32)
        a['var_node'] = np.array(vn + [0], np.uint32)
        a['var_t'] = np.array(vt + [0], np.uint64)
        a['var_range'] = np.array(vr + [0], np.uint32)
        self.n_problems = 1
        for r in vr:
            self.n_problems *= 3 if r == 3 else 2
        s = self.struct = _Space()
        s.n_nodes = n
        s.n_fixed, s.n_pert, s.n_var = len(fixed_nodes), len(pt), len(vk)
        for name in a:
            setattr(s, name, _ptr(a[name]))


def _cap(x):
    return ORC_INF if x is None or x == inf else int(x)


def codes_from_words(words):
    """[n, kw] uint64 -> list of Python ints."""
    words = np.asarray(words, np.uint64)
    out = []
    for row in words.reshape(-1, words.shape[-1]):
        v = 0
        for j, w in enumerate(row):
            v |= int(w) << (64 * j)
        out.append(v)
    return out


def words_from_code(code, kw):
    return np.array([(code >> (64 * j)) & (2 ** 64 - 1) for j in range(kw)], np.uint64)


def attract_records(space, max_t=inf, max_attractor_l=inf, idx_begin=0, idx_end=None, n_threads=0,
                    want_records=True):
    """-> dict(found u8[n], key u64[n,kw], lam u64[n], traj_l u64[n], stop_t u64[n], steps)"""
    idx_end = space.n_problems if idx_end is None else idx_end
    n = idx_end - idx_begin
    out = {}
    if want_records:
        out = {'found': np.zeros(n, np.uint8), 'key': np.zeros((n, space.kw), np.uint64),
               'lam': np.zeros(n, np.uint64), 'traj_l': np.zeros(n, np.uint64),
               'stop_t': np.zeros(n, np.uint64)}
    steps = C.c_uint64(0)
    rc = lib().orc_attract(C.byref(space.struct), idx_begin, idx_end, _cap(max_t),
                           _cap(max_attractor_l), _ptr(out.get('found')), _ptr(out.get('key')),
                           _ptr(out.get('lam')), _ptr(out.get('traj_l')), _ptr(out.get('stop_t')),
                           C.addressof(steps), n_threads)
    if rc:
        raise MemoryError('C oracle failed')
    out['steps'] = steps.value
    return out


def attract_table(rec, kw):
    """Aggregate per-problem records -> {key: [n, sum_l, sum_l2, lambda]} with exact integers."""
    table = {}
    idx = np.nonzero(rec['found'])[0]
    keys = codes_from_words(rec['key'][idx]) if len(idx) else []
    for key, lam, tl in zip(keys, rec['lam'][idx], rec['traj_l'][idx]):
        tl = int(tl)
        row = table.get(key)
        if row is None:
            table[key] = [1, tl, tl * tl, int(lam)]
        else:
            row[0] += 1
            row[1] += tl
            row[2] += tl * tl
            assert row[3] == int(lam)
    return table


def target_records(space, target_node_set, target_code, max_t=inf, idx_begin=0, idx_end=None,
                   n_threads=0):
    idx_end = space.n_problems if idx_end is None else idx_end
    n = idx_end - idx_begin
    mask = words_from_code(sum(1 << i for i in target_node_set), space.kw)
    code = words_from_code(target_code, space.kw)
    out = {'reached': np.zeros(n, np.uint8), 'stop_t': np.zeros(n, np.uint64)}
    steps = C.c_uint64(0)
    rc = lib().orc_target(C.byref(space.struct), idx_begin, idx_end, _cap(max_t), _ptr(mask),
                          _ptr(code), _ptr(out['reached']), _ptr(out['stop_t']),
                          C.addressof(steps), n_threads)
    if rc:
        raise MemoryError('C oracle failed')
    out['steps'] = steps.value
    return out


def simulate_states(space, max_t, idx_begin=0, idx_end=None, n_threads=0):
    """-> (states u64[n, max_t+1, kw], stop_t u64[n], steps)"""
    idx_end = space.n_problems if idx_end is None else idx_end
    n = idx_end - idx_begin
    states = np.zeros((n, max_t + 1, space.kw), np.uint64)
    stop_t = np.zeros(n, np.uint64)
    steps = C.c_uint64(0)
    rc = lib().orc_simulate(C.byref(space.struct), idx_begin, idx_end, int(max_t), _ptr(states),
                            _ptr(stop_t), C.addressof(steps), n_threads)
    if rc:
        raise MemoryError('C oracle failed')
    return states, stop_t, steps.value
