"""`attract` mode on the GPU path: the drop-in for reference `attract_master` (attract.py:67-230).

Same arguments in, same `db_conn.root.*` post-conditions out; between them the
work is done by libbsi (bit-sliced CUDA kernels).  Two ways to use several GPUs, both replacing the
MPI task farm (mpi.py):
  * ONE process (what `attract_master` does): a library group, one plan per visible GPU, one host
    thread each; the peers' export kernels write their tables into GPU 0's memory, one merge launch;
  * one process PER GPU under `torch.distributed` (`attract_sharded` with world > 1, what bench.py
    runs under torchrun): per-GPU tables merged by one NCCL all_gather.
"""
import logging
import time
from collections import OrderedDict
from fractions import Fraction
from math import inf

import numpy as np

from . import _lib
from .db import init_attractor_db_structure
from .problems import shard_range



class AggregatedAttractor:
    """Result record with the attributes the reference's consumers read (attract.py:18-45):
    `states` (min-code state first, cycle order), `frequency`, `trajectory_l_mean`,
    `trajectory_l_variation_sum` -- the last two derived from the exact integer triple.

    `states` may be given as the reference's list of lists of bool or PACKED -- `(uint64 words
    [n_states, key_words], n_nodes)`, what the device returns.  Packed states stay packed (8 bytes per
    64 nodes and state) until a consumer asks: `states_array` (bool array [n_states, n_nodes]) and
    `states` (list of lists of bool, what the reference's writers read) are made on first use;
    `n_states` needs neither."""

    def __init__(self, key, states, frequency, sum_l, sum_l2, node_on_counts=None):
        self.key = key
        self._words = self._array = self._states = None
        if isinstance(states, tuple):
            self._words, self._n_nodes = states
        else:
            self._states = states
        self.frequency = frequency
        # in how many states of the cycle each node is on, counted on the device (the input of the
        # node-correlation analysis, attractor_analysis.py:49-51), or None
        self.node_on_counts = node_on_counts
        self.sum_l, self.sum_l2 = sum_l, sum_l2
        self.trajectory_l_mean = float(Fraction(sum_l, frequency))
        self.trajectory_l_variation_sum = float(Fraction(sum_l2) - Fraction(sum_l * sum_l, frequency))

    @property
    def n_states(self):
        return len(self._words) if self._words is not None else len(self._states)

    @property
    def states_array(self):
        if self._array is None:
            self._array = _lib.state_array_from_word_rows(self._words, self._n_nodes) \
                if self._words is not None else np.asarray(self._states, dtype=bool)
        return self._array

    @property
    def states(self):
        if self._states is None:
            self._states = self.states_array.tolist()
        return self._states


class AttractEngine:
    """One plan on one GPU plus the rank's share of the index space."""

    def __init__(self, predecessor_node_lists, truth_tables, origin_simulation_problem,
                 simulation_problem_variations, device=0, rank=0, world=1, **plan_opts):
        self.packed = _lib.PackedProblem(predecessor_node_lists, truth_tables,
                                         origin_simulation_problem, simulation_problem_variations)
        self.rank, self.world, self.device = rank, world, device
        self.plan_opts = dict(plan_opts)
        self.plan = None
        self._open_plan()
        self.n_problems = self.packed.n_problems
        self.lo, self.hi = shard_range(self.n_problems, rank, world)
        self._gather = None
        self.totals = None

    def _open_plan(self):
        if self.plan is not None:
            self.plan.close()
        self.plan = _lib.Plan(self.packed, _lib.MODE_ATTRACT, device=self.device, **self.plan_opts)
        if self.world > 1:
            import torch
            # everything on torch's current stream so that the collective is ordered after the kernels
            self.plan.set_stream(torch.cuda.current_stream().cuda_stream)

    def _grow_table(self):
        """More attractors than the device table holds: the reference's store simply grows
        (attract.py:405-455), so the table is made eight times larger and the job run again.  Every
        rank takes this decision from the same gathered headers, so all of them grow together."""
        log2 = self.plan_opts.get('table_log2_slots') or 20
        if log2 >= 28:
            raise _lib.BsiError(-4, 'attractor table overflow at 2^28 slots')
        self.plan_opts['table_log2_slots'] = min(28, log2 + 3)
        self._open_plan()

    def close(self):
        self.plan.close()

    def run(self, max_t=inf, max_attractor_l=inf):
        """Clear the table, run this rank's problems, merge all ranks' tables (world > 1).
        Afterwards every rank holds the complete table on its device.
        :return: run stats of this rank (`steps` etc.) + job totals in `self.totals`"""
        while True:
            self.plan.table_clear()
            if self.world == 1:
                try:
                    st = self.plan.attract_run(self.lo, self.hi, max_t, max_attractor_l)
                except _lib.BsiError as e:
                    if e.status != -4:
                        raise
                    self._grow_table()
                    continue
                self.totals = {'steps': st['steps'], 'n_found': st['n_found'], 'n_without': st['n_without']}
                return st
            # kernels, table export and the collective are enqueued back to back; the first host
            # synchronisation is the read of the gathered headers.  A rank whose run fails still takes
            # part in the collective (its header carries an error mark), so that the others do not wait
            # for it forever; every rank then raises.
            failure = None
            try:
                self.plan.attract_run_async(self.lo, self.hi, max_t, max_attractor_l)
            except _lib.BsiError as e:
                failure = e
            totals = self._merge_over_ranks(failure)
            if totals is None:                  # some rank's table overflowed: all ranks grow and repeat
                try:
                    self.plan.attract_run_finish()
                except _lib.BsiError:
                    pass
                self._grow_table()
                continue
            self.totals = totals
            st = self.plan.attract_run_finish()
            st['launches'] += 1 + (self.rank > 0) + (self.rank < self.world - 1)    # export + merges
            return st

    def _merge_over_ranks(self, failure=None):
        from .sharding import exchange_tables
        if self._gather is None:
            self._gather = {}

        def fill(buf):
            if failure is None:
                self.plan.table_export_payload(buf.data_ptr(), buf.shape[0])
            else:
                buf.zero_()
                buf[0, 5] = 2           # header word 5: 0 ok, 1 table overflow, 2 the run failed on this rank
        everyone, headers = exchange_tables(fill, self.world, 'cuda', buffers=self._gather)
        if failure is not None:
            raise failure
        if int(headers[:, 5].max()) >= 2:
            bad = [r for r in range(self.world) if int(headers[r, 5]) >= 2]
            raise _lib.BsiError(-2, 'the run failed on rank(s) %s' % bad)
        if int(headers[:, 5].max()):
            return None                 # attractor table overflow on some rank: see run()
        # fold the OTHER ranks' records into this rank's table (header rows and unused rows have
        # n == 0 and are skipped by the merge kernel): two launches, no host synchronisation
        rows = everyone.shape[1]
        if self.rank > 0:
            self.plan.table_merge_device(everyone[0].data_ptr(), self.rank * rows)
        if self.rank < self.world - 1:
            self.plan.table_merge_device(everyone[self.rank + 1].data_ptr(),
                                         (self.world - 1 - self.rank) * rows)
        return {'steps': int(headers[:, 1].sum()), 'n_found': int(headers[:, 2].sum()),
                'n_without': int(headers[:, 3].sum())}

    def attractors(self):
        """-> list of AggregatedAttractor in the reference's output order."""
        return read_attractors(self.plan, self.packed.n_nodes)


_GROUPS = OrderedDict()          # (descriptor bytes, mode, devices, options) -> _lib.Group, most recent last
_GROUPS_MAX = 2
#: A job of at most this many problems runs on ONE device when the caller names none: a B200 holds
#: 37 888 words of 32 problems in flight, so fewer words than that finish in one wave on one GPU and
#: more GPUs cannot shorten it -- while every further device costs its CUDA context (0.35 s warm, over
#: a second on a cold box: profiles/README.md, small-job latency).
SMALL_JOB_PROBLEMS = 1 << 20


def cached_group(packed, mode, devices=None, target=None, **plan_opts):
    """The group (one plan per device) for this network / problem space, kept across calls: a second
    job on the same input skips kernel specialisation, module loading and buffer allocation -- the
    per-run set-up the reference repeats per batch (mpi.py:351-376, 525-528).  At most two groups
    stay alive; `release_cached_plans()` drops them."""
    if devices is None and packed.n_problems <= SMALL_JOB_PROBLEMS:
        devices = [0]
    key = (tuple((name, a.tobytes()) for name, a in sorted(packed.arrays.items())), mode,
           None if devices is None else tuple(int(d) for d in devices),
           None if target is None else bytes(target), tuple(sorted(plan_opts.items())))
    group = _GROUPS.pop(key, None)
    if group is None:
        group = _lib.Group(packed, mode, devices=devices, target=target, **plan_opts)
        while len(_GROUPS) >= _GROUPS_MAX:
            _GROUPS.popitem(last=False)[1].close()
    _GROUPS[key] = group
    return group


def release_cached_plans():
    while _GROUPS:
        _GROUPS.popitem()[1].close()
    while _ENGINES:
        _ENGINES.popitem()[1].close()
    _lib.release_cached_memory()


def read_attractors(plan, n_nodes):
    """Table of `plan` -> list of AggregatedAttractor in the reference's output order (attract.py:170-186:
    frequency descending, then key ascending), states = the orbit of the key, key first."""
    rows = [(_lib.code_from_words(r['key'][:plan.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']), int(r['lam']))
            for r in plan.table_read()]
    rows.sort(key=lambda x: (-x[1], x[0]))
    orbits, counts = plan.attractor_orbits([(key, lam) for key, _, _, _, lam in rows], with_activity=True)
    return [AggregatedAttractor(key, (words, n_nodes), n, s1, s2, on)
            for (key, n, s1, s2, lam), words, on in zip(rows, orbits, counts)]


def attract_devices(predecessor_node_lists, truth_tables, origin_simulation_problem,
                    simulation_problem_variations, max_t=inf, max_attractor_l=inf, devices=None, progress=None,
                    **plan_opts):
    """Host structures in, host results out, ONE process: the index space is split over `devices`
    (None = every visible GPU) by the library's group run; no torch, no launcher.  `progress(member,
    done, total)` is called after every finished wave of a member's problems (long runs).
    :return: dict with the sorted attractors and the job totals."""
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    group = cached_group(packed, _lib.MODE_ATTRACT, devices, **plan_opts)
    group.set_progress(progress)
    try:
        total, per_device = group.attract_run(0, packed.n_problems, max_t, max_attractor_l)
    finally:
        group.set_progress(None)
    plan = group.plan(0)
    attractors = read_attractors(plan, packed.n_nodes)
    h2d = sum(a.nbytes for a in packed.arrays.values()) + 8 * plan.kw * len(attractors)
    d2h = 64 * len(attractors) + sum(8 * plan.kw * a.n_states for a in attractors) + 128 * group.n_devices
    return {'attractors': attractors, 'n_attractors': len(attractors), 'n_found': total['n_found'],
            'n_without': total['n_without'], 'steps_total': total['steps'], 'h2d_bytes': h2d,
            'd2h_bytes': d2h, 'n_devices': group.n_devices, 'stats': total, 'per_device': per_device}


def attract_sharded(predecessor_node_lists, truth_tables, origin_simulation_problem,
                    simulation_problem_variations, max_t=inf, max_attractor_l=inf, device=0, rank=0,
                    world=1, **plan_opts):
    """One rank of a `torch.distributed` job (one process per GPU, e.g. under torchrun): this rank's
    slice on its GPU, tables merged over NCCL.  With `world == 1` it is `attract_devices` on one GPU.
    :return: dict with the sorted attractors and the job totals."""
    if world == 1:
        return attract_devices(predecessor_node_lists, truth_tables, origin_simulation_problem,
                               simulation_problem_variations, max_t, max_attractor_l, devices=[device],
                               **plan_opts)
    eng = cached_engine(predecessor_node_lists, truth_tables, origin_simulation_problem,
                        simulation_problem_variations, device, rank, world, **plan_opts)
    try:
        eng.run(max_t, max_attractor_l)
        attractors = eng.attractors()
    except Exception:
        # (a failed run leaves the engine in an unknown state: it is not kept)
        for key in [k for k, v in _ENGINES.items() if v is eng]:
            del _ENGINES[key]
        eng.close()
        raise
    h2d = sum(a.nbytes for a in eng.packed.arrays.values()) + 8 * eng.plan.kw * len(attractors)
    d2h = 64 * len(attractors) + sum(8 * eng.plan.kw * a.n_states for a in attractors) + 128
    return {'attractors': attractors, 'n_attractors': len(attractors),
            'n_found': eng.totals['n_found'], 'n_without': eng.totals['n_without'],
            'steps_total': eng.totals['steps'], 'h2d_bytes': h2d, 'd2h_bytes': d2h}


_ENGINES = OrderedDict()         # like _GROUPS, for the one-process-per-GPU form


def cached_engine(predecessor_node_lists, truth_tables, origin_simulation_problem,
                  simulation_problem_variations, device, rank, world, **plan_opts):
    """This rank's engine (plan, exchange buffers) for this network / problem space, kept across calls:
    the next job on the same input starts with its first kernel launch."""
    packed = _lib.PackedProblem(predecessor_node_lists, truth_tables, origin_simulation_problem,
                                simulation_problem_variations)
    key = (tuple((name, a.tobytes()) for name, a in sorted(packed.arrays.items())), device, rank, world,
           tuple(sorted(plan_opts.items())))
    eng = _ENGINES.pop(key, None)
    if eng is None:
        eng = AttractEngine(predecessor_node_lists, truth_tables, origin_simulation_problem,
                            simulation_problem_variations, device=device, rank=rank, world=world, **plan_opts)
        while len(_ENGINES) >= _GROUPS_MAX:
            _ENGINES.popitem(last=False)[1].close()
    _ENGINES[key] = eng
    return eng


def attract_master(mpi_comm, n_simulation_problem_batches_per_process, origin_simulation_problem,
                   simulation_problem_variations, predecessor_node_lists, truth_tables, max_t,
                   max_attractor_l, n_simulation_problems, storing_all_states, db_conn, packing_db,
                   output_dirpath, devices=None, device=None):
    """Drop-in for reference `attract_master` (attract.py:67-230): same positional arguments, same
    post-conditions on `db_conn.root` (aggregated_attractors keyed `(-frequency, key)`,
    aggregated_attractor_keys_by_batch_index, the counters), same return value.

    `mpi_comm`, the batch count, `packing_db` and `output_dirpath` are accepted and ignored (no
    task farm, no DB packing, no progress file).  Where the reference fans the batches out over the
    MPI ranks it was started with (mpi.py:193-340), this call splits the index space over every
    visible GPU from this one process (`devices`: a list of device numbers to restrict it to;
    `device`: shorthand for one).

    `storing_all_states=False` (`boolsi attract -r`, reference attract.py:305-371): detection by
    reference points is a memory-saving device of the CPU implementation, not a different result --
    without a time cap it reports exactly the attractors, frequencies and trajectory lengths of the
    default mode, and that is what this path returns (the device path stores no states in either
    mode).  With a finite `max_t` the reference's `-r` mode additionally needs the repeat to be
    DETECTED at a reference point by `max_t` and otherwise derives a pseudo-attractor from
    `max_t - last_reference_t` without checking that a repeat occurred (SURVEY.md section 7.2;
    tests/golden/rmode.json lists reference-generated cases of both kinds).  Those cap artefacts are
    not reproduced: a warning says so.
    """
    log = logging.getLogger()
    if not storing_all_states:
        if max_t is None or max_t == inf:
            log.info('-r: reference-point detection and the default detection report the same '
                     'attractors when there is no time cap; the GPU path stores no states in either mode.')
        else:
            log.warning('-r with a time cap (-t {}): results follow the DEFAULT detection mode (an attractor '
                        'counts when it is reached and closed by max_t); the reference\'s -r mode differs '
                        'for trajectories whose repeat is not detected at a reference point by max_t.'.format(max_t))
    log.info('GPU path will be used to find attractors from {} initial conditions...'.format(
        n_simulation_problems))
    if not hasattr(db_conn.root, 'aggregated_attractors'):
        init_attractor_db_structure(db_conn)
    if device is not None:
        devices = [device]
    # progress of long runs into the log, at most once a second (the reference feeds a progress bar
    # from its batch loop, utils.py:19-54 / mpi.py:164)
    last = [time.monotonic()]

    def report(member, done, total):
        now = time.monotonic()
        if now - last[0] >= 1.0:
            last[0] = now
            log.info('GPU {}: {} of its {} initial conditions done'.format(member, done, total))
    res = attract_devices(predecessor_node_lists, truth_tables, origin_simulation_problem,
                          simulation_problem_variations, max_t, max_attractor_l, devices=devices,
                          progress=report)
    log.info('Used {} GPU(s).'.format(res['n_devices']))
    root = db_conn.root
    keys = []
    for a in res['attractors']:
        sort_key = (-a.frequency, a.key)
        root.aggregated_attractors[sort_key] = a
        keys.append(sort_key)
    root.n_aggregated_attractors.set(len(keys))
    root.total_frequency.set(res['n_found'])
    root.aggregated_attractor_keys_by_batch_index.clear()
    if keys:
        root.aggregated_attractor_keys_by_batch_index[0] = keys
        root.n_aggregated_attractor_batches.set(1)
        root.aggregated_attractor_batch_sizes_lcm.set(len(keys))
    text = ['Found {} attractors.'.format(len(keys))]
    if res['n_found'] < n_simulation_problems:
        text.append('No attractor within the given limits from {:.2%} initial conditions.'.format(
            1 - res['n_found'] / n_simulation_problems))
    log.info(' '.join(text))
    return root.n_aggregated_attractors()
