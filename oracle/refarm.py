"""The reference arm of bench.py: the UNMODIFIED reference package (`baseline/_ref/boolsi`, installed
by `__graft_entry__.build()` with `pip install --target`; `/root/reference` where that exists) timed
on the host cores.  TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product.

What runs is the reference's own `attract_master` (attract.py:67-230) through
`generic_single_process_master` (mpi.py:131-190) -- batching, per-problem solver, aggregation and DB
writes included -- in P independent processes, one per host core, as `mpiexec -np P boolsi attract`
would fan out (no MPI launcher, mpi4py or ZODB on this image: oracle/refshim.py stands in for the
storage containers, with no path arithmetic).  Each process takes a contiguous slice of the bench
job's own index space: the low `m` free nodes vary, the higher free nodes are set to the bits of the
slice number.  Trajectory-steps of a slice are counted afterwards, untimed, by the C oracle (the
reference does not report them; the two are pinned to each other by tests/test_reference_live.py).
"""
import os
import time
from math import inf

from . import refshim


def _slice_problem(origin, variations, m, slice_id):
    initial, fixed, perturbed = origin
    free = list(variations[0])
    low, high = free[:m], free[m:]
    init = list(initial)
    for j, node in enumerate(high):
        init[node] = bool((slice_id >> j) & 1)
    return (init, fixed, perturbed), (low, variations[1], variations[2])


def run_slice(job):
    """One worker process: the reference's attract_master / target_master on one slice.
    -> (wall seconds, n results, total frequency / n simulations)"""
    root, preds, tts, origin, variations, m, slice_id, max_t, max_attractor_l, batches, target = job
    refshim.install(root)
    sub_origin, sub_variations = _slice_problem(origin, variations, m, slice_id)
    db = refshim.new_db_conn()
    if target is None:
        from boolsi.attract import attract_master, init_attractor_db_structure
        init_attractor_db_structure(db)
        t0 = time.perf_counter()
        n = attract_master(refshim.NoMpiComm(), batches, sub_origin, sub_variations, preds, tts, max_t,
                           max_attractor_l, 1 << m, True, db, False, None)
        return time.perf_counter() - t0, n, db.root.total_frequency()
    from boolsi.simulate import init_simulation_db_structure
    from boolsi.target import target_master
    target_nodes, target_code = target
    init_simulation_db_structure(db)
    t0 = time.perf_counter()
    n = target_master(refshim.NoMpiComm(), batches, sub_origin, sub_variations, target_code, target_nodes, preds,
                      tts, inf, max_t, 1 << m, db, None)
    return time.perf_counter() - t0, n, n


def slice_steps(preds, tts, origin, variations, m, slice_id, max_t, max_attractor_l=inf, target=None):
    """Trajectory-steps (sum of the reference's stop times, model.py:235) and found / hit count of a slice."""
    from . import coracle
    sub_origin, sub_variations = _slice_problem(origin, variations, m, slice_id)
    sp = coracle.Space(preds, tts, sub_origin, sub_variations)
    if target is None:
        rec = coracle.attract_records(sp, max_t, max_attractor_l, n_threads=os.cpu_count() or 1, want_records=True)
        return rec['steps'], int(rec['found'].sum())
    rec = coracle.target_records(sp, target[0], target[1], max_t, n_threads=os.cpu_count() or 1)
    return rec['steps'], int(rec['reached'].sum())


def reference_root():
    for root in (refshim.INSTALLED_ROOT, '/root/reference'):
        if refshim.reference_available(root):
            return root
    return None


def timed_sample(args, n_free, max_t, cores, m, first_slice=0, target=None):
    """`cores` processes x one slice of 2^m problems each, slices spread evenly over the index space.
    -> dict(value=trajectory-steps/s over the slowest process's wall time, ...)"""
    import multiprocessing as mp
    root = reference_root()
    if root is None:
        raise RuntimeError('the reference package is neither in baseline/_ref nor in /root/reference')
    preds, tts, origin, variations = args
    n_slices = 1 << (n_free - m)
    stride = max(1, n_slices // cores)
    ids = [(first_slice + c * stride) % n_slices for c in range(cores)]
    jobs = [(root, preds, tts, origin, variations, m, sid, max_t, inf, 100, target) for sid in ids]
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(run_slice, jobs)
    wall_pool = time.perf_counter() - t0
    steps = found = 0
    for sid, r in zip(ids, res):
        s, f = slice_steps(preds, tts, origin, variations, m, sid, max_t, target=target)
        assert f == r[2], 'reference and C oracle disagree on slice %d: %d vs %d found' % (sid, r[2], f)
        steps += s
        found += f
    wall = max(r[0] for r in res)
    return {'value': steps / wall, 'steps': steps, 'wall_s': wall, 'wall_with_startup_s': wall_pool, 'cores': cores,
            'problems': cores << m, 'found': found, 'root': root,
            'sample': 'unmodified reference boolsi.%s (generic_single_process_master) from %s, %d processes x 2^%d '
                      'problems (contiguous slices %s of the same index space), %d trajectory-steps, slowest '
                      'process %.1f s' % ('attract.attract_master' if target is None else 'target.target_master', root,
                                          cores, m, (ids[:4] + ['...']) if len(ids) > 4 else ids, steps, wall)}
