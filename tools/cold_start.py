"""Small-job latency (BASELINE.json configs[0] and configs[1]): wall time of a FRESH process from its
first line to the results of `attract example2` (3 nodes, 8 initial states) and `simulate glider -t 100`
-- this repo's drop-ins cold (no compiled units on disk), warm (units on disk from an earlier process)
-- beside the unmodified reference's masters in a fresh process on the host.  Measurement tool (it
runs oracle/refarm.py's reference shim for the CPU side); prints one JSON line.
usage: python tools/cold_start.py [attract_example2,simulate_glider_t100]"""
import json
import os
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import time
t0 = time.perf_counter()
import sys
from math import inf
sys.path[:0] = [%(repo)r, %(repo)r + '/tests', %(repo)r + '/tests/golden']
import goldenlib
case = [c for c in goldenlib.cases() if c['name'] == %(case)r][0]
a = goldenlib.case_args(case)
t_in = time.perf_counter()
if %(impl)r == 'b200':
    import boolsi_b200
    from boolsi_b200 import db, _lib
    t_a = time.perf_counter()
    n_dev = _lib.device_count()                 # (first driver call of the process: cuInit)
    t_b = time.perf_counter()
    print('LAP import_boolsi_b200 %%.4f' %% (t_a - t_in))
    print('LAP cuInit_and_device_count %%.4f (%%d devices)' %% (t_b - t_a, n_dev))
    conn = db.ResultStore() if hasattr(db, 'ResultStore') else db.new_db_conn()
    if case['mode'] == 'attract':
        n = boolsi_b200.attract_master(None, 1, a['origin'], a['variations'], a['preds'], a['tts'], a['max_t'],
                                       a['max_attractor_l'], a['n_problems'], True, conn, False, None)
    else:
        boolsi_b200.simulate_master(None, 1, a['origin'], a['variations'], a['preds'], a['tts'], a['max_t'],
                                    a['n_problems'], conn, None)
        n = conn.root.n_simulations()
else:
    from oracle import refarm, refshim
    refshim.install(refarm.reference_root())
    conn = refshim.new_db_conn()
    if case['mode'] == 'attract':
        from boolsi.attract import attract_master, init_attractor_db_structure
        init_attractor_db_structure(conn)
        n = attract_master(refshim.NoMpiComm(), 1, a['origin'], a['variations'], a['preds'], a['tts'], a['max_t'],
                           a['max_attractor_l'], a['n_problems'], True, conn, False, None)
    else:
        from boolsi.simulate import simulate_master, init_simulation_db_structure
        init_simulation_db_structure(conn)
        simulate_master(refshim.NoMpiComm(), 1, a['origin'], a['variations'], a['preds'], a['tts'], a['max_t'],
                        a['n_problems'], conn, None)
        n = conn.root.n_simulations()
t1 = time.perf_counter()
print('RESULT %%d %%.4f %%.4f' %% (n, t1 - t0, t1 - t_in))
'''


def child(case, impl, env):
    code = CHILD % {'repo': REPO, 'case': case, 'impl': impl}
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=90)
    for line in out.stdout.splitlines():
        if line.startswith('RESULT'):
            _, n, total, job = line.split()
            r = {'results': int(n), 'process_s': float(total), 'job_s': float(job)}
            r['python_laps'] = [ln[4:] for ln in out.stdout.splitlines() if ln.startswith('LAP ')]
            laps = [ln[6:].split() for ln in out.stderr.splitlines() if ln.startswith('[bsi] ')]
            if laps:        # BSI_DEBUG_TIMING=1: the library's own stages, ms
                r['stages_ms'] = [[' '.join(x[:-2]), float(x[-2])] for x in laps if len(x) >= 3 and float(x[-2]) >= 5.0]
            return r
    raise RuntimeError('child failed:\n' + out.stdout[-2000:] + out.stderr[-4000:])


def measure(cases):
    line = {}
    for case in cases:
        with tempfile.TemporaryDirectory() as cache:
            cold = child(case, 'b200', dict(os.environ, BSI_CACHE_DIR=cache, BSI_DEBUG_TIMING='1'))
            warm = child(case, 'b200', dict(os.environ, BSI_CACHE_DIR=cache, BSI_DEBUG_TIMING='1'))
        line[case] = {'b200_cold': cold, 'b200_warm_disk_cache': warm}
        try:
            line[case]['reference_cpu'] = child(case, 'reference', dict(os.environ))
        except Exception as e:      # noqa: BLE001
            line[case]['reference_cpu'] = {'unavailable': str(e)[-300:]}
    line['note'] = 'process_s: wall time of a fresh python process from its first line to the results (imports, ' \
                   'input decoding, CUDA context, kernel specialisation, run); job_s: the master call and its imports only'
    return line


def main():
    cases = (sys.argv[1] if len(sys.argv) > 1 else 'attract_example2,simulate_glider_t100').split(',')
    print(json.dumps(measure(cases)))


if __name__ == '__main__':
    main()
