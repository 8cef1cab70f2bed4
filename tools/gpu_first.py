"""First GPU visit: LOP3 peak, C4 sub-range parity vs the C oracle, full C4 timing."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boolsi_b200 import _lib, synth

out = {}
out['peak'] = _lib.measure_lop3_peak(0)
print('peak', out['peak'], flush=True)
n_free = int(sys.argv[1]) if len(sys.argv) > 1 else 26
cfg = synth.random_attract_problem(64, 3, 0, n_free=n_free)
pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                        cfg['origin simulation problem'], cfg['simulation problem variations'])
import ast
variants = [ast.literal_eval(v) for v in sys.argv[2:]] or [{}]
for opts in variants:
    t = time.time()
    plan = _lib.Plan(pk, _lib.MODE_ATTRACT, **opts)
    print('plan', opts, 'create s', round(time.time() - t, 2), plan.launch_config(), plan.op_counts(), flush=True)
    print('update-only rate', plan.measure_update_rate(), flush=True)
    for rep in range(3):
        plan.table_clear()
        st = plan.attract_run(0, pk.n_problems, 1000)
        tsps = st['steps'] / (st['ms_total'] * 1e-3)
        print(rep, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
              'steps/s %.3e' % tsps, 'lop3 alg frac %.3f' % (tsps * 4.0 / out['peak']['lop3_per_s']), flush=True)
    import hashlib
    rows = sorted((_lib.code_from_words(r['key'][:pk.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']),
                   int(r['lam'])) for r in plan.table_read())
    print('attractors', plan.table_count(), 'table sha', hashlib.sha1(repr(rows).encode()).hexdigest()[:12],
          'steps', st['steps'], flush=True)
    out[str(opts)] = st
    plan.close()
json.dump(out, open('gpurun_out/gpu_first.json', 'w'), indent=1, default=str)
