"""Timing across dynamical regimes: random N=64 nets with K=1,2,3 (ordered / critical / chaotic)."""
import sys, os, time
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boolsi_b200 import _lib, synth
n_free = int(sys.argv[1])
for k, seed, max_t in [(3, 0, 1000), (3, 2, 1000)] if len(sys.argv) > 2 else [(2, 0, 1000), (2, 1, inf), (1, 0, inf), (3, 1, 1000), (3, 2, 1000)]:
    cfg = synth.random_attract_problem(64, k, seed, n_free=n_free)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
                            cfg['simulation problem variations'])
    plan = _lib.Plan(pk, _lib.MODE_ATTRACT, **(eval(sys.argv[2]) if len(sys.argv) > 2 else {}))
    for rep in range(2):
        plan.table_clear()
        st = plan.attract_run(0, pk.n_problems, max_t, inf)
    ops = plan.op_counts()
    alg = st['steps'] * (ops['update_lop3'] + ops['compare_lop3']) / 32
    print('K=%d seed=%d max_t=%s' % (k, seed, max_t), {kk: (round(v, 2) if isinstance(v, float) else v) for kk, v in st.items()},
          'attractors', plan.table_count(), 'steps/s %.3e' % (st['steps'] / st['ms_total'] * 1e3),
          'alg LOP3/s %.3e' % (alg / st['ms_total'] * 1e3), flush=True)
    plan.close()
