import sys, os
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boolsi_b200 import _lib, synth
for n, k, nf in ((128, 4, 22), (96, 3, 22), (80, 3, 22)):
    cfg = synth.random_attract_problem(n, k, 0, n_free=nf)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'], cfg['simulation problem variations'])
    plan = _lib.Plan(pk, _lib.MODE_ATTRACT)
    for rep in range(2):
        plan.table_clear(); st = plan.attract_run(0, pk.n_problems, 500, inf)
    ops = plan.op_counts()
    alg = st['steps'] * (ops['update_lop3'] + ops['compare_lop3']) / 32
    print(n, k, plan.launch_config(), ops, {kk: (round(v, 2) if isinstance(v, float) else v) for kk, v in st.items()}, 'alg LOP3/s %.3e' % (alg / st['ms_total'] * 1e3), flush=True)
    plan.close()
