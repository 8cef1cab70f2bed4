"""Ordered regime (every lane closes within a few steps): random N=64, K=2 network, 2^n_free
initial states -- the case that is all resolution and no trajectory.  usage: gpu_ordered.py [n_free]"""
import sys, os
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boolsi_b200 import _lib, synth
n_free = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg = synth.random_attract_problem(64, 2, 0, n_free=n_free)
pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
                        cfg['simulation problem variations'])
with _lib.Plan(pk, _lib.MODE_ATTRACT) as plan:
    for rep in range(2):
        plan.table_clear()
        st = plan.attract_run(0, pk.n_problems, 1000, inf)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, 'attractors', plan.table_count())
