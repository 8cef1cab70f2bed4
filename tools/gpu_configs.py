"""Timing of the other BASELINE.json configurations (C3 cambium2 ordered regime, C5 target N=128)."""
import json, sys, time, os
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import codec
from boolsi_b200 import _lib, synth
from boolsi_b200.constants import NodeStateRange

which = sys.argv[1]
log2n = int(sys.argv[2])
if which == 'c3':
    pub = codec.load('published.json')
    _, preds, tts = codec.decode_network(codec.load('networks.json')['networks']['cambium'])
    origin, variations = codec.decode_problem(pub['cambium2_problem'], NodeStateRange)
    pk = _lib.PackedProblem(preds, tts, origin, variations)
    opts = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
    plan = _lib.Plan(pk, _lib.MODE_ATTRACT, **opts)
    print(plan.launch_config(), plan.op_counts())
    for rep in range(2):
        plan.table_clear()
        t = time.time()
        st = plan.attract_run(0, 1 << log2n, inf, inf)
        print(rep, 'wall', round(time.time() - t, 3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
              'steps/s %.3e' % (st['steps'] / (st['ms_total'] * 1e-3)), 'problems/s %.3e' % (st['n_problems'] / (st['ms_total'] * 1e-3)), flush=True)
    print('attractors', plan.table_count())
elif which == 'c5':
    cfg = synth.random_target_problem(128, 4, 0, n_free=log2n)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
                            cfg['simulation problem variations'])
    tgt = _lib.make_target(cfg['target node set'], cfg['target substate code'], pk.kw)
    plan = _lib.Plan(pk, _lib.MODE_TARGET, target=tgt)
    for rep in range(2):
        t = time.time()
        hi, ht, st = plan.target_run(0, pk.n_problems, 500, 10)
        print(rep, 'wall', round(time.time() - t, 3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
              'steps/s %.3e' % (st['steps'] / (st['ms_total'] * 1e-3)), flush=True)
    print('hits', st['n_found'], list(hi[:5]), list(ht[:5]))
