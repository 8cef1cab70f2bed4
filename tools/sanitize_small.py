"""A few tiny runs of every kernel, for compute-sanitizer."""
import sys, os
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import goldenlib
from boolsi_b200 import _lib
names = ['attract_example2', 'attract_rulesA_fixed_perturbed', 'attract_rnd64k3s0_f8_t200_a20', 'attract_rnd128k2s1_f9_t2000_ainf',
         'attract_cambium2_rand11', 'simulate_rulesB_all_ranges', 'simulate_rnd128k4s0', 'target_rulesB_var_perturbations',
         'target_rulesA_cycle_closure', 'target_rnd64k3s1']
for c in goldenlib.cases():
    if c['name'] not in names:
        continue
    a = goldenlib.case_args(c)
    pk = _lib.PackedProblem(a['preds'], a['tts'], a['origin'], a['variations'])
    if c['mode'] == 'attract':
        for fh in (0, 3):
            with _lib.Plan(pk, _lib.MODE_ATTRACT, first_horizon=fh, filter_nodes=4 if fh else 0) as plan:
                st = plan.attract_run(0, pk.n_problems, a['max_t'], a['max_attractor_l'])
                rows = plan.table_read()
                plan.attractor_states_batch([(_lib.code_from_words(r['key'][:pk.kw]), int(r['lam'])) for r in rows])
                assert st['steps'] == c['expected']['steps'], c['name']
    elif c['mode'] == 'simulate':
        with _lib.Plan(pk, _lib.MODE_SIMULATE) as plan:
            plan.simulate_run(0, pk.n_problems, a['max_t'])
    else:
        tgt = _lib.make_target(a['target_nodes'], a['target_code'], pk.kw)
        with _lib.Plan(pk, _lib.MODE_TARGET, target=tgt) as plan:
            plan.target_run(0, pk.n_problems, a['max_t'], a['n_to_find'])
    print('ok', c['name'], flush=True)
