"""Where the end-to-end time of one `attract` job goes (host structures -> host attractors), 1 GPU."""
import os
import sys
import time
from math import inf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boolsi_b200 import _lib, synth  # noqa: E402
from boolsi_b200.attract import attract_devices, cached_group, read_attractors  # noqa: E402

cfg = synth.random_attract_problem(64, 3, int(sys.argv[1]) if len(sys.argv) > 1 else 0, n_free=26)
args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'], cfg['simulation problem variations'])
for rep in range(4):
    t = [time.perf_counter()]
    pk = _lib.PackedProblem(*args); t.append(time.perf_counter())
    group = cached_group(pk, _lib.MODE_ATTRACT, [0]); t.append(time.perf_counter())
    total, _ = group.attract_run(0, pk.n_problems, 1000, inf); t.append(time.perf_counter())
    plan = group.plan(0)
    recs = plan.table_read(); t.append(time.perf_counter())
    rows = [(_lib.code_from_words(r['key'][:plan.kw]), int(r['lam'])) for r in recs]
    orbits, counts = plan.attractor_orbits(rows, with_activity=True); t.append(time.perf_counter())
    states = [_lib.states_from_word_rows(w, 64) for w in orbits]; t.append(time.perf_counter())
    print(rep, 'pack %.2f group %.2f run %.2f (device %.2f) table_read %.2f orbits+activity %.2f unpack %.2f ms; lambdas %s'
          % tuple([(t[i + 1] - t[i]) * 1e3 for i in range(3)] + [total['ms_total']] +
                  [(t[i + 1] - t[i]) * 1e3 for i in range(3, 6)] + [[lam for _, lam in rows]]), flush=True)
for rep in range(4):
    t0 = time.perf_counter()
    res = attract_devices(*args, max_t=1000, max_attractor_l=inf, devices=[0])
    print('attract_devices %.2f ms' % ((time.perf_counter() - t0) * 1e3), flush=True)
