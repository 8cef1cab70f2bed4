import sys, os, time
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boolsi_b200 import _lib, synth
from boolsi_b200.attract import AttractEngine
cfg = synth.random_attract_problem(64, 3, 0, n_free=26)
args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'], cfg['simulation problem variations'])
for rep in range(4):
    t = [time.perf_counter()]
    eng = AttractEngine(*args, device=0); t.append(time.perf_counter())
    st = eng.run(1000, inf); t.append(time.perf_counter())
    att = eng.attractors(); t.append(time.perf_counter())
    eng.close(); t.append(time.perf_counter())
    print(rep, 'create %.1f run %.1f (device %.1f) attractors %.1f close %.1f ms' % tuple(
        [(t[1]-t[0])*1e3, (t[2]-t[1])*1e3, st['ms_total'], (t[3]-t[2])*1e3, (t[4]-t[3])*1e3]))

# finer: the pieces of attractors() and of the sharded public call
from boolsi_b200.attract import attract_sharded
eng = AttractEngine(*args, device=0)
eng.run(1000, inf)
for rep in range(3):
    t = [time.perf_counter()]
    recs = eng.plan.table_read(); t.append(time.perf_counter())
    rows = [(_lib.code_from_words(r['key'][:eng.plan.kw]), int(r['lam'])) for r in recs]; t.append(time.perf_counter())
    orbits = eng.plan.attractor_states_batch(rows, as_words=True); t.append(time.perf_counter())
    states = [_lib.states_from_word_rows(w, 64) for w in orbits]; t.append(time.perf_counter())
    print('table_read %.2f rows %.2f states_batch %.2f unpack %.2f ms; lambdas' % tuple((t[i+1]-t[i])*1e3 for i in range(4)),
          [lam for _, lam in rows])
eng.close()
for rep in range(4):
    t0 = time.perf_counter()
    res = attract_sharded(*args, max_t=1000, max_attractor_l=inf, device=0)
    print('attract_sharded %.2f ms' % ((time.perf_counter() - t0) * 1e3))
