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
