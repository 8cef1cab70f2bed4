"""One process, every visible GPU: the headline job through the library's group run (what the drop-in
masters use), 1 device vs all devices.  usage: group_scaling.py [log2_states]"""
import os
import sys
import time
from math import inf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boolsi_b200 import _lib, synth  # noqa: E402
from boolsi_b200.attract import attract_devices  # noqa: E402

n_free = int(sys.argv[1]) if len(sys.argv) > 1 else 26
cfg = synth.random_attract_problem(64, 3, 0, n_free=n_free)
args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'], cfg['simulation problem variations'])
pk = _lib.PackedProblem(*args)
n_dev = _lib.device_count()
for devices in sorted({1, n_dev}):
    with _lib.Group(pk, _lib.MODE_ATTRACT, devices=list(range(devices))) as group:
        best = None
        for rep in range(5):
            t0 = time.perf_counter()
            total, per = group.attract_run(0, pk.n_problems, 1000, inf)
            wall = (time.perf_counter() - t0) * 1e3
            if best is None or wall < best[0]:
                best = (wall, total, per)
        wall, total, per = best
        print('group of %d: wall %.3f ms, slowest device %.3f ms (main %.3f), steps %d found %d attractors %d, per-device ms %s'
              % (devices, wall, total['ms_total'], total['ms_main'], total['steps'], total['n_found'],
                 group.plan(0).table_count(), [round(p['ms_total'], 3) for p in per]), flush=True)
    for rep in range(3):
        t0 = time.perf_counter()
        res = attract_devices(*args, max_t=1000, devices=list(range(devices)))
        e2e = (time.perf_counter() - t0) * 1e3
    print('attract_devices on %d device(s): e2e %.3f ms per call (host structures -> host attractors)' % (devices, e2e), flush=True)
