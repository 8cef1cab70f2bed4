"""torchrun script (one process per GPU): the `torch.distributed` forms against one-GPU runs of the same jobs.
  1. attract_sharded over all ranks == attract_devices on one GPU (seed-1 and seed-2 networks, 2^18 states);
  2. a table that overflows on every rank grows on all of them and the job is repeated (identity network);
  3. target_sharded == target_devices (hits, times, totals).
Prints 'torchrun checks ok' on rank 0; any mismatch raises on every rank.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/torchrun_checks.py"""
import os
import sys
from math import inf

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from boolsi_b200 import synth  # noqa: E402
from boolsi_b200.attract import AttractEngine, attract_devices, attract_sharded  # noqa: E402
from boolsi_b200.target import target_devices, target_sharded  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))


def rows(res):
    return [(a.key, a.frequency, a.sum_l, a.sum_l2, [tuple(s) for s in a.states], list(a.node_on_counts))
            for a in res['attractors']]


for seed, max_t in ((1, 1000), (2, 1000), (2, 3000)):
    cfg = synth.random_attract_problem(64, 3, seed, n_free=18)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    one = attract_devices(*args, max_t=max_t, max_attractor_l=inf, devices=[local])
    for _ in range(2):          # (the second call reuses the cached engine)
        many = attract_sharded(*args, max_t=max_t, max_attractor_l=inf, device=local, rank=rank, world=world)
        assert rows(many) == rows(one), ('attract', seed, max_t, rank)
        assert (many['steps_total'], many['n_found'], many['n_without']) == \
            (one['steps_total'], one['n_found'], one['n_without']), ('totals', seed, max_t, rank)

n = 14
preds = [[i] for i in range(n)]                         # identity network: every state is a fixed point
tts = [{(False,): False, (True,): True} for _ in range(n)]
eng = AttractEngine(preds, tts, ([False] * n, {}, {}), (list(range(n)), [], []), device=local, rank=rank, world=world,
                    table_log2_slots=9)
try:
    eng.run(inf, inf)
    assert eng.plan_opts['table_log2_slots'] >= 15, eng.plan_opts
    assert eng.plan.table_count() == 1 << n and eng.totals['n_found'] == 1 << n, (eng.plan.table_count(), eng.totals)
finally:
    eng.close()

cfg = synth.random_target_problem(128, 4, 0, n_free=18, n_target_nodes=5)
targs = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
         cfg['simulation problem variations'], cfg['target node set'], cfg['target substate code'])
one = target_devices(*targs, max_t=500, devices=[local])
many = target_sharded(*targs, max_t=500, device=local, rank=rank, world=world)
assert np.array_equal(one['hit_index'], many['hit_index']) and np.array_equal(one['hit_t'], many['hit_t'])
assert (one['n_hits_total'], one['steps_total']) == (many['n_hits_total'], many['steps_total'])
few = target_sharded(*targs, max_t=500, n_to_find=7, device=local, rank=rank, world=world)
assert np.array_equal(few['hit_index'], one['hit_index'][:7]) and few['n_hits_total'] == one['n_hits_total']

dist.barrier()
if rank == 0:
    print('torchrun checks ok: world %d' % world)
dist.destroy_process_group()
