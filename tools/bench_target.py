"""BASELINE.json configs[4]: synthetic random net N=128 K=4, target mode with fixed nodes and
per-step perturbations, 2^28 initial states, 500 steps.  Run under torchrun for N GPUs:
    python -m torch.distributed.run --nproc-per-node N tools/bench_target.py [log2_states]
Prints one JSON line (rank 0): trajectory-steps/s by CUDA-event time, max over ranks."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from boolsi_b200 import _lib, synth
from boolsi_b200.target import target_sharded

rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n_free = int(sys.argv[1]) if len(sys.argv) > 1 else 28
cfg = synth.random_target_problem(128, 4, 0, n_free=n_free)
args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
        cfg['simulation problem variations'], cfg['target node set'], cfg['target substate code'])
res = None
times = []
for rep in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = target_sharded(*args, max_t=500, n_to_find=10, device=local, rank=rank, world=world)
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0, res['stats']['ms_total']))
dev_ms = torch.tensor([min(t[1] for t in times[1:])], dtype=torch.float64, device='cuda')
wall = torch.tensor([min(t[0] for t in times[1:])], dtype=torch.float64, device='cuda')
if world > 1:
    dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
if rank == 0:
    pk = _lib.PackedProblem(*args[:4])
    with _lib.Plan(pk, _lib.MODE_TARGET, target=_lib.make_target(args[4], args[5], pk.kw), device=local) as plan:
        ops = plan.op_counts()
    lop3 = (ops['update_lop3'] + len(cfg['target node set'])) / 32.0
    peak = _lib.measure_lop3_peak(local)['lop3_per_s']
    v = res['steps_total'] / (float(dev_ms.item()) * 1e-3)
    print(json.dumps({'config': 'configs[4]: N=128 K=4 target, fixed nodes + perturbations, 2^%d initial states, max_t=500' % n_free,
                      'n_gpus': world, 'value_device': v, 'unit': 'trajectory-steps/s', 'device_ms': float(dev_ms.item()),
                      'e2e_ms': float(wall.item()) * 1e3, 'n_hits_total': res['n_hits_total'],
                      'first_hits': [int(i) for i in res['hit_index'][:5]],
                      'algorithmic_lop3_per_trajectory_step': lop3,
                      'frac_of_lop3_peak_per_gpu': v / world * lop3 / peak}))
if world > 1:
    dist.destroy_process_group()
