"""torchrun script: host-side time of each stage of AttractEngine.run at world > 1."""
import os, sys, time
from math import inf
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from boolsi_b200 import synth
from boolsi_b200.attract import AttractEngine
from boolsi_b200.sharding import exchange_tables
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg = synth.random_attract_problem(64, 3, 0, n_free=26)
eng = AttractEngine(cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
                    cfg['simulation problem variations'], device=local, rank=rank, world=world)
for _ in range(3):
    eng.run(1000, inf)
acc = {}
def lap(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0) + (t1 - t0) * 1e3; return t1
R = 10
for _ in range(R):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    eng.plan.table_clear(); t = lap('clear', t)
    st = eng.plan.attract_run(eng.lo, eng.hi, 1000, inf); t = lap('attract_run (device %.2f)' % st['ms_total'], t)
    buf = eng._gather
    mine, everyone = buf['mine'], buf['everyone']
    eng.plan.table_export_payload(mine.data_ptr(), mine.shape[0]); t = lap('export', t)
    dist.all_gather_into_tensor(everyone.view(-1, 8), mine); t = lap('all_gather', t)
    headers = everyone[:, 0, :].cpu(); t = lap('headers.cpu', t)
    rows = everyone.shape[1]
    if rank > 0: eng.plan.table_merge_device(everyone[0].data_ptr(), rank * rows)
    if rank < world - 1: eng.plan.table_merge_device(everyone[rank + 1].data_ptr(), (world - 1 - rank) * rows)
    t = lap('merge', t)
if rank == 0:
    print({k: round(v / R, 3) for k, v in acc.items()})
dist.destroy_process_group()
