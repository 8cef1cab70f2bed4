"""Benchmark of the hot path on BASELINE.json's headline configuration (configs[3]):
`attract` on a random Kauffman network, N=64, K=3, over 2^26 initial states, max_t = 1000.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference ...                   # the CPU restatement of the reference

One "step" = one pass of the whole job (all 2^26 initial states; under torchrun each rank takes a
contiguous, word-aligned 1/N of the index space, and the per-GPU attractor tables are merged with
one NCCL all_gather).  The metric is trajectory-steps/s with the reference's default-mode step
count (sum over problems of the stop time, model.py:235).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from math import inf

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = 'trajectory-steps/sec, attract on random N=64 K=3 net'
UNIT = 'trajectory-steps/s'
N_NODES, K_IN, SEED, N_FREE, MAX_T = 64, 3, 0, 26, 1000


def workload(n_free=N_FREE, seed=SEED):
    from boolsi_b200 import synth
    cfg = synth.random_attract_problem(N_NODES, K_IN, seed, n_free=n_free)
    return cfg, (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
                 cfg['simulation problem variations'])


def config_dict(n_free, n_gpus):
    return {'workload': 'configs[3]: synthetic random Kauffman net N=%d K=%d (numpy default_rng seed %d), '
                        'attract over 2^%d initial states (nodes 0..%d free), max_t=%d, no attractor-length cap'
                        % (N_NODES, K_IN, SEED, n_free, n_free - 1, MAX_T),
            'n_problems': 2 ** n_free, 'max_t': MAX_T,
            'parallelism': 'index space sharded over %d GPU(s), one NCCL all_gather of attractor tables' % n_gpus,
            'l2_policy': 'no input buffers exist: initial states are generated on the device from the '
                         'problem index; state lives in registers and checkpoints in shared memory; the found '
                         'queue (20 MB) and the attractor table are rewritten every step, so nothing is carried '
                         'between timed iterations'}


# ---- CPU side -----------------------------------------------------------------------------------
def _py_worker(job):
    from oracle import oracle
    args, lo, hi = job
    t0 = time.perf_counter()
    r = oracle.attract(*args, MAX_T, inf, lo, hi)
    return r['steps'], time.perf_counter() - t0


def cpu_python_port(args, n_problems, budget_s, cores):
    """oracle/oracle.py (the reference's algorithm, pure Python like the reference) in `cores`
    independent processes, each on a contiguous slice of a strided sample of the index space."""
    import multiprocessing as mp
    from oracle import oracle
    t0 = time.perf_counter()
    probe = oracle.attract(*args, MAX_T, inf, 0, 8)
    rate1 = probe['steps'] / (time.perf_counter() - t0)
    per_core = max(8, int(rate1 * budget_s / (MAX_T * 0.95)))
    stride = max(per_core, n_problems // cores)
    jobs = [(args, c * stride, min(n_problems, c * stride + per_core)) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(_py_worker, jobs)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    return {'value': steps / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/oracle.py (pure-Python restatement of model.py/attract.py), %d processes x %d '
                      'problems (contiguous slices at stride %d of the same index space), %d trajectory-steps in %.1f s'
                      % (cores, per_core, stride, steps, wall)}


def cpu_c_port(args, n_problems, budget_s, cores):
    """oracle/bsi_oracle.c (scalar C restatement, hash set of visited states), pthreads."""
    from oracle import coracle
    sp = coracle.Space(*args)
    t0 = time.perf_counter()
    rec = coracle.attract_records(sp, MAX_T, inf, 0, 512, n_threads=cores, want_records=False)
    rate = rec['steps'] / (time.perf_counter() - t0)
    n = int(min(n_problems, max(1024, rate * budget_s / (MAX_T * 0.95))))
    t0 = time.perf_counter()
    rec = coracle.attract_records(sp, MAX_T, inf, 0, n, n_threads=cores, want_records=False)
    wall = time.perf_counter() - t0
    return {'value': rec['steps'] / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/bsi_oracle.c (scalar C restatement), %d threads, first %d problems, '
                      '%d trajectory-steps in %.1f s' % (cores, n, rec['steps'], wall)}


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path = its pure-Python
    algorithm (the unmodified package cannot travel to the GPU box and has no compiled part, so
    the pinned restatement oracle/oracle.py stands in), on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg, args = workload(a.n_free)
    n = 2 ** a.n_free
    vals = []
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, a.steps + a.warmup)))
    t_steps = []
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        r = cpu_python_port(args, n, per_step_budget, cores)
        if i >= a.warmup:
            vals.append(r)
            t_steps.append(time.perf_counter() - t0)
    value = sum(v['value'] for v in vals) / len(vals)
    per_step_wall_ms = 1e3 * sum(t_steps) / len(t_steps)     # wall time of one bounded sample
    c_port = cpu_c_port(args, n, 5.0, cores)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': per_step_wall_ms, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'Python int state codes (the reference: lists of bool)',
            'data': 'synthetic', 'config': config_dict(a.n_free, a.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': vals[-1]['sample']},
            'cpu_baseline_c': c_port,
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every few ms DURING the timed region
    (the same quantities as the recipe's nvidia-smi line: clocks.sm, clocks.max.sm,
    clocks_event_reasons.*)."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap'}

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((mhz, reasons, power))
            except Exception as e:      # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.004)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling unavailable: %s' % self.err]}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(n for b, n in self.REASONS.items() if bits & b),
                'samples': len(sm), 'power_w_max': max(r[2] for r in self.rows)}


# ---- GPU side -------------------------------------------------------------------------------------
def run_gpu(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from boolsi_b200 import _lib
    from boolsi_b200.attract import attract_sharded, AttractEngine

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg, args = workload(a.n_free)
    n = 2 ** a.n_free

    t0 = time.perf_counter()
    eng = AttractEngine(*args, device=local, rank=rank, world=world)
    nvrtc_cold_s = time.perf_counter() - t0
    peak = _lib.measure_lop3_peak(local)
    ops = eng.plan.op_counts()
    lop3_per_step = (ops['update_lop3'] + ops['compare_lop3']) / 32.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        return eng.run(MAX_T, inf)          # kernels + all_gather + merge, results stay on the device

    for _ in range(a.warmup):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    stats = []
    for _ in range(a.steps):
        stats.append(one_step())
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    sums = torch.tensor([sum(s['steps'] for s in stats), sum(s['launches'] for s in stats),
                         sum(s['n_problems'] for s in stats)], dtype=torch.int64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums)
    ms = float(t.item())
    total_steps, launches, problems = (int(x) for x in sums.tolist())
    value = total_steps / (ms * 1e-3)

    # e2e: the public call (boolsi_b200.attract.attract_sharded) from HOST structures to HOST
    # results, every step: pack descriptors, create the plan (NVRTC result cached in-process),
    # run, all_gather, read the table back, regenerate the attractor states.
    # (one untimed call first: the orbit-regeneration and table-export kernels are launched for the
    # first time in this leg, and the driver loads a kernel's code lazily at its first launch)
    attract_sharded(*args, max_t=MAX_T, max_attractor_l=inf, device=local, rank=rank, world=world)
    barrier()
    e0 = time.perf_counter()
    e_steps = 0
    for _ in range(a.steps):
        res = attract_sharded(*args, max_t=MAX_T, max_attractor_l=inf, device=local, rank=rank, world=world)
        e_steps += res['steps_total']
    barrier()
    e_ms = (time.perf_counter() - e0) * 1e3
    et = torch.tensor([e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = e_steps / (float(et.item()) * 1e-3)

    if rank == 0:
        main_ms = sum(s['ms_main'] for s in stats) / len(stats)
        alg_lop3 = stats[-1]['steps'] * lop3_per_step         # this rank's launch
        achieved = alg_lop3 / (main_ms * 1e-3) / 1e12
        peak_t = peak['lop3_per_s'] / 1e12
        traffic = None
        tf = os.path.join(REPO, 'profiles', 'r01_main_kernel_dram_bytes.json')
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get('dram_bytes_per_launch')
        rec_bytes = res['d2h_bytes']
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'u32 (bit-sliced boolean, 32 trajectories per word; integer statistics u64)',
            'data': 'synthetic', 'config': config_dict(a.n_free, world),
            'roofline': {'bound': 'alu_lop3', 'achieved': achieved, 'peak': peak_t, 'unit': 'TLOP3/s',
                         'frac': achieved / peak_t, 'traffic': traffic,
                         'kernel': 'bsi_attract_dense', 'kernel_ms': main_ms,
                         'algorithmic_lop3_per_trajectory_step': lop3_per_step,
                         'peak_source': 'bsi_measure_lop3_peak on this GPU in this run (independent lop3.b32 '
                                        'chains; = 64 lop3/clk/SM x %d SMs x %.0f MHz)' % (peak['n_sms'], peak['implied_sm_mhz']),
                         'whole_step_frac': value / world * lop3_per_step / peak['lop3_per_s'],
                         'hbm_result_writeout_gbs': rec_bytes / (ms / a.steps * 1e-3) / 1e9},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': res['h2d_bytes'],
                    'd2h_bytes_per_step': res['d2h_bytes'], 'ms_per_step': float(et.item()) / a.steps,
                    'note': 'host structures in, host attractor records + states out; NVRTC cubin cached '
                            'in-process after the first plan (cold specialisation: %.2f s, reported here, not timed)' % nvrtc_cold_s},
            'gpu_launches': launches,
            'clocks': clocks,
            'device_ms': {'main': main_ms, 'resolve': sum(s['ms_resolve'] for s in stats) / len(stats),
                          'total': sum(s['ms_total'] for s in stats) / len(stats)},
            'result': {'attractors': res['n_attractors'], 'found': res['n_found'], 'without': res['n_without'],
                       'trajectory_steps': res['steps_total']},
        }
        if world == 1 and not a.no_cpu:
            # the same job on the other two seeded networks SURVEY.md section 8d names (not the headline)
            others = {}
            for seed in (1, 2):
                _, oargs = workload(a.n_free, seed)
                oeng = AttractEngine(*oargs, device=local)
                oeng.run(MAX_T, inf)
                ost = [oeng.run(MAX_T, inf) for _ in range(3)]
                oops = oeng.plan.op_counts()
                best = min(ost, key=lambda x: x['ms_total'])
                others['seed%d' % seed] = {
                    'value': best['steps'] / (best['ms_total'] * 1e-3), 'ms': best['ms_total'],
                    'found_fraction': best['n_found'] / best['n_problems'],
                    'whole_step_frac': best['steps'] * (oops['update_lop3'] + oops['compare_lop3']) / 32.0
                    / (best['ms_total'] * 1e-3) / peak['lop3_per_s']}
                oeng.close()
            line['other_seeds'] = others
            cores = os.cpu_count() or 1
            line['cpu_baseline'] = cpu_python_port(args, n, 12.0, cores)
            line['cpu_baseline_c'] = cpu_c_port(args, n, 8.0, cores)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n-free', type=int, default=N_FREE, help='log2 of the number of initial states')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == '__main__':
    main()
