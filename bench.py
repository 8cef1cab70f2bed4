"""Benchmark of the hot path.

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference ...                       # the reference itself on the host cores
    python bench.py --config c5 ...                            # BASELINE.json configs[4] instead of configs[3]

Default workload = BASELINE.json's headline configuration (configs[3], `--config c4`): `attract` on a
random Kauffman network, N=64, K=3, over 2^26 initial states, max_t = 1000.  `--config c5` is
configs[4]: `target` on a random N=128, K=4 network with fixed nodes and per-step perturbations, 2^28
initial states, 500 steps.

One "step" = one pass of the whole job.  Under torchrun each rank takes a contiguous, word-aligned 1/N
of the index space and the per-GPU results are merged with one NCCL all_gather; with one process the
library's group run does the same over its devices.  The metric is trajectory-steps/s with the
reference's default-mode step count (sum over problems of the stop time, model.py:235).  Prints ONE
JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time
from math import inf

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

UNIT = 'trajectory-steps/s'
CONFIGS = {
    'c4': dict(mode='attract', n_nodes=64, k=3, seed=0, n_free=26, max_t=1000,
               metric='trajectory-steps/sec, attract on random N=64 K=3 net'),
    'c5': dict(mode='target', n_nodes=128, k=4, seed=0, n_free=28, max_t=500,
               metric='trajectory-steps/sec, target on random N=128 K=4 net with fixed nodes and perturbations'),
}


def workload(cfg, n_free=None, seed=None):
    """-> (parsed-input dict, positional arguments of the per-mode calls)"""
    from boolsi_b200 import synth
    n_free = cfg['n_free'] if n_free is None else n_free
    seed = cfg['seed'] if seed is None else seed
    if cfg['mode'] == 'attract':
        w = synth.random_attract_problem(cfg['n_nodes'], cfg['k'], seed, n_free=n_free)
        return w, (w['incoming node lists'], w['truth tables'], w['origin simulation problem'],
                   w['simulation problem variations'])
    w = synth.random_target_problem(cfg['n_nodes'], cfg['k'], seed, n_free=n_free)
    return w, (w['incoming node lists'], w['truth tables'], w['origin simulation problem'],
               w['simulation problem variations'], w['target node set'], w['target substate code'])


def config_dict(name, cfg, n_free, n_gpus):
    what = 'attract over' if cfg['mode'] == 'attract' else \
        'target (4 fixed nodes, constant perturbations at 8 time steps <= 100, 16 target nodes) over'
    return {'workload': 'configs[%d]: synthetic random Kauffman net N=%d K=%d (numpy default_rng seed %d), %s 2^%d '
                        'initial states (nodes 0..%d free), max_t=%d%s'
                        % (3 if name == 'c4' else 4, cfg['n_nodes'], cfg['k'], cfg['seed'], what, n_free, n_free - 1,
                           cfg['max_t'], ', no attractor-length cap' if cfg['mode'] == 'attract' else ''),
            'n_problems': 2 ** n_free, 'max_t': cfg['max_t'],
            'parallelism': 'index space sharded over %d GPU(s)%s' % (
                n_gpus, ', one NCCL all_gather of the per-GPU result records' if n_gpus > 1 else ''),
            'l2_policy': 'no input buffers exist: initial states are generated on the device from the problem index; '
                         'state lives in registers and checkpoints in shared memory; the result queues, the '
                         'attractor table and the cycle dictionary are cleared and rewritten every step, so nothing '
                         'is carried between timed iterations'}


# ---- CPU side -----------------------------------------------------------------------------------
def reference_sample(cfg, args, n_free, budget_s, cores):
    """The UNMODIFIED reference (baseline/_ref, or /root/reference) on `cores` host processes, each on a
    contiguous slice of the job's own index space sized for about `budget_s` seconds (oracle/refarm.py)."""
    from oracle import refarm
    target = None if cfg['mode'] == 'attract' else (args[4], args[5])
    probe = refarm.timed_sample(args[:4], n_free, cfg['max_t'], 1, 2, target=target)       # 4 problems, one process
    per_core = probe['steps'] / probe['wall_s']
    per_problem = probe['steps'] / 4.0
    m = 2
    while m + 1 <= n_free - 4 and (2 << m) * per_problem / per_core <= budget_s:
        m += 1
    r = refarm.timed_sample(args[:4], n_free, cfg['max_t'], cores, m, first_slice=1, target=target)
    return {'value': r['value'], 'unit': UNIT, 'cores': cores, 'kind': 'reference', 'sample': r['sample']}


def _py_worker(job):
    from oracle import oracle
    args, lo, hi, max_t = job
    t0 = time.perf_counter()
    r = oracle.attract(*args, max_t, inf, lo, hi)
    return r['steps'], time.perf_counter() - t0


def cpu_python_port(cfg, args, n_problems, budget_s, cores):
    """oracle/oracle.py (the reference's algorithm restated, pure Python like the reference) in `cores`
    independent processes, each on a contiguous slice of a strided sample of the index space."""
    import multiprocessing as mp
    from oracle import oracle
    max_t = cfg['max_t']
    t0 = time.perf_counter()
    probe = oracle.attract(*args, max_t, inf, 0, 8)
    rate1 = probe['steps'] / (time.perf_counter() - t0)
    per_core = max(8, int(rate1 * budget_s / (max_t * 0.95)))
    stride = max(per_core, n_problems // cores)
    jobs = [(args, c * stride, min(n_problems, c * stride + per_core), max_t) for c in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(_py_worker, jobs)
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    return {'value': steps / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/oracle.py (pure-Python restatement of model.py/attract.py), %d processes x %d '
                      'problems (contiguous slices at stride %d of the same index space), %d trajectory-steps in %.1f s'
                      % (cores, per_core, stride, steps, wall)}


def cpu_c_port(cfg, args, n_problems, budget_s, cores):
    """oracle/bsi_oracle.c (scalar C restatement, hash set of visited states), pthreads."""
    from oracle import coracle
    sp = coracle.Space(*args[:4])
    max_t = cfg['max_t']

    def run(n):
        if cfg['mode'] == 'attract':
            return coracle.attract_records(sp, max_t, inf, 0, n, n_threads=cores, want_records=False)
        return coracle.target_records(sp, args[4], args[5], max_t, 0, n, n_threads=cores)
    t0 = time.perf_counter()
    rec = run(512)
    rate = rec['steps'] / (time.perf_counter() - t0)
    n = int(min(n_problems, max(1024, rate * budget_s / (max_t * 0.95))))
    t0 = time.perf_counter()
    rec = run(n)
    wall = time.perf_counter() - t0
    return {'value': rec['steps'] / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'oracle/bsi_oracle.c (scalar C restatement), %d threads, first %d problems, '
                      '%d trajectory-steps in %.1f s' % (cores, n, rec['steps'], wall)}


def run_reference(a, name, cfg):
    """--impl reference: the reference's own CPU implementation of the path -- the unmodified package
    (`baseline/_ref`, installed by __graft_entry__.build(); it travels to the GPU box with the repo),
    `attract_master` / `target_master` through `generic_single_process_master`, one process per host
    core.  Each step is a bounded sample of the job (contiguous slices of its index space)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import refarm
    cores = os.cpu_count() or 1
    w, args = workload(cfg, a.n_free)
    n = 2 ** a.n_free
    # (seconds of CPU work per step; BSI_BENCH_REF_BUDGET: the test-suite's short run)
    per_step_budget = float(os.environ.get('BSI_BENCH_REF_BUDGET', 0)) or \
        max(2.0, min(20.0, 150.0 / max(1, a.steps + a.warmup)))
    have_reference = refarm.reference_root() is not None
    vals, t_steps = [], []
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        if have_reference:
            r = reference_sample(cfg, args, a.n_free, per_step_budget, cores)
        elif cfg['mode'] == 'attract':
            r = cpu_python_port(cfg, args, n, per_step_budget, cores)
        else:
            r = cpu_c_port(cfg, args, n, per_step_budget, cores)
        if i >= a.warmup:
            vals.append(r)
            t_steps.append(time.perf_counter() - t0)
    value = sum(v['value'] for v in vals) / len(vals)
    line = {'impl': 'reference', 'metric': cfg['metric'], 'value': value, 'unit': UNIT, 'n_gpus': a.gpus,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': 1e3 * sum(t_steps) / len(t_steps),
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'Python lists of bool / int state codes (the reference)', 'data': 'synthetic',
            'config': config_dict(name, cfg, a.n_free, a.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': vals[-1]['kind'],
                             'sample': vals[-1]['sample']},
            'cpu_baseline_c': cpu_c_port(cfg, args, n, 5.0, cores),
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if not have_reference:
        line['note'] = 'baseline/_ref is absent (run __graft_entry__.build() where /root/reference exists): the ' \
                       'pinned restatement under oracle/ was timed instead'
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
def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest
    `ncu --set full` capture committed under profiles/ (a profiler cannot run inside a timed bench)."""
    best = None
    for fn in sorted(os.listdir(os.path.join(REPO, 'profiles'))):
        if fn.endswith('_dram_bytes.json'):
            d = json.load(open(os.path.join(REPO, 'profiles', fn)))
            if d.get('kernel', 'bsi_attract_dense') == kernel:
                best = (d.get('dram_bytes_per_launch'), 'profiles/' + fn)
    return best if best else (None, None)


def run_gpu(a, name, cfg):
    import torch
    import torch.distributed as dist
    from boolsi_b200 import _lib
    from boolsi_b200.attract import AttractEngine, attract_sharded
    from boolsi_b200.target import target_devices, target_sharded

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    w, args = workload(cfg, a.n_free)
    n = 2 ** a.n_free
    attract = cfg['mode'] == 'attract'
    max_t = cfg['max_t']

    t0 = time.perf_counter()
    if attract:
        eng = AttractEngine(*args, device=local, rank=rank, world=world)
        plan, kernel = eng.plan, 'bsi_attract_dense'
    else:
        from boolsi_b200.problems import shard_range
        pk = _lib.PackedProblem(*args[:4])
        plan = _lib.Plan(pk, _lib.MODE_TARGET, device=local, target=_lib.make_target(args[4], args[5], pk.kw))
        lo, hi = shard_range(pk.n_problems, rank, world)
        kernel = 'bsi_target_main'
    nvrtc_cold_s = time.perf_counter() - t0
    peak = _lib.measure_lop3_peak(local)
    ops = plan.op_counts()
    compare = ops['compare_lop3'] if attract else len(w['target node set'])
    # target: past the last perturbation time the kernel folds the fixed nodes into their readers; the
    # smaller update count is taken for EVERY step (a lower bound of the algorithmic work)
    update = ops['update_lop3'] if attract else min(ops['update_lop3'], ops['update_quiet_lop3'])
    lop3_per_step = (update + compare) / 32.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        if attract:
            # kernels + all_gather + merge, results stay on the device; the attractor table AND the cycle
            # dictionary start from empty in every step
            return eng.run(max_t, inf)
        # target: this rank's slice; the hit lists (first 10 per rank) come back to the host
        _, _, st = plan.target_run(lo, hi, max_t, 10)
        return st

    for _ in range(a.warmup):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # K steps between two events on torch's current stream.  world > 1: the plan launches on that very
    # stream.  world == 1: the plan has its own stream and every step ends with a host wait for it, so the
    # two events still bracket all device work of the K steps (and the host gaps between them).
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

    # e2e: the public call from HOST structures to HOST results, every step: pack the descriptors, look
    # the plan up (kernels specialised once per process; the plan object is reused between calls), run,
    # exchange, read the records back, orbit states of the attractors / trajectories are left out only
    # where the reference's caller would not ask for them either.
    def e2e_call():
        if attract:
            return attract_sharded(*args, max_t=max_t, max_attractor_l=inf, device=local, rank=rank, world=world)
        if world == 1:
            r = target_devices(*args, max_t=max_t, n_to_find=10, devices=[local])
        else:
            r = target_sharded(*args, max_t=max_t, n_to_find=10, device=local, rank=rank, world=world)
        r['h2d_bytes'] = sum(x.nbytes for x in _lib.PackedProblem(*args[:4]).arrays.values())
        r['d2h_bytes'] = 16 * len(r['hit_index']) + 128
        return r
    e2e_call()          # (untimed: first launch of the export / orbit kernels, group creation)
    barrier()
    e0 = time.perf_counter()
    e_steps = 0
    for _ in range(a.steps):
        res = e2e_call()
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
        traffic, traffic_src = measured_traffic(kernel)
        line = {
            'metric': cfg['metric'], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': ms / a.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'u32 (bit-sliced boolean, 32 trajectories per word; integer statistics u64)',
            'data': 'synthetic', 'config': config_dict(name, cfg, a.n_free, world),
            'roofline': {'bound': 'alu_lop3', 'achieved': achieved, 'peak': peak_t, 'unit': 'TLOP3/s',
                         'frac': achieved / peak_t, 'traffic': traffic, 'traffic_source': traffic_src,
                         'kernel': kernel, 'kernel_ms': main_ms,
                         'algorithmic_lop3_per_trajectory_step': lop3_per_step,
                         'peak_source': 'bsi_measure_lop3_peak on this GPU in this run (independent lop3.b32 '
                                        'chains; = 64 lop3/clk/SM x %d SMs x %.0f MHz; MEASURED_PEAKS.json holds no '
                                        'integer-logic figure)' % (peak['n_sms'], peak['implied_sm_mhz']),
                         'whole_step_frac': value / world * lop3_per_step / peak['lop3_per_s'],
                         'hbm_result_writeout_gbs': res['d2h_bytes'] / (ms / a.steps * 1e-3) / 1e9},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': res['h2d_bytes'],
                    'd2h_bytes_per_step': res['d2h_bytes'], 'ms_per_step': float(et.item()) / a.steps,
                    'note': 'host structures in, host result records%s out; kernels specialised once per process '
                            '(cold NVRTC specialisation: %.2f s, reported here, not timed) and the plan object '
                            'reused between calls' % (' + attractor states + node activity' if attract else '',
                                                      nvrtc_cold_s)},
            'gpu_launches': launches,
            'clocks': clocks,
            'device_ms': {'main': main_ms, 'resolve': sum(s['ms_resolve'] for s in stats) / len(stats),
                          'total': sum(s['ms_total'] for s in stats) / len(stats)},
        }
        if attract:
            line['result'] = {'attractors': res['n_attractors'], 'found': res['n_found'], 'without': res['n_without'],
                              'trajectory_steps': res['steps_total']}
        else:
            line['result'] = {'hits': int(res['n_hits_total']), 'first_hits': [int(i) for i in res['hit_index'][:5]],
                              'trajectory_steps': int(res['steps_total'])}
        if world == 1 and not a.no_cpu:
            if attract:
                # the same job on the other two seeded networks SURVEY.md section 8d names (not the headline)
                others = {}
                for seed in (1, 2):
                    _, oargs = workload(cfg, a.n_free, seed)
                    oeng = AttractEngine(*oargs, device=local)
                    oeng.run(max_t, inf)
                    ost = [oeng.run(max_t, inf) for _ in range(3)]
                    oops = oeng.plan.op_counts()
                    best = min(ost, key=lambda x: x['ms_total'])
                    others['seed%d' % seed] = {
                        'value': best['steps'] / (best['ms_total'] * 1e-3), 'ms': best['ms_total'],
                        'ms_main': best['ms_main'], 'ms_resolve': best['ms_resolve'],
                        'found_fraction': best['n_found'] / best['n_problems'],
                        'whole_step_frac': best['steps'] * (oops['update_lop3'] + oops['compare_lop3']) / 32.0
                        / (best['ms_total'] * 1e-3) / peak['lop3_per_s']}
                    oeng.close()
                line['other_seeds'] = others
            cores = os.cpu_count() or 1
            from oracle import refarm
            if refarm.reference_root() is not None:
                line['cpu_baseline'] = reference_sample(cfg, args, a.n_free, 12.0, cores)
                if attract:
                    line['cpu_baseline_port'] = cpu_python_port(cfg, args, n, 6.0, cores)
            elif attract:
                line['cpu_baseline'] = cpu_python_port(cfg, args, n, 12.0, cores)
            else:
                line['cpu_baseline'] = cpu_c_port(cfg, args, n, 8.0, cores)
            line['cpu_baseline_c'] = cpu_c_port(cfg, args, n, 8.0, cores)
            if attract:
                # BASELINE.json configs[0] / configs[1]: what a FRESH process pays for an 8-trajectory job --
                # cold, with the compiled kernels on disk, and the reference's masters on the host
                try:
                    sys.path.insert(0, os.path.join(REPO, 'tools'))
                    import cold_start
                    line['small_jobs'] = cold_start.measure(['attract_example2', 'simulate_glider_t100'])
                except Exception as e:      # noqa: BLE001  (an extra: never fails the bench line)
                    line['small_jobs'] = {'unavailable': repr(e)[-300:]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c4', choices=sorted(CONFIGS))
    ap.add_argument('--n-free', type=int, default=None, help='log2 of the number of initial states')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg and the other seeds')
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    if a.n_free is None:
        a.n_free = cfg['n_free']
    if a.impl == 'reference':
        run_reference(a, a.config, cfg)
    else:
        run_gpu(a, a.config, cfg)


if __name__ == '__main__':
    main()
