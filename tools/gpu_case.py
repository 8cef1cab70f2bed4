"""One named attract case, timed by the library's own CUDA events (stats of bsi_attract_run).
usage: gpu_case.py CASE[,CASE...] [log2_states] [plan-opts dict] [reps]
cases: s0 s1 s2 (random N=64 K=3, seeds 0/1/2, max_t 1000), k2s0 k2s1 k1s0 (ordered regime),
       c3 (cambium2, no caps), n128 (random N=128 K=3 seed 0, max_t 1000)"""
import os
import sys
from math import inf

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests', 'golden'))
from boolsi_b200 import _lib, synth  # noqa: E402

CASES = {'s0': (64, 3, 0, 1000), 's1': (64, 3, 1, 1000), 's2': (64, 3, 2, 1000), 'k2s0': (64, 2, 0, 1000),
         'k2s1': (64, 2, 1, inf), 'k1s0': (64, 1, 0, inf), 'n128': (128, 3, 0, 1000)}


def packed(case, n_free):
    if case == 'c3':
        import codec
        from boolsi_b200.constants import NodeStateRange
        pub = codec.load('published.json')
        _, preds, tts = codec.decode_network(codec.load('networks.json')['networks']['cambium'])
        origin, variations = codec.decode_problem(pub['cambium2_problem'], NodeStateRange)
        return _lib.PackedProblem(preds, tts, origin, variations), inf
    n, k, seed, max_t = CASES[case]
    cfg = synth.random_attract_problem(n, k, seed, n_free=n_free)
    return _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
                              cfg['simulation problem variations']), max_t


def main():
    cases = sys.argv[1].split(',')
    n_free = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    opts = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    peak = _lib.measure_lop3_peak(0)['lop3_per_s']
    for case in cases:
        pk, max_t = packed(case, n_free)
        n = min(pk.n_problems, 1 << n_free)
        with _lib.Plan(pk, _lib.MODE_ATTRACT, **opts) as plan:
            ops = plan.op_counts()
            best = None
            for _ in range(reps):
                plan.table_clear()
                st = plan.attract_run(0, n, max_t, inf)
                if best is None or st['ms_total'] < best['ms_total']:
                    best = st
            alg = best['steps'] * (ops['update_lop3'] + ops['compare_lop3']) / 32
            print(case, 'n=2^%d' % n_free, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in best.items()},
                  'attractors', plan.table_count(), 'steps/s %.3e' % (best['steps'] / best['ms_total'] * 1e3),
                  'whole_step_frac %.3f' % (alg / (best['ms_total'] * 1e-3) / peak), flush=True)


if __name__ == '__main__':
    main()
