"""GPU: the maximum network size (129..256 nodes, three and four key words per state) against the
C oracle -- attract, simulate, target.

Written after round 1's GPU time was spent: the plans specialise offline (NVRTC) and the two
oracles agree at this size (tests/test_coracle_golden.py), and the host emulation of the kernels
passes at 160 and 256 nodes (tests/test_kernel_emul.py), but these tests have NOT run on a GPU
yet, so they live here and not under tests/.  Next round: run
`python -m pytest tools/next_round/test_gpu_large.py -m gpu -q -p no:cacheprovider` on a B200 (from the
repo root, so that tests/conftest.py is not needed: the marker below is registered here), fix what
fails, then move the file to tests/."""
from math import inf

import numpy as np
import pytest

from boolsi_b200 import _lib, synth
from oracle import coracle

pytestmark = pytest.mark.gpu


def _problem(n_nodes, k, seed, n_free):
    cfg = synth.random_attract_problem(n_nodes, k, seed, n_free=n_free)
    args = (cfg['incoming node lists'], cfg['truth tables'], cfg['origin simulation problem'],
            cfg['simulation problem variations'])
    return args, coracle.Space(*args), _lib.PackedProblem(*args)


@pytest.mark.parametrize('n_nodes,k,seed,max_t', [(160, 2, 3, 300), (200, 2, 0, 300), (256, 2, 1, 400),
                                                  (256, 3, 0, 60)])
def test_attract_large_networks_vs_c_oracle(n_nodes, k, seed, max_t):
    args, sp, pk = _problem(n_nodes, k, seed, 8)
    assert pk.kw == (n_nodes + 63) // 64
    rec = coracle.attract_records(sp, max_t, inf)
    want = sorted((key, v[0], v[1], v[2], v[3]) for key, v in coracle.attract_table(rec, sp.kw).items())
    with _lib.Plan(pk, _lib.MODE_ATTRACT) as plan:
        st = plan.attract_run(0, pk.n_problems, max_t, inf)
        got = sorted((_lib.code_from_words(r['key'][:pk.kw]), int(r['n']), int(r['sum_l']), int(r['sum_l2']),
                      int(r['lam'])) for r in plan.table_read())
        assert got == want
        assert st['steps'] == rec['steps']
        assert st['n_without'] == int((rec['found'] == 0).sum())
        for key, _, _, _, lam in want[:2]:
            orbit = plan.attractor_states(key, lam)
            assert orbit[0] == key == min(orbit) and len(set(orbit)) == lam


def test_simulate_256_nodes_vs_c_oracle():
    args, sp, pk = _problem(256, 3, 2, 6)
    want, _, _ = coracle.simulate_states(sp, 40, 0, pk.n_problems)
    with _lib.Plan(pk, _lib.MODE_SIMULATE) as plan:
        got, _ = plan.simulate_run(0, pk.n_problems, 40)
    assert np.array_equal(np.asarray(got, np.uint64), np.asarray(want, np.uint64))
