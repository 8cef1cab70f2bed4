"""CPU: static checks of the generated sm_100a code for the headline network (NVRTC + cuobjdump,
no GPU): the four hot loops of the trajectory kernel stay free of register spills and close to
their LOP3 floor.  This is the guard for the roofline work: the numbers are what DESIGN.md 3.1
and profiles/README.md quote."""
import os
import shutil
import sys

import pytest

from boolsi_b200 import _lib, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import sass_hotpath  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which('cuobjdump') is None, reason='needs the CUDA toolkit')


def _headline_plan(**opts):
    cfg = synth.random_attract_problem(64, 3, 0, n_free=26)
    pk = _lib.PackedProblem(cfg['incoming node lists'], cfg['truth tables'],
                            cfg['origin simulation problem'], cfg['simulation problem variations'])
    return _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True, **opts)


def test_dense_kernel_hot_loops_have_no_spills_and_stay_at_the_lop3_floor(tmp_path):
    with _headline_plan() as plan:
        ops = plan.op_counts()
        cubin = tmp_path / 'dense.cubin'
        cubin.write_bytes(plan.cubin())
    loops = sass_hotpath.hot_loops(str(cubin), 'bsi_attract_dense')
    # one loop per number of armed checkpoint slots (0..3), two steps per iteration
    assert len(loops) == 4, loops
    update, per_slot = ops['update_lop3'], ops['filter_lop3']
    for armed, (_, length, mix) in enumerate(sorted(loops, key=lambda l: l[2]['LOP3'])):
        assert mix['LDL'] == 0 and mix['STL'] == 0, (armed, dict(mix))     # no spills on the hot path
        assert mix['LDS'] == 0 and mix['LDG'] == 0, (armed, dict(mix))     # state and filters in registers
        floor = 2 * (update + armed * per_slot)
        # (the compiler folds an op or two of the update; each armed slot adds its filter words
        # plus the merge of the two accumulators and the any-lane test)
        assert floor - 4 <= mix['LOP3'] <= floor + 2 * (2 * armed + 1), (armed, dict(mix))
        assert length - mix['LOP3'] <= 16, (armed, length, dict(mix))      # loop control + two exits


def test_third_warp_configuration_compiles_to_168_registers(tmp_path):
    """The documented tuning knob: smem_slots = 2, min_blocks = 6."""
    with _headline_plan(smem_slots=2, min_blocks=6) as plan:
        cubin = tmp_path / 'dense168.cubin'
        cubin.write_bytes(plan.cubin())
    import subprocess
    usage = subprocess.run(['cuobjdump', '--dump-resource-usage', str(cubin)], capture_output=True, text=True).stdout
    line = [l for l in usage.splitlines() if 'REG:' in l][0]
    assert 'REG:168' in line, line
    with pytest.raises(_lib.BsiError):
        _headline_plan(smem_slots=4)
