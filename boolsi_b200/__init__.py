"""boolsi_b200: the data-parallel core of BoolSi on B200.

`attract_master`, `simulate_master`, `target_master` take the reference's arguments and leave the
reference's post-conditions on `db_conn.root` (reference cli.py:200-204, 256-261, 314-319 are the
call sites they drop into); everything between is libbsi's CUDA path.  No CPU fallback.
"""
from .attract import AggregatedAttractor, AttractEngine, attract_master, attract_sharded  # noqa: F401
from .db import ResultStore, init_attractor_db_structure, init_simulation_db_structure  # noqa: F401
from .simulate import Simulation, simulate_master  # noqa: F401
from .target import merge_sharded_hits, target_master, target_sharded  # noqa: F401
