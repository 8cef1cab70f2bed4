"""The reference import shim lives in oracle/refshim.py (test / measurement infrastructure); this
module keeps the name the golden-fixture scripts import."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.refshim import *  # noqa: F401,F403,E402
from oracle.refshim import NoMpiComm, REFERENCE_ROOT, install, new_db_conn, reference_available  # noqa: F401,E402
