"""Compile a named case's kernel units offline (NVRTC, no GPU) and report per-kernel SASS size and
resources.  usage: offline_units.py CASE OUTDIR  (cases of tools/gpu_case.py)"""
import glob
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tools')]
case, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
for f in glob.glob(os.path.join(out, '*')):
    os.remove(f)
os.environ['BSI_CACHE_DIR'] = out
import gpu_case  # noqa: E402
import ncu_lines  # noqa: E402
from boolsi_b200 import _lib  # noqa: E402
pk, _ = gpu_case.packed(case, 26)
opts = eval(sys.argv[3]) if len(sys.argv) > 3 else {}
with _lib.Plan(pk, _lib.MODE_ATTRACT, offline=True, **opts):
    pass
for f in sorted(glob.glob(os.path.join(out, 'bsi-*.cubin'))):
    u = re.search(r'-u(\d)\.cubin', f).group(1)
    dst = os.path.join(out, 'u%s.cubin' % u)
    ncu_lines.extract(f, dst)
    res = subprocess.run(['cuobjdump', '-res-usage', dst], capture_output=True, text=True).stdout
    sass = subprocess.run(['cuobjdump', '-sass', dst], capture_output=True, text=True).stdout
    fn, counts = None, {}
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            counts[fn] = 0
        elif fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
            counts[fn] += 1
    for m in re.finditer(r'Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', res):
        name = m.group(1)
        if name in counts and not name.startswith('_Z'):
            print('unit %s %-24s REG %3s STACK %5s SHARED %5s  SASS %5d instr (%3d KB)' % (
                u, name, m.group(2), m.group(3), m.group(4), counts[name], counts[name] * 16 // 1024))
        elif name in counts:
            print('unit %s   %-40s SASS %5d' % (u, name[:40], counts[name]))
