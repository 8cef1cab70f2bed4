"""Summarise the loops of one kernel's SASS (instruction mix per backward branch).
usage: python tools/sass_loops.py <cubin> <kernel>"""
import re, subprocess, sys
from collections import Counter
out = subprocess.run(['cuobjdump', '-sass', '-fun', sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
idx = {a: i for i, (a, _) in enumerate(ins)}
def op(t):
    return re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]
print('total', len(ins), dict(Counter(op(t) for _, t in ins).most_common(12)))
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+.*?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in idx:
            body = ins[idx[tgt]:i + 1]
            print(hex(tgt), hex(a), len(body), dict(Counter(op(x[1]) for x in body).most_common(8)))
