"""Hot path of every loop of a kernel: the SHORTEST instruction path from each backward-branch
target back to itself in the control-flow graph (the cold handlers only ever add instructions),
with its instruction mix -- LDL/STL on it are register spills paid every iteration.
usage: python tools/sass_hotpath.py <cubin> <kernel> [min_lop3]"""
import heapq, re, subprocess, sys
from collections import Counter
out = subprocess.run(['cuobjdump', '-sass', '-fun', sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
minl = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ins = []
for l in out.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
idx = {a: i for i, (a, _) in enumerate(ins)}
def op(t):
    return re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]
succ = []
heads = set()
for i, (a, t) in enumerate(ins):
    s = []
    m = re.search(r'\bBRA(?:\.[A-Z]+)*\s+(?:U?P\d+,\s*|UR\d+,\s*|!U?P\d+,\s*)?0x([0-9a-f]+)', t)
    o = op(t)
    if m and o == 'BRA':
        tgt = int(m.group(1), 16)
        if tgt in idx:
            s.append(idx[tgt])
            if tgt < a:
                heads.add(tgt)
        if t.startswith('@') or '.DIV' in t or re.search(r'BRA\.U\s+!?UP', t):
            s.append(i + 1)
    elif o not in ('EXIT', 'RET'):
        s.append(i + 1)
    succ.append([x for x in s if x < len(ins)])
for h in sorted(heads):
    src = idx[h]
    dist = {}; prev = {}
    pq = [(1, j, src) for j in succ[src]]
    heapq.heapify(pq)
    found = None
    while pq:
        d, i, p = heapq.heappop(pq)
        if i in dist:
            continue
        dist[i] = d; prev[i] = p
        if i == src:
            found = d; break
        for j in succ[i]:
            if j not in dist:
                heapq.heappush(pq, (d + 1, j, i))
    if found is None:
        continue
    path = []; i = prev[src]
    while i != src:
        path.append(ins[i][1]); i = prev[i]
    path.append(ins[src][1])
    c = Counter(op(x) for x in path)
    if c['LOP3'] >= minl:
        print(hex(h), 'len', len(path), dict(c.most_common(10)))
