"""Hot path of every loop of a kernel: the SHORTEST instruction path from each backward-branch
target back to itself in the control-flow graph (the cold handlers only ever add instructions),
with its instruction mix -- LDL/STL on it are register spills paid every iteration.
usage: python tools/sass_hotpath.py <cubin> <kernel> [min_lop3]"""
import heapq
import re
import subprocess
import sys
from collections import Counter

_BRANCH = re.compile(r'\bBRA(?:\.[A-Z]+)*\s+(?:U?P\d+,\s*|UR\d+,\s*|!U?P\d+,\s*)?0x([0-9a-f]+)')


def _op(text):
    return re.sub(r'^@!?U?P\d+\s+', '', text).split()[0].split('.')[0]


def hot_loops(cubin_path, kernel, min_lop3=100):
    """-> [(loop head address, #instructions on the hot path, Counter of opcodes)] for the loops
    of `kernel` whose hot path has at least `min_lop3` LOP3."""
    sass = subprocess.run(['cuobjdump', '-sass', '-fun', kernel, cubin_path], capture_output=True,
                          text=True, check=True).stdout
    ins = []
    for line in sass.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    index = {address: i for i, (address, _) in enumerate(ins)}
    successors, heads = [], set()
    for i, (address, text) in enumerate(ins):
        nxt = []
        m = _BRANCH.search(text)
        opcode = _op(text)
        if m and opcode == 'BRA':
            target = int(m.group(1), 16)
            if target in index:
                nxt.append(index[target])
                if target < address:
                    heads.add(target)
            if text.startswith('@') or '.DIV' in text or re.search(r'BRA\.U\s+!?UP', text):
                nxt.append(i + 1)               # conditional: may fall through
        elif opcode not in ('EXIT', 'RET'):
            nxt.append(i + 1)
        successors.append([j for j in nxt if j < len(ins)])
    loops = []
    for head in sorted(heads):
        source = index[head]
        dist, prev = {}, {}
        queue = [(1, j, source) for j in successors[source]]
        heapq.heapify(queue)
        closed = False
        while queue:
            d, i, p = heapq.heappop(queue)
            if i in dist:
                continue
            dist[i], prev[i] = d, p
            if i == source:
                closed = True
                break
            for j in successors[i]:
                if j not in dist:
                    heapq.heappush(queue, (d + 1, j, i))
        if not closed:
            continue
        path, i = [ins[source][1]], prev[source]
        while i != source:
            path.append(ins[i][1])
            i = prev[i]
        mix = Counter(_op(text) for text in path)
        if mix['LOP3'] >= min_lop3:
            loops.append((head, len(path), mix))
    return loops


if __name__ == '__main__':
    for head, length, mix in hot_loops(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 100):
        print(hex(head), 'len', length, dict(mix.most_common(10)))
