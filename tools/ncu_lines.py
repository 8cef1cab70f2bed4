"""Per-source-line view of an ncu capture of an NVRTC-compiled kernel.
The kernels are compiled in memory, so ncu's own CUDA-source page has no metrics for them; this joins
ncu's SASS page (per-instruction counts and stall samples) with the line table of the same cubin
(nvdisasm -g), matched by instruction offset inside the kernel's .text section.

usage: ncu_lines.py REPORT.ncu-rep KERNEL UNIT.cubin [top_n]
The cubin is the unit file NVRTC produced for the same plan (BSI_CACHE_DIR keeps them: strip the
cache header with tools/ncu_lines.py --extract FILE OUT)."""
import csv
import io
import re
import struct
import subprocess
import sys
from collections import defaultdict


def extract(path, out):
    d = open(path, 'rb').read()
    n = struct.unpack('<3Q', d[8:32])
    off = 32 + n[0] + n[1]
    open(out, 'wb').write(d[off:off + n[2]])


def line_table(cubin, kernel):
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    table, line, on = {}, 0, False
    for s in txt.split('\n'):
        if s.startswith('.text.'):
            on = s.startswith('.text.%s:' % kernel)
            continue
        if not on:
            continue
        m = re.search(r'//## File "[^"]*", line (\d+)', s)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*)', s)
        if m:
            table[int(m.group(1), 16)] = (line, m.group(2))
    return table


def main():
    if sys.argv[1] == '--extract':
        return extract(sys.argv[2], sys.argv[3])
    rep, kernel, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name', kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hi]
    col = {k: h.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed', 'stall_no_inst', 'stall_long_sb',
                                   'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_branch_resolving', 'stall_membar')}
    body = [r for r in rows[hi + 1:] if len(r) == len(h)]
    base = int(body[0][col['Address']], 16)
    lt = line_table(cubin, kernel)
    agg = defaultdict(lambda: defaultdict(float))
    n_ins = 0
    for r in body:
        off = int(r[col['Address']], 16) - base
        line, text = lt.get(off, (0, '?'))
        op_ncu = r[col['Source']].split()[0] if r[col['Source']].split() else ''
        if op_ncu.startswith('@'):
            op_ncu = r[col['Source']].split()[1]
        if text != '?' and op_ncu.split('.')[0] not in text:
            raise SystemExit('cubin does not match the capture at offset %#x: %s / %s' % (off, r[col['Source']], text))
        a = agg[line]
        a['sass'] += 1
        for k in col:
            if k not in ('Address', 'Source'):
                a[k] += float(r[col[k]] or 0)
        n_ins += 1
    tot = defaultdict(float)
    for a in agg.values():
        for k, v in a.items():
            tot[k] += v
    print('%s: %d SASS instructions (%.0f KB), %.3g warp-instructions executed, %d samples' %
          (kernel, n_ins, n_ins * 16 / 1024, tot['Instructions Executed'], tot['# Samples']))
    print('stall samples: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * tot[k] / max(tot['# Samples'], 1))
                                        for k in col if k.startswith('stall_')))
    src = open('boolsi_b200/csrc/bsi_template.cuh').read().split('\n')
    print('%6s %6s %7s %7s %6s %6s  %s' % ('line', 'sass', 'exec%', 'samp%', 'noins%', 'lsb%', 'source'))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:top]:
        text = src[line - 1].strip()[:90] if 0 < line <= len(src) else ''
        print('%6d %6d %7.2f %7.2f %6.1f %6.1f  %s' % (
            line, a['sass'], 100 * a['Instructions Executed'] / max(tot['Instructions Executed'], 1),
            100 * a['# Samples'] / max(tot['# Samples'], 1), 100 * a['stall_no_inst'] / max(a['# Samples'], 1),
            100 * a['stall_long_sb'] / max(a['# Samples'], 1), text))


if __name__ == '__main__':
    main()
