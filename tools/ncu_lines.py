"""Per-source-line view of an ncu capture of an NVRTC-compiled kernel.
The kernels are compiled in memory, so ncu's own CUDA-source page has no metrics for them; this joins
ncu's SASS page (per-instruction counts and stall samples) with the line table of the same cubin
(nvdisasm -g), matched by instruction offset inside the kernel's .text section.

usage: ncu_lines.py REPORT.ncu-rep KERNEL UNIT.cubin [top_n]
       ncu_lines.py --summary REPORT.ncu-rep KERNEL UNIT.cubin N_WORDS   (what profiles/ holds: key
           metrics, executed LOP3, DRAM bytes, instructions per function and per word, top lines)
The cubin is the unit file NVRTC produced for the same plan (tools/offline_units.py CASE DIR
compiles them here, bit-identical to the GPU box's: same NVRTC, same source)."""
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


def sass_rows(rep, kernel):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name', kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    return rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]


def summary(rep, kernel, cubin, n_words):
    """Text for profiles/: everything a reader needs to check the numbers quoted in DESIGN.md."""
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, r = rows[0], rows[1], rows[2]
    get = lambda k: r[h.index(k)] if k in h else 'n/a'     # noqa: E731
    scale = {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9,
             'hz': 1.0, 'Khz': 1e3, 'Mhz': 1e6, 'Ghz': 1e9}

    def num(k):             # in ns / bytes / Hz whatever unit ncu chose for the column
        if get(k) in ('n/a', ''):
            return float('nan')
        return float(get(k).replace(',', '')) * scale.get(units[h.index(k)], 1.0)
    print('%s  (%s, ncu --set full --clock-control none, one launch)' % (kernel, rep.split('/')[-1]))
    print('duration %.3f ms   grid %s x block %s   registers/thread %s   dynamic smem %s B   local (stack) %s B' % (
        num('gpu__time_duration.sum') / 1e6, get('launch__grid_size'), get('launch__block_size'),
        get('launch__registers_per_thread'), get('launch__shared_mem_per_block_dynamic'), get('launch__stack_size')))
    print('SM clock %.0f MHz   achieved occupancy %.1f %%   IPC (active) %.2f   issue slots busy %.1f %%' % (
        num('gpc__cycles_elapsed.avg.per_second') / 1e6, num('sm__warps_active.avg.pct_of_peak_sustained_active'),
        num('sm__inst_executed.avg.per_cycle_active'), num('sm__inst_issued.avg.pct_of_peak_sustained_active')))
    print('ALU pipe %.1f %% of peak active   LSU pipe %.1f %%' % (
        num('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'), num('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active')))
    dram = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
    print('DRAM %.1f MB read + %.1f MB written = %.1f MB per launch (%.1f %% of peak throughput)   L2 hit rate %.1f %%' % (
        num('dram__bytes_read.sum') / 1e6, num('dram__bytes_write.sum') / 1e6, dram / 1e6,
        num('dram__bytes_read.sum.pct_of_peak_sustained_elapsed') + num('dram__bytes_write.sum.pct_of_peak_sustained_elapsed'), num('lts__t_sector_hit_rate.pct')))
    stalls = {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]: float(v)
              for k, v in zip(h, r) if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and v}
    print('stall cycles per issued instruction: ' + ', '.join('%s %.2f' % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
    hh, body = sass_rows(rep, kernel)
    ie, so = hh.index('Instructions Executed'), hh.index('Source')
    tie = hh.index('Thread Instructions Executed')
    tot = sum(float(x[ie]) for x in body)
    op = defaultdict(float)
    for x in body:
        w = x[so].split()
        if not w:
            continue
        o = w[1] if w[0].startswith('@') else w[0]
        op[o.split('.')[0]] += float(x[tie])
    ttot = sum(op.values())
    print('executed: %.4g warp instructions, %.4g thread instructions (%.0f per 32-lane word, %d words)' % (tot, ttot, ttot / n_words, n_words))
    print('by opcode (thread level): ' + ', '.join('%s %.4g (%.1f %%)' % (k, v, 100 * v / ttot) for k, v in sorted(op.items(), key=lambda kv: -kv[1])[:8]))
    lt = line_table(cubin, kernel)
    base = int(body[0][hh.index('Address')], 16)
    src = open('boolsi_b200/csrc/bsi_template.cuh').read().split('\n')
    funcs = [(i + 1, m.group(1)) for i, line in enumerate(src)
             for m in [re.match(r'(?:extern "C" )?__(?:device|global)__.*?\b(bsi_\w+)\(', line)] if m]

    def fn(line):
        name = '?'
        for st, n in funcs:
            if st <= line:
                name = n
        return name
    ns = hh.index('# Samples')
    reg = defaultdict(lambda: [0, 0.0, 0.0])
    for x in body:
        line, _ = lt.get(int(x[0], 16) - base, (0, '?'))
        a = reg[fn(line)]
        a[0] += 1
        a[1] += float(x[ie])
        a[2] += float(x[ns])
    tots = sum(a[2] for a in reg.values())
    print('by function of the template (SASS instructions, executed share, per word, stall-sample share):')
    for n, a in sorted(reg.items(), key=lambda kv: -kv[1][1])[:14]:
        print('  %-24s %5d  %5.1f %%  %7.0f  %5.1f %%' % (n, a[0], 100 * a[1] / tot, a[1] * 32 / n_words, 100 * a[2] / max(tots, 1)))
    print()


def main():
    if sys.argv[1] == '--extract':
        return extract(sys.argv[2], sys.argv[3])
    if sys.argv[1] == '--summary':
        summary(sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]))
        sys.argv = [sys.argv[0]] + sys.argv[2:5] + ['25']
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
