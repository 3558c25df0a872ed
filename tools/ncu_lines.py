#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel in an .ncu-rep.

ncu's CSV export of the source page carries SASS only; this merges it (by instruction order) with
`nvdisasm -g` line info of the same kernel in the object file, and prints the hottest source lines.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep rlcard_b200/csrc/tu_leduc.o 'k_rollout.*Leduc.*Philox.*Eh' [top]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    sections, cur, hdr, name = [], None, None, None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            name = r[1]
        elif r and r[0] == 'Address':
            hdr = r
            cur = []
            sections.append((name, hdr, cur))
        elif cur is not None and hdr and len(r) == len(hdr):
            cur.append(r)
    return sections


def disasm_lines(obj, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = glob.glob(os.path.join(tmp, '*.cubin'))[0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    cur_fn, line, out = None, None, []
    want = re.compile(pattern)
    for l in txt.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            cur_fn = m.group(1) if want.search(m.group(1)) else None
            continue
        if cur_fn is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
            line = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            out.append((line, m.group(2).strip()))
    return out


def main():
    rep, obj, pattern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    secs = sass_rows(rep)
    name, hdr, rows = secs[-1]
    ie, ns = hdr.index('Instructions Executed'), hdr.index('# Samples')
    dis = disasm_lines(obj, pattern)
    print('kernel:', name)
    print('sass rows %d, disasm instrs %d' % (len(rows), len(dis)))
    n = min(len(rows), len(dis))
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    by_outer = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for k in range(n):
        line = dis[k][0]
        i, s = int(rows[k][ie]), int(rows[k][ns])
        tot_i += i
        tot_s += s
        key = (line[0], line[1]) if line else ('?', 0)
        by_line[key][0] += i; by_line[key][1] += s; by_line[key][2] += 1
        outer = line[2][-1] if line and line[2] else key
        by_outer[outer][0] += i; by_outer[outer][1] += s; by_outer[outer][2] += 1
    print('total warp-instructions %d, samples %d' % (tot_i, tot_s))
    for title, d in (('innermost source line', by_line), ('outermost (kernel-level) line', by_outer)):
        print('\n== by %s: file:line  instr%%  samples%%  static-instrs' % title)
        for key, (i, s, c) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print('%-22s %6.2f%% %6.2f%% %5d' % ('%s:%d' % key, 100.0 * i / max(tot_i, 1), 100.0 * s / max(tot_s, 1), c))


if __name__ == '__main__':
    main()
