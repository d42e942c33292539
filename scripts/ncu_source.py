#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report (needs -lineinfo and --import-source on).
usage: ncu_source.py <report.ncu-rep> [top N] [--sass LINE]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sass_line = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == '--sass' else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = []      # (file, lineno, src, row)
cur_file = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if r[0] == 'Function Name' or hdr is None:
        continue
    if r[0] != '':
        lines.append([cur_file, r[0], r[1], r, []])
    elif lines:
        lines[-1][4].append(r)
# index from the end: unescaped quotes/commas in source text shift the front columns
cs = hdr.index('# Samples') - len(hdr); ci = hdr.index('Instructions Executed') - len(hdr)
stall_cols = [(k - len(hdr), h) for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def f(v):
    try:
        return float(v)
    except ValueError:
        return 0.0
tot_s = sum(f(l[3][cs]) for l in lines); tot_i = sum(f(l[3][ci]) for l in lines)
print("total samples %d  warp-instructions %d" % (tot_s, tot_i))
for l in sorted(lines, key=lambda l: -f(l[3][cs]))[:top]:
    r = l[3]; s = f(r[cs])
    if s == 0: break
    st = sorted(((f(r[k]), h) for k, h in stall_cols), reverse=True)[:3]
    print("%-12s:%4s %5.1f%% smp %5.1f%% ins | %-64s | %s" % (
        l[0].split('/')[-1], l[1], 100 * s / tot_s, 100 * f(r[ci]) / tot_i, l[2].strip()[:64],
        ' '.join('%s=%d' % (h[6:], v) for v, h in st if v > 0)))
if sass_line:
    for l in lines:
        if l[1] == sass_line:
            for r in l[4]:
                print("   %-60s smp %5s ins %8s" % (r[3].strip()[:60], r[cs], r[ci]))
