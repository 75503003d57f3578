#!/usr/bin/env python
"""Per-CUDA-source-line breakdown of an `ncu --page source --csv --print-source cuda,sass` export:
stall samples, executed warp instructions and shared-memory wavefronts per line and per file.
Usage: python profiles/ncu_lines.py gpurun_out/push_src_cuda.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == 'Line No')
iSm = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
iW = hdr.index('L1 Wavefronts Shared'); iWi = hdr.index('L1 Wavefronts Shared Ideal')
f = None; out = []
tot_s = tot_i = tot_w = 0
for r in rows:
    if r and r[0] == 'File Path':
        f = r[1].split('/')[-1]; continue
    if len(r) < iW or r[0] in ('Line No', 'Function Name', ''):
        continue
    try:
        s = int(r[iSm]); i = int(r[iE]); w = int(r[iW] or 0); wi = int(r[iWi] or 0)
    except ValueError:
        continue
    out.append((s, i, w, wi, f, r[0], r[1].strip()[:100])); tot_s += s; tot_i += i; tot_w += w
print('total samples', tot_s, 'inst', tot_i, 'shared wavefronts', tot_w)
byf = {}
for s, i, w, wi, f, l, src in out:
    a = byf.setdefault(f, [0, 0, 0, 0]); a[0] += s; a[1] += i; a[2] += w; a[3] += wi
for f, a in byf.items():
    print(f, 'samples %.1f%% inst %.1f%% wavefronts %.1f%% (ideal %.1f%%)' % (100 * a[0] / tot_s, 100 * a[1] / tot_i, 100 * a[2] / max(tot_w, 1), 100 * a[3] / max(tot_w, 1)))
out.sort(reverse=True)
for s, i, w, wi, f, l, src in out[:top]:
    print('%5.1f%% inst %5.1f%% wf %5.1f%% (ideal %4.1f%%) %s:%s  %s' % (100 * s / tot_s, 100 * i / tot_i, 100 * w / max(tot_w, 1), 100 * wi / max(tot_w, 1), f, l, src))
