#!/usr/bin/env python
"""Warp-instruction share per source function (by line range) from an ncu report."""
import csv, io, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
src = open('blazeseq_b200/csrc/bsq_device.cuh').read().splitlines()
# function start lines
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:template.*\n)?(?:__device__|__global__).*?(\w+)\(', l)
    if l.startswith('__device__') or l.startswith('__global__'):
        name = re.search(r'(\w+)\s*\(', l.split('__forceinline__')[-1])
        marks.append((i, name.group(1) if name else l[:40]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blk = [b for b in re.split(r'(?m)^"File Path",', out)[1:] if re.search(kern, b[:400])][0]
lines = blk.splitlines()
hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
rd = csv.reader(io.StringIO("\n".join(lines[hdr_i:])))
hdr = next(rd); ci = {h: i for i, h in enumerate(hdr)}
agg = {}
tot = 0
for r in rd:
    if len(r) < len(hdr) or not r[0]: continue
    try: ln = int(r[0]); inst = int(r[ci["Instructions Executed"]]); ti = int(r[ci["Thread Instructions Executed"]]); smp = int(r[ci["# Samples"]])
    except ValueError: continue
    fn = "?"
    for s, n in marks:
        if s <= ln: fn = n
    a = agg.setdefault(fn, [0, 0, 0]); a[0] += inst; a[1] += ti; a[2] += smp; tot += inst
tots = sum(a[2] for a in agg.values()) or 1
for fn, (inst, ti, smp) in sorted(agg.items(), key=lambda x: -x[1][2]):
    print(f"{fn:28s} inst {100*inst/tot:5.1f}%  time(samples) {100*smp/tots:5.1f}%  thr/inst {ti/max(inst,1):5.1f}")
print("total warp inst", tot)
