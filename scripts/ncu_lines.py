#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-regex] [top]
Prints, per CUDA source line: warp instructions executed, share, avg active threads, stall samples.
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
# the page is a sequence of kernels; take the first launch of each
blocks = re.split(r'(?m)^"File Path",', out)
seen = set()
for blk in blocks[1:]:
    lines = blk.splitlines()
    fn = [l for l in lines[:3] if l.startswith('"Function Name"')]
    name = fn[0] if fn else "?"
    if name in seen:
        continue
    seen.add(name)
    if len(sys.argv) > 2 and not re.search(sys.argv[2], name):
        continue
    hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
    rd = csv.reader(io.StringIO("\n".join(lines[hdr_i:])))
    hdr = next(rd)
    ci = {h: i for i, h in enumerate(hdr)}
    rows = []
    for r in rd:
        if len(r) < len(hdr) or not r[0]:
            continue  # SASS rows have an empty line number
        try:
            rows.append((int(r[0]), r[1], int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]]),
                         int(r[ci["# Samples"]])))
        except ValueError:
            pass
    tot = sum(x[2] for x in rows) or 1
    tots = sum(x[4] for x in rows) or 1
    print(name, "total warp inst", tot, "samples", tots)
    for ln, src, inst, tinst, smp in sorted(rows, key=lambda x: -(x[4] if len(sys.argv) > 4 else x[2]))[:top]:
        print(f"{ln:5d} {100 * inst / tot:5.1f}% inst  {100 * smp / tots:5.1f}% smp  thr/inst {tinst / max(inst, 1):5.1f}  {src.strip()[:110]}")
    # stall reason totals for the kernel (sampled, all samples)
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_st = {n: 0 for n in names}
    rd2 = csv.reader(io.StringIO("\n".join(lines[hdr_i + 1:])))
    for r in rd2:
        if len(r) < len(hdr) or not r[0]:
            continue
        for n in names:
            try:
                tot_st[n] += int(r[ci[n]])
            except ValueError:
                pass
    s = sum(tot_st.values()) or 1
    print("stalls:", ", ".join(f"{n[6:]} {100 * v / s:.1f}%" for n, v in sorted(tot_st.items(), key=lambda x: -x[1])[:8]))
