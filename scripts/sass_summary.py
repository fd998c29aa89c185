"""cuobjdump -sass of the built library -> per-kernel counts of the TMA / mbarrier / warp-level instructions (profiles/*_sass_summary.txt).
Runs without a GPU:  python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "blazeseq_b200", "lib", "libblazeseq_gpu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UBLKCP.S.G", "UBLKCP.G.S", "UBLKPF", "SYNCS", "REDUX", "MATCH", "VOTE", "SHFL", "LDS.128", "LDS", "STS", "LDG", "STG", "BAR", "WARPSYNC"]
print("# SASS of the shipped blazeseq_b200/lib/libblazeseq_gpu.so (cuobjdump -sass, sm_100a): the TMA / mbarrier / warp-level")
print("# instructions per kernel.  UBLKCP.S.G = cp.async.bulk global->shared (tile loads), UBLKCP.G.S = shared->global (SoA stores),")
print("# UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops, REDUX / MATCH / VOTE / SHFL = warp collectives.\n")
name, counts, n = None, collections.Counter(), 0
totals = collections.Counter()


def flush():
    if name and name.startswith("_ZN3bsq"):
        print(name)
        print("    instructions %d  " % n + "  ".join("%s %d" % (k, counts[k]) for k in keys if counts[k]))
        totals.update(counts)


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        n += 1
        for k in keys:
            if op == k or op.startswith(k + ".") or (k in ("UBLKCP.S.G", "UBLKCP.G.S", "LDS.128") and op.startswith(k)):
                counts[k] += 1
flush()
print("\ntotals over the bsq:: kernels: " + "  ".join("%s %d" % (k, totals[k]) for k in keys if totals[k]))
