#!/usr/bin/env python
"""profiles/traffic.json from the two `ncu --set full` captures of the bench (run in the build container).

    python scripts/traffic_from_ncu.py gpurun_out/r01_prof_resolve.ncu-rep gpurun_out/r01_prof_summarize.ncu-rep \
        RECORDS_IN_CAPTURED_LAUNCH > profiles/traffic.json

DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the captured launch of each kernel, and per record.
"""
import csv
import io
import json
import subprocess
import sys


def dram(rep, kern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if kern in r[hdr.index("Kernel Name")]:
            def val(k):
                i = hdr.index(k)
                v = float(r[i].replace(",", ""))
                u = units[i].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}[u]
            return val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
    raise SystemExit(f"{kern} not in {rep}")


res_rep, sum_rep, records = sys.argv[1], sys.argv[2], float(sys.argv[3])
rr, rw, _ = dram(res_rep, "k_resolve")
sr, sw, _ = dram(sum_rep, "k_summarize")
print(json.dumps({
    "source": "ncu --set full --clock-control none, one launch of each kernel inside bench.py (10 GiB config, full 2 GiB window)",
    "records_in_captured_launch": records,
    "k_resolve_dram_bytes_read": rr, "k_resolve_dram_bytes_write": rw,
    "k_resolve_bytes_per_record": (rr + rw) / records,
    "k_summarize_dram_bytes_read": sr, "k_summarize_dram_bytes_write": sw,
    "k_summarize_bytes_per_record": (sr + sw) / records,
}, indent=1))
