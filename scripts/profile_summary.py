#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/ (run in the build container).

    python scripts/profile_summary.py gpurun_out/prof.ncu-rep KERNEL_REGEX > profiles/rNN_kernel.txt
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
print(f"# ncu --set full --clock-control none --import-source on : {rep}")
n = 0
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    import re
    if not re.search(kern, name):
        continue
    n += 1
    print(f"\n## launch {n}")
    for k in KEYS:
        if k in hdr:
            print(f"{k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
print()
for tool in ("scripts/ncu_ranges.py", "scripts/ncu_lines.py"):
    out = subprocess.run([sys.executable, tool, rep, kern] + (["25"] if tool.endswith("lines.py") else []),
                         capture_output=True, text=True).stdout
    print(f"## {tool}\n{out}")
