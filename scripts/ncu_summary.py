#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu --set full --import-source on): key counters, stall reasons, hottest lines.

    python scripts/ncu_summary.py gpurun_out/X.ncu-rep "header line" > profiles/rNN_X_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, d = rows[0], dict(zip(rows[0], rows[1])), dict(zip(rows[0], rows[2]))
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
print("# " + header)
for w in want:
    if w in d:
        print(w, "=", d[w], units.get(w, ""))
print("# warp stall reasons (per issue-active cycle)")
st = {k: float(v) for k, v in d.items() if "smsp__average_warps_issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k and v}
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print(k.split("stalled_")[1].split("_per")[0], round(v, 3))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", "0", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
open("/tmp/_ncu_src.csv", "w").write(src)
print("# top source lines by stall samples (scripts/ncu_lines.py)")
print(subprocess.run([sys.executable, "scripts/ncu_lines.py", "/tmp/_ncu_src.csv", "14"], capture_output=True, text=True).stdout)
