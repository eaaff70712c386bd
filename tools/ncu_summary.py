"""Summarise an .ncu-rep into a small text file for profiles/ (run here, no GPU needed):
python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.summary.txt"""
import csv, io, subprocess, sys, json

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
]
lines = [f"ncu summary of {rep} (kernel b2_ensemble_kernel; --set full --clock-control none)"]
vals = {}
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        lines.append(f"{w:80s} {units[i]:14s} " + "  ".join(r[i] for r in data))
        vals[w] = data[0][i]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
reg = subprocess.run([sys.executable, "tools/sass_regions.py", "/tmp/_src.csv", "0.015"], capture_output=True, text=True).stdout
lines += ["", "SASS regions (consecutive instructions with equal execution count; share of all issued warp-instructions):", reg]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
