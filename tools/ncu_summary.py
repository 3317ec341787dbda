"""Summarise an .ncu-rep (ncu --set full) into profiles/<name>.txt (+ profiles/ncu_traffic.json).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_tile_kernel.txt [--traffic] "command line" """
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
traffic = "--traffic" in sys.argv
cmd = [a for a in sys.argv[3:] if a != "--traffic"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
lines = [f"# source: {rep}", f"# command: {' '.join(cmd)}",
         "# ncu --set full --clock-control none --import-source on (numbers under the profiler are NOT bench values)"]
tr = None
for k, row in enumerate(data):
    lines.append(f"--- launch {k}")
    vals = {}
    for i, h in enumerate(hdr):
        if h in want:
            lines.append(f"{h:86s} {row[i]:>22s} {units[i]}")
            vals[h] = (row[i], units[i])
    def to_bytes(key):
        v, u = vals[key]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tr = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    lines.append(f"{'dram bytes read+write per launch':86s} {tr:22.0f} byte")
open(out, "w").write("\n".join(lines) + "\n")
if traffic and tr is not None:
    json.dump({"dram_bytes_per_launch": tr, "source": out, "command": " ".join(cmd)},
              open("profiles/ncu_traffic.json", "w"), indent=1)
print("\n".join(lines[-34:]))
