"""Turn one `ncu --set full --import-source on` report of ONE kernel launch into
  * profiles/<out>.txt: selected raw metrics, dynamic SASS opcode histogram (instructions executed, warp-level),
    per-file and per-source-line shares of instructions and stall samples;
  * an entry of profiles/ncu_counts.json: per-chain-draw counts that bench.py reads for its roofline
    (fp64 flops = DADD + DMUL + 2 DFMA thread instructions + 512 per warp-level DMMA, all COUNTED, predicated-on).

Usage: python tools/ncu_counts.py REPORT.ncu-rep OUT.txt TAG CHAIN_DRAWS_PER_LAUNCH "command line"
(numbers under the profiler are never bench values; only counts and shares are used)"""
import collections, csv, io, json, os, re, subprocess, sys

rep, out, tag, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
cmd = " ".join(sys.argv[5:])


def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, unit_row, data = rows[0], rows[1], rows[2:]
assert len(data) == 1, f"expected one profiled launch, got {len(data)}"
raw = {h: (data[0][i], unit_row[i]) for i, h in enumerate(hdr)}
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
L = [f"# source: {os.path.basename(rep)}  (gpurun_out/, not tracked)", f"# command: {cmd}",
     "# ncu --set full --clock-control none --import-source on; ONE launch (the first timed one of the config)",
     f"# chain-draws in this launch: {units:.0f}"]
for k in want:
    if k in raw:
        L.append(f"{k:90s} {raw[k][0]:>22s} {raw[k][1]}")


def to_bytes(key):
    v, u = raw[key]
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
L.append(f"{'dram bytes read+write per launch':90s} {dram:22.0f} byte")

# ---- SASS + CUDA source page
rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--print-source", "sass,cuda", "--csv"))))
cur_file, col = None, None
ops = collections.Counter()          # warp-level instructions executed by opcode
ops_thr = collections.Counter()      # predicated-on thread instructions by opcode
ops_smp = collections.Counter()
files = collections.defaultdict(lambda: [0.0, 0.0])
lines = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if len(r) > 8 and r[0] == "Line No":
        col = {h: i for i, h in enumerate(r)}
        # two "Source" columns: the first is the CUDA line, the second the SASS text
        src_cols = [i for i, h in enumerate(r) if h == "Source"]
        continue
    if col is None or len(r) < len(col) or cur_file is None:
        continue
    inst, thr, smp = num(r[col["Instructions Executed"]]), num(r[col["Predicated-On Thread Instructions Executed"]]), num(r[col["# Samples"]])
    if r[col["Address"]] == "-":                     # aggregate of one CUDA source line
        files[cur_file][0] += inst
        files[cur_file][1] += smp
        key = (cur_file, r[0])
        if key in lines:
            lines[key][0] += inst
            lines[key][1] += smp
        else:
            lines[key] = [inst, smp, r[src_cols[0]].strip()]
    else:                                            # one SASS instruction
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", r[src_cols[1]])
        if m:
            full = m.group(2)
            base = full.split(".")[0]
            name = base
            if base in ("MUFU", "F2F", "I2F", "F2I", "IMAD", "DMMA", "LDS", "STS", "LDG", "STG", "SHFL", "LDGSTS", "UBLKCP", "SYNCS"):
                name = ".".join(full.split(".")[:3 if base in ("F2F", "I2F") else 2])
            ops[name] += inst
            ops_thr[name] += thr
            ops_smp[name] += smp
tot_i, tot_s = sum(ops.values()) or 1.0, sum(ops_smp.values()) or 1.0
L.append("")
L.append(f"# dynamic SASS opcode histogram (warp-level instructions executed; {tot_i:.0f} total = {tot_i / units:.1f} per chain-draw)")
for k, v in ops.most_common(28):
    L.append(f"  {k:22s} {v:16.0f}  {100 * v / tot_i:5.1f}%   {v / units:8.2f} /chain-draw   stall samples {100 * ops_smp[k] / tot_s:5.1f}%")
flops = sum(ops_thr[k] * w for k, w in (("DADD", 1), ("DMUL", 1), ("DFMA", 2)))
n_dmma = sum(v for k, v in ops.items() if k.startswith("DMMA"))
flops += 512.0 * n_dmma
L.append(f"# fp64 flops executed: DADD {ops_thr['DADD']:.0f} + DMUL {ops_thr['DMUL']:.0f} + 2 x DFMA {ops_thr['DFMA']:.0f} thread instructions"
         f" + 512 x DMMA {n_dmma:.0f} warp instructions = {flops:.4e} = {flops / units:.1f} per chain-draw")
feat = {k: v for k, v in ops.items() if any(k.startswith(p) for p in ("UTMA", "UBLKCP", "UTC", "LDTM", "STTM", "DMMA", "SYNCS", "LDGSTS", "UCGABAR", "CCTL"))}
L.append("# Blackwell / async-copy / tensor opcodes executed: " + (", ".join(f"{k} {v:.0f}" for k, v in sorted(feat.items())) or "none"))
L.append("")
L.append("# per source file: share of instructions executed / of stall samples")
ti, ts = sum(v[0] for v in files.values()) or 1.0, sum(v[1] for v in files.values()) or 1.0
for f, (i_, s_) in sorted(files.items(), key=lambda kv: -kv[1][0])[:10]:
    L.append(f"  {f:28s} inst {100 * i_ / ti:5.1f}%  samples {100 * s_ / ts:5.1f}%")
L.append("# top source lines by stall samples")
for (f, ln), (i_, s_, txt) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:22]:
    L.append(f"  {f:22s}:{ln:>4s} inst {100 * i_ / ti:5.2f}% smp {100 * s_ / ts:5.2f}%  {txt[:110]}")
open(out, "w").write("\n".join(L) + "\n")

cj = "profiles/ncu_counts.json"
allc = json.load(open(cj)) if os.path.exists(cj) else {}
allc[tag] = {
    "kernel": raw["Kernel Name"][0], "launch": cmd, "chain_draws_per_launch": units,
    "warp_inst_per_draw": num(raw["smsp__inst_executed.sum"][0]) / units,
    "fp64_flops_per_draw": flops / units,
    "issue_active_pct": num(raw["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
    "warps_active_pct": num(raw["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
    "fp64_pipe_pct": num(raw["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
    "xu_pipe_pct": num(raw["sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active"][0]),
    "dram_bytes_per_launch": dram, "duration_under_ncu_ms": num(raw["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(raw["gpu__time_duration.sum"][1], 1.0),
    "registers": num(raw["launch__registers_per_thread"][0]), "source": out,
}
json.dump(allc, open(cj, "w"), indent=1)
print("\n".join(L))
