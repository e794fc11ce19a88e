"""Per-opcode instruction / shared-wavefront / global-request breakdown of one kernel in an .ncu-rep, per block.
usage: python tools/ncu_ops.py rep.ncu-rep n_blocks"""
import csv, io, subprocess, sys
from collections import defaultdict
path, nb = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(r, suffix):
    c = [h for h in hdr if h.endswith(suffix)]
    return r[hdr.index(c[0])] if c else "?"
for r in rows[2:]:
    print(col(r, "Kernel Name")[:60])
    for w in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg",
              "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]:
        print(f"  {w:88s} {col(r, w)}")
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:60], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h, data = b["rows"][0], b["rows"][1:]
    iS, iW, iE, iG = h.index("Source"), h.index("L1 Wavefronts Shared"), h.index("Instructions Executed"), h.index("L1 Tag Requests Global")
    d, e, g = defaultdict(float), defaultdict(float), defaultdict(float)
    for r in data:
        t = r[iS].strip().split()
        if not t: continue
        op = t[1] if t[0].startswith("@") else t[0]
        d[op] += float(r[iW] or 0) / nb; e[op.split(".")[0]] += int(r[iE]) / nb; g[op] += float(r[iG] or 0) / nb
    print(b["name"])
    print("  shared wavefronts/block:", {k: round(v, 1) for k, v in d.items() if v > 0.5}, "sum", round(sum(d.values())))
    print("  global tag requests/block:", {k: round(v, 1) for k, v in g.items() if v > 0.5}, "sum", round(sum(g.values())))
    print("  instructions/block:", sorted(((round(v, 1), k) for k, v in e.items()), reverse=True)[:18], "total", round(sum(e.values())))
