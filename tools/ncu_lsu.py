"""LSU data-pipe view of an .ncu-rep (ncu --set full): what bounds these kernels is shared-memory wavefronts, so this
lists, per kernel, the data pipe's utilisation and its wavefronts by kind next to issue / ALU utilisation.
usage: python tools/ncu_lsu.py x.ncu-rep > profiles/rNN_ncu_lsu_x.txt"""
import csv, io, subprocess, sys
WANT = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print(f"# {sys.argv[1]}: LSU data pipe (shared-memory wavefronts) vs issue / ALU, per launch")
for r in rows[2:]:
    print(f"\n## {r[ki].split('(')[0]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:84s} {r[i]:>18s} {units[i]}")
