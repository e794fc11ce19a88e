"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output: python tools/ncu_pick.py file.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak",
                        "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "smsp__inst_executed.sum",
                        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_alu.", "sm__inst_executed_pipe_lsu.", "sm__inst_executed_pipe_fma.",
                        "pipe_alu_cycles_active.avg.pct", "pipe_fma_cycles_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts.sum ",
                        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "sm__throughput.avg.pct", "l1tex__lsu_writeback_active",
                        "lsu_mem_shared_op", "smsp__warps_eligible.avg.per_cycle", "launch__occupancy_per", "sm__maximum_warps"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:60])
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in want):
            print(f"  {h} [{u}] = {v}")
