"""Print a compact summary of an .ncu-rep (first profiled launch): the counters the roofline quotes."""
import csv, subprocess, sys, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__cluster_size", "launch__waves_per_multiprocessor", "launch__occupancy_cluster_pct",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__ctas_launched.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.avg.per_cycle_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "sass__inst_executed_shared_stores", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out = {}
for h, u, v in zip(hdr, units, val):
    if (h in want or "pipe_tensor" in h or h.startswith("smsp__average_warps_issue_stalled")) and "pred_on" not in h:
        out[h] = f"{v} {u}".strip()
for k in sorted(out):
    print(f"{k:100s} {out[k]}")
