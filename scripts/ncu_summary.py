import csv, sys
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
SHORT = ["grid", "blk", "ms", "rd", "wr", "sm%", "dram%", "warps%", "issue%", "regs", "occ_smem", "occ_reg", "st_long", "st_short", "st_bar", "st_mio", "st_wait", "st_math", "st_notsel", "bankconf", "tensor%", "fma%", "alu%", "lsu%"]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f, "units:", units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]])
    for r in data:
        out = [r[col["Kernel Name"]][:30]]
        for w, s in zip(WANT[1:], SHORT):
            v = r[col[w]] if w in col else "NA"
            try: v = "%.3g" % float(v.replace(",", ""))
            except Exception: pass
            out.append("%s=%s" % (s, v))
        print(" ".join(out))
