"""Print the headline metrics and the warp-stall breakdown of an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]
for r in rows[2:]:
    for w in want:
        if w in h:
            print(f"{w:75s} {r[h.index(w)]} {units[h.index(w)]}")
    print("--- stall reasons (pcsamp, warp-level) ---")
    st = [(float(r[i].replace(",", "")), h[i]) for i in range(len(h)) if h[i].startswith("smsp__pcsamp_warps_issue_stalled") and r[i] not in ("", "n/a")]
    tot = sum(v for v, _ in st) or 1
    for v, n in sorted(st, reverse=True)[:12]:
        print(f"  {100 * v / tot:5.1f}%  {n.replace('smsp__pcsamp_warps_issue_stalled_', '')}")
    print("--- pipes ---")
    for i, n in enumerate(h):
        if n.startswith("sm__inst_executed_pipe_") and n.endswith("pct_of_peak_sustained_active") and r[i] not in ("", "n/a") and float(r[i]) > 3:
            print(f"  {n}: {r[i]}")
