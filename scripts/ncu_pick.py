"""Reads `ncu --page raw --csv` on stdin and prints, per launch, the counters the profiles/ notes quote: pipe
utilisation (FP64, LSU, ALU), shared-memory wavefronts and bank conflicts, issue-stall reasons, DRAM/L2 traffic."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = [h for h in hdr if any(k in h for k in (
    "gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "per_issue_active.ratio"))]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print("== %s grid %s block %s" % (r[ix["Kernel Name"]].split("(")[0], r[ix["Grid Size"]], r[ix["Block Size"]]))
    for h in want:
        try:
            v = float(r[ix[h]].replace(",", ""))
        except ValueError:
            continue
        if v > 0.005:
            print("   %-86s %16.2f %s" % (h, v, units[ix[h]]))
