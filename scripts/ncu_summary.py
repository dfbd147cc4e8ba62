"""Turns `ncu -i X.ncu-rep --page raw --csv` (stdin or a file) into the compact `metric,unit,launch0,launch1,...` table kept under
profiles/ (one column per captured launch; the header row names the kernel of each column).
usage: ncu -i rep.ncu-rep --page raw --csv | python scripts/ncu_summary.py [kernel-name-substring] > profiles/rNN_prof_*_summary.csv"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
]


def main():
    sub = sys.argv[1] if len(sys.argv) > 1 else ""
    rows = [r for r in csv.reader(sys.stdin) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    data = [r for r in rows[hdr + 2:] if len(r) == len(names)]
    kcol = names.index("Kernel Name")
    data = [r for r in data if sub in r[kcol]]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(data))])
    w.writerow(["kernel", ""] + [r[kcol].split("(")[0] for r in data])
    for m in KEEP:
        if m in names:
            c = names.index(m)
            w.writerow([m, units[c]] + [r[c] for r in data])


if __name__ == "__main__":
    main()
