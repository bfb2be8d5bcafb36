"""Summarise `ncu -i X.ncu-rep --page raw --csv` (stdin or file) to the handful of metrics DESIGN.md cites:
one block per profiled launch."""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active", "gpu__dram_throughput", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__cluster", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem", "lts__t_sector_hit_rate.pct", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__mem_tensor_cycles_active", "sm__pipe_tensor_cycles_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "sm__pipe_fma_cycles_active", "sm__inst_executed_pipe_xu")


def main(path):
    rows = list(csv.reader(open(path) if path != "-" else sys.stdin))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    for r in rows[hi + 2:]:
        if len(r) != len(hdr):
            continue
        print("%-90s %s" % ("Kernel Name", r[hdr.index("Kernel Name")]))
        for name, unit, val in zip(hdr, units, r):
            if name.startswith(KEEP):
                print("%-90s %s%s" % (name, val, (" " + unit) if unit else ""))
        print()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "-")
