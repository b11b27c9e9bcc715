"""One line block per profiled launch from `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM bytes, pipe / issue utilisation.
    python scripts/ncu_raw_summary.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = hdr.index("Kernel Name")
    for k, r in enumerate(data):
        print(f"--- {r[name][:110]}  (id {k})")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:<72s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
