"""Condense an `ncu --page raw --csv` dump into the handful of metrics the roofline discussion uses (one column per
kernel launch).  usage: python scripts/ncu_summary.py <raw.csv> > profiles/<name>_summary.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max"]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
for name in want:
    if name in hdr:
        i = hdr.index(name)
        w.writerow([name, units[i]] + [r[i] for r in data])
