"""Key metrics of every kernel in an ncu report as metric,unit,value rows per kernel.
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_extract.py > profiles/xxx.csv"""
import csv
import sys

KEEP = ("gpu__time_duration.sum", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput", "dram__throughput", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster", "sm__warps_active.avg", "smsp__issue_active.avg", "sm__throughput.avg", "lts__throughput.avg",
        "lts__t_sector_hit_rate", "sm__cycles_elapsed.avg.per_second", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem", "sm__inst_executed.sum.per_cycle_elapsed")
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
w.writerow(["kernel#", "metric", "unit", "value"])
for n, r in enumerate(rows[2:]):
    w.writerow([n, "Kernel Name", "", r[hdr.index("Kernel Name")]])
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEEP) and r[i] not in ("", "n/a"):
            w.writerow([n, h, units[i], r[i]])
