"""Summarise an ncu report's raw page into a small per-kernel CSV (the metrics the DESIGN notes quote).
usage: python tools/ncu_raw_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
ki = hdr.index("Kernel Name")
names = [r[ki].split('(')[0].replace('void ba::', '').replace('void ', '') for r in data]
out = [["metric", "unit"] + names]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        out.append([w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", ""), units[i]] + [r[i] for r in data])
txt = "\n".join(",".join(o) for o in out) + "\n"
open(sys.argv[2], "w").write(txt)
print(txt)
