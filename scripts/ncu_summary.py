"""Summarise an .ncu-rep (read here, no GPU needed) into a small text table for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof_full.ncu-rep > profiles/rNN_xxx_ncu_summary.txt"""
import csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
           "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# source: {rep} (ncu --set full --clock-control none); one block per captured launch")
for r in rows[2:]:
    print("\n== " + r[hdr.index("Kernel Name")][:100])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"  {m:68s} {r[i]:>16s} {units[i]}")
