"""Stall reasons (per issued instruction) and a few pipe/cache counters of every launch in an .ncu-rep.
Usage: python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
EXTRA = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
         "l1tex__data_bank_conflicts_pipe_lsu.sum", "sm__cycles_active.avg"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for m in EXTRA:
        if m in hdr:
            print(f"   {m:70s} {r[hdr.index(m)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:9]:
        print(f"   stall {k:30s} {v:.3f}")
