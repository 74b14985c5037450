"""Instruction / stall-sample share per region of lm_umma_kernel (regions = line ranges given as name:start pairs).
Usage: python scripts/ncu_regions.py rep kernel-regex file.cu name:line name:line ..."""
import csv, io, subprocess, sys
rep, kre, fn = sys.argv[1:4]
marks = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[4:]] + [("end", 10**9)]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; cur = None; fname = ""; agg = {}
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] != "": cur = (fname, int(r[0])); continue
    if cur is None: continue
    a = agg.setdefault(cur, [0, 0]); a[0] += num(r[hdr.index("Instructions Executed")]); a[1] += num(r[hdr.index("# Samples")])
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
b = {}
for (f, ln), a in agg.items():
    key = f
    if f == fn:
        key = "before " + marks[0][0]
        for i in range(len(marks) - 1):
            if marks[i][1] <= ln < marks[i + 1][1]: key = marks[i][0]
    v = b.setdefault(key, [0, 0]); v[0] += a[0]; v[1] += a[1]
print("total warp inst", tot, "samples", ts)
for k, v in sorted(b.items(), key=lambda kv: -kv[1][0]): print(f"{k:34s} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/ts:5.1f}%")
