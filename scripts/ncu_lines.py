"""Per-source-line hot spots of one kernel in an .ncu-rep captured with --import-source on.
Usage: python scripts/ncu_lines.py rep.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


lines = {}
hdr = None
fname = ""
cur = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        d = lines.setdefault(cur, {"src": r[1].strip(), "samples": 0, "inst": 0, "stall": {}})
        continue
    if cur is None:
        continue
    d = lines[cur]
    d["samples"] += num(r[hdr.index("# Samples")])
    d["inst"] += num(r[hdr.index("Instructions Executed")])
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and num(r[i]) > 0:
            d["stall"][h[6:]] = d["stall"].get(h[6:], 0) + num(r[i])
ts = sum(d["samples"] for d in lines.values())
ti = sum(d["inst"] for d in lines.values())
print(f"total samples {ts}, warp instructions {ti}")
for (f, ln), d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(d["stall"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{f}:{ln:4d} samp {100*d['samples']/max(ts,1):5.1f}% inst {100*d['inst']/max(ti,1):5.1f}% "
          f"{','.join(f'{k}:{v}' for k, v in st):40s} | {d['src'][:90]}")
