"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump: samples, executed warp
instructions and local-memory sectors per CUDA source line (top N)."""
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = "?"
out = []
hdr = None


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {n: hdr.index(n) for n in ("# Samples", "Instructions Executed", "L2 Theoretical Sectors Local", "Avg. Threads Executed")}
        continue
    if hdr is None or not r[0].strip().isdigit():
        continue
    out.append((num(r[ci["# Samples"]]), num(r[ci["Instructions Executed"]]), num(r[ci["L2 Theoretical Sectors Local"]]),
                num(r[ci["Avg. Threads Executed"]]), cur_file, r[0], r[1].strip()[:110]))
tot_s = sum(o[0] for o in out) or 1
tot_i = sum(o[1] for o in out) or 1
print(f"total samples {tot_s:.0f}  warp instructions {tot_i:.3g}")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100 * o[0] / tot_s:5.1f}% smp {100 * o[1] / tot_i:5.1f}% inst  loc {o[2]:.2g}  thr {o[3]:4.1f}  {o[4]}:{o[5]}  {o[6]}")
