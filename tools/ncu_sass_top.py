"""Per-SASS-instruction summary of an `ncu --page source --csv` dump (one section per kernel launch):
    python tools/ncu_sass_top.py <source.csv> <section index> [min pct]
prints every instruction holding at least `min pct` of the section's stall samples, with its main stall reason."""
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2])
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}
        secs.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["ins"].append(r)
s = secs[want]
h = s["hdr"]
ci = {n: h.index(n) for n in h}
tot = sum(float(r[ci["# Samples"]]) for r in s["ins"]) or 1
toti = sum(float(r[ci["Instructions Executed"]]) for r in s["ins"])
print(s["name"][:60], "samples", tot, "warp-inst", toti, "n_sass", len(s["ins"]))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for k, r in enumerate(s["ins"]):
    p = 100 * float(r[ci["# Samples"]]) / tot
    if p >= minpct:
        top = max(stalls, key=lambda n: float(r[ci[n]]))
        print(f"{k:4d} {p:5.1f}% exec {float(r[ci['Instructions Executed']]) / 1e6:8.1f}M thr {r[ci['Avg. Threads Executed']]:>5} {top[6:]:12s} {r[ci['Source']].strip()[:80]}")
