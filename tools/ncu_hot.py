"""Top SASS lines by executed instructions / stall samples for one kernel of an .ncu-rep (needs -lineinfo + --import-source)."""
import csv, io, subprocess, sys
rep, kid, topn = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}" if False else "", ], capture_output=True, text=True).stdout if False else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for hi, r in enumerate(rows):
    if "Source" in r:
        break
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
si, ie = h.index("Source"), h.index("Instructions Executed")
ss = h.index("Warp Stall Sampling (All Samples)")
te = h.index("Thread Instructions Executed")
f = lambda x: float(x) if x.replace(".", "", 1).isdigit() else 0.0
ti, ts = sum(f(r[ie]) for r in body), sum(f(r[ss]) for r in body)
print(f"kernel launch {kid}: {len(body)} SASS lines, {ti:.0f} warp-instr, {ts:.0f} samples")
mode = sys.argv[4] if len(sys.argv) > 4 else "inst"
key = (lambda r: -f(r[ie])) if mode == "inst" else (lambda r: -f(r[ss]))
for r in sorted(body, key=key)[:topn]:
    print(f"inst {f(r[ie]) / ti * 100:5.2f}%  stall {f(r[ss]) / ts * 100:5.2f}%  thr/inst {f(r[te]) / max(f(r[ie]), 1):5.1f}  {r[si][:100]}")
