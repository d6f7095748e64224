"""Top CUDA-C source lines by executed instructions / stall samples for one kernel launch of an .ncu-rep."""
import csv, io, subprocess, sys
rep, kid, topn = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--launch-skip", kid, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r)
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
si, ie, ss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
ln = h.index("#") if "#" in h else 0
f = lambda x: float(x) if x.replace(".", "", 1).isdigit() else 0.0
ti, ts = sum(f(r[ie]) for r in body) or 1, sum(f(r[ss]) for r in body) or 1
print(f"launch {kid}: {ti:.0f} warp-instr, {ts:.0f} samples; columns {h[:4]}")
mode = sys.argv[4] if len(sys.argv) > 4 else "inst"
key = (lambda r: -f(r[ie])) if mode == "inst" else (lambda r: -f(r[ss]))
for r in sorted(body, key=key)[:topn]:
    print(f"inst {f(r[ie]) / ti * 100:5.2f}%  stall {f(r[ss]) / ts * 100:5.2f}%  L{r[ln]:>5s}  {r[si].strip()[:110]}")
