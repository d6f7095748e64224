"""Aggregate ncu per-SASS-instruction counters by CUDA source line (via nvdisasm line info).
usage: ncu_by_line.py <rep> <launch index> <kernel substring> [top n]"""
import csv, io, re, subprocess, sys, os, tempfile, collections
rep, kid, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-matching_b200", "pcr_b200", "libpcr_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
lines_of = None
cands = []  # every function whose section name contains kname: (instruction -> source line) lists
for cb in cubins:
    dis = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    if kname not in dis:
        continue
    cur, infn, seq = None, False, []
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if seq:
                cands.append(seq)
            seq = []
            infn = kname in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            seq.append(cur)
    if seq:
        cands.append(seq)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", kid, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r)
h = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r == h:  # ncu prints the view twice
        break
    if len(r) == len(h):
        body.append(r)
ie, ss = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
f = lambda x: float(x) if x.replace(".", "", 1).isdigit() else 0.0
# template instantiations share a name: take the one with the same number of SASS instructions as the profile
for c in cands:
    if len(c) == len(body):
        lines_of = c
if lines_of is None and cands:
    lines_of = min(cands, key=lambda c: abs(len(c) - len(body)))
print(f"sass lines: ncu {len(body)} nvdisasm {len(lines_of) if lines_of else None} (of {[len(c) for c in cands]})")
agg = collections.defaultdict(lambda: [0.0, 0.0])
for i, r in enumerate(body):
    key = lines_of[i] if lines_of and i < len(lines_of) else ("?", 0)
    agg[key][0] += f(r[ie]); agg[key][1] += f(r[ss])
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
srcs = {}
by = 1 if (len(sys.argv) > 5 and sys.argv[5] == "stall") else 0
for (fn, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][by])[:topn]:
    path = os.path.join(os.path.dirname(SO), "..", "csrc", fn) if fn else None
    text = ""
    try:
        if fn not in srcs: srcs[fn] = open(path).read().splitlines() if os.path.exists(path) else open(os.path.join(os.path.dirname(SO), "..", "..", "include", fn)).read().splitlines()
        text = srcs[fn][ln - 1].strip()[:90]
    except Exception: pass
    print(f"inst {v[0] / ti * 100:5.2f}%  stall {v[1] / ts * 100:5.2f}%  {fn}:{ln}  {text}")
