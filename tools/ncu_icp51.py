"""Merges the 51-pass ICP launches of tools/prof_icp.py (ncu --set full -k regex:k_icp_persist -c 2: launch 0 = 100k points,
launch 1 = 1M points) into profiles/r2_traffic.json as the per-PASS figures of the classes icp_pass / icp_pass_1m:
executed thread instructions, DRAM bytes, time under ncu, issue-active %.  usage: ncu_icp51.py <rep> <traffic.json>"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tmul = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
f = lambda r, n: float(r[col[n]].replace(",", ""))
d = json.load(open(out))
ls = [r for r in rows[2:] if len(r) == len(hdr) and "k_icp_persist" in r[col["Kernel Name"]]]
ls.sort(key=lambda r: f(r, "launch__grid_size"))
PASSES = 51.0
for r, key in zip(ls, ("icp_pass", "icp_pass_1m")):
    dram = sum(f(r, n) * mul.get(units[col[n]], 1.0) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    d["kernels"][key] = {
        "dram_bytes_per_launch": dram / PASSES, "passes_in_capture": int(PASSES), "launches": 1,
        "thread_inst_per_launch": f(r, "smsp__inst_executed.sum") * f(r, "smsp__thread_inst_executed_per_inst_executed.ratio") / PASSES,
        "warp_inst_per_launch": f(r, "smsp__inst_executed.sum") / PASSES,
        "avg_us_under_ncu": f(r, "gpu__time_duration.sum") * tmul.get(units[col["gpu__time_duration.sum"]], 1.0) / PASSES,
        "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "grid": int(f(r, "launch__grid_size")), "source": rep}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps({k: d["kernels"][k] for k in ("icp_pass", "icp_pass_1m")}, indent=1))
