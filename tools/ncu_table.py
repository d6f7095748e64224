"""Markdown table of the key counters of every kernel in an .ncu-rep (ncu --set full capture) -> profiles/."""
import csv, io, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
def g(r, n, f=float):
    try:
        return f(r[col[n]].replace(",", ""))
    except Exception:
        return float("nan")
spec = [("time us", lambda r: g(r, "gpu__time_duration.sum") * (1e-3 if rows[1][col["gpu__time_duration.sum"]] in ("ns", "nsecond") else (1e3 if rows[1][col["gpu__time_duration.sum"]] in ("ms", "msecond") else 1.0))),
        ("grid", lambda r: g(r, "launch__grid_size")), ("block", lambda r: g(r, "launch__block_size")),
        ("regs", lambda r: g(r, "launch__registers_per_thread")),
        ("warps active %", lambda r: g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")),
        ("issue active %", lambda r: g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
        ("thr/inst", lambda r: g(r, "smsp__thread_inst_executed_per_inst_executed.ratio")),
        ("warp inst M", lambda r: g(r, "smsp__inst_executed.sum") / 1e6),
        ("dram rd MB", lambda r: g(r, "dram__bytes_read.sum") * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(rows[1][col["dram__bytes_read.sum"]], 1.0)),
        ("dram wr MB", lambda r: g(r, "dram__bytes_write.sum") * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(rows[1][col["dram__bytes_write.sum"]], 1.0)),
        ("dram % peak", lambda r: g(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")),
        ("L1 hit %", lambda r: g(r, "l1tex__t_sector_hit_rate.pct")), ("L2 hit %", lambda r: g(r, "lts__t_sector_hit_rate.pct")),
        ("fp64 pipe %", lambda r: g(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")),
        ("tensor pipe %", lambda r: g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))]
# units differ per row in ncu raw csv only via the units row; dram units are per-column constants there
with open(out, "w") as o:
    o.write(f"# {title}\n\nSource: `{rep}` (`ncu --set full --clock-control none --import-source on`); one row per captured launch.\n\n")
    o.write("| kernel | " + " | ".join(n for n, _ in spec) + " |\n|---|" + "---:|" * len(spec) + "\n")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0][:40]
        o.write(f"| `{name}` | " + " | ".join(f"{f(r):.2f}" if f(r) == f(r) else "-" for _, f in spec) + " |\n")
print(open(out).read())
