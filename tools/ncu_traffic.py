"""profiles/r2_traffic.json: DRAM bytes (read + write) and executed thread instructions per launch for the kernel classes of bench.py's roofline,
from an `ncu --set full` capture.  usage: ncu_traffic.py <rep> <out.json>
A persistent kernel (k_icp_persist) runs many passes per launch: its bytes are divided by the passes of the capture
(ICP1M_ITERS + 1 for the 1M-point launch, 5 for the 100k launch of tools/prof_target.py)."""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def tinst(r):
    for n in ("smsp__thread_inst_executed.sum", "sm__thread_inst_executed.sum"):
        if n in col:
            try:
                return float(r[col[n]].replace(",", ""))
            except ValueError:
                pass
    try:  # --set full has the warp-instruction count and the average active threads per instruction
        return float(r[col["smsp__inst_executed.sum"]].replace(",", "")) * float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]].replace(",", ""))
    except (KeyError, ValueError):
        return None
def dram(r):
    t = 0.0
    for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        t += float(r[col[n]].replace(",", "")) * mul.get(units[col[n]], 1.0)
    return t
cls = {"k_ransac_validate": "ransac_validate", "k_knn_cov": "knn_cov", "k_icp_persist": "icp_pass", "k_fpfh": "fpfh",
       "k_match_tc": "nn_features", "k_match_final": "match_misc", "k_celllists_": "celllists_build", "k_knn_list": "knn_list",
       "k_ransac_generate": "ransac_generate", "k_spfh": "spfh"}
acc = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]]
    for k, c in cls.items():
        if k in name:
            grid = int(float(r[col["launch__grid_size"]]))
            acc.setdefault(c, []).append({"kernel": name.split("(")[0], "grid": grid, "dram_bytes": dram(r), "thread_inst": tinst(r),
                                          "us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                                          {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)})
res = {"source": rep, "kernels": {}}
for c, ls in acc.items():
    if c == "icp_pass":
        # largest launch = the 1M-point run (21 passes), the others the 100k alignments (5 passes)
        big = max(ls, key=lambda e: e["dram_bytes"])
        res["kernels"]["icp_pass_1m"] = {"dram_bytes_per_launch": big["dram_bytes"] / 21.0, "passes_in_capture": 21, "launches": 1,
                                         "thread_inst_per_launch": big["thread_inst"] / 21.0 if big["thread_inst"] else None,
                                         "avg_us_under_ncu": big["us"] / 21.0}
        small = [e for e in ls if e is not big]
        if small:
            ti = [e["thread_inst"] for e in small if e["thread_inst"] is not None]
            res["kernels"][c] = {"dram_bytes_per_launch": sum(e["dram_bytes"] for e in small) / len(small) / 5.0,
                                 "passes_in_capture": 5, "launches": len(small),
                                 "thread_inst_per_launch": (sum(ti) / len(ti) / 5.0) if ti else None,
                                 "avg_us_under_ncu": sum(e["us"] for e in small) / len(small) / 5.0}
    else:
        ti = [e["thread_inst"] for e in ls if e["thread_inst"] is not None]
        res["kernels"][c] = {"dram_bytes_per_launch": sum(e["dram_bytes"] for e in ls) / len(ls), "launches": len(ls),
                             "thread_inst_per_launch": (sum(ti) / len(ti)) if ti else None,
                             "avg_us_under_ncu": sum(e["us"] for e in ls) / len(ls)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
