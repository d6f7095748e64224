#!/usr/bin/env python
"""bench.py — headline benchmark of the registration hot path (contract in the task statement, §④).

    python bench.py --gpus N --steps K --warmup W          # this engine
    python bench.py --impl reference --steps K --warmup W  # the reference arm: the CPU oracle on the host cores

Metric (BASELINE.json): end-to-end align ms @100k points.  One step = one full alignment of the config-2 pair
(N_s = N_t = 100k synthetic points, voxel 0.005): voxel down-sampling, normals, FPFH, mutual feature matching,
RANSAC over 100k hypotheses, full-resolution normals, 50 point-to-plane ICP iterations.  FIXED WORK: RANSAC
confidence 1.0 (every hypothesis is scored, no early exit) and ICP relative criteria 0 (all 50 iterations run), so no
work is skipped inside the timed region; the reference's default criteria (confidence 0.999, 30 iterations, 1e-6)
are timed as well and reported under "aux".  At N > 1 every rank aligns its own pair (weak scaling, no data-path
collective); value = max-over-ranks step time / N = ms per aligned pair for the whole job.

`value`  : inputs already resident in HBM (pcr_align), CUDA events on the launching stream.
`e2e`    : the same step through the public API with pinned HOST buffers (pcr_align_host): H2D copies of both
           clouds and the D2H read of the result are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "3d-matching_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# Keep stdout to the single JSON line: libraries (NCCL's version banner, for one) write to file descriptor 1 directly.
# Everything this process prints goes to stderr; the JSON line is written to the saved descriptor at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1

N_POINTS = 100000
VOXEL = 0.005
RANSAC_ITERS = 100000
ICP_ITERS = 50
SEED_PAIR = 20242
WORKLOAD = "cfg2: 100k-point pair, voxel 0.005, RANSAC 100k hypotheses (confidence 1.0, fixed work) + 50 ICP iterations (fixed)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_tflops": d.get("bf16_tflops"),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, enabled=True):
        self.gpu = gpu_index
        self.enabled = enabled
        self.proc = None
        self.lines = []

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_pair(seed=SEED_PAIR, n=N_POINTS):
    from pcr_b200 import synth
    return synth.make_pair(n, VOXEL, seed)


# --------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (the reference's own arithmetic lives in the absent open3d wheel)
# --------------------------------------------------------------------------------------------------------------------
def oracle_step(orc, src, tgt, fixed=True):
    S = orc.preprocess(src, VOXEL)
    G = orc.preprocess(tgt, VOXEL)
    conf = 1.0 if fixed else 0.999
    r = orc.global_registration(S, G, VOXEL, RANSAC_ITERS, conf, 7)
    if fixed:
        i = orc.refine_registration(S, G, r.transformation, VOXEL, ICP_ITERS, 0.0, 0.0)
    else:
        i = orc.refine_registration(S, G, r.transformation, VOXEL)
    return r, i


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    from oracle import pcr_oracle as orc
    orc.build()
    orc.set_num_threads(host_threads())
    cores = orc.num_threads()
    src, tgt, _ = make_pair()
    for _ in range(args.warmup):
        oracle_step(orc, src, tgt)
    t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        r, i = oracle_step(orc, src, tgt)
        t.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.mean(t))
    line = {
        "impl": "reference", "metric": "end-to-end align ms @100k pts", "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle (C/OpenMP restatement of the Open3D 0.19.0 semantics; open3d itself "
                   "is absent and cannot be installed), all host threads, whole workload per step"},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": f"{args.steps} full alignments of the 100k pair"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result": {"fitness": i.fitness, "inlier_rmse": i.inlier_rmse, "ransac_best_hyp": r.best_hyp},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------------------------
# this engine
# --------------------------------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pcr_b200.engine import get_engine
    eng = get_engine(local)
    eng.comm_init()  # the context's own NCCL communicator (pcr_comm_init): the sharded legs exchange on the C side
    peaks = load_peaks()

    # Weak scaling = the SAME work on every GPU: all ranks align the same pair (different pairs differ by up to 2x in
    # RANSAC survivors — 3,660 to 8,463 for seeds 20242..20245 — and max-over-ranks would then measure the unluckiest
    # pair, not the scaling).  The batch leg below aligns different pairs per rank.
    src, tgt, T_true = make_pair(SEED_PAIR + int(os.environ.get("PCR_BENCH_SEED_OFFSET", "0")))
    ds, dt = eng.pack(src), eng.pack(tgt)
    src_pin = torch.from_numpy(src).pin_memory().numpy()
    tgt_pin = torch.from_numpy(tgt).pin_memory().numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.tdev)  # > 126 MB L2

    def params(fixed=True):
        p = eng.default_params(VOXEL)
        p.ransac_max_iter = RANSAC_ITERS
        p.ransac_confidence = 1.0 if fixed else 0.999
        p.seed = 7
        p.icp_max_iter = ICP_ITERS if fixed else 30
        p.icp_rel_fitness = 0.0 if fixed else 1e-6
        p.icp_rel_rmse = 0.0 if fixed else 1e-6
        p.source_normals = 1
        return p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        tot = 0.0
        res = None
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        barrier()
        ms = tot / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=eng.tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    p_fixed, p_default = params(True), params(False)

    # ---- device-resident arm (value) with per-kernel timing for the roofline --------------------------------------
    eng.set_profiling(True)
    eng.kernel_stats(reset=True)
    clocks = ClockSampler(local, enabled=(rank == 0))  # one nvidia-smi poller per job, not one per rank
    l0 = eng.launch_count()
    # warm-up outside the sampler
    for _ in range(args.warmup):
        eng.align_device(ds, dt, p_fixed)
    eng.kernel_stats(reset=True)
    l0 = eng.launch_count()
    clocks.start()
    ms_dev, res = timed(lambda: eng.align_device(ds, dt, p_fixed), args.steps, 0)
    clk = clocks.stop()
    launches = eng.launch_count() - l0
    kstats = eng.kernel_stats(reset=True)
    eng.set_profiling(False)

    # ---- end-to-end arm: pinned host buffers in, result on the host out ----------------------------------------------
    ms_e2e, res_e = timed(lambda: eng.align_host(src_pin, tgt_pin, p_fixed), args.steps, args.warmup)
    ms_dev_np, _ = timed(lambda: eng.align_device(ds, dt, p_fixed), args.steps, 1)       # without profiling events
    ms_default, res_d = timed(lambda: eng.align_host(src_pin, tgt_pin, p_default), args.steps, 1)
    # the same fixed-work step WITHOUT the full-resolution normals of the source, which Ply.__init__ estimates
    # (src/ply/ply.py:65) but point-to-plane ICP never reads: reported beside the headline, not as the headline
    p_lazy = params(True)
    p_lazy.source_normals = 0
    ms_lazy, res_l = timed(lambda: eng.align_host(src_pin, tgt_pin, p_lazy), args.steps, 1)
    lazy_same = bool(np.array_equal(np.array(res_l.icp.transformation), np.array(res_e.icp.transformation)))

    # ---- roofline of the dominant kernel class -----------------------------------------------------------------------
    # dominant kernel class of the CRITICAL PATH: time a class spends on pcr_align's helper context (full-resolution
    # normals, ICP search structures) overlaps the critical path and is time-sliced with it, so it is reported
    # separately (aux.kernel_ms_overlapped_per_step) and does not take part in the selection
    def crit(st):
        return st["ms"] - st.get("overlapped_ms", 0.0)
    dom = max(kstats.items(), key=lambda kv: crit(kv[1])) if kstats else None
    roof = None
    step_total = sum(v["ms"] for v in kstats.values()) or 1.0
    if dom:
        name, st = dom
        per_launch_ms = st["ms"] / max(st["launches"], 1)
        roof = roofline_of(name, st, peaks, clk)
        roof.update({"avg_launch_us": per_launch_ms * 1e3, "launches_per_step": st["launches"] / args.steps,
                     "share_of_step": st["ms"] / step_total, "peak_source": peaks["source"],
                     "algorithmic_bytes_per_launch": st["bytes"] / max(st["launches"], 1)})
    # the ICP pass kernel is the kernel the north star sets a bandwidth target for: always report it too
    icp = kstats.get("icp_pass")
    icp_roof = None
    if icp and icp["launches"]:
        ach = icp["bytes"] / (icp["ms"] * 1e-3) / 1e9
        icp_roof = {"achieved_gbs": ach, "frac": ach / peaks["hbm_gbs"], "avg_launch_us": icp["ms"] * 1e3 / icp["launches"],
                    "bytes_per_launch": icp["bytes"] / icp["launches"]}

    # ---- CPU baseline on rank 0 (bounded sample: full workload, few repetitions) ----------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import pcr_oracle as orc
        orc.build()
        orc.set_num_threads(host_threads())
        reps = 2
        oracle_step(orc, src, tgt)  # warm-up (page-in, thread pool)
        t0 = time.perf_counter()
        for _ in range(reps):
            ro, io = oracle_step(orc, src, tgt)
        cpu_ms = (time.perf_counter() - t0) * 1e3 / reps
        icp_c = res_e.icp
        same = (np.array_equal(np.array(icp_c.transformation).reshape(4, 4), io.transformation)
                and res_e.ransac.best_hyp == ro.best_hyp and icp_c.inlier_count == io.inlier_count)
        cpu = {"value": cpu_ms, "unit": "ms", "cores": orc.num_threads(), "kind": "port",
               "sample": f"{reps} full alignments of the same 100k pair (CPU oracle, OpenMP, all host threads)",
               "results_identical_to_gpu": bool(same)}

    # ---- aux: RANSAC hypotheses/s (sharded over ranks) and ICP iterations/s at 1M points -------------------------------
    aux = {"align_ms_reference_default_criteria_e2e": ms_default,
           "align_ms_e2e_without_unused_source_normals": ms_lazy, "without_source_normals_same_result": lazy_same,
           "align_ms_device_resident_no_profiling_events": ms_dev_np,
           "stage_ms_device": {k: float(v) for k, v in zip(
               ["preprocess_both_clouds", "_unused1", "_unused2", "match", "ransac",
                "wait_for_full_res_normals_overlapped_with_ransac", "icp", "total"], res.stage_ms) if not k.startswith("_")},
           "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(kstats.items(), key=lambda kv: -kv[1]["ms"])},
           "kernel_ms_overlapped_per_step": {k: v["overlapped_ms"] / args.steps for k, v in kstats.items() if v.get("overlapped_ms", 0.0) > 0.0},
           "icp_pass_roofline_100k": icp_roof}
    if not args.no_aux:
        aux.update(run_aux(eng, args, world, rank, peaks))
    if rank == 0 and not args.no_cpu and not args.no_cfg1:
        try:
            aux["cfg1"] = run_cfg1(eng, args)
        except Exception as e:  # an auxiliary leg never takes the headline line down
            aux["cfg1"] = {"error": f"{type(e).__name__}: {e}"}

    if world > 1:
        dist.barrier()
    if rank == 0:
        icp_r = res_e.icp
        line = {
            "metric": "end-to-end align ms @100k pts", "value": ms_dev / world, "unit": "ms", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 points / f64 solves / int64 fixed-point sums", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_src": N_POINTS, "n_tgt": N_POINTS, "voxel": VOXEL,
                       "n_src_down": res.n_src_down, "n_tgt_down": res.n_tgt_down, "n_corr": res.n_corr,
                       "parallelism": f"pair-parallel x{world} (every GPU aligns the same 100k pair: identical work per GPU, no "
                                      "collective on the data path)",
                       "l2": "256 MiB fill between timed steps (outside the per-step CUDA-event pair)",
                       "source_full_res_normals": "computed (as Ply.__init__ does), although point-to-plane ICP never reads them"},
            "e2e": {"value": ms_e2e / world, "unit": "ms", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(src.nbytes + tgt.nbytes), "d2h_bytes_per_step": int(ctypes_sizeof_result())},
            "gpu_launches": int(launches),
            # the paths that actually shard, whole job over all ranks, next to the replica headline
            "ransac_hyp_per_s": (aux.get("ransac") or {}).get("hyp_per_s"),
            "ransac_identical_to_oracle_golden_10m": (aux.get("ransac") or {}).get("identical_to_oracle_golden_10m"),
            "batch_pairs_per_s": (aux.get("batch") or {}).get("pairs_per_s"),
            "batch_identical_to_oracle": (aux.get("batch") or {}).get("identical_to_oracle"),
            "icp_iters_per_s_1m": (aux.get("icp_1m") or {}).get("iters_per_s_whole_job"),
            "icp_1m_identical_to_oracle": (aux.get("icp_1m") or {}).get("identical_to_oracle_10_iterations"),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "result": {"fitness": icp_r.fitness, "inlier_rmse": icp_r.inlier_rmse, "icp_iterations": icp_r.iterations,
                       "ransac_hyp_evaluated": res_e.ransac.hyp_evaluated, "ransac_survivors": res_e.ransac.survivors,
                       "max_abs_T_err_vs_truth": float(np.abs(np.array(icp_r.transformation).reshape(4, 4) - T_true).max())},
            "aux": aux,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# working set in L1/L2, DRAM traffic far below the algorithmic bytes; icp_pass: the per-point state of a certified
# pass is L2-resident (ncu: 1.9 MB of DRAM traffic per pass against 5.2 MB algorithmic at 100k points) and the pass is a
# chain of dependent latencies + instruction issue, so the HBM figure (the north star's target form) is kept as a
# secondary field ("report the ICP roofline with the bound that actually binds")
ISSUE_BOUND = ("ransac_validate", "ransac_generate", "knn_cov", "knn_list", "icp_pass")


def roofline_of(name, st, peaks, clk):
    """Roofline of one kernel class against the bound that actually binds it: tensor pipe for the
    descriptor GEMM, thread-instruction issue rate for the L2-resident search kernels (148 SMs x 128 lanes x f_SM;
    executed thread instructions per launch from the committed ncu capture of the same command), HBM otherwise."""
    secs = st["ms"] * 1e-3
    if name == "nn_features" and st["flops"] > 0:
        ach = st["flops"] / secs / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"]}
    elif name in ISSUE_BOUND and ncu_field(name, "thread_inst_per_launch"):
        f_mhz = (clk or {}).get("sm_mhz") or 1965.0
        peak = 148 * 128 * f_mhz * 1e6 / 1e9  # G thread-instructions / s
        ach = ncu_field(name, "thread_inst_per_launch") * st["launches"] / secs / 1e9
        roof = {"kernel": name, "bound": "issue", "achieved": ach, "peak": peak, "unit": "Gthread-inst/s", "frac": ach / peak,
                "l2_level_algorithmic_gbs": st["bytes"] / secs / 1e9,
                "note": "thread instructions per launch: smsp__thread_inst_executed.sum of the committed ncu capture; time: live CUDA events"}
        if name == "icp_pass":
            roof.update({"hbm_algorithmic_gbs": st["bytes"] / secs / 1e9, "hbm_frac": st["bytes"] / secs / 1e9 / peaks["hbm_gbs"],
                         "issue_active_pct_ncu": ncu_field(name, "issue_active_pct"),
                         "note": "a launch = one ICP pass (the persistent kernel's time divided by the passes it ran); thread instructions "
                                 "per pass from the committed 51-pass ncu capture (tools/prof_icp.py), time from live CUDA events; "
                                 "hbm_* = 52 B per point and pass against the measured HBM peak (the north star's form)"})
    else:
        ach = st["bytes"] / secs / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
    roof["traffic"], roof["traffic_source"] = ncu_traffic(name)
    return roof


def run_cfg1(eng, args):
    """BASELINE.json configs[0] / SURVEY 8d cfg 1: the reference's benchmark_ransac.py phases (benchmark_ransac.py:31-220,
    layout of benchmark_results.txt:6-12) on one synthetic 20k-point pair at the reference's default voxel 0.3
    (benchmark_ransac.py:46-47 ignores --voxel-size), timed on the host cores beside the GPU:
      oracle_all / oracle_1core : the CPU oracle (C/OpenMP restatement of the Open3D semantics; open3d is absent);
      reference_numpy           : the reference's OWN NumPy compute_step_transformation + evaluate_inlier_ratio
                                  (src/matcher/ransac.py:104-277, unmodified copy in the git-ignored baseline/_ref);
      gpu                       : this engine through the reference-facing mirror (ply.Ply, matcher.ransac, matcher.icp).
    Wall clock (perf_counter, device synchronised), best of 3 after one warm-up; ms per call."""
    import torch
    from oracle import pcr_oracle as orc
    from pcr_b200 import synth
    from matcher.icp import refine_registration
    from matcher.ransac import (compute_feature_correspondences, compute_step_transformation, compute_step_transformations,
                                evaluate_inlier_ratio, evaluate_inlier_ratios, global_registration)
    from ply import Ply
    orc.build()
    v, n, iters = 0.3, 20000, 100
    src, tgt, _ = synth.make_pair(n, v, 20241)

    def best_of(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts), r

    table = {}

    def oracle_phases(threads, tag):
        orc.set_num_threads(threads)
        ms, (S, G) = best_of(lambda: (orc.preprocess(src, v), orc.preprocess(tgt, v)))
        table.setdefault("ply_loading", {})[tag] = ms
        ms, corr = best_of(lambda: orc.match_features(S.pcd_fpfh, G.pcd_fpfh, False))
        table.setdefault("correspondence_computation", {})[tag] = ms

        def steps():
            for h in range(iters):
                T, _ = orc.ransac_step(S.pcd_down, G.pcd_down, corr, 0, h)
                orc.inlier_count(S.pcd_down, G.pcd_down, corr, T, 1.5 * v)
        ms, _ = best_of(steps)
        table.setdefault("ransac_iteration", {})[tag] = ms / iters
        ms, r = best_of(lambda: orc.global_registration(S, G, v, 30, 0.999, 0))
        table.setdefault("full_ransac", {})[tag] = ms
        ms, _ = best_of(lambda: orc.refine_registration(S, G, r.transformation, v))
        table.setdefault("icp_refinement", {})[tag] = ms
        return S, G, corr, r

    cores = host_threads()
    S, G, ocorr, oran = oracle_phases(cores, "oracle_all_cores_ms")
    oracle_phases(1, "oracle_1_core_ms")
    orc.set_num_threads(cores)

    # the reference's own NumPy path on the same clouds and correspondences
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_numpy
    ref = ref_numpy.load()

    class _Cloud:
        def __init__(self, pts):
            self.points = np.asarray(pts, np.float64)

    class _MockPly:  # the reference's duck type, test_ransac_crash.py:92-96
        def __init__(self, pts):
            self.pcd = self.pcd_down = _Cloud(pts)
            self.pcd_fpfh = None

    if ref is not None:
        ms_, mt_ = _MockPly(S.pcd_down), _MockPly(G.pcd_down)
        t_k = t_e = 0.0
        np.random.seed(0)
        for _ in range(iters):
            t0 = time.perf_counter()
            r_ = ref.compute_step_transformation(ms_, mt_, ocorr)
            t1 = time.perf_counter()
            ref.evaluate_inlier_ratio(ms_, mt_, ocorr, r_.transformation, v)
            t_e += time.perf_counter() - t1
            t_k += t1 - t0
        table["ransac_iteration"]["reference_numpy_ms"] = (t_k + t_e) * 1e3 / iters
        table["compute_transformation"] = {"reference_numpy_ms": t_k * 1e3 / iters}
        table["evaluate_inliers"] = {"reference_numpy_ms": t_e * 1e3 / iters}

    # this engine, same phases through the mirror of the reference's API
    ms, (ps, pt) = best_of(lambda: (Ply.from_points(src, v, noise_sigma=0.0), Ply.from_points(tgt, v, noise_sigma=0.0)))
    table["ply_loading"]["gpu_ms"] = ms  # full-resolution normals are lazy in the mirror: forced below, inside ICP
    ms, corr = best_of(lambda: compute_feature_correspondences(ps, pt, noise_ratio=0.0))
    table["correspondence_computation"]["gpu_ms"] = ms
    same_corr = bool(np.array_equal(corr, ocorr))

    def gpu_steps():
        for h in range(iters):
            r_ = compute_step_transformation(ps, pt, corr, seed=0, index=h)
            evaluate_inlier_ratio(ps, pt, corr, r_.transformation, v)
    ms, _ = best_of(gpu_steps)
    table["ransac_iteration"]["gpu_one_at_a_time_ms"] = ms / iters
    nb = 1 << 20

    def gpu_batched():
        Ts = compute_step_transformations(ps, pt, corr, nb, seed=0, start=0)
        return evaluate_inlier_ratios(ps, pt, corr, Ts, v)
    ms, _ = best_of(gpu_batched)
    hyp_s_gpu = nb / (ms * 1e-3)
    ms, rg = best_of(lambda: global_registration(ps, pt, v, 30))
    table["full_ransac"]["gpu_ms"] = ms
    _ = pt.pcd.normals_xyzw  # estimate_normals on the full-resolution target (part of ply_loading in the reference)
    ms, ri = best_of(lambda: refine_registration(ps, pt, rg.transformation, v))
    table["icp_refinement"]["gpu_ms"] = ms
    oi = orc.refine_registration(S, G, oran.transformation, v)
    out = {"workload": "cfg1: 20k-point pair, voxel 0.3 (reference default), benchmark_ransac.py phases; ms per call",
           "host_cores": cores, "phases": table, "n_down": [int(len(S.pcd_down)), int(len(G.pcd_down))], "n_corr": int(len(ocorr)),
           "gpu_equals_oracle": {"correspondences": same_corr,
                                 "ransac_T": bool(np.array_equal(rg.transformation, oran.transformation)),
                                 "icp_T": bool(np.array_equal(ri.transformation, oi.transformation)),
                                 "icp_fitness": bool(ri.fitness == oi.fitness)},
           "manual_step_hyp_per_s": {"gpu_batched_pcr_ransac_step_plus_pcr_inlier_count": hyp_s_gpu,
                                     "gpu_one_at_a_time": 1e3 / table["ransac_iteration"]["gpu_one_at_a_time_ms"],
                                     "oracle_all_cores": 1e3 / table["ransac_iteration"]["oracle_all_cores_ms"],
                                     "reference_numpy": (1e3 / table["ransac_iteration"]["reference_numpy_ms"]) if ref is not None else None,
                                     "reference_published_benchmark_results_txt": 1e3 / 0.76},
           "reference_published_ms": {"ply_loading": 791.23, "ransac_iteration": 0.76, "evaluate_inliers": 0.50,
                                      "compute_transformation": 0.24, "full_ransac": 21.12, "correspondence_computation": 8.98,
                                      "source": "benchmark_results.txt:6-12 (unknown CPU, unknown clouds)"}}
    return out


def ncu_field(kernel_class, field):
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                e = json.load(f)["kernels"].get(kernel_class)
            if e and e.get(field) is not None:
                return e[field]
        except (OSError, ValueError, KeyError):
            pass
    return None


def ncu_traffic(kernel_class):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel class, from the committed ncu --set full
    capture (profiles/r2_traffic.json, else r1; written by tools/ncu_traffic.py); None when the class was not captured."""
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                t = json.load(f)
            e = t["kernels"].get(kernel_class)
            if e:
                return e["dram_bytes_per_launch"], f"profiles/{fn} ({e.get('source') or t['source']})"
        except (OSError, ValueError, KeyError):
            pass
    return None, None


def ctypes_sizeof_result():
    import ctypes
    from pcr_b200 import _capi
    return ctypes.sizeof(_capi.AlignResult)


def run_aux(eng, args, world, rank, peaks):
    """RANSAC hypotheses/s over all ranks (config 4 style, confidence 1.0) and ICP iterations/s at 1M points (config 3)."""
    import torch
    import torch.distributed as dist
    from pcr_b200 import synth
    out = {}
    v = VOXEL
    src, tgt, _ = make_pair(SEED_PAIR)  # every rank holds the SAME clouds for the sharded RANSAC
    ds, dt = eng.pack(src), eng.pack(tgt)
    sd, td = eng.voxel_downsample(ds, v).contiguous(), eng.voxel_downsample(dt, v).contiguous()
    sn, tn = eng.estimate_normals(sd, 2 * v, 30), eng.estimate_normals(td, 2 * v, 30)
    sf, tf = eng.compute_fpfh(sd, sn, 5 * v, 100), eng.compute_fpfh(td, tn, 5 * v, 100)
    corr = eng.match_features(sf, tf, True).contiguous()
    H = args.ransac_hyps
    # pcr_ransac_multi: wave loop, per-wave ncclAllGather and replay all on the C side of the boundary
    for _ in range(2):  # warm-up at full size: the scratch arena and the exchange buffers grow here, NCCL sets its channels up
        eng.ransac_multi(sd, td, corr, 1.5 * v, H, 1.0, 7)
    runs = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r, n_waves = eng.ransac_multi(sd, td, corr, 1.5 * v, H, 1.0, 7)
        b.record()
        torch.cuda.synchronize()
        t_ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([t_ms], dtype=torch.float64, device=eng.tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        runs.append(t_ms)
    ms = float(np.median(runs))  # max over ranks per run, median of three runs (all three are reported)
    out["ransac"] = {"hypotheses": H, "ms": ms, "hyp_per_s": H / (ms * 1e-3), "survivors": r.survivors,
                     "checker_pass_rate": r.survivors / H, "best_hyp": r.best_hyp, "inlier_count": r.inlier_count,
                     "waves": n_waves, "n_gpus": world, "driver": "pcr_ransac_multi (C, NCCL all-gather per wave)",
                     "ms_of_each_run_max_over_ranks": runs}
    # cfg4 parity: the CPU oracle's sequential loop over the same 10M hypotheses, run once offline (368 s on 8 cores) and
    # committed (tests/golden/cfg4_ransac_10m.json, made by tests/golden/make_cfg4_golden.py); the sharded run must
    # reproduce it at every GPU count — winner, counts, fixed-point sum, every bit of T, survivors, consumed iterations
    try:
        with open(os.path.join(ROOT, "tests", "golden", "cfg4_ransac_10m.json")) as f:
            gold = json.load(f)
        if gold["hypotheses"] == H and gold["pair_seed"] == SEED_PAIR:
            import hashlib
            Tg = np.array([float.fromhex(x) for x in gold["transformation_hex"]]).reshape(4, 4)
            same = (r.best_hyp == gold["best_hyp"] and r.inlier_count == gold["inlier_count"]
                    and r.sum_d2_fixed == gold["sum_d2_fixed"] and r.survivors == gold["survivors"]
                    and r.hyp_evaluated == gold["hyp_evaluated"] and int(corr.shape[0]) == gold["n_corr"]
                    and hashlib.sha256(corr.cpu().numpy().tobytes()).hexdigest() == gold["corr_sha256"]
                    and np.array_equal(np.asarray(r.transformation).reshape(4, 4), Tg))
            out["ransac"]["identical_to_oracle_golden_10m"] = bool(same)
    except (OSError, ValueError, KeyError) as e:
        out["ransac"]["identical_to_oracle_golden_10m"] = f"golden unreadable: {e}"
    # batch of independent pairs (config 5 style: 50k-point pairs, full pipeline, reference-default criteria), pair i ->
    # rank i mod world, no communication until the final all-gather of 18 doubles per pair
    B = args.batch_pairs
    if B > 0:
        pairs, host_pairs = [], {}
        for i in range(rank, B * world, world):  # this rank's pairs: global index rank, rank + world, ...
            s_i, t_i, _ = synth.make_pair(50000, v, 30000 + i)
            pairs.append((eng.pack(s_i), eng.pack(t_i)))
            if rank == 0 and len(host_pairs) < args.batch_check and not args.no_cpu:
                host_pairs[i] = (s_i, t_i)
        pb = eng.default_params(v)
        pb.ransac_max_iter = RANSAC_ITERS
        pb.seed = 7
        batch = {"pairs": B * world, "points_per_cloud": 50000, "criteria": "confidence 0.999, ICP 30 iterations / 1e-6",
                 "driver": "pcr_align_batch (C: native worker threads, one final ncclAllGather)"}
        for workers in (1, 3, 6):
            eng.align_batch(pairs, pb, B * world, workers=workers)  # warm-up (worker contexts, arenas)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tab = eng.align_batch(pairs, pb, B * world, workers=workers)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=eng.tdev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            batch[f"pairs_per_s_workers{workers}"] = B * world / dt
            batch[f"min_fitness_workers{workers}"] = float(tab[:, 16].min())
        best_w = max((1, 3, 6), key=lambda w: batch[f"pairs_per_s_workers{w}"])
        batch["pairs_per_s"] = batch[f"pairs_per_s_workers{best_w}"]
        batch["workers_of_pairs_per_s"] = best_w
        if rank == 0 and not args.no_cpu:  # per-pair parity with the oracle on a sample (the checker, not the thing measured)
            from oracle import pcr_oracle as orc
            orc.build()
            orc.set_num_threads(host_threads())
            same = True
            for i, (s_i, t_i) in host_pairs.items():
                S_, G_ = orc.preprocess(s_i, v), orc.preprocess(t_i, v)
                ro = orc.global_registration(S_, G_, v, RANSAC_ITERS, 0.999, 7)
                io = orc.refine_registration(S_, G_, ro.transformation, v)
                same = same and np.array_equal(tab[i, :16].reshape(4, 4), io.transformation) and tab[i, 16] == io.fitness
            batch["oracle_checked_pairs"] = len(host_pairs)
            batch["identical_to_oracle"] = bool(same)
        out["batch"] = batch
        del pairs
    # ICP at 1M points: replicas only (single-pair ICP does not shard); aggregate = world x per-GPU rate
    n = args.icp_points
    s1, t1, _ = synth.make_icp_pair(n, v, 20243)
    d1s, d1t = eng.pack(s1), eng.pack(t1)
    nrm = eng.estimate_normals(d1t, 2 * v, 30)
    for _ in range(2):  # warm-up: arena growth + consolidation happen here, not in the timed call
        eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 5, 0.0, 0.0, want_corr=True)
    eng.set_profiling(True)
    eng.kernel_stats(reset=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g, _ = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), ICP_ITERS, 0.0, 0.0, want_corr=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    ks = eng.kernel_stats(reset=True).get("icp_pass")
    eng.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    passes = g.iterations + 1
    icp = {"points": n, "passes": passes, "call_ms": ms, "iters_per_s_whole_job": world * passes / (ms * 1e-3),
           "fitness": g.fitness, "n_gpus": world, "parallelism": "replicas only"}
    if ks and ks["launches"]:
        us = ks["ms"] * 1e3 / ks["launches"]
        gbs = ks["bytes"] / (ks["ms"] * 1e-3) / 1e9
        icp.update({"pass_kernel_us": us, "pass_kernel_gbs_algorithmic_52B_per_point": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                    "kernel_iters_per_s_per_gpu": 1e6 / us})
        ti = ncu_field("icp_pass_1m", "thread_inst_per_launch")
        if ti and n == 1000000:  # the bound that binds: thread-instruction issue (148 SMs x 128 lanes x 1.965 GHz)
            icp.update({"issue_gthread_inst_per_s": ti / us / 1e3, "frac_of_issue_peak": ti / us / 1e3 / (148 * 128 * 1.965),
                        "issue_active_pct_ncu": ncu_field("icp_pass_1m", "issue_active_pct"),
                        "dram_bytes_per_pass_ncu": ncu_field("icp_pass_1m", "dram_bytes_per_launch")})
    # cfg3 parity: the first 10 iterations of the same 1M-point run against the CPU oracle (src/matcher/icp.py:42-48):
    # every correspondence index, the inlier count, the fixed-point sum of d2, every bit of T
    if rank == 0 and not args.no_cpu:
        from oracle import pcr_oracle as orc
        orc.build()
        orc.set_num_threads(host_threads())
        g10, c10 = eng.icp_point_to_plane(d1s, d1t, nrm, 0.4 * v, np.eye(4), 10, 0.0, 0.0, want_corr=True)
        t0 = time.perf_counter()
        o10 = orc.icp_point_to_plane(s1, t1, nrm[:, :3].cpu().numpy(), 0.4 * v, np.eye(4), 10, 0.0, 0.0)
        icp["oracle_10_iterations_s"] = time.perf_counter() - t0
        icp["identical_to_oracle_10_iterations"] = bool(
            np.array_equal(c10.cpu().numpy(), o10.correspondence) and g10.inlier_count == o10.inlier_count
            and g10.sum_d2_fixed == o10.sum_d2_fixed and np.array_equal(np.asarray(g10.transformation).reshape(4, 4), o10.transformation)
            and g10.iterations == o10.iterations)
    out["icp_1m"] = icp
    # file -> result on the host (the reference's Ply() starts from PLY files, src/ply/ply.py:80):
    # native pcr_ply_read into pinned packed float4, one H2D copy, the whole alignment.  Wall clock, rank 0 only.
    if rank == 0:
        try:
            import tempfile
            from pcr_b200 import align
            from pcr_b200.plyio import read_ply_xyzw, write_ply
            leg = {"points_per_cloud": int(len(src)), "criteria": "as the headline step (confidence 1.0, 50 fixed ICP iterations)"}
            with tempfile.TemporaryDirectory() as td_:
                for fmt, binary in (("binary_little_endian", True), ("ascii", False)):
                    ps, pt = os.path.join(td_, f"s_{fmt}.ply"), os.path.join(td_, f"t_{fmt}.ply")
                    write_ply(ps, src, binary=binary)
                    write_ply(pt, tgt, binary=binary)
                    kw = dict(ransac_iteration=RANSAC_ITERS, confidence=1.0, seed=7, icp_max_iteration=ICP_ITERS,
                              relative_fitness=0.0, relative_rmse=0.0)
                    for _ in range(2):
                        res = align(ps, pt, v, **kw)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(5):
                        res = align(ps, pt, v, **kw)
                    leg[f"{fmt}_file_to_result_ms"] = (time.perf_counter() - t0) / 5 * 1e3
                    t0 = time.perf_counter()
                    for _ in range(5):
                        read_ply_xyzw(ps), read_ply_xyzw(pt)
                    leg[f"{fmt}_read_both_files_ms"] = (time.perf_counter() - t0) / 5 * 1e3
                    leg[f"{fmt}_fitness"] = float(res[1])
            out["from_ply_files"] = leg
        except Exception as e:  # an auxiliary leg never takes the headline line down
            out["from_ply_files"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-aux", action="store_true", help="skip the RANSAC hyp/s and 1M-point ICP legs")
    ap.add_argument("--ransac-hyps", type=int, default=10000000)
    ap.add_argument("--icp-points", type=int, default=1000000)
    ap.add_argument("--batch-pairs", type=int, default=128, help="pairs per GPU of the batch leg (cfg5: 128 x 8 GPUs = 1024; 0 = skip)")
    ap.add_argument("--batch-check", type=int, default=16, help="pairs of the batch re-run through the CPU oracle (cfg5: 16)")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the cfg1 leg (20k pair, per-phase CPU table beside the GPU)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
