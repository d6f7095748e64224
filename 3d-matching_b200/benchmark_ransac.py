"""RANSAC benchmark suite on the B200 engine — the counterpart of the reference's benchmark_ransac.py
(benchmark_ransac.py:223-343: same phases, same command-line flags, same report table as src/utils/profiler.py:151-215),
timed on the device with CUDA events instead of time.perf_counter.

    python 3d-matching_b200/benchmark_ransac.py --source sample.ply --target target.ply --voxel-size 0.005
    python 3d-matching_b200/benchmark_ransac.py --synthetic 100000 --voxel-size 0.005 --icp --export-dir out/

Phases (reference line in brackets):
    ply_loading                  Ply() x 2: load, voxel grid, normals, FPFH                         [:31-60]
    correspondence_computation   compute_feature_correspondences(noise_ratio)                       [:63-84]
    ransac_iteration             one manual step = compute_transformation + evaluate_inliers        [:87-125]
    ransac_iterations_batched    the same `test_iterations` steps in two launches (no reference twin)
    full_ransac                  global_registration(voxel, ransac_iterations)                      [:177-202]
    icp_refinement (--icp)       refine_registration on the full-resolution clouds                  [src/matcher/icp.py:17-48]
The reference's deep_copy / sleep phases time GUI costs (benchmark_ransac.py:128-174) and have no counterpart here.
--export-dir writes the aligned source cloud (PLY) and the transforms (JSON): the headless stand-in for the reference's
blocking viewers (src/visualization/draw_registration_result.py:20-49).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
from collections import OrderedDict
from pathlib import Path

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

DATA_DIRECTORY = Path(HERE).parent / "3d_data"  # where the reference looks for sample.ply / target.ply


class EventProfiler:
    """Named sections timed with CUDA event pairs on the current stream (device time, host launch overhead included
    between the two records).  `report()` prints the reference Profiler's table."""

    def __init__(self) -> None:
        self._pending: "OrderedDict[str, list]" = OrderedDict()

    class _Section:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()
            return self

        def __exit__(self, *exc):
            self.b.record()
            self.owner._pending.setdefault(self.name, []).append((self.a, self.b))
            return False

    def section(self, name: str) -> "EventProfiler._Section":
        return EventProfiler._Section(self, name)

    def stats(self) -> "OrderedDict[str, list]":
        torch.cuda.synchronize()
        return OrderedDict((k, [a.elapsed_time(b) * 1e-3 for a, b in v]) for k, v in self._pending.items())

    def report(self) -> str:
        st = self.stats()
        rows = sorted(st.items(), key=lambda kv: -sum(kv[1]))
        out = ["=" * 100, "PROFILING REPORT (CUDA events)", "=" * 100,
               f"{'Name':<40} {'Calls':>8} {'Total (s)':>12} {'Avg (ms)':>12} {'Median (ms)':>12} {'Min (ms)':>12} {'Max (ms)':>12}",
               "-" * 100]
        for name, t in rows:
            out.append(f"{name:<40} {len(t):>8} {sum(t):>12.4f} {sum(t) / len(t) * 1e3:>12.2f} "
                       f"{statistics.median(t) * 1e3:>12.2f} {min(t) * 1e3:>12.2f} {max(t) * 1e3:>12.2f}")
        out += ["-" * 100, f"{'TOTAL':<40} {'':<8} {sum(sum(t) for _, t in rows):>12.4f}", "=" * 100]
        return "\n".join(out)

    def save_report(self, path) -> None:
        Path(path).write_text(self.report() + "\n")


def run_comprehensive_benchmark(src, tgt, voxel_size: float, noise_ratio: float, test_iterations: int,
                                ransac_iterations: int, *, icp: bool = False, export_dir=None, seed: int = 0,
                                report_path="benchmark_results.txt", quiet: bool = False):
    """src / tgt: PLY paths or (n,3) arrays.  Returns (profiler, results dict)."""
    from matcher.icp import refine_registration
    from matcher.ransac import (compute_feature_correspondences, compute_step_transformation, compute_step_transformations,
                                evaluate_inlier_ratio, evaluate_inlier_ratios, global_registration)
    from pcr_b200.plyio import write_ply
    from ply import Ply

    prof = EventProfiler()
    say = (lambda *a: None) if quiet else (lambda *a: print(*a, file=sys.stderr))

    def make(x):
        return Ply(x, voxel_size, noise_sigma=0.0) if isinstance(x, (str, Path)) else Ply.from_points(x, voxel_size)

    make(src)  # warm-up: library load, arena growth (the reference times a cold start; a cold GPU start times the driver)
    with prof.section("ply_loading"):
        source, target = make(src), make(tgt)
    say(f"down-sampled: {len(source.pcd_down.points)} / {len(target.pcd_down.points)} points")

    with prof.section("correspondence_computation"):
        corres = compute_feature_correspondences(source, target, noise_ratio=noise_ratio, seed=seed)
    say(f"correspondences: {len(corres)}")

    best = 0.0
    for i in range(test_iterations):  # the reference's Python loop, one step per iteration (benchmark_ransac.py:105-113)
        with prof.section("ransac_iteration"):
            with prof.section("compute_transformation"):
                step = compute_step_transformation(source, target, corres, seed=seed, index=i)
            with prof.section("evaluate_inliers"):
                w = evaluate_inlier_ratio(source, target, corres, step.transformation, voxel_size)
        best = max(best, w)
    with prof.section("ransac_iterations_batched"):
        Ts = compute_step_transformations(source, target, corres, test_iterations, seed=seed, start=0)
        ws = evaluate_inlier_ratios(source, target, corres, Ts, voxel_size)
    assert test_iterations == 0 or abs(float(ws.max()) - best) < 1e-12, "batched and per-step paths disagree"

    with prof.section("full_ransac"):
        reg = global_registration(source, target, voxel_size, ransac_iterations, seed=seed)
    say(f"full RANSAC: fitness {reg.fitness:.4f}  inlier_rmse {reg.inlier_rmse:.6f}")
    results = {"n_correspondences": int(len(corres)), "best_manual_inlier_ratio": float(best),
               "ransac": {"transformation": np.asarray(reg.transformation).tolist(), "fitness": float(reg.fitness),
                          "inlier_rmse": float(reg.inlier_rmse)}}
    final_T = np.asarray(reg.transformation)
    if icp:
        with prof.section("icp_refinement"):
            ref = refine_registration(source, target, reg.transformation, voxel_size)
        say(f"ICP: fitness {ref.fitness:.4f}  inlier_rmse {ref.inlier_rmse:.6f}")
        results["icp"] = {"transformation": np.asarray(ref.transformation).tolist(), "fitness": float(ref.fitness),
                          "inlier_rmse": float(ref.inlier_rmse)}
        final_T = np.asarray(ref.transformation)

    per_iter = prof.stats().get("ransac_iteration", [0.0])
    results["estimated_10k_iterations_s"] = float(sum(per_iter) / max(len(per_iter), 1) * 10000)  # benchmark_ransac.py:205-220
    if export_dir is not None:
        out = Path(export_dir)
        out.mkdir(parents=True, exist_ok=True)
        pts = np.asarray(source.pcd.points, np.float64)
        write_ply(out / "source_aligned.ply", pts @ final_T[:3, :3].T + final_T[:3, 3])
        (out / "registration.json").write_text(json.dumps(results, indent=1))
    if report_path:
        prof.save_report(report_path)
    if not quiet:
        print(prof.report())
    return prof, results


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="RANSAC performance benchmark suite (B200 engine)")
    ap.add_argument("--source", type=str, default="sample.ply", help="Source PLY file name (default: sample.ply)")
    ap.add_argument("--target", type=str, default="target.ply", help="Target PLY file name (default: target.ply)")
    ap.add_argument("--voxel-size", type=float, default=0.3, help="Voxel size for downsampling (default: 0.3)")
    ap.add_argument("--noise-ratio", type=float, default=0.0, help="Noise ratio for correspondence (default: 0.0)")
    ap.add_argument("--test-iterations", type=int, default=100, help="Number of iterations for testing (default: 100)")
    ap.add_argument("--ransac-iterations", type=int, default=30, help="Number of RANSAC iterations for full pipeline (default: 30)")
    ap.add_argument("--synthetic", type=int, default=0, metavar="N",
                    help="no files: generate an N-point synthetic pair with a known SE(3) (3d_data/ ships no clouds)")
    ap.add_argument("--icp", action="store_true", help="also run refine_registration (point-to-plane ICP)")
    ap.add_argument("--export-dir", type=str, default=None, help="write source_aligned.ply and registration.json here")
    ap.add_argument("--report", type=str, default="benchmark_results.txt", help="report file (default: benchmark_results.txt)")
    args = ap.parse_args(argv)
    if args.synthetic > 0:
        from pcr_b200 import synth
        src, tgt, _ = synth.make_pair(args.synthetic, args.voxel_size, 20240)
    else:
        src, tgt = DATA_DIRECTORY / args.source, DATA_DIRECTORY / args.target
        for p in (src, tgt):
            if not p.exists():
                print(f"file not found: {p}", file=sys.stderr)
                return 1
    run_comprehensive_benchmark(src, tgt, args.voxel_size, args.noise_ratio, args.test_iterations, args.ransac_iterations,
                                icp=args.icp, export_dir=args.export_dir, report_path=args.report)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
