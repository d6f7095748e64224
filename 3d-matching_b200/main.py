"""Pipeline entry point on the B200 engine — counterpart of the reference's src/main.py:24-39.

Same flow, same four calls:
    1. Ply(src), Ply(tgt)                    load + preprocess (down-sample, normals, FPFH)      src/main.py:30-31
    2. global_registration(src, tgt)         FPFH-feature RANSAC, coarse alignment               src/main.py:34
    3. refine_registration(src, tgt, init)   point-to-plane ICP                                  src/main.py:38
    4. draw_registration_result(...)         after each step — written to files here (no display) src/main.py:35,39

With `3d-matching_b200/` on PYTHONPATH in place of the reference's `src/`, the reference's own main.py runs unchanged
as well (its imports `matcher`, `ply`, `utils.setup_logging`, `visualization` all resolve here).  This file adds what a
headless box needs: paths, voxel size and iteration counts on the command line, a synthetic pair when 3d_data/ holds
no clouds (the reference ships none, 3d_data/.gitignore), and a printed summary.

    python 3d-matching_b200/main.py --source 3d_data/sample.ply --target 3d_data/target.ply --voxel-size 0.005
    python 3d-matching_b200/main.py --synthetic 100000 --voxel-size 0.005 --ransac-iterations 100000
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
if str(HERE) not in sys.path:
    sys.path.insert(0, str(HERE))

from matcher.icp import refine_registration  # noqa: E402
from matcher.ransac import global_registration  # noqa: E402
from ply import Ply  # noqa: E402
from utils.setup_logging import setup_logging  # noqa: E402
from visualization.draw_registration_result import draw_registration_result  # noqa: E402

logger = setup_logging(__name__)

DATA_DIRECTORY = (HERE / ".." / "3d_data").resolve()  # src/main.py:21


def run(src_ply, tgt_ply, voxel_size=None, ransac_iterations: int = 30, *, export_dir=None, seed: int = 0):
    """Steps 2-4 on two preprocessed clouds; returns (ransac_result, icp_result, [exported files])."""
    files = []
    init = global_registration(src_ply, tgt_ply, voxel_size, ransac_iterations, seed=seed)
    logger.info("RANSAC: fitness %.4f inlier_rmse %.6f", init.fitness, init.inlier_rmse)
    files.append(draw_registration_result(src_ply, tgt_ply, init.transformation, out_dir=export_dir))
    icp = refine_registration(src_ply, tgt_ply, init.transformation, voxel_size)
    logger.info("ICP: fitness %.4f inlier_rmse %.6f", icp.fitness, icp.inlier_rmse)
    files.append(draw_registration_result(src_ply, tgt_ply, icp, out_dir=export_dir))
    return init, icp, files


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="FPFH-RANSAC global registration + point-to-plane ICP on one B200")
    ap.add_argument("--source", type=Path, default=DATA_DIRECTORY / "sample.ply")  # src/main.py:26
    ap.add_argument("--target", type=Path, default=DATA_DIRECTORY / "target.ply")  # src/main.py:27
    ap.add_argument("--voxel-size", type=float, default=0.3, help="Ply default (src/ply/ply.py:32)")
    ap.add_argument("--ransac-iterations", type=int, default=30, help="global_registration default (src/matcher/ransac.py:24)")
    ap.add_argument("--noise-sigma", type=float, default=0.05, help="noise on pcd_down after FPFH (src/ply/ply.py:61-62); 0 disables")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--synthetic", type=int, default=0, metavar="N", help="generate an N-point pair with a known SE(3) instead of reading files")
    ap.add_argument("--export-dir", type=Path, default=Path("registration_out"))
    args = ap.parse_args(argv)

    if args.synthetic > 0:
        from pcr_b200 import synth
        s, t, T_true = synth.make_pair(args.synthetic, args.voxel_size, 20240 + args.seed)
        src_ply = Ply.from_points(s, args.voxel_size, noise_sigma=0.0)
        tgt_ply = Ply.from_points(t, args.voxel_size, noise_sigma=0.0)
    else:
        T_true = None
        try:
            src_ply = Ply(args.source, args.voxel_size, noise_sigma=args.noise_sigma, seed=args.seed)
            tgt_ply = Ply(args.target, args.voxel_size, noise_sigma=args.noise_sigma, seed=args.seed + 1)
        except (FileNotFoundError, TypeError, ValueError) as e:  # src/ply/ply.py:46-51, 81-84
            print(f"error: {e}", file=sys.stderr)
            return 1
    init, icp, files = run(src_ply, tgt_ply, args.voxel_size, args.ransac_iterations, export_dir=args.export_dir, seed=args.seed)
    import numpy as np
    np.set_printoptions(precision=6, suppress=True)
    print("transformation (ICP):")
    print(icp.transformation)
    print(f"fitness {icp.fitness:.6f}  inlier_rmse {icp.inlier_rmse:.6g}  (RANSAC: fitness {init.fitness:.6f})")
    if T_true is not None:
        print(f"max |T - T_true| = {np.abs(icp.transformation - T_true).max():.3g}")
    print("written:", ", ".join(str(f) for f in files))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
