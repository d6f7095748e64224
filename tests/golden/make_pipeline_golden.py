"""Writes tests/golden/pipeline_golden.npz: the CPU oracle's outputs for one small whole-pipeline case (the pair that
__graft_entry__.smoke() aligns: synth.make_pair(8000, 0.005, 123), RANSAC 20,000 iterations, confidence 0.999, seed 5,
ICP defaults), stage by stage.  The fixture pins (i) the oracle itself against drift — tests/test_oracle_golden.py
re-runs it and demands identical bits — and (ii) the CUDA path against a committed file rather than only against a
live oracle run (tests/test_gpu_parity.py::test_pipeline_matches_the_committed_golden).

Regenerate ONLY together with a deliberate change of the arithmetic specification:
    python tests/golden/make_pipeline_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "3d-matching_b200"), ROOT]

CASE = dict(n=8000, voxel=0.005, pair_seed=123, ransac_iter=20000, confidence=0.999, ransac_seed=5)


def compute():
    from oracle import pcr_oracle as orc
    from pcr_b200 import synth
    orc.build()
    v = CASE["voxel"]
    src, tgt, T_true = synth.make_pair(CASE["n"], v, CASE["pair_seed"])
    S, G = orc.preprocess(src, v), orc.preprocess(tgt, v)
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    ro = orc.global_registration(S, G, v, CASE["ransac_iter"], CASE["confidence"], CASE["ransac_seed"])
    io = orc.refine_registration(S, G, ro.transformation, v)
    crc = lambda a: np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).astype(np.uint64).dot(  # noqa: E731
        (np.arange(a.nbytes, dtype=np.uint64) % np.uint64(65521)) + np.uint64(1)) % np.uint64(2**61 - 1)
    return dict(
        src_down=S.pcd_down, tgt_down=G.pcd_down, src_down_normals=S.down_normals,
        src_fpfh_checksum=np.uint64(crc(S.pcd_fpfh)), tgt_fpfh_checksum=np.uint64(crc(G.pcd_fpfh)),
        tgt_normals_checksum=np.uint64(crc(G.normals)),
        corr=corr, ransac_T=ro.transformation, ransac_best_hyp=np.int64(ro.best_hyp), ransac_hyp_evaluated=np.int64(ro.hyp_evaluated),
        ransac_inlier_count=np.int64(ro.inlier_count), ransac_sum_d2_fixed=np.int64(ro.sum_d2_fixed),
        icp_T=io.transformation, icp_fitness=np.float64(io.fitness), icp_inlier_rmse=np.float64(io.inlier_rmse),
        icp_iterations=np.int64(io.iterations), icp_inlier_count=np.int64(io.inlier_count), T_true=T_true)


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "pipeline_golden.npz"), **compute())
    print("written", os.path.join(HERE, "pipeline_golden.npz"))
