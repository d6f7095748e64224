"""Golden record for BASELINE.json configs[3] ("RANSAC 10M hypotheses", SURVEY §8d cfg 4): the CPU oracle's sequential
RANSAC loop (oracle/pcr_oracle.c: orc_ransac, restating src/matcher/ransac.py:42-59 / SURVEY A.6) run ONCE offline over
10,000,000 hypotheses at confidence 1.0 on the cfg-2 pair (seed 20242, voxel 0.005, RANSAC seed 7) — about 4.6e9
KD-tree queries, minutes of CPU time, so it is committed instead of recomputed.  bench.py's RANSAC leg and
tests/test_gpu_parity.py compare the sharded GPU run with this record at every GPU count.

    python tests/golden/make_cfg4_golden.py            # writes tests/golden/cfg4_ransac_10m.json
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-matching_b200")]

import numpy as np  # noqa: E402

from oracle import pcr_oracle as orc  # noqa: E402
from pcr_b200 import synth  # noqa: E402

V, SEED_PAIR, H, SEED = 0.005, 20242, int(os.environ.get("CFG4_H", "10000000")), 7


def main():
    orc.build()
    src, tgt, _ = synth.make_pair(100000, V, SEED_PAIR)
    S, G = orc.preprocess(src, V, full_normals=False), orc.preprocess(tgt, V, full_normals=False)
    corr = orc.match_features(S.pcd_fpfh, G.pcd_fpfh, True)
    t0 = time.time()
    r = orc.ransac(S.pcd_down, G.pcd_down, corr, 1.5 * V, H, 1.0, seed=SEED)
    out = {"pair_seed": SEED_PAIR, "n_points": 100000, "voxel": V, "hypotheses": H, "confidence": 1.0, "ransac_seed": SEED,
           "ms": int(len(S.pcd_down)), "mt": int(len(G.pcd_down)), "n_corr": int(len(corr)),
           "corr_sha256": hashlib.sha256(np.ascontiguousarray(corr, np.int32).tobytes()).hexdigest(),
           "best_hyp": int(r.best_hyp), "inlier_count": int(r.inlier_count), "sum_d2_fixed": int(r.sum_d2_fixed),
           "k_d": int(r.k_d), "hyp_evaluated": int(r.hyp_evaluated), "survivors": int(r.survivors),
           "transformation": [float(x) for x in np.asarray(r.transformation).reshape(-1)],
           "transformation_hex": [float(x).hex() for x in np.asarray(r.transformation).reshape(-1)],
           "oracle_seconds": round(time.time() - t0, 1), "oracle_threads": orc.num_threads()}
    name = "cfg4_ransac_10m.json" if H == 10000000 else f"cfg4_ransac_{H}.json"
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), name), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
