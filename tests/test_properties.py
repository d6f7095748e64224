"""Property tests (hypothesis) — SURVEY §8c item 4: size-independent properties of the oracle and of the host code.
No GPU.  Each property is one the CUDA path is also held to in tests/test_gpu_*.py on fixed inputs."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st
from hypothesis.extra import numpy as hnp

from pcr_b200 import synth
from pcr_b200.plyio import read_ply, write_ply

angles = st.floats(-np.pi / 2 + 0.05, np.pi / 2 - 0.05)
shifts = st.floats(-10.0, 10.0)
finite32 = st.floats(width=32, allow_nan=False, allow_infinity=False)
COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(max_examples=60, **COMMON)
@given(ax=angles, ay=angles, az=angles, tx=shifts, ty=shifts, tz=shifts, seed=st.integers(0, 2**31 - 1))
def test_kabsch_recovers_any_rigid_motion_from_three_pairs(orc, ax, ay, az, tx, ty, tz, seed):
    """compute_step_transformation's contract (src/matcher/ransac.py:143-181): three non-degenerate exact pairs
    determine the rigid motion."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((3, 3))
    # keep the triangle well conditioned: area not tiny relative to its size
    area = np.linalg.norm(np.cross(s[1] - s[0], s[2] - s[0]))
    if area < 0.05 * np.linalg.norm(s - s.mean(0)) ** 2:
        return
    R = synth.euler_zyx(ax, ay, az)
    t = s @ R.T + np.array([tx, ty, tz])
    T = orc.kabsch3(s, t)
    assert np.abs(T[:3, :3] - R).max() < 1e-7 and np.abs(T[:3, 3] - [tx, ty, tz]).max() < 1e-6
    assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-9


@settings(max_examples=25, **COMMON)
@given(ax=angles, ay=angles, az=angles, tx=shifts, ty=shifts, tz=shifts)
def test_inlier_count_is_invariant_under_a_common_rigid_motion(orc, ax, ay, az, tx, ty, tz):
    """evaluate_inlier_ratio (src/matcher/ransac.py:195-236): moving BOTH clouds by the same motion M and conjugating
    the hypothesis (M T M^-1) leaves the count unchanged, up to pairs that sit on the threshold within fp32 rounding."""
    rng = np.random.default_rng(5)
    n = 400
    src = rng.random((n, 3)).astype(np.float32)
    T = np.eye(4)
    T[:3, :3] = synth.euler_zyx(0.2, -0.1, 0.3)
    T[:3, 3] = [0.05, -0.02, 0.01]
    tgt = (src.astype(np.float64) @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 0.02, (n, 3))).astype(np.float32)
    corr = np.stack([np.arange(n), np.arange(n)], 1).astype(np.int32)
    thr = 0.04
    base = orc.inlier_count(src, tgt, corr, T, thr)
    M = np.eye(4)
    M[:3, :3] = synth.euler_zyx(ax, ay, az)
    M[:3, 3] = [tx, ty, tz]
    s2 = (src.astype(np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    t2 = (tgt.astype(np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    T2 = M @ T @ np.linalg.inv(M)
    moved = orc.inlier_count(s2, t2, corr, T2, thr)
    d = np.linalg.norm(src.astype(np.float64) @ T[:3, :3].T + T[:3, 3] - tgt, axis=1)
    borderline = int((np.abs(d - thr) < 2e-5).sum())  # fp32 coordinates of magnitude ~10: 1e-6 each
    assert 0.2 * n < base < 0.98 * n and abs(moved - base) <= borderline


@settings(max_examples=40, **COMMON)
@given(pts=hnp.arrays(np.float32, st.tuples(st.integers(0, 40), st.just(3)), elements=finite32), binary=st.booleans())
def test_ply_round_trip_for_any_finite_fp32_cloud(tmp_path_factory, pts, binary):
    p = tmp_path_factory.mktemp("ply") / "h.ply"
    write_ply(p, pts, binary=binary)
    back, nrm = read_ply(p)
    assert nrm is None and back.astype(np.float32).tobytes() == pts.tobytes()


@settings(max_examples=30, **COMMON)
@given(seed=st.integers(0, 2**31 - 1), voxel=st.floats(0.05, 0.5))
def test_voxel_downsample_partitions_the_cloud(orc, seed, voxel):
    """VoxelDownSample (src/ply/ply.py:106, A.1): every output is the mean of the inputs of one voxel; the outputs'
    count-weighted mean is the cloud's mean; no two outputs share a voxel; a second pass changes nothing."""
    rng = np.random.default_rng(seed)
    pts = rng.random((300, 3)).astype(np.float32)
    out = orc.voxel_downsample(pts, voxel)
    assert 1 <= len(out) <= len(pts)
    org = pts.min(0).astype(np.float64) - voxel / 2
    key = lambda a: [tuple(k) for k in np.floor((a.astype(np.float64) - org) / voxel).astype(np.int64)]  # noqa: E731
    kin = key(pts)
    assert len(set(kin)) == len(out)
    # each output is the mean of its voxel's inputs (fp32 rounding of the stored mean)
    sums = {}
    for k, p in zip(kin, pts.astype(np.float64)):
        s = sums.setdefault(k, [np.zeros(3), 0])
        s[0] += p
        s[1] += 1
    means = np.array(sorted([s[0] / s[1] for s in sums.values()], key=lambda m: tuple(m)))
    got = np.array(sorted(out.astype(np.float64), key=lambda m: tuple(m)))
    assert np.abs(means - got).max() < 1e-6
