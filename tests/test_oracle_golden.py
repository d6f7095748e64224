"""The CPU oracle reproduces the committed whole-pipeline fixture bit for bit (tests/golden/pipeline_golden.npz, written
by tests/golden/make_pipeline_golden.py): guards the checker itself against drift — a compiler, flag or code change that
moved one bit of any stage would show up here before it could silently move the target the CUDA path is held to."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_the_committed_pipeline_fixture(orc):
    spec = importlib.util.spec_from_file_location("make_pipeline_golden", os.path.join(HERE, "golden", "make_pipeline_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.compute()
    want = np.load(os.path.join(HERE, "golden", "pipeline_golden.npz"))
    assert sorted(got) == sorted(want.files)
    for k in want.files:
        a, b = np.asarray(got[k]), want[k]
        assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), k
    # and the fixture is a sensible registration: the known SE(3) is recovered
    assert want["icp_fitness"] > 0.95 and np.abs(want["icp_T"] - want["T_true"]).max() < 5e-3
    assert len(want["corr"]) > 100 and want["ransac_best_hyp"] >= 0
