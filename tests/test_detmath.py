"""include/pcr_detmath.h against libm: the deterministic functions must be accurate elementary functions."""
import math

import numpy as np


def ulp_err(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / max(np.spacing(abs(b)), 5e-324)


def test_sincos_accuracy(orc):
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-10, 10, 4000), rng.uniform(-1e-3, 1e-3, 500), rng.uniform(-1000, 1000, 500),
                         [0.0, math.pi / 4, -math.pi / 4, math.pi / 2, math.pi, 1e-300]])
    worst = 0.0
    for x in xs:
        s, c, _, _ = orc.detmath(float(x), 0.0)
        # absolute error relative to 1 ulp of 1.0 near zeros of the function, relative elsewhere
        worst = max(worst, abs(s - math.sin(x)) / max(abs(math.sin(x)), 1e-3), abs(c - math.cos(x)) / max(abs(math.cos(x)), 1e-3))
    assert worst < 1e-15


def test_atan2_accuracy(orc):
    rng = np.random.default_rng(2)
    worst = 0.0
    for _ in range(6000):
        y, x = rng.normal(), rng.normal()
        if rng.random() < 0.1:
            y *= 1e-9
        if rng.random() < 0.1:
            x *= 1e-9
        a = orc.detmath(float(x), float(y))[2]
        worst = max(worst, ulp_err(a, math.atan2(y, x)))
    assert worst <= 4.0
    assert orc.detmath(0.0, 0.0)[2] == 0.0
    assert orc.detmath(-1.0, 0.0)[2] == math.pi
    assert orc.detmath(0.0, 1.0)[2] == math.pi / 2
    assert orc.detmath(0.0, -1.0)[2] == -math.pi / 2


def test_acos_accuracy(orc):
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-1, 1, 4000), 1 - np.abs(rng.normal(0, 1e-8, 300)), -1 + np.abs(rng.normal(0, 1e-8, 300)),
                         [1.0, -1.0, 0.0, 0.5, -0.5]])
    worst = 0.0
    for x in xs:
        a = orc.detmath(float(x), 0.0)[3]
        worst = max(worst, ulp_err(a, math.acos(x)))
    assert worst <= 4.0
    assert math.isnan(orc.detmath(1.0000001, 0.0)[3])
    assert orc.detmath(1.0, 0.0)[3] == 0.0
