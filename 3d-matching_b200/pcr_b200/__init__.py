"""pcr_b200 — B200-native point-cloud registration engine (host layer).

Public surface: `align` (the north-star call), `Engine` (one per device) and `synth` (synthetic clouds).
The heavy lifting is in libpcr_b200.so (hand-written sm_100a CUDA behind the C ABI of include/pcr.h).
"""
from . import synth  # noqa: F401  (numpy only; importable without a GPU)

__all__ = ["synth", "Engine", "get_engine", "align"]


def __getattr__(name):
    # engine/api import torch and load the CUDA library: defer so that `import pcr_b200` works on CPU-only hosts
    if name in ("Engine", "get_engine"):
        from . import engine
        return getattr(engine, name)
    if name == "align":
        from .api import align
        return align
    raise AttributeError(name)
