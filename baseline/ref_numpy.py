"""Loader of the reference's own NumPy RANSAC-step functions from baseline/_ref (see make_ref.py).  Returns None when
the copy is absent (a fresh clone without /root/reference)."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    path = os.path.join(HERE, "_ref", "refsrc", "matcher", "ransac.py")
    if not os.path.exists(path):
        return None
    sys.path.insert(0, HERE)
    from o3d_stub import install_stub
    saved = {k: sys.modules.get(k) for k in ("ply", "open3d")}
    install_stub()
    try:
        spec = importlib.util.spec_from_file_location("_ref_matcher_ransac", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        # the stub `ply` module must not shadow this repo's own `ply` package for later imports
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            elif k == "ply":
                sys.modules.pop(k, None)
    return mod
