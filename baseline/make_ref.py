"""Populate baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box) with an UNMODIFIED copy of the
one reference file whose arithmetic can run without open3d: src/matcher/ransac.py (the NumPy 3-point Kabsch and the
inlier-ratio functions, :104-277).  The reference project cannot be pip-installed (pyproject.toml:1-17 has no build
backend and depends on the absent open3d wheel), so this copy is the reference arm for those functions; it is made by
__graft_entry__.build() whenever /root/reference is present and is never committed.

    python baseline/make_ref.py
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/matcher/ransac.py"


def main() -> bool:
    if not os.path.exists(SRC):
        return False
    dst_dir = os.path.join(HERE, "_ref", "refsrc", "matcher")
    os.makedirs(dst_dir, exist_ok=True)
    shutil.copyfile(SRC, os.path.join(dst_dir, "ransac.py"))
    open(os.path.join(dst_dir, "__init__.py"), "w").close()
    with open(SRC, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": sha, "note": "unmodified copy; not committed"}, f)
    return True


if __name__ == "__main__":
    print("baseline/_ref", "written" if main() else "skipped (no /root/reference)")
