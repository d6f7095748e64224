"""Where the file -> result time goes (diagnostic; prints a small table)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from pcr_b200 import align, synth
from pcr_b200.engine import get_engine
from pcr_b200.plyio import probe_ply, read_ply_xyzw, write_ply
eng = get_engine(0)
v = 0.005
src, tgt, _ = synth.make_pair(100000, v, 20241)
def bench(fn, n=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with tempfile.TemporaryDirectory() as d:
    ps, pt = os.path.join(d, "s.ply"), os.path.join(d, "t.ply")
    write_ply(ps, src); write_ply(pt, tgt)
    kw = dict(ransac_iteration=100000, confidence=1.0, seed=7, icp_max_iteration=50, relative_fitness=0.0, relative_rmse=0.0)
    p = eng.default_params(v); p.ransac_max_iter = 100000; p.ransac_confidence = 1.0; p.seed = 7; p.icp_max_iter = 50
    p.icp_rel_fitness = 0.0; p.icp_rel_rmse = 0.0
    print("probe x2            %.3f ms" % bench(lambda: (probe_ply(ps), probe_ply(pt))))
    print("read pinned x2      %.3f ms" % bench(lambda: (read_ply_xyzw(ps), read_ply_xyzw(pt))))
    print("read pageable x2    %.3f ms" % bench(lambda: (read_ply_xyzw(ps, pin=False), read_ply_xyzw(pt, pin=False))))
    for th in (1, 2, 4, 8):
        print("read pinned x2 t=%d  %.3f ms" % (th, bench(lambda: (read_ply_xyzw(ps, threads=th), read_ply_xyzw(pt, threads=th)))))
    hs, ht = read_ply_xyzw(ps)[0], read_ply_xyzw(pt)[0]
    print("pack (H2D) x2       %.3f ms" % bench(lambda: (eng.pack(hs), eng.pack(ht))))
    ds, dt = eng.pack(hs), eng.pack(ht)
    print("align_device        %.3f ms" % bench(lambda: eng.align_device(ds, dt, p)))
    print("pack + align_device %.3f ms" % bench(lambda: eng.align_device(eng.pack(hs), eng.pack(ht), p)))
    print("align_host (arrays) %.3f ms" % bench(lambda: eng.align_host(src, tgt, p)))
    print("align(paths)        %.3f ms" % bench(lambda: align(ps, pt, v, **kw)))
    print("align(arrays)       %.3f ms" % bench(lambda: align(src, tgt, v, **kw)))
