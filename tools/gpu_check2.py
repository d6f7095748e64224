"""Stage-by-stage GPU vs oracle parity check (run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "3d-matching_b200"))
import numpy as np, torch
from oracle import pcr_oracle as orc
from pcr_b200 import synth
from pcr_b200.engine import Engine

eng = Engine(0)
def xyz(t): return t[:, :3].cpu().numpy()
def ev():
    return torch.cuda.Event(enable_timing=True)
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); a, b = ev(), ev(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return r, best

N = int(os.environ.get("N", "100000")); v = 0.005
src, tgt, Ttrue = synth.make_pair(N, v, 20242)
ds, dt = eng.pack(src), eng.pack(tgt)

# voxel
sd, ms_v = timed(lambda: eng.voxel_downsample(ds, v)); td = eng.voxel_downsample(dt, v)
t0 = time.time(); osd = orc.voxel_downsample(src, v); otd = orc.voxel_downsample(tgt, v); t_or = time.time() - t0
print("voxel: M", len(osd), len(otd), "dev", sd.shape[0], td.shape[0], "equal", np.array_equal(xyz(sd), osd), np.array_equal(xyz(td), otd), "dev ms %.3f oracle(2 clouds) %.1f ms" % (ms_v, t_or * 1e3))
# knn lists
(idx, d2, cnt), ms_k = timed(lambda: eng.knn_hybrid(sd, sd, 5 * v, 100))
t0 = time.time(); oi, od, oc = orc.knn_hybrid(osd, osd, 5 * v, 100); t_or = time.time() - t0
print("knn(5v,100): idx eq", np.array_equal(idx.cpu().numpy(), oi), "d2 eq", np.array_equal(d2.cpu().numpy(), od), "cnt eq", np.array_equal(cnt.cpu().numpy(), oc), "mean cnt %.1f" % oc.mean(), "dev ms %.3f oracle %.1f ms" % (ms_k, t_or * 1e3))
(idx, d2, cnt), ms_k = timed(lambda: eng.knn_hybrid(ds, ds[:20000].contiguous(), 2 * v, 30))
oi, od, oc = orc.knn_hybrid(src, src[:20000], 2 * v, 30)
print("knn full(2v,30) 20k queries: idx eq", np.array_equal(idx.cpu().numpy(), oi), "d2 eq", np.array_equal(d2.cpu().numpy(), od), "dev ms %.3f" % ms_k)
# normals
sn, ms_n = timed(lambda: eng.estimate_normals(sd, 2 * v, 30)); tn = eng.estimate_normals(td, 2 * v, 30)
t0 = time.time(); osn = orc.estimate_normals(osd, 2 * v, 30); otn = orc.estimate_normals(otd, 2 * v, 30); t_or = time.time() - t0
print("normals down: bit-equal", np.array_equal(xyz(sn), osn), np.array_equal(xyz(tn), otn), "max diff", np.abs(xyz(sn) - osn).max(), "dev ms %.3f oracle(2) %.1f ms" % (ms_n, t_or * 1e3))
fn_t, ms_nf = timed(lambda: eng.estimate_normals(dt, 2 * v, 30))
t0 = time.time(); ofn_t = orc.estimate_normals(tgt, 2 * v, 30); t_or = time.time() - t0
print("normals full: bit-equal", np.array_equal(xyz(fn_t), ofn_t), "n mismatch rows", int((xyz(fn_t) != ofn_t).any(1).sum()), "dev ms %.3f oracle %.1f ms" % (ms_nf, t_or * 1e3))
# fpfh
sf, ms_f = timed(lambda: eng.compute_fpfh(sd, sn, 5 * v, 100)); tf = eng.compute_fpfh(td, tn, 5 * v, 100)
t0 = time.time(); osf = orc.fpfh(osd, osn, 5 * v, 100); otf = orc.fpfh(otd, otn, 5 * v, 100); t_or = time.time() - t0
print("fpfh: bit-equal", np.array_equal(sf.cpu().numpy(), osf), np.array_equal(tf.cpu().numpy(), otf), "max diff", np.abs(sf.cpu().numpy() - osf).max(), "dev ms %.3f oracle(2) %.1f ms" % (ms_f, t_or * 1e3))
# match
corr, ms_m = timed(lambda: eng.match_features(sf, tf, True))
t0 = time.time(); ocorr = orc.match_features(osf, otf, True); t_or = time.time() - t0
print("match mutual: C", len(ocorr), corr.shape[0], "equal", np.array_equal(corr.cpu().numpy(), ocorr), "dev ms %.3f oracle %.1f ms" % (ms_m, t_or * 1e3))
corr1 = eng.match_features(sf, tf, False); ocorr1 = orc.match_features(osf, otf, False)
print("match one-way equal", np.array_equal(corr1.cpu().numpy(), ocorr1))
# ransac
for conf, iters in ((0.999, 100000), (1.0, 20000)):
    r, ms_r = timed(lambda: eng.ransac(sd, td, corr, 1.5 * v, iters, conf, seed=7), reps=2)
    t0 = time.time(); o = orc.ransac(osd, otd, ocorr, 1.5 * v, iters, conf, seed=7); t_or = time.time() - t0
    print("ransac conf", conf, ": best_hyp", r.best_hyp, o.best_hyp, "count", r.inlier_count, o.inlier_count, "sumq eq", r.sum_d2_fixed == o.sum_d2_fixed,
          "evaluated", r.hyp_evaluated, o.hyp_evaluated, "est_k", r.est_k, o.est_k, "surv", r.survivors, o.survivors,
          "T bit-eq", np.array_equal(r.transformation, o.transformation), "maxdT %.2e" % np.abs(r.transformation - o.transformation).max(),
          "fit %.4f rmse %.6f" % (r.fitness, r.inlier_rmse), "dev ms %.3f oracle %.1f ms" % (ms_r, t_or * 1e3))
# manual-step twins
Ts = eng.ransac_step(sd, td, corr, 3, 0, 1000)
cnts = eng.inlier_count(sd, td, corr, Ts, 1.5 * v).cpu().numpy(); Tsn = Ts.cpu().numpy()
ok = True
for h in (0, 1, 17, 999):
    oT, smp = orc.ransac_step(osd, otd, ocorr, 3, h)
    ok &= np.array_equal(oT, Tsn[h]) and orc.inlier_count(osd, otd, ocorr, oT, 1.5 * v) == cnts[h]
print("ransac_step / inlier_count bit-equal on samples:", ok)
# icp from ransac result
fn_tt = fn_t
g, ms_i = timed(lambda: eng.icp_point_to_plane(ds, dt, fn_tt, 0.4 * v, r.transformation, 30)[0])
t0 = time.time(); oi = orc.icp_point_to_plane(src, tgt, ofn_t, 0.4 * v, o.transformation, 30); t_or = time.time() - t0
print("icp: T bit-eq", np.array_equal(g.transformation, oi.transformation), "maxdT %.2e" % np.abs(g.transformation - oi.transformation).max(), "count", g.inlier_count, oi.inlier_count, "iters", g.iterations, oi.iterations,
      "fit %.4f rmse %.6f" % (g.fitness, g.inlier_rmse), "err vs truth %.2e" % np.abs(g.transformation - Ttrue).max(), "dev ms %.3f oracle %.1f ms" % (ms_i, t_or * 1e3))
# end to end
p = eng.default_params(v); p.ransac_max_iter = 100000; p.seed = 7; p.icp_max_iter = 30
for rep in range(3):
    t0 = time.perf_counter(); res = eng.align_host(src, tgt, p); t1 = time.perf_counter()
    print("align_host wall %.2f ms; stages(ms):" % ((t1 - t0) * 1e3), ["%.2f" % x for x in res.stage_ms], "fit %.4f rmse %.6f" % (res.icp.fitness, res.icp.inlier_rmse), "M", res.n_src_down, res.n_tgt_down, "C", res.n_corr, "ransac eval", res.ransac.hyp_evaluated)
print("align T == staged T:", np.array_equal(np.array(res.icp.transformation).reshape(4, 4), g.transformation))
print("launches", eng.launch_count())
