"""Debug helper: teacher-forced C3 replay (GPU follows the CPU oracle's trajectory); at the given scans compare the
first linearize of the S2M align (H, b, error, correspondences) and the first LM step between the two.
    python benchmarks/debug_mismatch.py 1117 1118"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks import configs as cf  # noqa: E402
from direct_lidar_odometry_b200 import NanoGICP  # noqa: E402
from oracle import oracle as O  # noqa: E402

targets = sorted(int(a) for a in sys.argv[1:]) or [1117]
last = targets[-1]
scans = cf.gen_scans(list(range(last + 1)))
vox = NanoGICP(0)


def make_gpu(cfg):
    g = NanoGICP(0); cf.configure(g, cfg); return g


rc = cf.Replay(lambda cfg: cf.OracleGicp(O, cfg, os.cpu_count()), lambda p, l: O.voxel_filter(p, l), None)
rg = cf.Replay(make_gpu, lambda p, l: vox.voxel_filter(p, l), None)


def probe(tag, rp, is_gpu):
    # called right before s2m.align(T_s2s): linearize at the guess on the live S2M object
    s2m = rp.s2m if is_gpu else rp.s2m.g
    T = np.asarray(rp._T_s2s, dtype=np.float64)
    L = s2m.linearize(T)
    H, b, e = (L["H"], L["b"], L["err"]) if isinstance(L, dict) else L[:3]
    return dict(H=np.asarray(H), b=np.asarray(b), err=float(e), corr=np.asarray(L["corr"]) if isinstance(L, dict) and "corr" in L else None)


for i in range(last + 1):
    T_true, raw = scans[i]
    sc_c = O.voxel_filter(raw, 0.25); sc_g = vox.voxel_filter(raw, 0.25)
    if i == 0:
        rc.first(sc_c, T_true); rg.first(sc_g, T_true); continue
    if i in targets:
        # re-implement step() up to the S2M align to probe
        import ctypes as C
        from direct_lidar_odometry_b200 import _lib
        for rp, sc, is_gpu in ((rc, sc_c, False), (rg, sc_g, True)):
            rp.s2s.setInputSource(sc); rp.s2m.registerInputSource(sc)
            rp.s2m.source_kdtree_ = rp.s2s.source_kdtree_; rp.s2m.source_covs_.clear()
            s2s_obj = rp.s2s if is_gpu else rp.s2s.g
            rp.s2s.calculateSourceCovariances()
            rp._L0 = s2s_obj.linearize(np.eye(4))
            rp._src_covs = np.asarray(rp.s2s.getSourceCovariances())
            rp._tgt_covs = np.asarray(s2s_obj.getTargetCovariances() if is_gpu else s2s_obj.get_target_covs())
        # manual LM iteration 0 with the host helper, on both implementations' linearize / compute_error
        Lh = _lib.load()
        H = np.ascontiguousarray(rc._L0["H"].T).reshape(36); bb = np.ascontiguousarray(rc._L0["b"])
        lam = 1e-9 * np.abs(np.diag(rc._L0["H"])).max()
        x0 = np.ascontiguousarray(np.eye(4).T).reshape(16)
        d6 = np.zeros(6); delta = np.zeros(16); xi = np.zeros(16)
        dp = C.POINTER(C.c_double)
        Lh.ngicp_lm_trial(H.ctypes.data_as(dp), bb.ctypes.data_as(dp), C.c_double(lam), x0.ctypes.data_as(dp), d6.ctypes.data_as(dp), delta.ctypes.data_as(dp), xi.ctypes.data_as(dp))
        X1 = xi.reshape(4, 4).T.copy()
        print("  manual trial 0: d6", d6, "cond(H) %.3e" % np.linalg.cond(rc._L0["H"]))
        ec, eg = rc.s2s.g.compute_error(X1), rg.s2s.compute_error(X1)
        print("  compute_error(x1): cpu %.12g gpu %.12g (y0 %.12g)" % (ec, eg, rc._L0["err"]))
        L1c, L1g = rc.s2s.g.linearize(X1), rg.s2s.linearize(X1)
        nd = np.nonzero(L1c["corr"] != L1g["corr"])[0]
        print("  linearize(x1): err cpu %.12g gpu %.12g; H rel %.3e b rel %.3e corr differing %d matched %d/%d" % (L1c["err"], L1g["err"],
              np.abs(L1c["H"] - L1g["H"]).max() / np.abs(L1c["H"]).max(), np.abs(L1c["b"] - L1g["b"]).max() / np.abs(L1c["b"]).max(), nd.size,
              int((L1c["corr"] >= 0).sum()), int((L1g["corr"] >= 0).sum())))
        for j in nd[:6]:
            print("    point", int(j), "corr cpu/gpu", int(L1c["corr"][j]), int(L1g["corr"][j]), "sqd cpu/gpu %.9g %.9g" % (float(L1c["sqd"][j]), float(L1g["sqd"][j])))
        rg.s2s.setAlignMode(1); rg.s2s.align(); rs = rg.s2s.result
        print("  s2s gpu STEPPED iters", rs.nr_iterations, "lin", rs.n_linearize, "err-evals", rs.n_compute_error, "t", np.asarray(rg.s2s.getFinalTransformation())[:3, 3])
        rg.s2s.setAlignMode(0)
        for rp, sc, is_gpu in ((rc, sc_c, False), (rg, sc_g, True)):
            rp.s2s.align()
            rr = rp.s2s.result if is_gpu else rp.s2s._r
            print("  s2s", "gpu" if is_gpu else "cpu", "iters", rr.nr_iterations, "lin", rr.n_linearize, "err-evals", rr.n_compute_error, "conv", rr.converged,
                  "t", np.asarray(rp.s2s.getFinalTransformation())[:3, 3], "n_src", sc.shape[0])
            rp._s2s_T = np.asarray(rp.s2s.getFinalTransformation(), dtype=np.float64)
            rp._T_s2s = rp.T_prev @ rp.s2s.getFinalTransformation()
            rp.s2m.source_covs_ = rp.s2s.source_covs_
            rp.s2s.swapSourceAndTarget()
            sel = tuple(rp.selector.select([kf[0] for kf in rp.keyframes], rp._T_s2s[:3, 3])[0])
            if sel != rp.prev_set:
                rp._set_submap(sel); rp.prev_set = sel
        print("scan", i, "voxel outputs identical", np.array_equal(sc_c.view(np.uint32), sc_g.view(np.uint32)), "T_prev identical", np.array_equal(np.asarray(rc.T_prev), np.asarray(rg.T_prev)),
              "s2s final max diff", float(np.abs(rc._s2s_T - rg._s2s_T).max()))
        Lc, Lg = rc._L0, rg._L0
        print(" S2S linearize at identity: err cpu %.12g gpu %.12g; H rel %.3e b rel %.3e; corr differing %d; matched %d/%d" % (
            Lc["err"], Lg["err"], np.abs(Lc["H"] - Lg["H"]).max() / np.abs(Lc["H"]).max(), np.abs(Lc["b"] - Lg["b"]).max() / np.abs(Lc["b"]).max(),
            int((Lc["corr"] != Lg["corr"]).sum()), int((Lc["corr"] >= 0).sum()), int((Lg["corr"] >= 0).sum())))
        print(" src covs: shapes", rc._src_covs.shape, rg._src_covs.shape, "max abs diff", float(np.abs(rc._src_covs - rg._src_covs).max()) if rc._src_covs.shape == rg._src_covs.shape else None)
        print(" tgt covs: shapes", rc._tgt_covs.shape, rg._tgt_covs.shape, "max abs diff", float(np.abs(rc._tgt_covs - rg._tgt_covs).max()) if rc._tgt_covs.shape == rg._tgt_covs.shape else None)
        if rc._tgt_covs.shape == rg._tgt_covs.shape:
            dd = np.abs(rc._tgt_covs - rg._tgt_covs).reshape(rc._tgt_covs.shape[0], -1).max(axis=1)
            print(" tgt covs: points with diff > 1e-6:", int((dd > 1e-6).sum()), "of", dd.size)
        dsc = np.abs(rc._src_covs - rg._src_covs).reshape(rc._src_covs.shape[0], -1).max(axis=1)
        worst = np.argsort(-dsc)[:3]
        print(" src covs: points with diff > 1e-6:", int((dsc > 1e-6).sum()))
        np.set_printoptions(precision=17, linewidth=200)
        for j in worst[:2]:
            print("  src point", int(j), sc_c[j, :3], "diff", dsc[j])
            print("   cpu cov\n", rc._src_covs[j][:3, :3], "\n   eig", np.linalg.eigvalsh(rc._src_covs[j][:3, :3]))
            print("   gpu cov\n", rg._src_covs[j][:3, :3], "\n   eig", np.linalg.eigvalsh(rg._src_covs[j][:3, :3]))
            idx, d2 = vox.knn if False else (None, None)
            cc = O.Cloud(sc_c)
            ki, kd = cc.knn(sc_c[j:j + 1, :3].copy(), 10)
            nbp = sc_c[ki[0], :3].astype(np.float64)
            print("   neighbours (cpu kNN) d2", kd[0])
            print(nbp)
            cm = nbp - nbp.mean(0)
            print("   raw cov eig", np.linalg.eigvalsh(cm.T @ cm / 10.0))
        bad = np.nonzero(Lc["corr"] != Lg["corr"])[0][:5]
        for j in bad:
            print("  point", int(j), "corr cpu/gpu", int(Lc["corr"][j]), int(Lg["corr"][j]), "sqd cpu/gpu", float(Lc["sqd"][j]), float(Lg["sqd"][j]))
        print("scan", i, "T_s2s max diff", float(np.abs(np.asarray(rc._T_s2s, np.float64) - np.asarray(rg._T_s2s, np.float64)).max()))
        pc, pg = probe("cpu", rc, False), probe("gpu", rg, True)
        print(" err cpu %.12g gpu %.12g rel %.3e" % (pc["err"], pg["err"], abs(pc["err"] - pg["err"]) / abs(pc["err"])))
        print(" H rel diff %.3e  b rel diff %.3e" % (np.abs(pc["H"] - pg["H"]).max() / np.abs(pc["H"]).max(), np.abs(pc["b"] - pg["b"]).max() / np.abs(pc["b"]).max()))
        if pc["corr"] is not None and pg["corr"] is not None:
            print(" corr differing:", int((pc["corr"] != pg["corr"]).sum()), "of", pc["corr"].size, " matched cpu/gpu", int((pc["corr"] >= 0).sum()), int((pg["corr"] >= 0).sum()))
        for rp, is_gpu in ((rc, False), (rg, True)):
            rp.s2m.align(rp._T_s2s)
            r = rp.s2m.result if is_gpu else rp.s2m._r
            print("  ", "gpu" if is_gpu else "cpu", "iters", r.nr_iterations, "lin", r.n_linearize, "err-evals", r.n_compute_error, "conv", r.converged)
        break
    rc.step(sc_c)
    rg.step(sc_g, force_T=rc.T)
