"""ctypes driver for the CPU oracle (oracle/gicp_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under direct_lidar_odometry_b200/ imports this.

Two builds of the same source exist:
  oracle/liboracle.so            restated kd-tree (always buildable)
  oracle/_ref/liboracle_ref.so   kNN through the reference's own vendored nanoflann header
                                 (compiled here from /root/reference; git-ignored, travels to the GPU box)
`load(prefer_ref=True)` returns the _ref build when present.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BACKEND_DEFAULT, BACKEND_RESTATED, BACKEND_REF, BACKEND_BRUTE = 0, 1, 2, 3
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)


class AlignResult(C.Structure):
    _fields_ = [("final_transformation", C.c_float * 16), ("final_x", C.c_double * 16),
                ("final_hessian", C.c_double * 36), ("lm_lambda", C.c_double), ("last_error", C.c_double),
                ("nr_iterations", C.c_int), ("converged", C.c_int), ("n_linearize", C.c_int),
                ("n_compute_error", C.c_int), ("lm_failed", C.c_int), ("reserved", C.c_int)]

    def T(self) -> np.ndarray:
        return np.array(self.final_transformation, dtype=np.float32).reshape(4, 4).T.copy()

    def Tx(self) -> np.ndarray:
        return np.array(self.final_x, dtype=np.float64).reshape(4, 4).T.copy()

    def H(self) -> np.ndarray:
        return np.array(self.final_hessian, dtype=np.float64).reshape(6, 6).T.copy()


_LIBS: dict = {}


def lib_path(ref: bool) -> str:
    return os.path.join(_HERE, "_ref", "liboracle_ref.so") if ref else os.path.join(_HERE, "liboracle.so")


def load(prefer_ref: bool = True):
    want_ref = prefer_ref and os.path.exists(lib_path(True))
    key = "ref" if want_ref else "plain"
    if key in _LIBS:
        return _LIBS[key]
    path = lib_path(want_ref)
    if not os.path.exists(path):
        raise RuntimeError(f"oracle library missing: {path} (run `make -C oracle` or __graft_entry__.build())")
    L = C.CDLL(path)
    fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.orc_has_ref_nanoflann.restype = C.c_int
    L.orc_max_threads.restype = C.c_int
    L.orc_voxel_filter.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_float, fp, C.POINTER(C.c_size_t), ip]
    L.orc_cloud_create.restype = C.c_void_p
    L.orc_cloud_create.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
    L.orc_cloud_destroy.argtypes = [C.c_void_p]
    L.orc_cloud_knn.argtypes = [C.c_void_p, fp, C.c_size_t, C.c_size_t, C.c_int, ip, fp, C.c_int]
    L.orc_cloud_covariances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, ip, fp]
    L.orc_gicp_create.restype = C.c_void_p
    L.orc_gicp_destroy.argtypes = [C.c_void_p]
    L.orc_gicp_set_params.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.c_double, C.c_int, C.c_int, C.c_int]
    for n in ("orc_gicp_set_source", "orc_gicp_set_target"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
    for n in ("orc_gicp_set_source_covs", "orc_gicp_set_target_covs"):
        getattr(L, n).argtypes = [C.c_void_p, dp, C.c_size_t]
    for n in ("orc_gicp_calc_source_covs", "orc_gicp_calc_target_covs"):
        getattr(L, n).argtypes = [C.c_void_p]
    for n in ("orc_gicp_get_source_covs", "orc_gicp_get_target_covs"):
        getattr(L, n).argtypes = [C.c_void_p, dp]
        getattr(L, n).restype = C.c_size_t
    L.orc_gicp_swap.argtypes = [C.c_void_p]
    L.orc_gicp_linearize.argtypes = [C.c_void_p, dp, dp, dp, dp, ip, fp, dp]
    L.orc_gicp_compute_error.argtypes = [C.c_void_p, dp, dp]
    L.orc_gicp_align.argtypes = [C.c_void_p, fp, C.POINTER(AlignResult)]
    L.orc_svd3.argtypes = [dp, dp, dp, dp]
    L.orc_ldlt6_solve.argtypes = [dp, dp, dp]
    L.orc_so3_exp.argtypes = [dp, dp]
    L.orc_inverse4.argtypes = [dp, dp]
    _LIBS[key] = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def as_xyzi(pts: np.ndarray) -> np.ndarray:
    """(n,3|4|8) -> contiguous (n,8) float32 PointXYZI records."""
    pts = np.asarray(pts, dtype=np.float32)
    if pts.ndim == 2 and pts.shape[1] == 8:
        return np.ascontiguousarray(pts)
    out = np.zeros((pts.shape[0], 8), dtype=np.float32)
    out[:, :3] = pts[:, :3]
    out[:, 3] = 1.0
    if pts.shape[1] >= 4:
        out[:, 4] = pts[:, 3]
    return out


def from_ros_msg(msg) -> np.ndarray:
    """pcl::fromROSMsg(msg, pcl::PointCloud<pcl::PointXYZI>) restated (reference src/dlo/odom.cc:636-637; PCL
    conversions.h: createMapping + fromPCLPointCloud2): every PointXYZI member (x, y, z, intensity — all FLOAT32) is
    copied from the message field of the same name, the same datatype and count 1; a member without such a field
    keeps the default-constructed value (0; data[3] = 1).  Point (row, col) starts at row * row_step + col *
    point_step.  Byte order is taken as the host's (PCL ignores is_bigendian).  Returns (n, 8) float32 records
    {x, y, z, 1, intensity, 0, 0, 0}."""
    n = int(msg.width) * int(msg.height)
    out = np.zeros((n, 8), dtype=np.float32)
    out[:, 3] = 1.0
    raw = np.frombuffer(msg.data, dtype=np.uint8)
    member = {"x": 0, "y": 1, "z": 2, "intensity": 4}
    done = set()
    for f in msg.fields:
        if f.name not in member or f.name in done or int(f.datatype) != 7 or int(f.count) not in (0, 1):
            continue
        done.add(f.name)
        col = np.empty(n, dtype=np.float32)
        for r in range(int(msg.height)):
            base = r * int(msg.row_step) + int(f.offset)
            idx = base + np.arange(int(msg.width), dtype=np.int64)[:, None] * int(msg.point_step) + np.arange(4)[None, :]
            col[r * int(msg.width):(r + 1) * int(msg.width)] = raw[idx].copy().view("<f4").reshape(-1)
        out[:, member[f.name]] = col
    return out


def integrate_imu(stamps, ang_vel, prev_frame_stamp, curr_frame_stamp) -> np.ndarray:
    """Restatement of OdomNode::integrateIMU (reference src/dlo/odom.cc:859-919) with numpy scalars of the reference's
    types: Eigen::Quaternionf state (float32), double gyro samples and time steps, first-order update, normalisation by a
    double norm, Quaternionf::toRotationMatrix in float32.  Returns imu_SE3 (4x4 float32)."""
    f32, f64 = np.float32, np.float64
    stamps = np.asarray(stamps, dtype=f64)
    ang_vel = np.asarray(ang_vel, dtype=f64).reshape(-1, 3)
    sel = [i for i in range(stamps.shape[0]) if curr_frame_stamp - stamps[i] >= 0.0 and prev_frame_stamp - stamps[i] <= 0.0]
    sel.sort(key=lambda i: stamps[i])
    q = [f32(1), f32(0), f32(0), f32(0)]   # w x y z
    prev = f64(0.0)
    for i in sel:
        if prev == 0.0:
            prev = stamps[i]
            continue
        dt = stamps[i] - prev
        prev = stamps[i]
        w, x, y, z = q
        ax, ay, az = ang_vel[i]
        q[0] = f32(f64(w) - f64(0.5) * (f64(x) * ax + f64(y) * ay + f64(z) * az) * dt)
        q[1] = f32(f64(x) + f64(0.5) * (f64(w) * ax - f64(z) * ay + f64(y) * az) * dt)
        q[2] = f32(f64(y) + f64(0.5) * (f64(z) * ax + f64(w) * ay - f64(x) * az) * dt)
        q[3] = f32(f64(z) + f64(0.5) * (f64(x) * ay - f64(y) * ax + f64(w) * az) * dt)
    w, x, y, z = q
    norm = np.sqrt(f64(f32(f32(f32(w * w) + f32(x * x)) + f32(y * y)) + f32(z * z)))
    w, x, y, z = [f32(f64(v) / norm) for v in (w, x, y, z)]
    tx, ty, tz = f32(2) * x, f32(2) * y, f32(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    T = np.eye(4, dtype=f32)
    T[:3, :3] = [[f32(1) - (tyy + tzz), txy - twz, txz + twy], [txy + twz, f32(1) - (txx + tzz), tyz - twx], [txz - twy, tyz + twx, f32(1) - (txx + tyy)]]
    return T


def preprocess_points(pts: np.ndarray, crop_size, leaf: float, lib=None):
    """OdomNode::preprocessPoints restated (reference src/dlo/odom.cc:443-465): removeNaNFromPointCloud (:451), the
    negative pcl::CropBox of +-crop_size (:122-124,454-457; PCL keeps a point as "inside" when min <= p <= max on all
    axes and setNegative(true) drops those), then the scan voxel grid (:460-463) unless leaf <= 0."""
    pts = as_xyzi(pts)
    keep = np.isfinite(pts[:, :3]).all(axis=1)
    if crop_size is not None:
        inside = ((pts[:, :3] >= -np.float32(crop_size)) & (pts[:, :3] <= np.float32(crop_size))).all(axis=1)
        keep &= ~inside
    kept = np.ascontiguousarray(pts[keep])
    if leaf <= 0:
        out = kept.copy()
        out[:, 3] = 1.0
        out[:, 5:] = 0.0
        return out
    return voxel_filter(kept, leaf, lib=lib)


def voxel_filter(pts: np.ndarray, leaf: float, lib=None, return_assignment: bool = False):
    L = lib or load()
    pts = as_xyzi(pts)
    n = pts.shape[0]
    out = np.zeros((max(n, 1), 8), dtype=np.float32)
    m = C.c_size_t(0)
    assign = np.zeros(max(n, 1), dtype=np.int32)
    rc = L.orc_voxel_filter(_f(pts), n, 8, C.c_float(leaf), _f(out), C.byref(m), _i(assign))
    res = out[: m.value].copy()
    if return_assignment:
        return res, assign[:n].copy(), rc
    return res


class Cloud:
    def __init__(self, pts: np.ndarray, backend: int = BACKEND_DEFAULT, build_index: bool = True, lib=None):
        self.L = lib or load()
        self.pts = as_xyzi(pts)
        self.n = self.pts.shape[0]
        self.h = self.L.orc_cloud_create(_f(self.pts), self.n, 8, backend, 1 if build_index else 0)
        if not self.h:
            raise RuntimeError("oracle backend unavailable in this build")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_cloud_destroy(self.h)
            self.h = None

    def knn(self, q: np.ndarray, k: int, nthreads: int = 0):
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float32))
        nq, st = q.shape
        idx = np.zeros((nq, k), dtype=np.int32)
        d2 = np.zeros((nq, k), dtype=np.float32)
        rc = self.L.orc_cloud_knn(self.h, _f(q), nq, st, k, _i(idx), _f(d2), nthreads or self.L.orc_max_threads())
        if rc:
            raise RuntimeError(f"orc_cloud_knn rc={rc}")
        return idx, d2

    def covariances(self, k: int, method: int = REG_PLANE, nthreads: int = 0, with_knn: bool = False):
        covs = np.zeros((self.n, 16), dtype=np.float64)
        idx = np.zeros((self.n, k), dtype=np.int32) if with_knn else None
        d2 = np.zeros((self.n, k), dtype=np.float32) if with_knn else None
        rc = self.L.orc_cloud_covariances(self.h, k, method, nthreads or self.L.orc_max_threads(), _d(covs),
                                          _i(idx) if with_knn else None, _f(d2) if with_knn else None)
        if rc:
            raise RuntimeError(f"orc_cloud_covariances rc={rc}")
        covs = covs.reshape(self.n, 4, 4).transpose(0, 2, 1).copy()  # col-major -> [r,c]
        return (covs, idx, d2) if with_knn else covs


class Gicp:
    """Oracle NanoGICP (reference include/nano_gicp/nano_gicp.hpp:58-137)."""

    def __init__(self, lib=None, **params):
        self.L = lib or load()
        self.h = self.L.orc_gicp_create()
        self.p = dict(k=20, max_corr_dist=float(np.finfo(np.float32).max), max_iter=64, trans_eps=5e-4, rot_eps=2e-3,
                      lm_max_iter=10, lm_init_lambda_factor=1e-9, reg_method=REG_PLANE, optimizer=1, num_threads=0)
        self.set_params(**params)
        self._src = self._tgt = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_gicp_destroy(self.h)
            self.h = None

    def set_params(self, **kw):
        self.p.update(kw)
        p = self.p
        self.L.orc_gicp_set_params(self.h, p["k"], p["max_corr_dist"], p["max_iter"], p["trans_eps"], p["rot_eps"],
                                   p["lm_max_iter"], p["lm_init_lambda_factor"], p["reg_method"], p["optimizer"],
                                   p["num_threads"])

    def set_source(self, cloud: Cloud):
        self._src = cloud
        self.L.orc_gicp_set_source(self.h, cloud.h)

    def set_target(self, cloud: Cloud):
        self._tgt = cloud
        self.L.orc_gicp_set_target(self.h, cloud.h)

    @staticmethod
    def _covs_in(covs):
        covs = np.asarray(covs, dtype=np.float64)
        return np.ascontiguousarray(covs.transpose(0, 2, 1)).reshape(-1, 16)

    def set_source_covs(self, covs):
        c = self._covs_in(covs)
        self.L.orc_gicp_set_source_covs(self.h, _d(c), c.shape[0])

    def set_target_covs(self, covs):
        c = self._covs_in(covs)
        self.L.orc_gicp_set_target_covs(self.h, _d(c), c.shape[0])

    def calc_source_covs(self):
        rc = self.L.orc_gicp_calc_source_covs(self.h)
        if rc:
            raise RuntimeError(f"rc={rc}")

    def calc_target_covs(self):
        rc = self.L.orc_gicp_calc_target_covs(self.h)
        if rc:
            raise RuntimeError(f"rc={rc}")

    def _get(self, fn):
        n = fn(self.h, None)
        out = np.zeros((n, 16), dtype=np.float64)
        fn(self.h, _d(out))
        return out.reshape(n, 4, 4).transpose(0, 2, 1).copy()

    def get_source_covs(self):
        return self._get(self.L.orc_gicp_get_source_covs)

    def get_target_covs(self):
        return self._get(self.L.orc_gicp_get_target_covs)

    def swap(self):
        self._src, self._tgt = self._tgt, self._src
        self.L.orc_gicp_swap(self.h)

    def linearize(self, T: np.ndarray, per_point: bool = False):
        Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16)
        H = np.zeros(36)
        b = np.zeros(6)
        e = C.c_double(0)
        n = self._src.n
        corr = np.zeros(n, dtype=np.int32)
        sqd = np.zeros(n, dtype=np.float32)
        mah = np.zeros((n, 16)) if per_point else None
        rc = self.L.orc_gicp_linearize(self.h, _d(Tc), _d(H), _d(b), C.byref(e), _i(corr), _f(sqd),
                                       _d(mah) if per_point else None)
        if rc:
            raise RuntimeError(f"orc_gicp_linearize rc={rc}")
        out = dict(H=H.reshape(6, 6).T.copy(), b=b, err=e.value, corr=corr, sqd=sqd)
        if per_point:
            out["mahalanobis"] = mah.reshape(n, 4, 4).transpose(0, 2, 1).copy()
        return out

    def compute_error(self, T: np.ndarray) -> float:
        Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16)
        e = C.c_double(0)
        rc = self.L.orc_gicp_compute_error(self.h, _d(Tc), C.byref(e))
        if rc:
            raise RuntimeError(f"rc={rc}")
        return e.value

    def align(self, guess: np.ndarray | None = None) -> AlignResult:
        g = np.eye(4, dtype=np.float32) if guess is None else np.asarray(guess, dtype=np.float32)
        gc = np.ascontiguousarray(g.T).reshape(16)
        res = AlignResult()
        rc = self.L.orc_gicp_align(self.h, _f(gc), C.byref(res))
        if rc:
            raise RuntimeError(f"orc_gicp_align rc={rc}")
        return res
