"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The shipped GPU path (better_fastlio2_b200/) never imports this module.

Two map back ends with the same method names:
  * RefIkdTree  -> oracle/_ref/libikd_ref.so : the REFERENCE's own ikd-Tree compiled unmodified ("reference").
  * PortMap     -> oracle/liblio_oracle.so   : our logical restatement of it ("port").
and OracleLIO: the restated h_share_model / ESIKF update / map_incremental / fov segment (lio_oracle.cpp).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIO = os.path.join(_HERE, "liblio_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libikd_ref.so")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

KNN_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)


def build(force=False):
    """Compile the oracle (and, when /root/reference is present, oracle/_ref). Building the checker is not using it."""
    if force or not os.path.exists(_LIO) or (os.path.isdir("/root/reference") and not os.path.exists(_REF)):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def have_ref():
    return os.path.exists(_REF)


_lio = None
_ref = None


def lio():
    global _lio
    if _lio is None:
        build()
        L = C.CDLL(_LIO)
        L.orc_esti_plane.argtypes = [f32p, C.c_float, f32p]
        L.orc_esti_plane.restype = C.c_int
        L.orc_transform.argtypes = [f64p, f32p, C.c_int, f32p]
        L.orc_residual_pass.argtypes = [f64p, f32p, f32p, C.c_int, f32p, f32p, i32p, C.c_int, u8p, C.c_int, f32p, f64p,
                                        f64p, f64p]
        L.orc_residual_pass.restype = C.c_int
        L.orc_esikf_update.argtypes = [f64p, f64p, C.c_double, C.c_int, f64p, f32p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, f32p, f32p, f32p, i32p, u8p, f32p, i32p, C.c_void_p]
        L.orc_map_incremental_classify.argtypes = [f64p, f32p, C.c_int, f32p, i32p, C.c_int, C.c_double, f32p, u8p]
        L.orc_fov_segment.argtypes = [f64p, C.c_double, C.c_float, f32p, i32p, f32p]
        L.orc_fov_segment.restype = C.c_int
        L.orc_boxplus.argtypes = [f64p, f64p]
        L.orc_boxminus.argtypes = [f64p, f64p, f64p]
        L.orc_A_matrix.argtypes = [f64p, f64p]
        L.orc_invert.argtypes = [f64p, f64p, C.c_int]
        L.orc_invert.restype = C.c_int
        L.orc_s2_mats.argtypes = [f64p, f64p, f64p, f64p, f64p]
        # front-end rows (frontend_oracle.cpp)
        L.orc_undistort.argtypes = [f32p, f32p, C.c_int, f64p, C.c_int, f64p, f32p, i32p]
        L.orc_voxel_grid.argtypes = [f32p, C.c_void_p, C.c_int, C.c_float, C.c_int, f32p, C.c_void_p]
        L.orc_voxel_grid.restype = C.c_int
        L.orc_rpy_matrix.argtypes = [f32p, f32p]
        L.orc_transform_cloud_rpy.argtypes = [f32p, C.c_int, f32p, f32p]
        L.orc_body_to_world4.argtypes = [f64p, f32p, C.c_int, f32p]
        for pre in ("mapport",):
            _bind_map(L, pre)
        L.mapport_create.argtypes = [C.c_float]
        L.mapport_create.restype = C.c_void_p
        _lio = L
    return _lio


def ref():
    global _ref
    if _ref is None:
        build()
        if not have_ref():
            raise RuntimeError("oracle/_ref/libikd_ref.so not built (reference tree absent at build time)")
        L = C.CDLL(_REF)
        _bind_map(L, "ikdref")
        L.ikdref_create.argtypes = [C.c_float, C.c_float, C.c_float]
        L.ikdref_create.restype = C.c_void_p
        L.ikdref_set_threads.argtypes = [C.c_int]
        L.ikdref_has_root.argtypes = [C.c_void_p]
        L.ikdref_has_root.restype = C.c_int
        L.ikdref_nearest_md.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_double, f32p, f32p, i32p]
        L.ikdref_delete_points.argtypes = [C.c_void_p, f32p, C.c_int]
        L.ikdref_build_i.argtypes = [C.c_void_p, f32p, C.c_int]
        L.ikdref_add_points_i.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int]
        L.ikdref_add_points_i.restype = C.c_int
        L.ikdref_flatten_i.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ikdref_flatten_i.restype = C.c_int
        L.ikdref_nearest_i.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, f32p, f32p, i32p]
        _ref = L
    return _ref


def _bind_map(L, pre):
    g = lambda n: getattr(L, pre + "_" + n)
    g("destroy").argtypes = [C.c_void_p]
    g("set_downsample").argtypes = [C.c_void_p, C.c_float]
    g("build").argtypes = [C.c_void_p, f32p, C.c_int]
    g("reconstruct").argtypes = [C.c_void_p, f32p, C.c_int]
    g("nearest").argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, f32p, f32p, i32p, C.c_int]
    g("add_points").argtypes = [C.c_void_p, f32p, C.c_int, C.c_int]
    g("add_points").restype = C.c_int
    g("delete_boxes").argtypes = [C.c_void_p, f32p, C.c_int]
    g("delete_boxes").restype = C.c_int
    g("size").argtypes = [C.c_void_p]
    g("size").restype = C.c_int
    g("validnum").argtypes = [C.c_void_p]
    g("validnum").restype = C.c_int
    g("flatten").argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    g("flatten").restype = C.c_int


def _xyz(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


class _MapBase:
    _pre = None
    _L = None

    def _f(self, n):
        return getattr(self._L, self._pre + "_" + n)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_downsample_param(self, ds):
        self._f("set_downsample")(self.h, float(ds))

    def Build(self, pts):
        pts = _xyz(pts)
        self._f("build")(self.h, pts, len(pts))

    def reconstruct(self, pts):
        pts = _xyz(pts)
        self._f("reconstruct")(self.h, pts, len(pts))

    def Nearest_Search(self, q, k=5, threads=0):
        q = _xyz(q)
        n = len(q)
        xyz = np.empty((n, k, 3), np.float32)
        d2 = np.empty((n, k), np.float32)
        cnt = np.empty(n, np.int32)
        self._f("nearest")(self.h, q, n, k, xyz.reshape(-1), d2.reshape(-1), cnt, threads)
        return xyz, d2, cnt

    def Add_Points(self, pts, downsample_on):
        pts = _xyz(pts)
        if len(pts) == 0:
            return 0
        return self._f("add_points")(self.h, pts, len(pts), 1 if downsample_on else 0)

    def Delete_Point_Boxes(self, boxes):
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        if len(b) == 0:
            return 0
        return self._f("delete_boxes")(self.h, b.reshape(-1), len(b))

    def size(self):
        return self._f("size")(self.h)

    def validnum(self):
        return self._f("validnum")(self.h)

    def flatten(self):
        n = self._f("flatten")(self.h, None, 0)
        out = np.empty((max(n, 1), 3), np.float32)
        n2 = self._f("flatten")(self.h, out.ctypes.data_as(C.c_void_p), n)
        return out[:n2].copy()

    @property
    def knn5_ptr(self):
        return C.cast(self._f("knn5"), C.c_void_p)


class RefIkdTree(_MapBase):
    """The reference's own KD_TREE<PointXYZINormal> (ikd_Tree.h:225-249), default ctor args of laserMapping.cpp:116."""
    _pre = "ikdref"
    kind = "reference"

    def __init__(self, ds=0.2, delete_param=0.5, balance_param=0.6, threads=0):
        self._L = ref()
        self.h = self._L.ikdref_create(delete_param, balance_param, ds)
        self._L.ikdref_set_threads(threads)

    def set_threads(self, t):
        self._L.ikdref_set_threads(int(t))

    def Nearest_Search_md(self, q, k, max_dist):
        q = _xyz(q)
        n = len(q)
        xyz = np.empty((n, k, 3), np.float32)
        d2 = np.empty((n, k), np.float32)
        cnt = np.empty(n, np.int32)
        self._L.ikdref_nearest_md(self.h, q, n, k, float(max_dist), xyz.reshape(-1), d2.reshape(-1), cnt)
        return xyz, d2, cnt

    def Delete_Points(self, pts):
        pts = _xyz(pts)
        self._L.ikdref_delete_points(self.h, pts, len(pts))

    # whole-PointType behaviour of the reference tree: x, y, z, intensity records in and out
    def Build_xyzi(self, pts4):
        a = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
        self._L.ikdref_build_i(self.h, a.reshape(-1), len(a))

    def Add_Points_xyzi(self, pts4, downsample_on):
        a = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
        return self._L.ikdref_add_points_i(self.h, a.reshape(-1), len(a), 1 if downsample_on else 0)

    def flatten_xyzi(self):
        n = self._L.ikdref_flatten_i(self.h, None, 0)
        out = np.empty((max(n, 1), 4), np.float32)
        n2 = self._L.ikdref_flatten_i(self.h, out.ctypes.data_as(C.c_void_p), n)
        return out[:min(n, n2)].copy()

    def Nearest_Search_xyzi(self, q, k=5):
        q = _xyz(q)
        n = len(q)
        out = np.empty((n, k, 4), np.float32)
        d2 = np.empty((n, k), np.float32)
        cnt = np.empty(n, np.int32)
        self._L.ikdref_nearest_i(self.h, q.reshape(-1), n, k, out.reshape(-1), d2.reshape(-1), cnt)
        return out, d2, cnt


class PortMap(_MapBase):
    _pre = "mapport"
    kind = "port"

    def __init__(self, ds=0.2, **_):
        self._L = lio()
        self.h = self._L.mapport_create(ds)

    def set_threads(self, t):
        pass


def make_map(ds=0.2, prefer_ref=True, threads=0):
    if prefer_ref and have_ref():
        return RefIkdTree(ds=ds, threads=threads)
    return PortMap(ds=ds)


class ScanScratch:
    """Per-scan persistent buffers of laserMapping.cpp (Nearest_Points, point_selected_surf, normvec, feats_down_world)."""

    def __init__(self, n):
        self.n = n
        self.world = np.zeros((n, 3), np.float32)
        self.nbr = np.full((n, 5, 3), np.nan, np.float32)
        self.nbr_d2 = np.full((n, 5), np.inf, np.float32)
        self.nbr_cnt = np.zeros(n, np.int32)
        self.sel = np.ones(n, np.uint8)  # memset(point_selected_surf, true), laserMapping.cpp:2131
        self.normvec = np.zeros((n, 4), np.float32)


def esikf_update(state26, P, body, map_obj, max_iter=3, R=0.001, extrinsic_est_en=False, limit=None, scratch=None,
                 want_trace=False):
    """update_iterated_dyn_share_modified (esekfom.hpp:1620) with h_share_model; returns (state, P, scratch, stats, trace)."""
    L = lio()
    body = _xyz(body)
    n = len(body)
    st = np.array(state26, np.float64).copy()
    Pm = np.ascontiguousarray(np.array(P, np.float64).reshape(23, 23)).copy()
    lim = np.full(23, 0.001) if limit is None else np.asarray(limit, np.float64)
    sc = scratch or ScanScratch(n)
    stats = np.zeros(4, np.int32)
    trace = np.zeros((max_iter + 1, 26), np.float64) if want_trace else None
    L.orc_esikf_update(st, Pm.reshape(-1), R, max_iter, lim, body, n, 1 if extrinsic_est_en else 0, map_obj.knn5_ptr,
                       map_obj.h, sc.world.reshape(-1), sc.nbr.reshape(-1), sc.nbr_d2.reshape(-1), sc.nbr_cnt, sc.sel,
                       sc.normvec.reshape(-1), stats, trace.ctypes.data_as(C.c_void_p) if want_trace else None)
    return st, Pm, sc, stats, trace


def residual_pass(state26, body, world, nbr, nbr_d2, nbr_cnt, search, sel, extrinsic_est_en=False):
    L = lio()
    n = len(body)
    normvec = np.zeros((n, 4), np.float32)
    hx = np.zeros((max(n, 1), 12), np.float64)
    h = np.zeros(max(n, 1), np.float64)
    tot = np.zeros(1, np.float64)
    M = L.orc_residual_pass(np.asarray(state26, np.float64), _xyz(body).reshape(-1), _xyz(world).reshape(-1), n,
                            np.ascontiguousarray(nbr, np.float32).reshape(-1),
                            np.ascontiguousarray(nbr_d2, np.float32).reshape(-1),
                            np.ascontiguousarray(nbr_cnt, np.int32), 1 if search else 0, sel,
                            1 if extrinsic_est_en else 0, normvec.reshape(-1), hx.reshape(-1), h, tot)
    return M, hx[:M], h[:M], normvec, float(tot[0])


def transform(state26, body):
    body = _xyz(body)
    out = np.empty_like(body)
    lio().orc_transform(np.asarray(state26, np.float64), body.reshape(-1), len(body), out.reshape(-1))
    return out


def map_incremental_classify(state26, body, nbr, nbr_cnt, flg_EKF_inited=True, filter_size_map_min=0.2):
    body = _xyz(body)
    n = len(body)
    world = np.empty((n, 3), np.float32)
    cls = np.zeros(n, np.uint8)
    lio().orc_map_incremental_classify(np.asarray(state26, np.float64), body.reshape(-1), n,
                                       np.ascontiguousarray(nbr, np.float32).reshape(-1),
                                       np.ascontiguousarray(nbr_cnt, np.int32), 1 if flg_EKF_inited else 0,
                                       float(filter_size_map_min), world.reshape(-1), cls)
    return world, cls


def map_incremental(state26, body, sc, map_obj, flg_EKF_inited=True, filter_size_map_min=0.2):
    """map_incremental (laserMapping.cpp:1440-1496): classify, then Add_Points(ToAdd,true), Add_Points(NoNeed,false)."""
    world, cls = map_incremental_classify(state26, body, sc.nbr, sc.nbr_cnt, flg_EKF_inited, filter_size_map_min)
    to_add = world[cls == 1]
    no_ds = world[cls == 2]
    map_obj.Add_Points(to_add, True)
    map_obj.Add_Points(no_ds, False)
    return len(to_add), len(no_ds)


class FovSegment:
    """lasermap_fov_segment state (laserMapping.cpp:1132-1200)."""

    def __init__(self, cube_len=200.0, det_range=100.0):
        self.cube_len = float(cube_len)
        self.det_range = float(det_range)
        self.local_map = np.zeros(6, np.float32)
        self.init = np.zeros(1, np.int32)

    def step(self, pos_lid):
        boxes = np.zeros(18, np.float32)
        nb = lio().orc_fov_segment(np.asarray(pos_lid, np.float64), self.cube_len, self.det_range, self.local_map,
                                   self.init, boxes)
        return boxes.reshape(3, 6)[:nb].copy()


def esti_plane(nn5x3, thr=0.1):
    out = np.zeros(4, np.float32)
    ok = lio().orc_esti_plane(np.ascontiguousarray(nn5x3, np.float32).reshape(-1), thr, out)
    return bool(ok), out


# ------------------------------------------------------------------------------------------------ front-end rows
def undistort(xyz, curvature, poses, state26):
    """UndistortPcl backward pass (IMU_Processing.hpp:241-243,334-386). Returns (xyz_sorted[n,3], perm[n])."""
    xyz = _xyz(xyz)
    cur = np.ascontiguousarray(curvature, dtype=np.float32)
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 22)
    out = np.empty_like(xyz)
    perm = np.empty(len(xyz), dtype=np.int32)
    lio().orc_undistort(xyz, cur, len(xyz), poses, len(poses), np.ascontiguousarray(state26, dtype=np.float64), out, perm)
    return out, perm


def voxel_grid(pts4, leaf, curvature=None, order="pcl"):
    """pcl::VoxelGrid centroid filter (PCL 1.10 restated). pts4 = [n,4] x,y,z,intensity. order 'pcl' | 'stable'.
    Returns (out4[m,4], out_curv[m] or None, overflow flag)."""
    pts4 = np.ascontiguousarray(pts4, dtype=np.float32)
    n = len(pts4)
    out = np.empty((max(n, 1), 4), dtype=np.float32)
    cur = None if curvature is None else np.ascontiguousarray(curvature, dtype=np.float32)
    oc = None if cur is None else np.empty(max(n, 1), dtype=np.float32)
    m = lio().orc_voxel_grid(pts4, None if cur is None else cur.ctypes.data, n, float(leaf), 0 if order == "pcl" else 1, out,
                             None if oc is None else oc.ctypes.data)
    ovf = m < 0
    m = n if ovf else m
    return out[:m].copy(), (None if oc is None else oc[:m].copy()), ovf


def rpy_matrix(pose6):
    t = np.empty(12, dtype=np.float32)
    lio().orc_rpy_matrix(np.ascontiguousarray(pose6, dtype=np.float32), t)
    return t.reshape(3, 4)


def transform_cloud_rpy(pts4, pose6):
    pts4 = np.ascontiguousarray(pts4, dtype=np.float32)
    out = np.empty_like(pts4)
    lio().orc_transform_cloud_rpy(pts4, len(pts4), np.ascontiguousarray(pose6, dtype=np.float32), out)
    return out


def body_to_world4(state26, pts4):
    pts4 = np.ascontiguousarray(pts4, dtype=np.float32)
    out = np.empty_like(pts4)
    lio().orc_body_to_world4(np.ascontiguousarray(state26, dtype=np.float64), pts4, len(pts4), out)
    return out
