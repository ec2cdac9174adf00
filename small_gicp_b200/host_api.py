"""ctypes front-end of libsgicp_b200_host.so: the C++ host mirror of the reference's template surface
(Registration<Factor, ParallelReductionCUDA, ...>::align, KdTree<PointCloud>) instantiated behind one C entry point.
Used by the tests; C++ users include small_gicp_b200/host/include/small_gicp_b200/*.hpp directly."""
import ctypes as C
import os

import numpy as np

from . import capi

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_LIB = None

TREE_HOST_KDTREE, TREE_DEVICE_KDTREE, TREE_VOXELMAP = 0, 1, 2
OPT_GN, OPT_LM = 0, 1


def library_path():
    return os.path.join(os.path.dirname(capi.library_path()), "libsgicp_b200_host.so")


def _lib():
    global _LIB
    if _LIB is None:
        capi._lib()  # libsgicp_b200.so first (the host library links against it)
        path = library_path()
        if not os.path.exists(path):
            raise capi.SgbError(f"{path} is missing: build it with __graft_entry__.build()")
        L = C.CDLL(path)
        L.sgbh_last_error.restype = C.c_char_p
        L.sgbh_align.restype = C.c_int
        L.sgbh_align.argtypes = [C.c_size_t, _dp, _dp, _dp, C.c_size_t, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.sgbh_kdtree_knn.restype = C.c_int
        L.sgbh_kdtree_knn.argtypes = [C.c_size_t, _dp, C.c_size_t, _dp, C.c_int, _u64p, _dp]
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Result:
    pass


def align(
    target_points,
    source_points,
    target_normals=None,
    target_covs=None,
    source_covs=None,
    init_T=None,
    factor=capi.FACTOR_GICP,
    robust=capi.ROBUST_NONE,
    robust_c=1.0,
    rejector=capi.REJECT_DISTANCE,
    max_dist_sq=1.0,
    optimizer=OPT_LM,
    max_iterations=20,
    rotation_eps=0.1 * np.pi / 180.0,
    translation_eps=1e-3,
    tree=TREE_HOST_KDTREE,
    voxel_resolution=1.0,
    voxel_search_offsets=1,
    dof_mask=None,
    device=0,
):
    """Registration<Factor, ParallelReductionCUDA, GeneralFactor, Rejector, Optimizer>::align (C++ host mirror)."""
    L = _lib()
    tp, sp = capi._points4(target_points), capi._points4(source_points)
    tn, tc, sc = _f64(target_normals), _f64(target_covs), _f64(source_covs)
    opts = np.zeros(21)
    opts[:13] = [factor, robust, robust_c, rejector, max_dist_sq, optimizer, max_iterations, rotation_eps, translation_eps, tree, voxel_resolution, voxel_search_offsets, 0 if dof_mask is None else 1]
    opts[13:19] = 1.0 if dof_mask is None else np.asarray(dof_mask, dtype=float)
    opts[19] = device
    T0 = np.ascontiguousarray((np.eye(4) if init_T is None else np.asarray(init_T, dtype=float)).T)
    T = np.empty(16)
    sc4 = np.empty(4)
    H = np.empty(36)
    b = np.empty(6)
    rc = L.sgbh_align(tp.shape[0], _d(tp), _d(tn), _d(tc), sp.shape[0], _d(sp), _d(sc), _d(opts), _d(T0), _d(T), _d(sc4), _d(H), _d(b))
    if rc != 0:
        raise capi.SgbError(L.sgbh_last_error().decode())
    r = Result()
    r.T_target_source = T.reshape(4, 4).T.copy()
    r.converged = bool(sc4[0])
    r.iterations = int(sc4[1])
    r.num_inliers = int(sc4[2])
    r.error = float(sc4[3])
    r.H = H.reshape(6, 6).copy()
    r.b = b.copy()
    return r


ICP, PLANE_ICP, GICP, VGICP = 0, 1, 2, 3  # RegistrationSetting::RegistrationType


def helper_align(target_points, source_points, init_T=None, type=GICP, voxel_resolution=1.0, downsampling_resolution=0.25, max_correspondence_distance=1.0,
                 rotation_eps=0.1 * np.pi / 180.0, translation_eps=1e-3, max_iterations=20, device=0):
    """small_gicp::align(target, source, init_T, RegistrationSetting) on raw points (registration_helper.cpp:58-69):
    device voxel-grid sampling (0.25 m), k = 10 normals + covariances, kd-tree, LM registration with the CUDA reduction."""
    L = _lib()
    if not hasattr(L.sgbh_helper_align, "_typed"):
        L.sgbh_helper_align.restype = C.c_int
        L.sgbh_helper_align.argtypes = [C.c_size_t, _dp, C.c_size_t, _dp, _dp, _dp, _dp, _dp, _dp]
        L.sgbh_helper_align._typed = True
    tp, sp = capi._points4(target_points), capi._points4(source_points)
    s = np.array([type, voxel_resolution, downsampling_resolution, max_correspondence_distance, rotation_eps, translation_eps, max_iterations, device], dtype=np.float64)
    T0 = np.ascontiguousarray((np.eye(4) if init_T is None else np.asarray(init_T, dtype=float)).T)
    T, sc4, sizes = np.empty(16), np.empty(4), np.empty(2)
    rc = L.sgbh_helper_align(tp.shape[0], _d(tp), sp.shape[0], _d(sp), _d(s), _d(T0), _d(T), _d(sc4), _d(sizes))
    if rc != 0:
        raise capi.SgbError(L.sgbh_last_error().decode())
    r = Result()
    r.T_target_source = T.reshape(4, 4).T.copy()
    r.converged, r.iterations, r.num_inliers, r.error = bool(sc4[0]), int(sc4[1]), int(sc4[2]), float(sc4[3])
    r.target_size, r.source_size = int(sizes[0]), int(sizes[1])
    return r


def kdtree_knn(points, queries, k):
    """KdTree<PointCloud>(points).knn_search for each query (host mirror, CPU)."""
    L = _lib()
    p, q = capi._points4(points), capi._points4(queries)
    idx = np.empty((q.shape[0], k), dtype=np.uint64)
    d2 = np.empty((q.shape[0], k))
    rc = L.sgbh_kdtree_knn(p.shape[0], _d(p), q.shape[0], _d(q), k, idx.ctypes.data_as(_u64p), _d(d2))
    if rc != 0:
        raise capi.SgbError(L.sgbh_last_error().decode())
    return idx, d2


def read_points_cpp(filename, kind="ply"):
    """read_ply / read_points of the C++ host mirror (small_gicp_b200/host/include/small_gicp_b200/read_points.hpp) -> (N, 4) float32."""
    L = _lib()
    n = C.c_size_t(0)
    k = 0 if kind == "ply" else 1
    fn = str(filename).encode()
    L.sgbh_read_points.restype = C.c_int
    L.sgbh_read_points.argtypes = [C.c_char_p, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
    if L.sgbh_read_points(fn, k, 0, None, C.byref(n)) != 0:
        raise capi.SgbError(L.sgbh_last_error().decode())
    out = np.empty((n.value, 4), dtype=np.float32)
    if n.value and L.sgbh_read_points(fn, k, n.value, out.ctypes.data_as(C.c_void_p), C.byref(n)) != 0:
        raise capi.SgbError(L.sgbh_last_error().decode())
    return out


def write_points_cpp(filename, points):
    """write_points of the C++ host mirror: (N, 4) float32 -> KITTI-style .bin"""
    L = _lib()
    p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 4)
    L.sgbh_write_points.restype = C.c_int
    L.sgbh_write_points.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    if L.sgbh_write_points(str(filename).encode(), p.shape[0], p.ctypes.data_as(C.c_void_p)) != 0:
        raise capi.SgbError(L.sgbh_last_error().decode())
