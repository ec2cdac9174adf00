"""ctypes front-end of the CPU oracle (oracle/sgicp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (small_gicp_b200) must never import this.

All matrices are numpy float64, row-major (T is the usual 4x4 homogeneous matrix).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None
_SIGNATURES = {}

FACTOR_ICP, FACTOR_PLANE, FACTOR_GICP = 0, 1, 2
ROBUST_NONE, ROBUST_HUBER, ROBUST_CAUCHY = 0, 1, 2
REJECT_NONE, REJECT_DISTANCE = 0, 1
OPT_GN, OPT_LM = 0, 1
FEAT_NORMAL, FEAT_COV, FEAT_NORMAL_COV = 1, 2, 3
NO_INDEX = np.uint64(0xFFFFFFFFFFFFFFFF)

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)


def build(native=False):
    target = "native" if native else "all"
    subprocess.check_call(["make", "-s", "-C", _HERE, target])


def lib(native=False):
    global _LIB
    if _LIB is not None and not native:
        return _LIB
    name = "libsgicp_oracle_native.so" if native else "libsgicp_oracle.so"
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "sgicp_oracle.cpp")
    if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
        build(native)
    L = C.CDLL(path)
    vp = C.c_void_p
    sig = {
        "orc_max_threads": (C.c_int, []),
        "orc_cloud_create": (vp, [C.c_size_t, _dp, C.c_int]),
        "orc_cloud_destroy": (None, [vp]),
        "orc_cloud_size": (C.c_size_t, [vp]),
        "orc_cloud_get": (None, [vp, _dp, _dp, _dp]),
        "orc_cloud_set_features": (None, [vp, _dp, _dp]),
        "orc_cloud_transformed": (vp, [vp, _dp]),
        "orc_voxelgrid_sampling": (vp, [vp, C.c_double]),
        "orc_kdtree_create": (vp, [vp]),
        "orc_kdtree_destroy": (None, [vp]),
        "orc_kdtree_num_nodes": (C.c_size_t, [vp]),
        "orc_kdtree_export": (None, [vp, vp, _u64p]),
        "orc_kdtree_knn": (None, [vp, C.c_size_t, _dp, C.c_int, _u64p, _dp, _u64p, C.c_int]),
        "orc_estimate_features": (None, [vp, vp, C.c_int, C.c_int, C.c_int]),
        "orc_voxelmap_create": (vp, [vp, C.c_double, C.c_int]),
        "orc_voxelmap_destroy": (None, [vp]),
        "orc_voxelmap_size": (C.c_size_t, [vp]),
        "orc_voxelmap_export": (None, [vp, _i32p, _dp, _dp, _u64p]),
        "orc_voxelmap_nn": (None, [vp, C.c_size_t, _dp, _u64p, _dp, _u64p]),
        "orc_reg_create": (vp, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int]),
        "orc_reg_destroy": (None, [vp]),
        "orc_reg_set_optimizer": (None, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
        "orc_reg_linearize": (None, [vp, vp, vp, vp, vp, _dp, _dp]),
        "orc_reg_error": (C.c_double, [vp, vp, vp, vp, _dp]),
        "orc_reg_correspondences": (None, [vp, _u64p]),
        "orc_reg_align": (None, [vp, vp, vp, vp, vp, _dp, C.c_int, _dp, _dp, _dp, _dp]),
        "orc_reg_trace_rows": (C.c_size_t, [vp]),
        "orc_reg_trace_get": (None, [vp, _dp]),
        "orc_se3_exp": (None, [_dp, _dp]),
        "orc_ldlt_solve6": (None, [_dp, _dp, _dp]),
        "orc_eigen_sym3": (None, [_dp, _dp, _dp]),
    }
    _SIGNATURES.update(sig)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if not native:
        _LIB = L
    return L


def reference_lib():
    """oracle/_ref/libsmallgicp_ref.so: the reference's OWN headers compiled against the Eigen API shim (oracle/ref_build), exposing the
    same C API.  Built only where /root/reference exists; returns None when it is absent."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libsmallgicp_ref.so")
        mk = os.path.join(_HERE, "ref_build", "Makefile")
        if not os.path.exists(path) and os.path.isdir("/root/reference") and os.path.exists(mk):
            subprocess.check_call(["make", "-s", "-C", os.path.dirname(mk)])
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        lib()  # fills _SIGNATURES
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _REF = L
    return _REF


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return lib().orc_max_threads()


class Cloud:
    """points/point_cloud.hpp PointCloud: points (N,4) w=1, normals (N,4) w=0, covs (N,4,4)."""

    def __init__(self, xyz=None, _handle=None, _lib=None):
        self._L = _lib or lib()
        if _handle is not None:
            self._h = _handle
        else:
            xyz = _f64(xyz)
            assert xyz.ndim == 2 and xyz.shape[1] in (3, 4)
            self._h = self._L.orc_cloud_create(xyz.shape[0], _d(xyz), xyz.shape[1])

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_cloud_destroy(self._h)
            self._h = None

    def __len__(self):
        return self._L.orc_cloud_size(self._h)

    def _get(self, which):
        n = len(self)
        p = np.empty((n, 4)) if which == 0 else None
        nr = np.empty((n, 4)) if which == 1 else None
        cv = np.empty((n, 4, 4)) if which == 2 else None
        self._L.orc_cloud_get(self._h, _d(p) if p is not None else None, _d(nr) if nr is not None else None, _d(cv) if cv is not None else None)
        return (p, nr, cv)[which]

    @property
    def points(self):
        return self._get(0)

    @property
    def normals(self):
        return self._get(1)

    @property
    def covs(self):
        return self._get(2)

    def set_features(self, normals=None, covs=None):
        n = _f64(normals) if normals is not None else None
        c = _f64(covs) if covs is not None else None
        self._L.orc_cloud_set_features(self._h, _d(n) if n is not None else None, _d(c) if c is not None else None)

    def transformed(self, T):
        T = _f64(T)
        return Cloud(_handle=self._L.orc_cloud_transformed(self._h, _d(T)), _lib=self._L)

    def voxelgrid_sampling(self, leaf):
        """util/downsampling.hpp:22-78"""
        return Cloud(_handle=self._L.orc_voxelgrid_sampling(self._h, float(leaf)), _lib=self._L)


class KdTree:
    """ann/kdtree.hpp KdTree<PointCloud> with the serial KdTreeBuilder (leaf <= 20)."""

    def __init__(self, cloud):
        self._L = cloud._L
        self.cloud = cloud
        self._h = self._L.orc_kdtree_create(cloud._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_kdtree_destroy(self._h)
            self._h = None

    def export(self):
        """(nodes: uint8 (n_nodes,24) raw reference layout, indices: uint64 (N,))"""
        nn = self._L.orc_kdtree_num_nodes(self._h)
        nodes = np.zeros((nn, 24), dtype=np.uint8)
        idx = np.zeros(len(self.cloud), dtype=np.uint64)
        self._L.orc_kdtree_export(self._h, nodes.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(_u64p))
        return nodes, idx

    def knn(self, queries, k, num_threads=1):
        q = _f64(queries)
        if q.shape[1] == 3:
            q = np.concatenate([q, np.ones((q.shape[0], 1))], axis=1)
        q = _f64(q)
        n = q.shape[0]
        idx = np.empty((n, k), dtype=np.uint64)
        d2 = np.empty((n, k), dtype=np.float64)
        cnt = np.empty(n, dtype=np.uint64)
        self._L.orc_kdtree_knn(self._h, n, _d(q), k, idx.ctypes.data_as(_u64p), _d(d2), cnt.ctypes.data_as(_u64p), num_threads)
        return idx, d2, cnt

    def estimate(self, k=20, mode=FEAT_NORMAL_COV, num_threads=1):
        """util/normal_estimation.hpp estimate_*(cloud, kdtree, k); writes into self.cloud"""
        self._L.orc_estimate_features(self.cloud._h, self._h, k, mode, num_threads)


class GaussianVoxelMap:
    """ann/gaussian_voxelmap.hpp GaussianVoxelMap(leaf).insert(cloud)"""

    def __init__(self, cloud, leaf=1.0, search_offsets=1):
        self._L = cloud._L
        self.leaf = float(leaf)
        self.search_offsets = int(search_offsets)
        self._h = self._L.orc_voxelmap_create(cloud._h, self.leaf, self.search_offsets)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_voxelmap_destroy(self._h)
            self._h = None

    def __len__(self):
        return self._L.orc_voxelmap_size(self._h)

    def export(self):
        n = len(self)
        coords = np.empty((n, 3), dtype=np.int32)
        means = np.empty((n, 4))
        covs = np.empty((n, 4, 4))
        cnt = np.empty(n, dtype=np.uint64)
        self._L.orc_voxelmap_export(self._h, coords.ctypes.data_as(_i32p), _d(means), _d(covs), cnt.ctypes.data_as(_u64p))
        return coords, means, covs, cnt

    def nn(self, queries):
        q = _f64(queries)
        n = q.shape[0]
        idx = np.empty(n, dtype=np.uint64)
        d2 = np.empty(n)
        cnt = np.empty(n, dtype=np.uint64)
        self._L.orc_voxelmap_nn(self._h, n, _d(q), idx.ctypes.data_as(_u64p), _d(d2), cnt.ctypes.data_as(_u64p))
        return idx, d2, cnt


class Result:
    pass


class Registration:
    """registration/registration.hpp Registration<Factor, Reduction, NullFactor, Rejector, Optimizer>.

    num_threads = 0 -> SerialReduction ; > 0 -> ParallelReductionOMP(num_threads).
    """

    def __init__(self, factor=FACTOR_GICP, robust=ROBUST_NONE, robust_c=1.0, rejector=REJECT_DISTANCE, max_dist_sq=1.0, num_threads=0, native=False, _lib=None):
        self._L = _lib or lib(native)
        self._h = self._L.orc_reg_create(factor, robust, float(robust_c), rejector, float(max_dist_sq), num_threads)
        self.set_optimizer()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_reg_destroy(self._h)
            self._h = None

    def set_optimizer(self, type=OPT_LM, max_iterations=20, gn_lambda=1e-6, max_inner_iterations=10, init_lambda=1e-3, lambda_factor=10.0, rotation_eps=0.1 * np.pi / 180.0, translation_eps=1e-3):
        self._L.orc_reg_set_optimizer(self._h, type, max_iterations, gn_lambda, max_inner_iterations, init_lambda, lambda_factor, rotation_eps, translation_eps)

    @staticmethod
    def _split(target, tree):
        if isinstance(target, GaussianVoxelMap):
            return None, None, target._h
        return target._h, (tree._h if tree is not None else None), None

    def linearize(self, target, tree, source, T):
        """Reduction::linearize -> (H 6x6, b 6, e)"""
        T = _f64(T)
        out = np.empty(43)
        th, kh, vh = self._split(target, tree)
        self._L.orc_reg_linearize(self._h, th, kh, vh, source._h, _d(T), _d(out))
        return out[:36].reshape(6, 6).copy(), out[36:42].copy(), float(out[42])

    def error(self, target, source, T):
        T = _f64(T)
        th, _, vh = self._split(target, None)
        return self._L.orc_reg_error(self._h, th, vh, source._h, _d(T))

    def correspondences(self, n):
        out = np.empty(n, dtype=np.uint64)
        self._L.orc_reg_correspondences(self._h, out.ctypes.data_as(_u64p))
        return out

    def align(self, target, tree, source, init_T=None, trace=False):
        init_T = _f64(np.eye(4) if init_T is None else init_T)
        T = np.empty((4, 4))
        sc = np.empty(4)
        H = np.empty((6, 6))
        b = np.empty(6)
        th, kh, vh = self._split(target, tree)
        self._L.orc_reg_align(self._h, th, kh, vh, source._h, _d(init_T), int(trace), _d(T), _d(sc), _d(H), _d(b))
        r = Result()
        r.T_target_source = T
        r.converged = bool(sc[0])
        r.iterations = int(sc[1])
        r.num_inliers = int(sc[2])
        r.error = float(sc[3])
        r.H = H
        r.b = b
        if trace:
            n = self._L.orc_reg_trace_rows(self._h)
            rows = np.empty((n, 59))
            if n:
                self._L.orc_reg_trace_get(self._h, _d(rows))
            r.trace = rows
        return r


def se3_exp(a):
    a = _f64(a)
    T = np.empty((4, 4))
    lib().orc_se3_exp(_d(a), _d(T))
    return T


def ldlt_solve6(A, b):
    A = _f64(A)
    b = _f64(b)
    x = np.empty(6)
    lib().orc_ldlt_solve6(_d(A), _d(b), _d(x))
    return x


def eigen_sym3(A):
    A = _f64(A)
    w = np.empty(3)
    V = np.empty((3, 3))
    lib().orc_eigen_sym3(_d(A), _d(w), _d(V))
    return w, V


def preprocess_points(xyz, downsampling_resolution, num_neighbors=10, num_threads=1):
    """registration_helper.cpp:22-33 preprocess_points (serial voxelgrid + serial tree + normals/covs)"""
    cloud = Cloud(xyz).voxelgrid_sampling(downsampling_resolution)
    tree = KdTree(cloud)
    tree.estimate(num_neighbors, FEAT_NORMAL_COV, num_threads)
    return cloud, tree
