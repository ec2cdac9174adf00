"""ctypes binding of include/sgicp_b200.h.  numpy in, numpy out; poses are 4x4 row-major numpy arrays
(transposed to Eigen's column-major at the boundary)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

FACTOR_ICP, FACTOR_PLANE_ICP, FACTOR_GICP = 0, 1, 2
ROBUST_NONE, ROBUST_HUBER, ROBUST_CAUCHY = 0, 1, 2
REJECT_NONE, REJECT_DISTANCE = 0, 1
NO_CORRESPONDENCE = np.uint64(0xFFFFFFFFFFFFFFFF)

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/sgicp_b200.h declares
_SIGNATURES = {
    "sgb_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "sgb_destroy": (None, [_vp]),
    "sgb_last_error": (C.c_char_p, [_vp]),
    "sgb_set_stream": (C.c_int, [_vp, _vp]),
    "sgb_synchronize": (C.c_int, [_vp]),
    "sgb_kernel_launches": (C.c_uint64, [_vp]),
    "sgb_comm_handle": (C.c_int, [_vp, _vp]),
    "sgb_comm_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "sgb_comm_mailbox": (C.c_int, [_vp, C.POINTER(_vp)]),
    "sgb_comm_connect_ptrs": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "sgb_comm_disconnect": (C.c_int, [_vp]),
    "sgb_comm_set_timeout_ms": (C.c_int, [_vp, C.c_int]),
    "sgb_comm_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "sgb_comm_wait_ns": (C.c_int, [_vp, _u64p, _u64p]),
    "sgb_drop_seeds": (C.c_int, [_vp]),
    "sgb_target_set_points": (C.c_int, [_vp, C.c_size_t, _dp, _dp, _dp]),
    "sgb_target_set_kdtree": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, _u64p]),
    "sgb_target_build_kdtree": (C.c_int, [_vp, C.c_int]),
    "sgb_target_set_voxelmap": (C.c_int, [_vp, C.c_double, C.c_size_t, _i32p, _dp, _dp, C.c_int]),
    "sgb_target_build_voxelmap": (C.c_int, [_vp, C.c_size_t, _dp, _dp, C.c_double, C.c_int]),
    "sgb_target_batch_knn": (C.c_int, [_vp, C.c_size_t, _dp, C.c_int, _u64p, _dp]),
    "sgb_target_size": (C.c_size_t, [_vp]),
    "sgb_source_set_points": (C.c_int, [_vp, C.c_size_t, _dp, _dp]),
    "sgb_source_size": (C.c_size_t, [_vp]),
    "sgb_linearize": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, _dp, _dp]),
    "sgb_error": (C.c_int, [_vp, _dp, _dp]),
    "sgb_linearize_device": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, _dp, _vp]),
    "sgb_error_device": (C.c_int, [_vp, _dp, _vp]),
    "sgb_correspondences": (C.c_int, [_vp, _u64p]),
    "sgb_num_inliers": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "sgb_estimate_features": (C.c_int, [_vp, C.c_size_t, _dp, C.c_int, _dp, _dp]),
    "sgb_target_estimate_features": (C.c_int, [_vp, C.c_int]),
    "sgb_source_estimate_features": (C.c_int, [_vp, C.c_int]),
    "sgb_target_adopt_source": (C.c_int, [_vp]),
    "sgb_voxelgrid_sampling": (C.c_int, [_vp, C.c_size_t, _dp, C.c_double, _dp, C.POINTER(C.c_size_t)]),
}


class SgbError(RuntimeError):
    pass


def library_path(profiling=False):
    """the product library; profiling=True: the same sources built with -DSGB_PROFILING (experiment switches read from SGB_*
    environment variables in sgb_create + the superseded A/B kernels) -- A/B scripts and the structure-agreement test only"""
    if os.environ.get("SGB_LIBRARY", "") == "prev" and not profiling:  # A/B against a saved build of the product library (scripts/gpu_ab.sh)
        return os.path.join(_HERE, "lib", "libsgicp_b200_prev.so")
    return os.path.join(_HERE, "lib", "libsgicp_b200_prof.so" if profiling else "libsgicp_b200.so")


def exported_symbols():
    return sorted(_SIGNATURES)


_LIBS = {}


def _lib(profiling=False):
    if profiling not in _LIBS:
        path = library_path(profiling)
        if not os.path.exists(path):
            raise SgbError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C small_gicp_b200/csrc`). There is no CPU fallback."
            )
        L = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _LIBS[profiling] = L
    return _LIBS[profiling]


def _is_dev(a):
    """a torch tensor (host or CUDA): passed by address -- the library copies with cudaMemcpyDefault, so device-resident producers
    hand their buffers over without a host round trip"""
    return hasattr(a, "data_ptr") and hasattr(a, "is_contiguous")


def _f64(a):
    if _is_dev(a):
        import torch

        assert a.dtype == torch.float64 and a.is_contiguous(), "torch inputs must be contiguous float64"
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    if a is None:
        return None
    if _is_dev(a):
        return C.cast(C.c_void_p(a.data_ptr()), _dp)
    return a.ctypes.data_as(_dp)


def _pose(T):
    T = _f64(T)
    assert T.shape == (4, 4)
    return np.ascontiguousarray(T.T)  # row-major transpose == column-major original


def _points4(p):
    if _is_dev(p):
        assert p.dim() == 2 and p.shape[1] == 4, "torch point inputs must be (N, 4) = (x, y, z, 1)"
        return _f64(p)
    p = _f64(p)
    assert p.ndim == 2 and p.shape[1] in (3, 4)
    if p.shape[1] == 3:
        p = np.concatenate([p, np.ones((p.shape[0], 1))], axis=1)
    return np.ascontiguousarray(p)


class Context:
    """One GPU, one stream (sgb_ctx)."""

    def __init__(self, device=0, profiling=None):
        if profiling is None:  # A/B scripts select the profiling build for a whole process
            profiling = os.environ.get("SGB_LIBRARY", "") == "prof"
        self._L = _lib(bool(profiling))
        h = _vp()
        rc = self._L.sgb_create(int(device), C.byref(h))
        if rc != 0:
            raise SgbError(self._L.sgb_last_error(None).decode())
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sgb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SgbError(self._L.sgb_last_error(self._h).decode())

    # ---- plumbing ----
    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.sgb_set_stream(self._h, _vp(cuda_stream_ptr) if cuda_stream_ptr else None))

    def synchronize(self):
        self._check(self._L.sgb_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(self._L.sgb_kernel_launches(self._h))

    # ---- target ----
    def set_target(self, points, normals=None, covs=None):
        p = _points4(points)
        n = _f64(normals) if normals is not None else None
        c = _f64(covs) if covs is not None else None
        if n is not None:
            assert tuple(n.shape) == (p.shape[0], 4)
        if c is not None:
            assert tuple(c.shape) == (p.shape[0], 4, 4)
        self._check(self._L.sgb_target_set_points(self._h, p.shape[0], _d(p), _d(n), _d(c)))

    def set_target_kdtree(self, nodes24, indices, root=0):
        nodes24 = np.ascontiguousarray(nodes24, dtype=np.uint8)
        indices = np.ascontiguousarray(indices, dtype=np.uint64)
        self._check(self._L.sgb_target_set_kdtree(self._h, nodes24.ctypes.data_as(_vp), nodes24.shape[0], int(root), indices.ctypes.data_as(_u64p)))

    def build_target_kdtree(self, max_leaf_size=0):
        self._check(self._L.sgb_target_build_kdtree(self._h, int(max_leaf_size)))

    def set_target_voxelmap(self, leaf_size, coords, means, covs, search_offsets=1):
        coords = np.ascontiguousarray(coords, dtype=np.int32)
        m = _points4(means)
        c = _f64(covs) if covs is not None else None
        self._check(self._L.sgb_target_set_voxelmap(self._h, float(leaf_size), m.shape[0], coords.ctypes.data_as(_i32p), _d(m), _d(c), int(search_offsets)))

    def build_target_voxelmap(self, points, covs, leaf_size, search_offsets=1):
        """GaussianVoxelMap(leaf_size).insert(points) built on the device (incremental_voxelmap.hpp:55-92)"""
        p = _points4(points)
        c = _f64(covs) if covs is not None else None
        if c is not None:
            assert tuple(c.shape) == (p.shape[0], 4, 4)
        self._check(self._L.sgb_target_build_voxelmap(self._h, p.shape[0], _d(p), _d(c), float(leaf_size), int(search_offsets)))

    def target_batch_knn(self, queries, k=1):
        """KdTree.batch_knn_search(queries, k) on the device -> (indices (N,k) uint64, squared distances (N,k))"""
        q = _points4(queries)
        idx = np.empty((q.shape[0], int(k)), dtype=np.uint64)
        d = np.empty((q.shape[0], int(k)))
        self._check(self._L.sgb_target_batch_knn(self._h, q.shape[0], _d(q), int(k), idx.ctypes.data_as(_u64p), _d(d)))
        return idx, d

    @property
    def target_size(self):
        return int(self._L.sgb_target_size(self._h))

    # ---- source ----
    def set_source(self, points, covs=None):
        p = _points4(points)
        c = _f64(covs) if covs is not None else None
        if c is not None:
            assert tuple(c.shape) == (p.shape[0], 4, 4)
        self._check(self._L.sgb_source_set_points(self._h, p.shape[0], _d(p), _d(c)))

    @property
    def source_size(self):
        return int(self._L.sgb_source_size(self._h))

    # ---- hot path ----
    def linearize(self, T, factor=FACTOR_GICP, robust=ROBUST_NONE, robust_c=1.0, rejector=REJECT_DISTANCE, max_dist_sq=1.0):
        """Reduction::linearize -> (H 6x6, b 6, e)"""
        Tc = _pose(T)
        out = np.empty(43)
        self._check(self._L.sgb_linearize(self._h, factor, robust, float(robust_c), rejector, float(max_dist_sq), _d(Tc), _d(out)))
        return out[:36].reshape(6, 6).copy(), out[36:42].copy(), float(out[42])

    def error(self, T):
        Tc = _pose(T)
        e = C.c_double(0.0)
        self._check(self._L.sgb_error(self._h, _d(Tc), C.byref(e)))
        return e.value

    def linearize_device(self, T, d_out_ptr, factor=FACTOR_GICP, robust=ROBUST_NONE, robust_c=1.0, rejector=REJECT_DISTANCE, max_dist_sq=1.0):
        """asynchronous; d_out_ptr = device address of >= 44 doubles (H | b | e | num_inliers)"""
        Tc = _pose(T)
        self._check(self._L.sgb_linearize_device(self._h, factor, robust, float(robust_c), rejector, float(max_dist_sq), _d(Tc), _vp(d_out_ptr)))

    def error_device(self, T, d_out_ptr):
        Tc = _pose(T)
        self._check(self._L.sgb_error_device(self._h, _d(Tc), _vp(d_out_ptr)))

    def correspondences(self):
        out = np.empty(self.source_size, dtype=np.uint64)
        self._check(self._L.sgb_correspondences(self._h, out.ctypes.data_as(_u64p)))
        return out

    # ---- multi-GPU: all-reduce of H|b|e fused into the reduction kernel (peer memory over NVLink) ----
    COMM_HANDLE_BYTES = 64

    def comm_handle(self):
        """64-byte CUDA IPC handle of this context's mailbox (to be all-gathered across the ranks)"""
        buf = (C.c_ubyte * self.COMM_HANDLE_BYTES)()
        self._check(self._L.sgb_comm_handle(self._h, C.cast(buf, _vp)))
        return bytes(buf)

    def comm_connect(self, rank, world, handles):
        """handles: world x 64 bytes in rank order; afterwards linearize()/error() return the sum over all ranks"""
        blob = b"".join(handles) if not isinstance(handles, (bytes, bytearray)) else bytes(handles)
        assert len(blob) == world * self.COMM_HANDLE_BYTES
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self._L.sgb_comm_connect(self._h, int(rank), int(world), C.cast(buf, _vp)))

    def comm_mailbox(self):
        """device address of this context's mailbox (wiring contexts of one process: tests)"""
        p = _vp()
        self._check(self._L.sgb_comm_mailbox(self._h, C.byref(p)))
        return int(p.value)

    def comm_connect_ptrs(self, rank, world, mailboxes):
        arr = (_vp * world)(*[_vp(int(m)) for m in mailboxes])
        self._check(self._L.sgb_comm_connect_ptrs(self._h, int(rank), int(world), arr))

    def comm_disconnect(self):
        self._check(self._L.sgb_comm_disconnect(self._h))

    def comm_set_timeout_ms(self, ms):
        self._check(self._L.sgb_comm_set_timeout_ms(self._h, int(ms)))

    def comm_status(self):
        """0 = fine, bit 0 = an exchange gave up on a peer, bit 1 = a peer reported a failed call (sticky until the next connect)"""
        st = C.c_int(0)
        self._check(self._L.sgb_comm_status(self._h, C.byref(st)))
        return int(st.value)

    def comm_wait_ns(self):
        """(ring of the last 64 exchanges' waiting times in ns indexed by call number % 64, number of exchanges so far)"""
        ring = np.zeros(64, dtype=np.uint64)
        calls = C.c_uint64(0)
        self._check(self._L.sgb_comm_wait_ns(self._h, ring.ctypes.data_as(_u64p), C.byref(calls)))
        return ring, int(calls.value)

    def drop_seeds(self):
        self._check(self._L.sgb_drop_seeds(self._h))

    def num_inliers(self):
        n = C.c_size_t(0)
        self._check(self._L.sgb_num_inliers(self._h, C.byref(n)))
        return int(n.value)

    # ---- per-cloud preparation on the device ----
    def estimate_features(self, points, num_neighbors=20, normals=True, covs=True):
        """estimate_normals_covariances (util/normal_estimation.hpp): returns (normals (N,4) | None, covs (N,4,4) | None)"""
        p = _points4(points)
        if _is_dev(p) and p.is_cuda:  # device in, device out
            import torch

            n_out = torch.empty((p.shape[0], 4), dtype=torch.float64, device=p.device) if normals else None
            c_out = torch.empty((p.shape[0], 4, 4), dtype=torch.float64, device=p.device) if covs else None
        else:
            n_out = np.empty((p.shape[0], 4)) if normals else None
            c_out = np.empty((p.shape[0], 4, 4)) if covs else None
        self._check(self._L.sgb_estimate_features(self._h, p.shape[0], _d(p), int(num_neighbors), _d(n_out), _d(c_out)))
        return n_out, c_out

    def estimate_target_features(self, num_neighbors=20):
        self._check(self._L.sgb_target_estimate_features(self._h, int(num_neighbors)))

    def estimate_source_features(self, num_neighbors=20):
        self._check(self._L.sgb_source_estimate_features(self._h, int(num_neighbors)))

    def adopt_source_as_target(self):
        """frame streams: the current source (points, its device-built tree, covariances) becomes the target; no source afterwards"""
        self._check(self._L.sgb_target_adopt_source(self._h))

    def voxelgrid_sampling(self, points, leaf_size):
        """voxelgrid_sampling (util/downsampling.hpp:22-78): returns the (M,4) down-sampled points"""
        p = _points4(points)
        out = np.empty_like(p)
        m = C.c_size_t(0)
        self._check(self._L.sgb_voxelgrid_sampling(self._h, p.shape[0], _d(p), float(leaf_size), _d(out), C.byref(m)))
        return out[: int(m.value)].copy()
