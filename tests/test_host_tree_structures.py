"""The product's host-side tree code (small_gicp_b200/csrc/sgb_kdtree_host.cpp) compiled stand-alone by tests/host_math and
checked without a GPU: adoption of a reference-built kd-tree (what sgb_target_set_kdtree does before the upload), the library's own
host builder (SGB_TREE=host), and the 64-byte packet records derived from either -- each walked on the host by the traversal rules
of the device kernels and held against brute force (kdtree_test.cpp:81-105: exact indices; kdtree_synthetic_test.cpp:177-193:
distances where ties allow another index).  Malformed trees must be refused with a message, not walked."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial import cKDTree

import oracle as O
from conftest import ROOT

HM_DIR = os.path.join(ROOT, "tests", "host_math")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def ht():
    subprocess.check_call(["make", "-s", "-C", HM_DIR])
    lib = C.CDLL(os.path.join(HM_DIR, "libsgb_host_tree.so"))
    lib.sgbt_last_error.restype = C.c_char_p
    lib.sgbt_adopt_and_search.restype = C.c_int
    lib.sgbt_build_and_search.restype = C.c_int
    return lib


def _p4(a):
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(np.c_[a[:, :3], np.ones(len(a))])


def _d(a):
    return a.ctypes.data_as(_dp)


def adopt(ht, nodes, idx, pts4, q4, mode, root=0):
    centre = np.ascontiguousarray(0.5 * (pts4[:, :3].min(0) + pts4[:, :3].max(0))) if len(pts4) else np.zeros(3)
    out_i = np.empty(len(q4), dtype=np.uint64)
    out_d = np.empty(len(q4), dtype=np.float32)
    info = np.zeros(4, dtype=np.int32)
    rc = ht.sgbt_adopt_and_search(nodes.ctypes.data_as(C.c_void_p), C.c_size_t(len(nodes)), C.c_uint32(root), idx.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(len(pts4)),
                                  _d(pts4), _d(centre), C.c_size_t(len(q4)), _d(q4), mode, out_i.ctypes.data_as(C.POINTER(C.c_uint64)), out_d.ctypes.data_as(C.POINTER(C.c_float)),
                                  info.ctypes.data_as(C.POINTER(C.c_int)))
    return rc, out_i, out_d, info


def build(ht, pts4, q4, mode, max_leaf):
    centre = np.ascontiguousarray(0.5 * (pts4[:, :3].min(0) + pts4[:, :3].max(0)))
    out_i = np.empty(len(q4), dtype=np.uint64)
    out_d = np.empty(len(q4), dtype=np.float32)
    info = np.zeros(4, dtype=np.int32)
    rc = ht.sgbt_build_and_search(C.c_size_t(len(pts4)), _d(pts4), _d(centre), max_leaf, C.c_size_t(len(q4)), _d(q4), mode, out_i.ctypes.data_as(C.POINTER(C.c_uint64)),
                                  out_d.ctypes.data_as(C.POINTER(C.c_float)), info.ctypes.data_as(C.POINTER(C.c_int)))
    return rc, out_i, out_d, info


def check_exact(pts4, q4, got_i, got_d, exact_indices=True):
    d, j = cKDTree(pts4[:, :3]).query(q4[:, :3])
    scale = max(1.0, float(np.abs(pts4[:, :3]).max()))
    # FP32 search on centred coordinates: same point unless two candidates tie to FP32 accuracy; the distance always agrees
    dist_got = np.linalg.norm(pts4[got_i.astype(np.int64), :3] - q4[:, :3], axis=1)
    assert np.all(np.abs(dist_got - d) <= 1e-6 * scale + 1e-6 * d), float(np.abs(dist_got - d).max())
    if exact_indices:
        assert (got_i.astype(np.int64) != j).mean() <= 2e-3
    assert np.all(np.abs(np.sqrt(got_d.astype(np.float64)) - d) <= 2e-6 * scale + 1e-5 * d)


@pytest.mark.parametrize("mode", [0, 1])
def test_adopted_reference_tree_is_an_exact_search_structure(ht, golden_prepared, mode):
    g = golden_prepared
    pts4 = _p4(g["target"].points)
    nodes, idx = g["target_tree"].export()
    rng = np.random.default_rng(3)
    q4 = _p4(np.concatenate([g["source"].points[:1500, :3], pts4[:500, :3], rng.uniform(-60, 60, (200, 3))]))
    rc, got_i, got_d, info = adopt(ht, nodes, idx, pts4, q4, mode)
    assert rc == 0, ht.sgbt_last_error()
    assert info[0] == len(nodes) and info[1] >= 1
    if mode == 1:
        assert info[2] == (len(nodes) - 1) // 2  # one packet record per inner node of a full binary tree
        assert info[3] <= info[1] + 1  # pending subtrees never exceed the depth the kernels size their stacks from
    check_exact(pts4, q4, got_i, got_d)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("max_leaf", [1, 20, 32, 64])
def test_own_host_builder_is_an_exact_search_structure(ht, golden_prepared, mode, max_leaf):
    g = golden_prepared
    pts4 = _p4(g["source"].points)
    q4 = _p4(g["target"].points[:2000, :3])
    rc, got_i, got_d, info = build(ht, pts4, q4, mode, max_leaf)
    assert rc == 0, ht.sgbt_last_error()
    check_exact(pts4, q4, got_i, got_d)


@pytest.mark.parametrize("mode", [0, 1])
def test_degenerate_clouds(ht, mode):
    """kdtree_synthetic_test.cpp:26-93: integer lattice with exact ties, duplicates, a +-1e6 range, collinear points, tiny clouds."""
    rng = np.random.default_rng(5)
    lattice = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(float)
    cases = {
        "lattice": (lattice, np.r_[lattice[::7] + 0.5, rng.uniform(-2, 14, (300, 3))]),
        "duplicates": (np.repeat(rng.normal(size=(300, 3)), 5, axis=0), rng.normal(size=(400, 3))),
        "huge-range": (rng.uniform(-1e6, 1e6, (5000, 3)), rng.uniform(-1e6, 1e6, (500, 3))),
        "collinear": (np.c_[np.linspace(0, 100, 3000), np.zeros(3000), np.zeros(3000)], np.c_[rng.uniform(0, 100, 300), rng.normal(size=(300, 2))]),
        "three-points": (rng.normal(size=(3, 3)), rng.normal(size=(20, 3))),
        "one-point": (np.array([[1.0, 2.0, 3.0]]), rng.normal(size=(5, 3))),
    }
    for name, (pts, q) in cases.items():
        pts4, q4 = _p4(pts.astype(np.float32)), _p4(q.astype(np.float32))
        rc, got_i, got_d, info = build(ht, pts4, q4, mode, 0)
        assert rc == 0, (name, ht.sgbt_last_error())
        check_exact(pts4, q4, got_i, got_d, exact_indices=False)
        tc = O.Cloud(pts4[:, :3])
        nodes, idx = O.KdTree(tc).export()
        rc, got_i, got_d, info = adopt(ht, nodes, idx, pts4, q4, mode)
        assert rc == 0, (name, ht.sgbt_last_error())
        check_exact(pts4, q4, got_i, got_d, exact_indices=False)


def test_empty_tree_is_accepted(ht):
    rc, got_i, got_d, info = adopt(ht, np.zeros((0, 24), dtype=np.uint8), np.zeros(0, dtype=np.uint64), np.zeros((0, 4)), _p4(np.zeros((3, 3))), 0)
    assert rc == 0 and np.all(got_i == np.uint64(0xFFFFFFFFFFFFFFFF))


def test_malformed_trees_are_refused(ht, golden_prepared):
    """sgb_target_set_kdtree takes raw memory from the caller: a root / child / point index out of range, an inverted leaf range, a bad
    axis or a cycle must produce an error message, never a walk through foreign memory."""
    g = golden_prepared
    pts4 = _p4(g["target"].points)
    nodes, idx = g["target_tree"].export()
    q4 = pts4[:4]
    view = lambda a: a.view(np.uint32).reshape(len(a), 6)  # [first|axis, last|pad, thresh lo, thresh hi? ...] -- raw words of the 24-byte record
    inner = int(np.nonzero(view(nodes)[:, 4] != 0xFFFFFFFF)[0][0])  # word 4 = left child (0xFFFFFFFF marks a leaf, ann/kdtree.hpp:197)
    leaf = int(np.nonzero(view(nodes)[:, 4] == 0xFFFFFFFF)[0][0])

    def expect_refusal(mut_nodes, mut_idx, root=0, what=""):
        rc, *_ = adopt(ht, mut_nodes, mut_idx, pts4, q4, 0, root)
        assert rc == 1 and what in ht.sgbt_last_error().decode(), (what, ht.sgbt_last_error())

    expect_refusal(nodes, idx, root=len(nodes), what="root")
    bad = idx.copy()
    bad[7] = len(pts4)
    expect_refusal(nodes, bad, what="point index")
    bad = nodes.copy()
    view(bad)[inner, 5] = len(nodes) + 3  # right child
    expect_refusal(bad, idx, what="inner node")
    bad = nodes.copy()
    view(bad)[inner, 0] = 7  # axis
    expect_refusal(bad, idx, what="inner node")
    bad = nodes.copy()
    view(bad)[leaf, 1] = len(pts4) + 1  # last > N
    expect_refusal(bad, idx, what="leaf range")
    bad = nodes.copy()
    view(bad)[inner, 4] = 0 if inner != 0 else inner  # left child points back to the root: a cycle
    view(bad)[inner, 5] = 0
    expect_refusal(bad, idx, what="cycle")
