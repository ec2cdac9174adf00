import numpy as np
import pytest

from conftest import load_golden_xyz
from small_gicp_b200 import io


def test_read_ply_roundtrip(tmp_path):
    xyz = load_golden_xyz("target")[:1000].astype(np.float32)
    inten = np.arange(1000, dtype=np.float32)[:, None]
    header = "ply\nformat binary_little_endian 1.0\ncomment test\nelement vertex 1000\nproperty float x\nproperty float y\nproperty float z\nproperty float scalar_intensity\nend_header\n"
    p = tmp_path / "a.ply"
    with open(p, "wb") as f:
        f.write(header.encode())
        f.write(np.hstack([xyz, inten]).astype("<f4").tobytes())
    pts = io.read_ply(p)
    assert pts.shape == (1000, 4) and np.all(pts[:, 3] == 1.0)
    np.testing.assert_array_equal(pts[:, :3], xyz)
    with open(tmp_path / "bad.ply", "wb") as f:
        f.write(header.replace("property float x", "property double x").encode())
    with pytest.raises(ValueError):
        io.read_ply(tmp_path / "bad.ply")


def test_kitti_bin_roundtrip(tmp_path):
    pts = np.random.default_rng(0).normal(size=(77, 4)).astype(np.float32)
    io.write_points(tmp_path / "f.bin", pts)
    got = io.read_points(tmp_path / "f.bin")
    np.testing.assert_array_equal(got[:, :3], pts[:, :3])
    assert np.all(got[:, 3] == 1.0)


# ---- the C++ readers of the host mirror (small_gicp_b200/host/include/small_gicp_b200/read_points.hpp) ----
PLY_HEADER = "ply\nformat binary_little_endian 1.0\ncomment test\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n{extra}end_header\n"


def _write_ply(path, xyz, extra_cols=0, header=None):
    extra = "".join(f"property float p{k}\n" for k in range(extra_cols))
    body = np.hstack([xyz, np.arange(len(xyz) * extra_cols, dtype=np.float32).reshape(len(xyz), extra_cols)]).astype("<f4")
    with open(path, "wb") as f:
        f.write((header or PLY_HEADER.format(n=len(xyz), extra=extra)).encode())
        f.write(body.tobytes())


@pytest.mark.parametrize("extra_cols", [0, 1, 3])
def test_cpp_read_ply_matches_python_reader(tmp_path, extra_cols):
    """benchmark/read_points.hpp:52-109 semantics: all-float vertex properties, the first three x y z, w := 1 -- for the bundled
    layout (x, y, z, intensity) and for other property counts (the reference's fixed 4-float read is the extra_cols == 1 case)."""
    from small_gicp_b200 import host_api

    xyz = load_golden_xyz("source")[:2500].astype(np.float32)
    p = tmp_path / "c.ply"
    _write_ply(p, xyz, extra_cols)
    got = host_api.read_points_cpp(p, "ply")
    assert got.shape == (2500, 4) and np.all(got[:, 3] == 1.0)
    np.testing.assert_array_equal(got, io.read_ply(p))
    np.testing.assert_array_equal(got[:, :3], xyz)


def test_cpp_read_ply_rejects_bad_files_without_throwing(tmp_path, capfd):
    """The reference prints to std::cerr and returns an empty vector (read_points.hpp:55-58,73-76,84-87); so does the mirror."""
    from small_gicp_b200 import host_api

    xyz = load_golden_xyz("source")[:10].astype(np.float32)
    assert host_api.read_points_cpp(tmp_path / "missing.ply", "ply").shape == (0, 4)
    bad = tmp_path / "double.ply"
    _write_ply(bad, xyz, header=PLY_HEADER.format(n=10, extra="").replace("property float x", "property double x"))
    assert host_api.read_points_cpp(bad, "ply").shape == (0, 4)
    wrong = tmp_path / "face.ply"
    _write_ply(wrong, xyz, header=PLY_HEADER.format(n=10, extra="").replace("element vertex", "element face"))
    assert host_api.read_points_cpp(wrong, "ply").shape == (0, 4)
    short = tmp_path / "short.ply"
    _write_ply(short, xyz, header=PLY_HEADER.format(n=11, extra=""))  # header promises one vertex more than the file holds
    assert host_api.read_points_cpp(short, "ply").shape == (0, 4)
    assert "error" in capfd.readouterr().err


def test_cpp_kitti_bin_roundtrip(tmp_path):
    from small_gicp_b200 import host_api

    pts = np.random.default_rng(1).normal(size=(129, 4)).astype(np.float32)
    host_api.write_points_cpp(tmp_path / "g.bin", pts)
    np.testing.assert_array_equal(np.fromfile(tmp_path / "g.bin", dtype="<f4").reshape(-1, 4), pts)  # raw float4 records
    got = host_api.read_points_cpp(tmp_path / "g.bin", "bin")
    np.testing.assert_array_equal(got, io.read_points(tmp_path / "g.bin"))  # intensity overwritten by w = 1 (read_points.hpp:28-30)
    assert np.all(got[:, 3] == 1.0)
    io.write_points(tmp_path / "h.bin", pts)  # and the Python writer's files read back through the C++ reader
    np.testing.assert_array_equal(host_api.read_points_cpp(tmp_path / "h.bin", "bin")[:, :3], pts[:, :3])


def test_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/registration_on_b200.cpp (the three entry levels of the reference's public API against this backend) compiles against the header-only
    mirror; on a machine without a CUDA device it reads its inputs and then stops with the backend's message -- no CPU fallback."""
    import os
    import subprocess

    from conftest import ROOT

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    exe = os.path.join(ROOT, "examples", "registration_on_b200")
    tgt, src = tmp_path / "t.ply", tmp_path / "s.ply"
    _write_ply(tgt, load_golden_xyz("target")[:3000].astype(np.float32), 1)
    _write_ply(src, load_golden_xyz("source")[:3000].astype(np.float32), 1)
    r = subprocess.run([exe, str(tgt), str(src)], capture_output=True, text=True, timeout=120)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert r.returncode == 0 and "inliers" in r.stdout, r.stderr[-500:]
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr[-500:])
    assert subprocess.run([exe, str(tmp_path / "none.ply"), str(src)], capture_output=True).returncode == 1
