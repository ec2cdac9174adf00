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
