"""On-disk readers of the reference's benchmark inputs (include/small_gicp/benchmark/read_points.hpp):
  read_points : KITTI velodyne .bin, N x float32 (x, y, z, intensity)  -> (N, 4) float32 with w := 1   (:15-33)
  write_points: the inverse                                                                           (:35-43)
  read_ply    : binary little-endian PLY whose vertex properties are all `float`, first three x y z   (:52-109)
Numpy only; the arrays feed small_gicp_b200.Context / host_api directly."""
import numpy as np


def read_points(filename):
    raw = np.fromfile(filename, dtype="<f4")
    pts = raw[: (raw.size // 4) * 4].reshape(-1, 4).copy()
    pts[:, 3] = 1.0
    return pts


def write_points(filename, points):
    np.ascontiguousarray(points, dtype="<f4").reshape(-1, 4).tofile(filename)


def read_ply(filename):
    with open(filename, "rb") as f:
        props, n = [], None
        while True:
            raw = f.readline()
            if not raw:
                raise ValueError(f"{filename}: no end_header")
            line = raw.decode("ascii", "replace").strip()
            if line == "end_header":
                break
            if line.startswith("element"):
                tok = line.split()
                if len(tok) != 3 or tok[1] != "vertex":
                    raise ValueError(f"invalid ply format (line={line})")
                n = int(tok[2])
            elif line.startswith("property"):
                tok = line.split()
                if tok[1] != "float":
                    raise ValueError(f"only float properties are supported (line={line})")
                props.append(tok[2])
        if n is None or len(props) < 3 or [p.lower() for p in props[:3]] != ["x", "y", "z"]:
            raise ValueError(f"invalid properties {props}")
        buf = np.frombuffer(f.read(4 * len(props) * n), dtype="<f4")
    if buf.size != len(props) * n:
        raise ValueError(f"{filename}: truncated vertex data")
    out = np.ones((n, 4), dtype=np.float32)
    out[:, :3] = buf.reshape(n, len(props))[:, :3]
    return out
