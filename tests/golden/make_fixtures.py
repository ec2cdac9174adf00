#!/usr/bin/env python3
"""Generate the golden input fixtures from the reference's bundled test data.

Run HERE (the container that has /root/reference); the outputs are committed
so nothing on the GPU box ever reads /root/reference.

Source of truth (reference, read-only):
  data/target.ply, data/source.ply   69,088 / 69,792 vertices, 4 float props
                                     (x, y, z, scalar_intensity), binary LE
  data/T_target_source.txt           ground-truth 4x4 (row major text)
used by src/test/registration_test.cpp:29-44,90-97, helper_test.cpp and
src/test/python_test.py.  PLY parsing follows
include/small_gicp/benchmark/read_points.hpp:52-109 (float props, stride =
number of props, w := 1).

Outputs (this directory):
  target_xyz.f32 / source_xyz.f32    raw little-endian float32, N x 3
  T_target_source.txt                16 numbers, row major
"""
import os
import sys
import numpy as np

REF = os.environ.get("SGB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def read_ply_xyz(path):
    with open(path, "rb") as f:
        props = []
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            if line.startswith("element"):
                tok = line.split()
                assert tok[1] == "vertex", line
                n = int(tok[2])
            elif line.startswith("property"):
                tok = line.split()
                assert tok[1] == "float", line
                props.append(tok[2])
        assert [p.lower() for p in props[:3]] == ["x", "y", "z"], props
        buf = np.frombuffer(f.read(4 * len(props) * n), dtype="<f4").reshape(n, len(props))
    return np.ascontiguousarray(buf[:, :3])


def main():
    for name in ("target", "source"):
        xyz = read_ply_xyz(os.path.join(REF, "data", name + ".ply"))
        out = os.path.join(HERE, name + "_xyz.f32")
        xyz.astype("<f4").tofile(out)
        print(name, xyz.shape, "->", out)
    T = np.loadtxt(os.path.join(REF, "data", "T_target_source.txt")).reshape(4, 4)
    np.savetxt(os.path.join(HERE, "T_target_source.txt"), T, fmt="%.10g")
    print(T)


if __name__ == "__main__":
    sys.exit(main())
