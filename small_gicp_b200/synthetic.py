"""Fixed-seed synthetic clouds for the benchmark configurations (BASELINE.json configs).

The reference ships no generator; this one is ours and is frozen:
  world  : axis-aligned room  side x side x 10 m (6 faces) + boxes (side U[2,8] m) standing on the floor,
           64 boxes per 100 x 100 m of floor, seed 42
  cloud  : N points area-uniform over all surfaces + N(0, 0.01^2 m) isotropic noise
  pair   : target (seed 43) and an INDEPENDENT resample (seed 44) moved by T_gt^-1, with
           T_gt = AngleAxis(1.5 deg, normalize(0.2, 0.3, 0.93)) (+) t = (0.35, -0.20, 0.05)
  sizes  : density is kept at 100 pts / m^2 of floor: side = 100 m * sqrt(N / 1e6)
Coordinates are rounded to float32 (what a LiDAR driver / PLY / KITTI .bin delivers) and returned as float64.
"""
import numpy as np


def gt_transform():
    axis = np.array([0.2, 0.3, 0.93])
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(1.5)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = [0.35, -0.20, 0.05]
    return T


def make_world(n_points, seed=42):
    side = 100.0 * np.sqrt(n_points / 1.0e6)
    height = 10.0
    rng = np.random.default_rng(seed)
    n_boxes = max(1, int(round(64 * (side / 100.0) ** 2)))
    dims = rng.uniform(2.0, 8.0, size=(n_boxes, 3))
    dims = np.minimum(dims, [side / 2, side / 2, height * 0.8])
    cxy = rng.uniform(0.0, side, size=(n_boxes, 2))
    # faces: (origin, edge u, edge v); area = |u||v|
    faces = []
    S, Hh = side, height
    faces += [((0, 0, 0), (S, 0, 0), (0, S, 0)), ((0, 0, Hh), (S, 0, 0), (0, S, 0))]  # floor, ceiling
    faces += [((0, 0, 0), (S, 0, 0), (0, 0, Hh)), ((0, S, 0), (S, 0, 0), (0, 0, Hh))]  # walls y=0, y=S
    faces += [((0, 0, 0), (0, S, 0), (0, 0, Hh)), ((S, 0, 0), (0, S, 0), (0, 0, Hh))]  # walls x=0, x=S
    for (dx, dy, dz), (cx, cy) in zip(dims, cxy):
        x0, y0 = cx - dx / 2, cy - dy / 2
        faces += [((x0, y0, dz), (dx, 0, 0), (0, dy, 0))]  # top
        faces += [((x0, y0, 0), (dx, 0, 0), (0, 0, dz)), ((x0, y0 + dy, 0), (dx, 0, 0), (0, 0, dz))]
        faces += [((x0, y0, 0), (0, dy, 0), (0, 0, dz)), ((x0 + dx, y0, 0), (0, dy, 0), (0, 0, dz))]
    F = np.array(faces, dtype=np.float64)  # (nf, 3, 3)
    area = np.linalg.norm(F[:, 1], axis=1) * np.linalg.norm(F[:, 2], axis=1)
    return F, area


def sample_cloud(world, n_points, seed, noise=0.01, return_normals=False):
    F, area = world
    rng = np.random.default_rng(seed)
    fid = rng.choice(len(F), size=n_points, p=area / area.sum())
    u = rng.random(n_points)[:, None]
    v = rng.random(n_points)[:, None]
    p = F[fid, 0] + u * F[fid, 1] + v * F[fid, 2]
    p += rng.normal(0.0, noise, size=p.shape)
    if return_normals:
        nrm = np.cross(F[:, 1], F[:, 2])
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        return p, nrm[fid]
    return p


def plane_covariances(normals, eps=1e-3):
    """Regularised plane covariance V diag(eps,1,1) V^T = I - (1-eps) n n^T as 4x4 zero-padded (N,4,4)
    (the closed form of normal_estimation.hpp:40-45 for an exactly known surface normal)."""
    n = normals
    C = np.zeros((len(n), 4, 4))
    C[:, :3, :3] = np.eye(3)[None] - (1.0 - eps) * n[:, :, None] * n[:, None, :]
    return C


def make_pair(n_points, n_source=None, seed_world=42, seed_target=43, seed_source=44):
    """Returns (target_xyz, source_xyz, T_gt) with T_gt * source ~ target."""
    n_source = n_points if n_source is None else n_source
    world = make_world(n_points, seed_world)
    T = gt_transform()
    tgt = sample_cloud(world, n_points, seed_target)
    src_w = sample_cloud(world, n_source, seed_source)
    Ti = np.linalg.inv(T)
    src = src_w @ Ti[:3, :3].T + Ti[:3, 3]
    return tgt.astype(np.float32).astype(np.float64), src.astype(np.float32).astype(np.float64), T
