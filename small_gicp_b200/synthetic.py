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


def gt_transform_scaled(side):
    """T_gt for clouds of any size: the same 1.5 deg / (0.35, -0.20, 0.05) m offset as gt_transform(), but turned about the CENTRE of
    the room and with the angle scaled by 100 m / side for rooms larger than the 100 m one, so that the displacement of the farthest
    point (what decides whether a 1 m correspondence distance / a 1 m voxel can still find its neighbour) stays what it is at 1M points
    instead of growing to tens of metres at 100M (bench.py c4 / c5; the headline 1M pair keeps gt_transform())."""
    axis = np.array([0.2, 0.3, 0.93])
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(1.5) * min(1.0, 100.0 / side)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    c = np.array([0.5 * side, 0.5 * side, 0.0])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = c - R @ c + np.array([0.35, -0.20, 0.05])
    return T


def world_side(n_points):
    return 100.0 * np.sqrt(n_points / 1.0e6)


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


# ------------------------------------------------------------------------------------------------------------------
# Device-side generators (torch) for the large / streaming benchmark configurations: the same world definition as above,
# sampled on the GPU so that 10M - 100M point clouds never exist on the host (BASELINE configs[2..4]; bench.py c3 / c4 / c5).
# They use torch's generator, so the clouds differ from the numpy ones point by point (same distribution, fixed seeds).
# ------------------------------------------------------------------------------------------------------------------
def sample_cloud_torch(world, n_points, seed, device, noise=0.01, transform=None):
    """(N, 4) float64 CUDA tensor (x, y, z, 1): N points area-uniform over the world's faces + isotropic noise, rounded through float32;
    `transform` (4x4 numpy) is applied before the rounding (the source of a pair = resample moved by T_gt^-1)."""
    import torch

    F, area = world
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    Ft = torch.as_tensor(F, dtype=torch.float64, device=device)
    prob = torch.as_tensor(area / area.sum(), dtype=torch.float64, device=device)
    out = torch.empty((n_points, 4), dtype=torch.float64, device=device)
    step = 8_000_000  # bounds the transient memory at 100M points
    for lo in range(0, n_points, step):
        m = min(step, n_points - lo)
        fid = torch.multinomial(prob, m, replacement=True, generator=g)
        u = torch.rand((m, 1), dtype=torch.float64, device=device, generator=g)
        v = torch.rand((m, 1), dtype=torch.float64, device=device, generator=g)
        p = Ft[fid, 0] + u * Ft[fid, 1] + v * Ft[fid, 2]
        p += noise * torch.randn((m, 3), dtype=torch.float64, device=device, generator=g)
        if transform is not None:
            Tt = torch.as_tensor(transform, dtype=torch.float64, device=device)
            p = p @ Tt[:3, :3].T + Tt[:3, 3]
        out[lo : lo + m, :3] = p.to(torch.float32).to(torch.float64)
        out[lo : lo + m, 3] = 1.0
    return out


def lidar_pose(frame, side):
    """sensor pose of frame f: 1.0 m forward + 1 deg yaw per frame, starting near the middle of the room"""
    T = np.eye(4)
    x, y, yaw = 0.5 * side - 40.0, 0.5 * side - 20.0, 0.0
    for _ in range(frame):
        x += np.cos(yaw) * 1.0
        y += np.sin(yaw) * 1.0
        yaw += np.deg2rad(1.0)
    T[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
    T[:3, 3] = [x, y, 1.7]
    return T


def lidar_frame_torch(world, pose, seed, device, beams=64, azimuths=1875, max_range=100.0, range_noise=0.02):
    """One sweep of a spinning LiDAR (64 beams, -24.8 .. +2 deg, x 1875 azimuths = 120,000 rays) cast into the world from `pose`
    (4x4 numpy, sensor -> world): nearest hit per ray over all faces (axis-aligned rectangles), range noise N(0, 0.02^2), rays
    without a hit within max_range dropped.  Returns an (M, 4) float64 CUDA tensor in the SENSOR frame, rounded through float32."""
    import torch

    F, _ = world
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    el = torch.deg2rad(torch.linspace(-24.8, 2.0, beams, dtype=torch.float64, device=device))
    az = torch.linspace(0.0, 2.0 * np.pi, azimuths + 1, dtype=torch.float64, device=device)[:-1]
    ce, se = torch.cos(el)[:, None], torch.sin(el)[:, None]
    d_s = torch.stack([(ce * torch.cos(az)[None, :]).reshape(-1), (ce * torch.sin(az)[None, :]).reshape(-1), (se * torch.ones_like(az)[None, :]).reshape(-1)], dim=1)
    R = torch.as_tensor(pose[:3, :3], dtype=torch.float64, device=device)
    o = torch.as_tensor(pose[:3, 3], dtype=torch.float64, device=device)
    d = (d_s @ R.T).to(torch.float32)
    of = o.to(torch.float32)
    Ft = torch.as_tensor(F, dtype=torch.float32, device=device)
    lo = Ft[:, 0]
    hi = Ft[:, 0] + Ft[:, 1] + Ft[:, 2]
    ext = (hi - lo).abs()
    axis = ext.argmin(dim=1)  # the normal axis of an axis-aligned rectangle is the one it has no extent along
    best = torch.full((d.shape[0],), float("inf"), dtype=torch.float32, device=device)
    for a in range(3):
        sel = (axis == a).nonzero().squeeze(1)
        if sel.numel() == 0:
            continue
        b1, b2 = (a + 1) % 3, (a + 2) % 3
        c = lo[sel, a]
        l1, h1 = torch.minimum(lo[sel, b1], hi[sel, b1]), torch.maximum(lo[sel, b1], hi[sel, b1])
        l2, h2 = torch.minimum(lo[sel, b2], hi[sel, b2]), torch.maximum(lo[sel, b2], hi[sel, b2])
        for r0 in range(0, d.shape[0], 20000):  # rays x faces of one orientation, in slabs
            dd = d[r0 : r0 + 20000]
            t = (c[None, :] - of[a]) / dd[:, a, None]
            p1 = of[b1] + t * dd[:, b1, None]
            p2 = of[b2] + t * dd[:, b2, None]
            ok = (t > 0.5) & (p1 >= l1[None, :]) & (p1 <= h1[None, :]) & (p2 >= l2[None, :]) & (p2 <= h2[None, :])
            t = torch.where(ok, t, torch.full_like(t, float("inf")))
            best[r0 : r0 + 20000] = torch.minimum(best[r0 : r0 + 20000], t.min(dim=1).values)
    keep = best < max_range
    rng = best[keep].to(torch.float64) + range_noise * torch.randn(int(keep.sum()), dtype=torch.float64, device=device, generator=g)
    pts = d_s[keep] * rng[:, None]
    out = torch.ones((pts.shape[0], 4), dtype=torch.float64, device=device)
    out[:, :3] = pts.to(torch.float32).to(torch.float64)
    return out
