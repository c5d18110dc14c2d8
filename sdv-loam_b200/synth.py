"""Seeded synthetic KITTI-shape data: grey images + 64-beam LiDAR depth pixels + trajectories.

There is no dataset and no network in this environment, so every test and bench input comes from here.
World = a corridor of textured planes (ground, two walls, ceiling, far wall) in KITTI camera axes
(x right, y down, z forward); texture = band-limited sum of sinusoids in plane coordinates, attenuated by
pixel footprint (analytic anti-aliasing) so photometric alignment between consecutive frames is well posed.

Shapes follow the reference's inputs: cropped 1200x360 image (calib/KITTI/00.txt:2-4), 64 beams x 1800 azimuth
LiDAR with the vertical layout of src/main.cpp:103-107 (ang_bottom 24.9 deg, ang_res_y 0.427 deg), pixel projection
keep-rule [4,w-5) x [4,h-4] (src/main.cpp:810-848).
"""
from __future__ import annotations
import numpy as np

KITTI_K = (718.856, 718.856, 607.1928 - 20.5, 185.2157 - 8.0)  # crop of (1241x376 -> 1200x360): principal point shifts by the crop offset
KITTI_WH = (1200, 360)
K360_K = (552.554261, 552.554261, 682.049453 + (1400 - 1408) / 2.0, 238.769549 - 8.0)
K360_WH = (1400, 360)
STRESS_K = (1100.0, 1100.0, 959.5, 599.5)
STRESS_WH = (1920, 1200)


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class World:
    """Corridor of 5 textured planes. plane k: n_k . p = d_k, with in-plane axes (e1,e2)."""

    def __init__(self, seed: int = 1000, n_waves: int = 48):
        rng = np.random.default_rng(seed)
        self.planes = [  # (normal, offset, e1, e2)
            (np.array([0.0, 1.0, 0.0]), 1.65, np.array([1.0, 0, 0]), np.array([0, 0, 1.0])),   # ground
            (np.array([1.0, 0.0, 0.0]), -7.0, np.array([0, 0, 1.0]), np.array([0, 1.0, 0])),   # left wall
            (np.array([1.0, 0.0, 0.0]), 9.0, np.array([0, 0, 1.0]), np.array([0, 1.0, 0])),    # right wall
            (np.array([0.0, 1.0, 0.0]), -6.0, np.array([1.0, 0, 0]), np.array([0, 0, 1.0])),   # ceiling
            (np.array([0.0, 0.0, 1.0]), 150.0, np.array([1.0, 0, 0]), np.array([0, 1.0, 0])),  # far wall
        ]
        # per plane: wavelengths log-uniform 0.15 m .. 6 m, 1/f amplitude
        self.waves = []
        for _ in self.planes:
            lam = np.exp(rng.uniform(np.log(0.15), np.log(6.0), n_waves))
            ang = rng.uniform(0, 2 * np.pi, n_waves)
            kx, ky = 2 * np.pi / lam * np.cos(ang), 2 * np.pi / lam * np.sin(ang)
            amp = 14.0 * (lam / 6.0) ** 0.35
            ph = rng.uniform(0, 2 * np.pi, n_waves)
            self.waves.append((kx, ky, amp, ph, lam))

    def cast(self, o, d):
        """o (3,), d (...,3) world rays -> (lambda, plane index); lambda=inf when nothing is hit."""
        best = np.full(d.shape[:-1], np.inf)
        idx = np.full(d.shape[:-1], -1, dtype=np.int32)
        for k, (n, off, _, _) in enumerate(self.planes):
            den = d @ n
            with np.errstate(divide="ignore", invalid="ignore"):
                lam = (off - o @ n) / den
            ok = (lam > 1e-3) & (lam < best) & np.isfinite(lam)
            best = np.where(ok, lam, best)
            idx = np.where(ok, k, idx)
        return best, idx

    def shade(self, p, idx, footprint):
        """p (...,3) hit points, idx plane index, footprint (...) world-space pixel size -> grey value."""
        out = np.full(idx.shape, 128.0)
        for k, (n, off, e1, e2) in enumerate(self.planes):
            m = idx == k
            if not m.any():
                continue
            a, b = p[m] @ e1, p[m] @ e2
            kx, ky, amp, ph, lam = self.waves[k]
            fp = footprint[m][:, None]
            att = np.exp(-0.5 * (2.2 * fp / lam[None, :]) ** 2)
            val = (amp[None, :] * att * np.sin(a[:, None] * kx[None, :] + b[:, None] * ky[None, :] + ph[None, :])).sum(1)
            out[m] = 128.0 + val
        return np.clip(out, 0.0, 255.0)


def trajectory(n: int, seed: int = 1000, step: float = 1.0, yaw_sigma: float = 0.01):
    """camToWorld poses (R (n,3,3), t (n,3)): ~1 m/frame forward, yaw random walk, small pitch/roll jitter."""
    rng = np.random.default_rng(seed + 7)
    Rs, ts = [], []
    yaw, pos = 0.0, np.zeros(3)
    for i in range(n):
        R = _rot(rng.normal(0, 0.002), yaw, rng.normal(0, 0.002))
        Rs.append(R)
        ts.append(pos.copy())
        yaw += rng.normal(0, yaw_sigma)
        pos = pos + R @ np.array([rng.normal(0, 0.02), rng.normal(0, 0.01), step * (1.0 + rng.normal(0, 0.03))])
    return np.stack(Rs), np.stack(ts)


def render(world: World, R, t, K=KITTI_K, wh=KITTI_WH, gain: float = 1.0, bias: float = 0.0, noise: float = 0.0, seed: int = 0):
    """Grey float32 image (h,w) in 0..255 and z-depth map for camToWorld (R,t)."""
    fx, fy, cx, cy = K
    w, h = wh
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    dc = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
    dw = dc @ R.T
    lam, idx = world.cast(t, dw)
    lamc = np.where(np.isfinite(lam), lam, 0.0)
    p = t + dw * lamc[..., None]
    n_all = np.stack([pl[0] for pl in world.planes])
    cosang = np.abs(np.einsum("...k,...k->...", dw, n_all[np.clip(idx, 0, None)])) / np.linalg.norm(dw, axis=-1)
    fp = lamc / fx / np.clip(cosang, 0.05, None)
    img = world.shade(p, idx, fp)
    img = np.where(idx >= 0, img, 128.0)
    img = gain * img + bias
    if noise > 0:
        img = img + np.random.default_rng(seed).normal(0, noise, img.shape)
    # mono8 like the reference's ingest (sensor_msgs/Image mono8 -> float, DatasetReader.h:152-155)
    return np.rint(np.clip(img, 0, 255)).astype(np.float32), np.where(idx >= 0, lam, 0.0)


def lidar_pixels(world: World, R, t, K=KITTI_K, wh=KITTI_WH, beams: int = 64, az: int = 1800, seed: int = 0, range_noise: float = 0.0):
    """64-beam sweep from the camera centre -> rows {Ku, Kv, depth} kept by the rule of src/main.cpp:810-848."""
    fx, fy, cx, cy = K
    w, h = wh
    elev = np.deg2rad(-24.9 + 0.427 * np.arange(beams) * (64.0 / beams))   # main.cpp:103-107 (128-beam stress: same span)
    azim = np.linspace(-np.pi, np.pi, az, endpoint=False)
    E, A = np.meshgrid(elev, azim, indexing="ij")
    # camera axes: z forward, x right, y down ; elevation up = -y
    dc = np.stack([np.cos(E) * np.sin(A), -np.sin(E), np.cos(E) * np.cos(A)], -1).reshape(-1, 3)
    dc = dc[dc[:, 2] > 0.05]
    lam, idx = world.cast(t, dc @ R.T)
    ok = idx >= 0
    pc = dc[ok] * lam[ok, None]
    if range_noise > 0:
        pc = pc * (1.0 + np.random.default_rng(seed).normal(0, range_noise, (pc.shape[0], 1)))
    Ku, Kv = fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy
    keep = (Ku >= 4) & (Ku < w - 5) & (Kv >= 4) & (Kv <= h - 4) & (pc[:, 2] > 0.5) & (pc[:, 2] < 120.0)
    return np.stack([Ku[keep], Kv[keep], pc[keep, 2]], -1)


def select_points(img, cloud_px, n_target: int, seed: int = 0, cell: int = 12):
    """Stand-in for the LiDAR-aware pixel selector (PixelSelector2.cpp:354-622, out of scope): keep, per grid cell,
    the LiDAR pixels with the largest image gradient until ~n_target remain. Returns rows {u,v,idepth}."""
    h, w = img.shape
    gx = np.zeros_like(img); gy = np.zeros_like(img)
    gx[:, 1:-1] = 0.5 * (img[:, 2:] - img[:, :-2]); gy[1:-1, :] = 0.5 * (img[2:, :] - img[:-2, :])
    g2 = gx * gx + gy * gy
    ui, vi = cloud_px[:, 0].astype(np.int64), cloud_px[:, 1].astype(np.int64)
    score = g2[vi, ui]
    cellid = (vi // cell) * ((w + cell - 1) // cell) + (ui // cell)
    order = np.lexsort((-score, cellid))
    cs = cellid[order]
    first = np.ones(len(cs), bool); first[1:] = cs[1:] != cs[:-1]
    rank = np.arange(len(cs)) - np.maximum.accumulate(np.where(first, np.arange(len(cs)), 0))
    per_cell = 1
    ncell = int(first.sum())
    while ncell * per_cell < n_target and per_cell < 64:
        per_cell += 1
    sel = order[rank < per_cell]
    rng = np.random.default_rng(seed)
    if len(sel) > n_target:
        sel = rng.choice(sel, n_target, replace=False)
    sel = np.sort(sel)
    return np.stack([cloud_px[sel, 0], cloud_px[sel, 1], 1.0 / cloud_px[sel, 2]], -1).astype(np.float32)


def rel_pose(Ra, ta, Rb, tb):
    """T_b<-a for camToWorld poses a (reference KF) and b (new frame): x_b = R x_a + t."""
    R = Rb.T @ Ra
    t = Rb.T @ (ta - tb)
    return R, t


class Sequence:
    """n frames of one synthetic drive: images, LiDAR pixels, ground-truth camToWorld poses."""

    def __init__(self, n: int, seed: int = 1000, K=KITTI_K, wh=KITTI_WH, beams: int = 64, step: float = 1.0,
                 gain_jitter: float = 0.0, bias_jitter: float = 0.0, noise: float = 0.0):
        self.K, self.wh, self.n, self.seed = K, wh, n, seed
        self.world = World(seed)
        self.R, self.t = trajectory(n, seed, step)
        rng = np.random.default_rng(seed + 13)
        self.images, self.clouds = [], []
        for i in range(n):
            g = 1.0 + (rng.normal(0, gain_jitter) if gain_jitter > 0 else 0.0)
            b = rng.normal(0, bias_jitter) if bias_jitter > 0 else 0.0
            img, _ = render(self.world, self.R[i], self.t[i], K, wh, gain=g, bias=b, noise=noise, seed=seed * 1000 + i)
            self.images.append(img)
            self.clouds.append(lidar_pixels(self.world, self.R[i], self.t[i], K, wh, beams=beams))
