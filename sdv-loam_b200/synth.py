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



# LiDAR -> camera extrinsics of a KITTI-like rig (Velodyne: x forward, y left, z up; camera: z forward, x right, y down): p_cam = RLC p_lidar + TLC   (sensor/00.txt role)
RLC = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
TLC = np.array([0.0, -0.08, -0.27])


def lidar_sweep(world: World, R, t, beams: int = 64, az: int = 1800, seed: int = 0, Rlc=RLC, tlc=TLC, objects: int = 40, dropout: float = 0.03, speckle: int = 60, range_noise: float = 0.004):
    """One raw sweep in the LiDAR frame as the node receives it (sensor_msgs/PointCloud2 XYZI rows, main.cpp:785-792): `beams` rings x `az` azimuths cast into the world,
    plus what makes the front-end's branches run: box-like objects in front of the planes (range steps -> segment borders), isolated speckle returns (infeasible
    segments), dropouts (empty cells), a few NaN rows and near returns (< 0.1 m), jittered firing angles (two returns in one cell: the later one wins)."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(-24.9 + 0.427 * (np.arange(beams) + 0.5) * (64.0 / beams))          # ring centres (main.cpp:103-107: ang_bottom, ang_res_y)
    azim = np.deg2rad(-180.0 + 0.2 * (np.arange(az) + 0.5) * (1800.0 / az))
    E, A = np.meshgrid(elev, azim, indexing="ij")
    E = E + rng.normal(0, np.deg2rad(0.03), E.shape); A = A + rng.normal(0, np.deg2rad(0.03), A.shape)
    dl = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)   # LiDAR frame
    o = t + R @ tlc; dw = (dl @ Rlc.T) @ R.T                                               # world rays from the LiDAR origin
    lam, idx = world.cast(o, dw); lam = np.where(idx >= 0, lam, np.nan)
    Ef, Af = E.reshape(-1), A.reshape(-1)
    for _ in range(objects):                                                               # angular boxes at a shorter range
        a0, e0 = rng.uniform(-np.pi, np.pi), np.deg2rad(rng.uniform(-20, 0)); da, de = np.deg2rad(rng.uniform(0.5, 8)), np.deg2rad(rng.uniform(1, 8)); r = rng.uniform(3, 35)
        m = (np.abs(np.angle(np.exp(1j * (Af - a0)))) < da) & (np.abs(Ef - e0) < de) & ~(lam < r)
        lam = np.where(m, r / np.maximum(np.cos(Ef - e0) * np.cos(np.angle(np.exp(1j * (Af - a0)))), 0.3), lam)
    lam = lam * (1.0 + rng.normal(0, range_noise, lam.shape))
    sp = rng.choice(len(lam), speckle, replace=False); lam[sp] = rng.uniform(2, 60, speckle)   # isolated returns
    lam[rng.uniform(size=len(lam)) < dropout] = np.nan
    keep = np.isfinite(lam) & (lam < 120.0)
    pts = (dl[keep] * lam[keep, None]).astype(np.float32); inten = rng.uniform(0, 1, len(pts)).astype(np.float32)
    cloud = np.concatenate([pts, inten[:, None]], 1)
    extra = np.array([[np.nan, 1, 1, 0], [1, np.inf, 0, 0], [0.01, 0.02, -0.01, 0.5], [0.03, -0.05, 0.0, 0.5]], np.float32)
    cloud = np.concatenate([cloud[: len(cloud) // 2], extra, cloud[len(cloud) // 2:]])
    return np.ascontiguousarray(cloud[rng.permutation(len(cloud))] if seed % 2 else cloud)

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


# ------------------------------------------------------------------------------------------------ back-end (BA) window
PATTERN8 = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]])   # src/util/settings.cpp:250


def _quat_from_R(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2; return np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    i = int(np.argmax(np.diag(R))); j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    q = np.zeros(4); q[0] = (R[k, j] - R[j, k]) / s; q[1 + i] = 0.25 * s; q[1 + j] = (R[j, i] + R[i, j]) / s; q[1 + k] = (R[k, i] + R[i, k]) / s
    return q


def make_ba_window(seq: "Sequence", kf_idx, n_per_frame: int = 250, seed: int = 0, sensor_frac: float = 0.7, pose_noise=(0.02, 0.002),
                   match_noise: float = 0.3, idepth_noise: float = 0.03, prior_scale: float = 0.0, no_matcher_frac: float = 0.03):
    """Flattened sliding window in the reference's iteration order (frames = ef->frames, points = ef->allPoints,
    residuals grouped per point).  Stands in for the host-side bookkeeping that is out of scope (point activation,
    Reprojector matches, marginalisation prior):
      * keyframes kf_idx of `seq`, evaluation-point poses = ground truth perturbed by pose_noise (m, rad) except frame 0
      * points: integer host pixels on LiDAR hits (isFromSensor, fixed depth) or vision-only (noisy idepth)
      * one residual per (point, other keyframe in view); matcher = ground-truth reprojection + match_noise px
    Returns a dict of numpy arrays consumed by both the oracle (orc.BAWindow) and the CUDA path (api.BAWindow)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = seq.K; w, h = seq.wh; nF = len(kf_idx)
    Rw2c = [seq.R[k].T for k in kf_idx]; tw2c = [-seq.R[k].T @ seq.t[k] for k in kf_idx]
    T_eval = np.zeros((nF, 7)); state = np.zeros((nF, 10)); state_zero = np.zeros((nF, 10))
    for i in range(nF):
        R, t = Rw2c[i], tw2c[i]
        if i > 0:
            dR = _rot(*rng.normal(0, pose_noise[1], 3)); R = dR @ R; t = dR @ t + rng.normal(0, pose_noise[0], 3)
        T_eval[i, :4] = _quat_from_R(R); T_eval[i, 4:] = t
    uv, idp, color, weights, host, hasPrior, fromSensor, res_begin = [], [], [], [], [], [], [], [0]
    r_point, r_host, r_target, r_hasM, r_match, r_new = [], [], [], [], [], []
    for hi, k in enumerate(kf_idx):
        img = seq.images[k]; cloud = seq.clouds[k]
        sel = select_points(img, cloud, n_per_frame, seed=seed + hi)
        gx = np.zeros_like(img); gy = np.zeros_like(img)
        gx[:, 1:-1] = 0.5 * (img[:, 2:] - img[:, :-2]); gy[1:-1, :] = 0.5 * (img[2:, :] - img[:-2, :])
        for (pu, pv, pid) in sel:
            u, v = int(pu), int(pv)                                      # ImmaturePoint truncates LiDAR sub-pixel coordinates (ImmaturePoint.cpp:8)
            if u < 4 or v < 4 or u >= w - 5 or v >= h - 5:
                continue
            d_true = 1.0 / pid
            sensor = rng.uniform() < sensor_frac
            idepth = pid if sensor else pid * (1.0 + rng.normal(0, idepth_noise))
            Xc = np.array([(u - cx) / fx, (v - cy) / fy, 1.0]) * d_true
            Xw = seq.R[k] @ Xc + seq.t[k]
            pidx = len(uv); nres = 0
            for ti in range(nF):
                if ti == hi:
                    continue
                Xt = Rw2c[ti] @ Xw + tw2c[ti]
                if Xt[2] < 0.5:
                    continue
                Ku, Kv = fx * Xt[0] / Xt[2] + cx, fy * Xt[1] / Xt[2] + cy
                if not (Ku > 6 and Kv > 6 and Ku < w - 7 and Kv < h - 7):
                    continue
                r_point.append(pidx); r_host.append(hi); r_target.append(ti)
                r_hasM.append(0 if rng.uniform() < no_matcher_frac else 1)
                r_match.append([Ku + rng.normal(0, match_noise), Kv + rng.normal(0, match_noise)]); r_new.append(1); nres += 1
            if nres == 0:
                continue
            uv.append([u, v]); idp.append(idepth)
            pu8, pv8 = u + PATTERN8[:, 0], v + PATTERN8[:, 1]
            color.append(img[pv8, pu8]); g2 = gx[pv8, pu8] ** 2 + gy[pv8, pu8] ** 2
            weights.append(np.sqrt(2500.0 / (2500.0 + g2)))              # setting_outlierTHSumComponent weighting of the host pattern
            host.append(hi); hasPrior.append(1 if sensor else 0); fromSensor.append(1 if sensor else 0)
            res_begin.append(len(r_point))
    n = 4 + 6 * nF
    if prior_scale > 0:
        A = rng.normal(size=(n + 4, n)); HM = prior_scale * (A.T @ A); bM = prior_scale * rng.normal(size=n)
    else:
        HM = np.zeros((n, n)); bM = np.zeros(n)
    return dict(nF=nF, kf_idx=list(kf_idx), K=np.array(seq.K, np.float64), wh=(w, h), T_eval=T_eval, state=state, state_zero=state_zero,
                ab_exposure=np.ones(nF, np.float32), frameID=np.arange(nF, dtype=np.int32), frameEnergyTH=np.full(nF, 8 * 8 * 8, np.float32),
                uv=np.array(uv, np.float32), idepth=np.array(idp, np.float32), idepth_zero=np.array(idp, np.float32),
                color=np.array(color, np.float32), weights=np.array(weights, np.float32), host=np.array(host, np.int32),
                hasDepthPrior=np.array(hasPrior, np.int32), isFromSensor=np.array(fromSensor, np.int32), res_begin=np.array(res_begin, np.int32),
                r_point=np.array(r_point, np.int32), r_host=np.array(r_host, np.int32), r_target=np.array(r_target, np.int32),
                r_hasMatcher=np.array(r_hasM, np.int32), r_matcher=np.array(r_match, np.float32), r_isNew=np.array(r_new, np.int32),
                HM=HM, bM=bM, T_gt=np.array([np.concatenate([_quat_from_R(Rw2c[i]), tw2c[i]]) for i in range(nF)]))


# ------------------------------------------------------------------------------------------------ overlap points (structPoseEstimation input)
def make_overlap_points(n: int, nH: int = 5, seed: int = 0, K=KITTI_K, wh=KITTI_WH, match_noise: float = 0.4, outlier_frac: float = 0.05,
                        pose_noise=(0.05, 0.004), step: float = 1.0):
    """Stand-in for Reprojector::reprojectMap output (FullSystem.cpp:483-485): n map points hosted in nH keyframes, each with the pixel
    it was matched to in the current frame (ground-truth projection + match_noise px, outlier_frac gross mismatches), plus the
    photometric tracker's pose estimate of the current frame (ground truth perturbed by pose_noise (m, rad)).  Geometry only."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K; w, h = wh
    R, t = trajectory(nH + 1, seed + 7, step)
    host_T7 = np.array([np.concatenate([_quat_from_R(R[i]), t[i]]) for i in range(nH)])
    Rc, tc = R[nH], t[nH]
    pts = np.zeros(n, np.dtype([("u", np.float32), ("v", np.float32), ("idepth", np.float32), ("host", np.int32), ("obs_x", np.float32), ("obs_y", np.float32)]))
    k = 0
    while k < n:
        hi = int(rng.integers(0, nH)); u = float(rng.integers(4, w - 5)); v = float(rng.integers(4, h - 5)); d = rng.uniform(4.0, 60.0)
        Xw = R[hi] @ (np.array([(u - cx) / fx, (v - cy) / fy, 1.0]) * d) + t[hi]
        Xc = Rc.T @ (Xw - tc)
        if Xc[2] < 1.0:
            continue
        Ku, Kv = fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy
        if not (Ku > 8 and Kv > 8 and Ku < w - 9 and Kv < h - 9):
            continue
        if rng.uniform() < outlier_frac:
            ox, oy = Ku + rng.normal(0, 25.0), Kv + rng.normal(0, 25.0)
        else:
            ox, oy = Ku + rng.normal(0, match_noise), Kv + rng.normal(0, match_noise)
        pts[k] = (u, v, 1.0 / d, hi, ox, oy); k += 1
    dR = _rot(*rng.normal(0, pose_noise[1], 3))
    T_gt = np.concatenate([_quat_from_R(Rc), tc])
    T_init = np.concatenate([_quat_from_R(dR @ Rc), tc + rng.normal(0, pose_noise[0], 3)])
    return dict(pts=pts, host_T7=host_T7, T_init=T_init, T_gt=T_gt, K=np.array(K, np.float64), wh=(w, h))


# ------------------------------------------------------------------------------------------------ map points (Reprojector input)
MAP_PT_DTYPE = np.dtype([("u", np.float32), ("v", np.float32), ("idepth", np.float32), ("host", np.int32), ("type", np.int32)])


def make_map(seq: "Sequence", kf_idx, n_per_frame: int = 300, seed: int = 0, edgelet_frac: float = 0.3, idepth_noise: float = 0.0):
    """Active map points of the keyframes kf_idx of `seq` (grouped by host, window order): integer pixels on LiDAR hits with their
    measured inverse depth (ImmaturePoint truncation, ImmaturePoint.cpp:8), point type drawn at random (the Shi-Tomasi classification of
    FullSystem.cpp:1326-1332 is point selection = out of scope).  Returns (pts[MAP_PT_DTYPE], host_T7 (camToWorld), host_ab)."""
    rng = np.random.default_rng(seed); w, h = seq.wh
    out = []
    for hi, k in enumerate(kf_idx):
        sel = select_points(seq.images[k], seq.clouds[k], n_per_frame, seed=seed + 17 * hi)
        for (pu, pv, pid) in sel:
            u, v = int(pu), int(pv)
            if u < 10 or v < 10 or u >= w - 11 or v >= h - 11:
                continue
            out.append((u, v, pid * (1.0 + rng.normal(0, idepth_noise)) if idepth_noise > 0 else pid, hi, 1 if rng.uniform() < edgelet_frac else 0))
    pts = np.array(out, MAP_PT_DTYPE)
    host_T7 = np.array([np.concatenate([_quat_from_R(seq.R[k]), seq.t[k]]) for k in kf_idx])
    return pts, host_T7, np.zeros((len(kf_idx), 2))


# ------------------------------------------------------------------------------------------------ small SE(3) helpers on T7 = {qw,qx,qy,qz,tx,ty,tz} (numpy, for data generation)
def _qmul(a, b):
    return np.array([a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3], a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
                     a[0]*b[2] + a[2]*b[0] + a[3]*b[1] - a[1]*b[3], a[0]*b[3] + a[3]*b[0] + a[1]*b[2] - a[2]*b[1]])


def _qrot(q, v):
    qv = q[1:]; uv = 2.0 * np.cross(qv, v); return v + q[0] * uv + np.cross(qv, uv)


def se3_mul7(a, b):
    q = _qmul(a[:4], b[:4]); q = q / np.linalg.norm(q); return np.concatenate([q, a[4:] + _qrot(a[:4], b[4:])])


def se3_inv7(a):
    q = np.array([a[0], -a[1], -a[2], -a[3]]); return np.concatenate([q, _qrot(q, -a[4:])])


def se3_exp7(xi):
    """xi = [upsilon(3); omega(3)] (Sophus order)."""
    ups, om = np.asarray(xi[:3], np.float64), np.asarray(xi[3:], np.float64); th = np.linalg.norm(om)
    Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0.0]])
    if th < 1e-10:
        q = np.array([1.0, 0.5 * om[0], 0.5 * om[1], 0.5 * om[2]]); V = np.eye(3) + 0.5 * Om
    else:
        q = np.concatenate([[np.cos(th / 2)], np.sin(th / 2) / th * om]); V = np.eye(3) + (1 - np.cos(th)) / th**2 * Om + (th - np.sin(th)) / th**3 * (Om @ Om)
    return np.concatenate([q / np.linalg.norm(q), V @ ups])


def se3_log7(T):
    q, t = T[:4], T[4:]; n = np.linalg.norm(q[1:])
    om = np.zeros(3) if n < 1e-12 else 2.0 * np.arctan2(n, q[0]) / n * q[1:]
    if q[0] < 0 and n >= 1e-12:
        om = 2.0 * np.arctan2(-n, -q[0]) / (-n) * (-q[1:])
    th = np.linalg.norm(om); Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0.0]])
    Vi = np.eye(3) - 0.5 * Om + ((1.0 / 12.0) if th < 1e-8 else (1 - th / (2 * np.tan(th / 2))) / th**2) * (Om @ Om)
    return np.concatenate([Vi @ t, om])
