"""Seeded synthetic Oxford/MulRan-shape polar radar scans.

There is no radar data in the reference repository and no network, so every test and benchmark in this repo
feeds the oracle and the CUDA path the *same bytes* produced here:

* a 600 m x 600 m world of wall segments and point reflectors,
* a figure-8 trajectory driven at 10 m/s and sampled at 4 Hz (Tsensor = 0.25 s,
  cfear_radarodometry/include/cfear_radarodometry/odometrykeyframefuser.h:213),
* per scan: 400 azimuths, theta_b = (b+1)/N_az * 2*pi as in radar_filters.cpp:317, rays cast from the
  sensor pose at the instant that azimuth is sampled (the inverse of the reference's motion compensation,
  utils.h:28-32 / utils.cpp:96-107), Gaussian noise floor, triangular returns, multipath ghosts, dropout.

Numpy only (host side); nothing here touches the GPU.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

OXFORD = dict(n_az=400, n_range=3768, range_res=0.0438, ccw=False)
MULRAN = dict(n_az=400, n_range=3360, range_res=0.0595238, ccw=True)


@dataclasses.dataclass
class World:
    seg_a: np.ndarray  # [S,2] wall start
    seg_b: np.ndarray  # [S,2] wall end
    refl: np.ndarray   # [R,2] point reflectors
    seg_amp: np.ndarray
    refl_amp: np.ndarray


def make_world(seed: int = 20230417, n_walls: int = 300, n_refl: int = 400, half: float = 300.0) -> World:
    rng = np.random.default_rng(seed)
    c = rng.uniform(-half, half, size=(n_walls, 2))
    length = rng.uniform(5.0, 40.0, size=n_walls)
    ang = rng.integers(0, 2, size=n_walls) * (math.pi / 2) + rng.normal(0.0, math.radians(5.0), size=n_walls)
    d = np.stack([np.cos(ang), np.sin(ang)], axis=1) * (length[:, None] / 2)
    refl = rng.uniform(-half, half, size=(n_refl, 2))
    return World(c - d, c + d, refl, rng.normal(110.0, 12.0, size=n_walls), rng.normal(120.0, 15.0, size=n_refl))


def figure8(n_frames: int, speed: float = 10.0, hz: float = 4.0, a: float = 150.0, s0: float = 0.0) -> np.ndarray:
    """Poses (x, y, yaw) [n_frames,3] along a lemniscate x=a sin t, y=a sin t cos t at constant speed."""
    t = np.linspace(0.0, 2 * math.pi, 20001)
    x, y = a * np.sin(t), a * np.sin(t) * np.cos(t)
    s = np.concatenate([[0.0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))])
    total = s[-1]
    sq = (s0 + np.arange(n_frames) * speed / hz) % total
    tq = np.interp(sq, s, t)
    px, py = a * np.sin(tq), a * np.sin(tq) * np.cos(tq)
    dx, dy = a * np.cos(tq), a * np.cos(2 * tq)
    yaw = np.unwrap(np.arctan2(dy, dx))
    return np.stack([px, py, yaw], axis=1)


def se2_mul(a, b):
    ca, sa = math.cos(a[2]), math.sin(a[2])
    return np.array([a[0] + ca * b[0] - sa * b[1], a[1] + sa * b[0] + ca * b[1], a[2] + b[2]])


def se2_inv(a):
    ca, sa = math.cos(a[2]), math.sin(a[2])
    return np.array([-(ca * a[0] + sa * a[1]), -(-sa * a[0] + ca * a[1]), -a[2]])


def render_scan(world: World, pose, motion, rng: np.random.Generator, n_az=400, n_range=3768, range_res=0.0438,
                ccw=False, noise_mean=32.0, noise_std=8.0, dropout=0.2) -> np.ndarray:
    """One polar scan u8 [n_az, n_range]. `motion` = frame-to-frame (x, y, yaw) used for the rolling distortion."""
    max_r = n_range * range_res
    b = np.arange(n_az)
    theta = (b + 1) / n_az * 2 * math.pi
    a = np.arctan2(np.sin(theta), np.cos(theta))
    d = np.where(a > 1e-5, a, 2 * math.pi + a) / (2 * math.pi) - 0.5
    if ccw:
        d = -d
    # sensor pose when azimuth b is sampled: T_b = T_center * (R(d*yaw_m), d*t_m)
    cy, sy = math.cos(pose[2]), math.sin(pose[2])
    ox = pose[0] + cy * (d * motion[0]) - sy * (d * motion[1])
    oy = pose[1] + sy * (d * motion[0]) + cy * (d * motion[1])
    phi = pose[2] + d * motion[2] + theta
    dx, dy = np.cos(phi), np.sin(phi)
    img = rng.normal(noise_mean, noise_std, size=(n_az, n_range)).astype(np.float32)

    hits_b, hits_r, hits_amp = [], [], []
    # walls: ray/segment intersection, all hits along the ray (attenuated by order)
    A, B = world.seg_a, world.seg_b
    ex, ey = (B - A)[:, 0], (B - A)[:, 1]
    den = dx[:, None] * ey[None, :] - dy[:, None] * ex[None, :]
    wx, wy = A[None, :, 0] - ox[:, None], A[None, :, 1] - oy[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (wx * ey[None, :] - wy * ex[None, :]) / den
        u = (wx * dy[:, None] - wy * dx[:, None]) / den
    ok = (np.abs(den) > 1e-9) & (t > 1.0) & (t < max_r) & (u >= 0.0) & (u <= 1.0)
    bi, si = np.nonzero(ok)
    if bi.size:
        tr = t[bi, si]
        order = np.lexsort((tr, bi))
        bi, si, tr = bi[order], si[order], tr[order]
        first = np.concatenate([[True], bi[1:] != bi[:-1]])
        rank = np.arange(bi.size) - np.maximum.accumulate(np.where(first, np.arange(bi.size), 0))
        amp = world.seg_amp[si] * (0.75 ** rank)
        hits_b.append(bi); hits_r.append(tr); hits_amp.append(amp)
    # point reflectors: visible in azimuths within +-0.9 deg
    rx, ry = world.refl[:, 0], world.refl[:, 1]
    for k in range(0, n_az, 100):  # chunk to bound memory
        sl = slice(k, k + 100)
        vx, vy = rx[None, :] - ox[sl, None], ry[None, :] - oy[sl, None]
        rr = np.hypot(vx, vy)
        dang = np.abs(np.arctan2(vy * dx[sl, None] - vx * dy[sl, None], vx * dx[sl, None] + vy * dy[sl, None]))
        okr = (dang < math.radians(0.9)) & (rr > 1.0) & (rr < max_r)
        bj, rj = np.nonzero(okr)
        if bj.size:
            hits_b.append(bj + k); hits_r.append(rr[bj, rj]); hits_amp.append(world.refl_amp[rj])
    if hits_b:
        hb, hr, ha = np.concatenate(hits_b), np.concatenate(hits_r), np.concatenate(hits_amp)
        # multipath ghost at twice the range, 30 % amplitude
        g = 2 * hr < max_r
        hb = np.concatenate([hb, hb[g]]); ha = np.concatenate([ha, 0.3 * ha[g]]); hr = np.concatenate([hr, 2 * hr[g]])
        keep = rng.random(hb.size) >= dropout
        hb, hr, ha = hb[keep], hr[keep], ha[keep]
        ha = ha + rng.normal(0.0, 20.0, size=ha.size)
        hw = rng.integers(2, 6, size=hb.size).astype(np.float32)  # half width -> 3..9 bins support
        rb = np.floor(hr / range_res).astype(np.int64)
        for off in range(-5, 6):
            w = np.maximum(0.0, 1.0 - abs(off) / hw)
            q = rb + off
            m = (w > 0) & (q >= 0) & (q < n_range)
            np.add.at(img, (hb[m], q[m]), (ha[m] * w[m]).astype(np.float32))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


@dataclasses.dataclass
class Stream:
    scans: np.ndarray     # [n, n_az, n_range] u8
    gt: np.ndarray        # [n, 3] ground-truth (x, y, yaw) of each scan's centre pose
    cfg: dict


def make_stream(n_frames: int, seed: int = 20230417, dataset: dict = OXFORD, s0: float = 0.0, world: World | None = None,
                speed: float = 10.0) -> Stream:
    world = world or make_world(seed)
    gt = figure8(n_frames + 1, speed=speed, s0=s0)
    rng = np.random.default_rng(seed + 7919)
    scans = np.empty((n_frames, dataset["n_az"], dataset["n_range"]), dtype=np.uint8)
    for i in range(n_frames):
        motion = se2_mul(se2_inv(gt[i]), gt[i + 1])  # constant-velocity proxy for the motion during scan i
        scans[i] = render_scan(world, gt[i], motion, rng, n_az=dataset["n_az"], n_range=dataset["n_range"],
                               range_res=dataset["range_res"], ccw=dataset["ccw"])
    return Stream(scans, gt[:n_frames], dict(dataset))


def stress_image(kind: str, n_az: int = 400, n_range: int = 3768, seed: int = 1) -> np.ndarray:
    """K1 stress inputs: 'uniform' iid u8, 'equal' all-equal rows, 'zeros', 'ramp'."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, size=(n_az, n_range), dtype=np.uint8)
    if kind == "equal":
        v = rng.integers(0, 256, size=(n_az, 1), dtype=np.uint8)
        return np.repeat(v, n_range, axis=1)
    if kind == "zeros":
        return np.zeros((n_az, n_range), dtype=np.uint8)
    if kind == "ramp":
        return (np.arange(n_range)[None, :] + np.arange(n_az)[:, None]).astype(np.uint8)
    if kind == "sparse":
        img = np.zeros((n_az, n_range), dtype=np.uint8)
        m = rng.random((n_az, n_range)) < 0.002
        img[m] = rng.integers(60, 256, size=int(m.sum()), dtype=np.uint8)
        return img
    raise ValueError(kind)
