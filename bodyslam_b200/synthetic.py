"""Seeded synthetic endoscopic scenes: analytic surfaces rendered to depth maps.

There is no dataset on the build or GPU boxes, so the bench and the parity tests run on these:
  * `cavity`  -- laparoscopy: ellipsoidal cavity seen from a trocar point, raster sweep (config 4)
  * `lumen`   -- colonoscopy: curved tube with haustral folds, camera on the centre line (config 2)
  * `gastro`  -- gastroscopy: the lumen opening into an ellipsoidal stomach (config 5)
Everything is torch, device-agnostic and deterministic (integer-hash dropout, no RNG state), so the
same call gives the same u16 depth on the CPU (oracle side) and on the GPU.
Poses are Open3D-style extrinsics: 4x4 float64 world->camera, camera looks down +z, y down.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# intrinsics used by the reference (N/3DM/slam.py:25-28)
K_640 = (957.411, 959.386, 282.192, 170.731)          # 640x480 (slam.py:26-27)
K_600 = (383.1901395, 383.1901395, 276.4727783203125, 124.3335933685303)  # 600x480 (slam.py:25)


def look_at(eye, forward, up_hint=(0.0, -1.0, 0.0)):
    """camera->world 4x4 with z = forward, y = down-ish; returns the world->camera extrinsic (f64)."""
    f = np.asarray(forward, dtype=np.float64)
    f = f / np.linalg.norm(f)
    down = -np.asarray(up_hint, dtype=np.float64)
    x = np.cross(down, f)
    if np.linalg.norm(x) < 1e-9:
        x = np.cross(np.array([1.0, 0.0, 0.0]), f)
    x /= np.linalg.norm(x)
    y = np.cross(f, x)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, f, np.asarray(eye, dtype=np.float64)
    return np.linalg.inv(c2w)


def _rays(K, W, H, c2w, device):
    fx, fy, cx, cy = K
    u = torch.arange(W, device=device, dtype=torch.float64)
    v = torch.arange(H, device=device, dtype=torch.float64)
    dx = ((u - cx) / fx)[None, :].expand(H, W)
    dy = ((v - cy) / fy)[:, None].expand(H, W)
    d_cam = torch.stack([dx, dy, torch.ones_like(dx)], -1)  # z = 1 -> ray parameter == depth
    R = torch.as_tensor(c2w[:3, :3], device=device, dtype=torch.float64)
    o = torch.as_tensor(c2w[:3, 3], device=device, dtype=torch.float64)
    return o, d_cam @ R.T


# ---------------------------------------------------------------- surfaces
class Cavity:
    """Ellipsoid x^2/a^2 + y^2/b^2 + z^2/c^2 = 1 seen from inside."""

    def __init__(self, semi_axes=(0.24, 0.22, 0.20)):
        self.axes = tuple(float(a) for a in semi_axes)

    def depth(self, o, d):
        A = torch.tensor([1.0 / a for a in self.axes], device=d.device, dtype=torch.float64)
        oa, da = o * A, d * A
        a = (da * da).sum(-1)
        b = 2.0 * (da * oa).sum(-1)
        c = (oa * oa).sum(-1) - 1.0
        disc = (b * b - 4 * a * c).clamp_min(0.0)
        t = (-b + disc.sqrt()) / (2 * a)
        return t


class Lumen:
    """Tube around a circular arc of radius Rc in the x-z plane; lumen radius with folds."""

    def __init__(self, Rc=0.4, radius=0.025, fold_amp=0.004, fold_period=0.03, stomach=None, max_depth=0.45):
        self.Rc, self.r0, self.amp, self.period, self.max_depth = Rc, radius, fold_amp, fold_period, max_depth
        self.stomach = stomach  # optional (centre xyz, semi axes) ellipsoid joined to the tube

    def centre(self, s):
        th = s / self.Rc
        return np.array([self.Rc * (1.0 - math.cos(th)), 0.0, self.Rc * math.sin(th)])

    def tangent(self, s):
        th = s / self.Rc
        return np.array([math.sin(th), 0.0, math.cos(th)])

    def _inside(self, p):
        """> 0 inside the lumen (approximate distance to the wall)."""
        x, y, z = p[..., 0], p[..., 1], p[..., 2]
        rho = torch.sqrt((x - self.Rc) ** 2 + z * z)
        th = torch.atan2(z, self.Rc - x)
        s = th * self.Rc
        dist = torch.sqrt((rho - self.Rc) ** 2 + y * y)
        r = self.r0 + self.amp * torch.sin(2 * math.pi * s / self.period)
        f = r - dist
        if self.stomach is not None:
            c, ax = self.stomach
            c = torch.as_tensor(c, device=p.device, dtype=torch.float64)
            ax = torch.as_tensor(ax, device=p.device, dtype=torch.float64)
            q = (p - c) / ax
            k = torch.sqrt((q * q).sum(-1))
            g = (1.0 - k) * ax.min()  # > 0 inside the stomach
            f = torch.maximum(f, g)
        return f

    def depth(self, o, d):
        n = torch.sqrt((d * d).sum(-1))
        t = torch.zeros(d.shape[:-1], device=d.device, dtype=torch.float64)
        for _ in range(96):
            f = self._inside(o + t[..., None] * d)
            t = t + (0.6 * f.clamp_min(0.0) / n)
        f = self._inside(o + t[..., None] * d)
        t = torch.where((f < 2e-4) & (t < self.max_depth), t, torch.zeros_like(t))
        return t


# ---------------------------------------------------------------- trajectories
def cavity_sweep(n_frames, eye=(0.0, 0.0, -0.17), max_angle_deg=35.0, cols=40):
    """raster sweep of the viewing direction around +z: yaw back and forth, pitch row by row."""
    rows = max(1, int(math.ceil(n_frames / cols)))
    E = []
    for i in range(n_frames):
        r, c = divmod(i, cols)
        cc = c if r % 2 == 0 else cols - 1 - c
        yaw = math.radians(max_angle_deg) * (2.0 * cc / max(cols - 1, 1) - 1.0)
        pitch = math.radians(max_angle_deg) * (2.0 * r / max(rows - 1, 1) - 1.0) if rows > 1 else 0.0
        f = np.array([math.sin(yaw) * math.cos(pitch), math.sin(pitch), math.cos(yaw) * math.cos(pitch)])
        E.append(look_at(eye, f))
    return np.stack(E)


def lumen_trajectory(lumen: Lumen, n_frames, step=0.001, wobble_deg=2.0, seed=0, s0=0.0):
    rng = np.random.default_rng(seed)
    E = []
    for i in range(n_frames):
        s = s0 + i * step
        f = lumen.tangent(s)
        w = np.radians(wobble_deg) * rng.uniform(-1, 1, size=2)
        side = np.array([math.cos(s / lumen.Rc), 0.0, -math.sin(s / lumen.Rc)])
        f = f + math.tan(w[0]) * side + math.tan(w[1]) * np.array([0.0, 1.0, 0.0])
        E.append(look_at(lumen.centre(s), f))
    return np.stack(E)


# ---------------------------------------------------------------- rendering
def _dropout_mask(shape, seed, frac, device):
    n = int(np.prod(shape))
    idx = torch.arange(n, device=device, dtype=torch.int64)
    h = (idx * 2654435761 + (seed + 1) * 40503) & 0xFFFFFFFF
    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
    h = (h ^ (h >> 13)) & 0xFFFFFFFF
    return (h.to(torch.float64) / 4294967296.0 < frac).view(shape)


def render(surface, extrinsics, K=K_640, W=640, H=480, device="cpu", invalid_frac=0.02, depth_scale=1000.0,
           with_color=True, seed=0, chunk=16, first_frame=0):
    """-> depth_u16 [F,H,W] (3DM units: metres * depth_scale, 0 = invalid), color u8 [F,H,W,3] | None.
    first_frame: trajectory index of extrinsics[0] -- the invalid-pixel pattern of a frame depends on its index in the
    trajectory, so a trajectory rendered in pieces (one piece per rank) equals the one rendered in one go."""
    device = torch.device(device)
    E = np.asarray(extrinsics, dtype=np.float64).reshape(-1, 4, 4)
    F = E.shape[0]
    depth = torch.empty((F, H, W), dtype=torch.uint16, device=device)
    color = torch.empty((F, H, W, 3), dtype=torch.uint8, device=device) if with_color else None
    for f in range(F):
        c2w = np.linalg.inv(E[f])
        o, d = _rays(K, W, H, c2w, device)
        t = surface.depth(o, d)
        q = torch.floor(t * depth_scale + 0.5).clamp(0, 65535)
        if invalid_frac > 0:
            q = torch.where(_dropout_mask((H, W), seed * 100003 + first_frame + f, invalid_frac, device), torch.zeros_like(q), q)
        depth[f] = q.to(torch.int32).to(torch.uint16)
        if with_color:
            p = o + t[..., None] * d
            ph = torch.stack([p[..., 0] * 90.0, p[..., 1] * 70.0 + 1.0, p[..., 2] * 50.0 + 2.0], -1)
            c = 128.0 + 100.0 * torch.sin(ph) + 20.0 * torch.sin(ph.flip(-1) * 7.0)
            color[f] = c.clamp(0, 255).to(torch.uint8)
    return depth, color


# ---------------------------------------------------------------- BASELINE.json configs
def config(name: str):
    """Geometry of the BASELINE.json workloads -> dict(surface, extrinsics fn, volume params)."""
    if name == "colonoscopy256":      # config 2: 300 frames, 256^3 @ 2 mm
        lum = Lumen()
        n = 256
        vl = 0.002
        mid = lum.centre(0.15)
        origin = (mid[0] - 0.5 * n * vl, -0.5 * n * vl, mid[2] - 0.5 * n * vl)
        return dict(surface=lum, frames=300, resolution=n, voxel_length=vl, sdf_trunc=0.01, origin=origin,
                    extrinsics=lambda F=300: lumen_trajectory(lum, F), K=K_640, W=640, H=480)
    if name == "laparoscopy512":      # config 4: 1000 frames, 512^3 @ 1 mm
        n, vl = 512, 0.001
        return dict(surface=Cavity(), frames=1000, resolution=n, voxel_length=vl, sdf_trunc=0.005,
                    origin=(-0.5 * n * vl,) * 3, extrinsics=lambda F=1000: cavity_sweep(F), K=K_640, W=640, H=480)
    if name == "gastroscopy1024":     # config 5: 5000 frames, 1024^3 @ 1 mm over 8 slabs
        n, vl = 1024, 0.001
        lum = Lumen(Rc=0.6, radius=0.012, fold_amp=0.0015, fold_period=0.02, max_depth=0.6)
        s_end = 0.35
        c_end, t_end = lum.centre(s_end), lum.tangent(s_end)
        lum.stomach = (tuple(c_end + 0.13 * t_end), (0.11, 0.09, 0.14))
        mid = lum.centre(0.25)
        origin = (mid[0] - 0.5 * n * vl, -0.5 * n * vl, mid[2] - 0.5 * n * vl)
        return dict(surface=lum, frames=5000, resolution=n, voxel_length=vl, sdf_trunc=0.005, origin=origin,
                    extrinsics=lambda F=5000: lumen_trajectory(lum, F, step=0.4 / 5000, wobble_deg=3.0), K=K_640, W=640, H=480)
    raise KeyError(name)
