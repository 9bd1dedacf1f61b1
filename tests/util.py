"""Shared helpers of the parity tests: seeded scenes and mesh canonicalisation."""
import numpy as np
import torch

from bodyslam_b200 import synthetic as S
from bodyslam_b200.geometry import PinholeCameraIntrinsic


def small_scene(name="laparoscopy512", res=128, frames=6, W=640, H=480, seed=0, with_color=True, frame_ids=None):
    """BASELINE config geometry at a resolution the CPU oracle finishes in seconds."""
    cfg = S.config(name)
    scale = cfg["resolution"] / res
    E_all = cfg["extrinsics"](cfg["frames"])
    if frame_ids is None:
        frame_ids = np.linspace(0, cfg["frames"] - 1, frames).astype(int)
    E = E_all[frame_ids]
    K = cfg["K"]
    if (W, H) != (cfg["W"], cfg["H"]):
        sx, sy = W / cfg["W"], H / cfg["H"]
        K = (K[0] * sx, K[1] * sy, K[2] * sx, K[3] * sy)
    depth_u16, color = S.render(cfg["surface"], E, K=K, W=W, H=H, device="cpu", seed=seed, with_color=with_color)
    return dict(E=E, K=K, W=W, H=H, depth_u16=depth_u16.numpy(), color=None if color is None else color.numpy(),
                resolution=res, voxel_length=cfg["voxel_length"] * scale, sdf_trunc=cfg["sdf_trunc"] * scale,
                origin=np.asarray(cfg["origin"], dtype=np.float64),
                intrinsic=PinholeCameraIntrinsic(W, H, *K))


def canon_mesh(vertices, keys, triangles, dims):
    """Order-independent form: vertices sorted by edge key, triangles as sorted rows of key ids."""
    vertices, keys, triangles = np.asarray(vertices), np.asarray(keys).astype(np.int64), np.asarray(triangles).astype(np.int64)
    nx, ny, nz = dims
    code = ((keys[:, 0] * (ny + 1) + keys[:, 1]) * (nz + 1) + keys[:, 2]) * 4 + keys[:, 3]
    order = np.argsort(code, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    tri = rank[triangles] if len(triangles) else triangles.reshape(0, 3)
    # rotate each triangle so its smallest id comes first (keeps orientation), then sort rows
    if len(tri):
        k = np.argmin(tri, axis=1)
        tri = np.stack([np.take_along_axis(tri, ((k + i) % 3)[:, None], 1)[:, 0] for i in range(3)], 1)
        tri = tri[np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))]
    return code[order], vertices[order], tri


def mesh_is_closed_and_oriented(triangles):
    """every directed edge appears exactly once and its reverse exactly once"""
    t = np.asarray(triangles, dtype=np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    code = e[:, 0] * (t.max() + 1) + e[:, 1]
    rev = e[:, 1] * (t.max() + 1) + e[:, 0]
    u, c = np.unique(code, return_counts=True)
    return bool(np.all(c == 1) and np.array_equal(np.sort(code), np.sort(rev)))
