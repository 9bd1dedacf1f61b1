"""3DM frame container and map rebuild -- drop-in for the hot-path parts of `N/3DM/slam_utils.py`
and `N/3DM/scaling_system.py:72-77`.

Kept signatures: `RGBD(color_path, depth_path, device, depth_scale=1000, depth_trunc=3.0)`,
`update_map_after_pg(global_extrinsic, list_of_rgb, list_of_depth, depth_scale, device, intrinsic)`,
`get_o3d_intrinsic(...)`, `compute_curr_estimate_global_pose(...)`, `ensure_so3_v2`, `pixel_to_3d`.
Out of scope (not on the path): the Open3D device pickers, Umeyama alignment.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _lib, ops
from .geometry import PinholeCameraIntrinsic, RGBDImage, to_numpy
from .tsdf import TSDF


def get_o3d_intrinsic(frame_width: int, frame_height: int, fx: float, fy: float, cx: float, cy: float):
    '''
    This function build the intrinsic matrix used by open3d (slam_utils.py:48-68)
    :return: (intrinsic, 3x3 float64 matrix)
    '''
    o3d_intrinsic = PinholeCameraIntrinsic()
    o3d_intrinsic.set_intrinsics(width=frame_width, height=frame_height, fx=fx, fy=fy, cx=cx, cy=cy)
    o3d_intrinsic_t = np.array(o3d_intrinsic.intrinsic_matrix, dtype=np.float64)
    return o3d_intrinsic, o3d_intrinsic_t


def ensure_so3_v2(matrix: np.ndarray) -> np.ndarray:
    """Projects a 3x3 matrix to the closest SO(3) matrix (slam_utils.py:93-108)."""
    U, _, Vt = np.linalg.svd(matrix)
    D = np.eye(3)
    D[2, 2] = np.linalg.det(U) * np.linalg.det(Vt)
    return np.dot(U, np.dot(D, Vt))


def compute_curr_estimate_global_pose(global_extrinsic: np.ndarray, transformation: np.ndarray) -> np.ndarray:
    '''
    Compute the global current pose from the relative motion (slam_utils.py:110-122): E_prev @ T, then
    the rotation block is re-projected onto SO(3).  The result is used as Open3D's world->camera
    extrinsic unchanged (slam.py:148,179).
    '''
    curr_global_pose = np.dot(global_extrinsic, transformation)
    curr_global_pose[:3, :3] = ensure_so3_v2(curr_global_pose[:3, :3])
    return curr_global_pose


def pixel_to_3d(u, v, depth, fx, fy, cx, cy):
    """Pinhole back-projection of one pixel (scaling_system.py:72-77).  The dense form is
    `ops.backproject` (K2); this scalar keeps the reference helper's signature and value."""
    x = (u - cx) * depth / fx
    y = (v - cy) * depth / fy
    z = depth
    return np.array([x, y, z])


class RGBD:
    """Frame container (slam_utils.py:172-264): loads a colour + 16-bit depth pair and exposes the
    representations the SLAM loop reads.  Depth scaling (`/depth_scale`, `>= depth_trunc -> 0`),
    min/max and the JET preview run on the GPU; decoded images stay as CUDA tensors."""

    def __init__(self, color_path: str, depth_path: str, device=None, depth_scale: int = 1000, depth_trunc: float = 3.0):
        import cv2
        from PIL import Image

        torch = _lib.require_cuda()
        self.color_path = color_path
        self.depth_path = depth_path
        self.depth_scale = depth_scale
        self.depth_trunc = depth_trunc
        self.device = ops._device(str(device).lower() if (device is not None and "cuda" in str(device).lower()) else None)

        depth_u16 = cv2.imread(depth_path, cv2.IMREAD_ANYDEPTH)
        if depth_u16 is None or depth_u16.dtype != np.uint16:
            raise RuntimeError(f"[RGBD] cannot read a 16-bit depth image from {depth_path}")
        bgr = cv2.imread(color_path)
        if bgr is None:
            raise RuntimeError(f"[RGBD] cannot read a colour image from {color_path}")
        self.cv2_color = bgr
        rgb = np.ascontiguousarray(bgr[:, :, ::-1])

        self.o3d_depth = ops.as_cuda(depth_u16, torch.uint16, self.device)       # raw u16
        self.o3d_color = ops.as_cuda(rgb, torch.uint8, self.device)              # RGB u8
        self.o3d_t_depth, self.o3d_t_color = self.o3d_depth, self.o3d_color
        # create_from_color_and_depth(..., convert_rgb_to_intensity=False) -- slam_utils.py:216-220
        self.rgbd_tsdf = RGBDImage(self.o3d_color, ops.depth_from_u16(self.o3d_depth, depth_scale, depth_trunc, self.device))
        self.rgbd = self.rgbd_tsdf
        self.rgbd_t = self.rgbd_tsdf
        # cv2 flavour: astype(float32) / depth_scale, no truncation -- slam_utils.py:231-233
        self.cv2_depth = ops.depth_from_u16(self.o3d_depth, depth_scale, 0.0, self.device)
        self.colored_depth = self._compute_colored_depth()
        self.pil_color = Image.open(color_path)
        self.depth_min, self.depth_max = self._compute_min_max_depth()
        self.height, self.width = int(depth_u16.shape[0]), int(depth_u16.shape[1])

    def _compute_min_max_depth(self) -> Tuple[float, float]:
        mn, mx = self.cv2_depth.min(), self.cv2_depth.max()
        return float(mn.item()), float(mx.item())

    def _compute_colored_depth(self):
        """min-max normalised JET preview (slam_utils.py:250-264), BGR like cv2.applyColorMap."""
        import cv2

        lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_JET).reshape(256, 3)
        _, bgr = ops.minmax_colormap(self.o3d_depth, lut_bgr=lut, device=self.device)
        return bgr


def load_frames(list_of_rgb, list_of_depth, n, depth_scale, depth_trunc, device):
    """decode n colour/depth pairs -> (depth u16 [n,H,W], rgb u8 [n,H,W,3]) pinned host tensors"""
    import cv2

    torch = _lib.require_cuda()
    d0 = cv2.imread(list_of_depth[0], cv2.IMREAD_ANYDEPTH)
    H, W = d0.shape
    depth = torch.empty((n, H, W), dtype=torch.uint16).pin_memory()
    rgb = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
    for i in range(n):
        d = cv2.imread(list_of_depth[i], cv2.IMREAD_ANYDEPTH)
        c = cv2.imread(list_of_rgb[i])
        if d is None or c is None or d.shape != (H, W) or c.shape[:2] != (H, W):
            raise RuntimeError("[update_map_after_pg] Unsupported image format.")
        depth[i] = torch.from_numpy(d)
        rgb[i] = torch.from_numpy(np.ascontiguousarray(c[:, :, ::-1]))
    return depth, rgb


def update_map_after_pg(global_extrinsic, list_of_rgb, list_of_depth, depth_scale, device, intrinsic, **tsdf_kwargs):
    """Re-integrate frames 0..n-1 with the optimised poses into a fresh TSDF (slam_utils.py:124-135).

    Where the reference loops `RGBD(...)` + `build_3D_map` per frame, the replay here is batched:
    frames are decoded into pinned memory, copied once, converted by the a4 kernel and integrated
    by the multi-frame K3 launch (the volume is read and written once per 256 frames).
    """
    torch = _lib.require_cuda()
    tsdf = TSDF(**tsdf_kwargs)
    n = len(global_extrinsic)
    if n == 0:
        return tsdf
    dev = tsdf.tsdf.device
    depth_u16, rgb = load_frames(list_of_rgb, list_of_depth, n, depth_scale, 3.0, dev)
    E = np.stack([np.asarray(to_numpy(e), dtype=np.float64) for e in global_extrinsic])
    if "origin" not in tsdf_kwargs:     # no box given: centre it on what the first frame sees (TSDF.auto_centre)
        tsdf.auto_centre(ops.depth_from_u16(depth_u16[0], depth_scale, 3.0, dev), intrinsic, E[0])
        dev = tsdf.tsdf.device
    tsdf.tsdf.integrate_host(depth_u16, rgb if tsdf.tsdf.color else None, intrinsic, E, depth_scale=depth_scale, depth_trunc=3.0)
    return tsdf
