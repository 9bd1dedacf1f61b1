"""Tensor-level entry points of the hot path (torch tensors in, torch CUDA tensors out).

Each function is a thin call into libbodyslam_b200.so on the current CUDA stream.  torch is
plumbing here (device memory + streams); all arithmetic happens in the hand-written kernels.
No function has a CPU path: without a CUDA device they raise RuntimeError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .geometry import inverse4x4, to_numpy


def _device(device=None):
    torch = _lib.require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(f"bodyslam_b200 has no CPU path (device={device!r})")
    return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())


def as_cuda(x, dtype, device=None):
    """contiguous CUDA tensor of `dtype` from numpy / torch (CPU or CUDA) / array-like."""
    torch = _lib.require_cuda()
    dev = _device(device)
    if not hasattr(x, "is_cuda"):
        x = torch.from_numpy(np.ascontiguousarray(to_numpy(x)))
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.to(dev, non_blocking=True).contiguous()


def _aligned(x, nbytes=16):
    """the image kernels read 4 pixels per access: a contiguous view whose base is not `nbytes`-aligned (a slice of a
    larger buffer) is re-packed into its own allocation; torch allocations themselves are 512-byte aligned."""
    return x.clone() if x.data_ptr() % nbytes else x


def _dtype_name(x):
    return str(x.dtype if hasattr(x, "dtype") else to_numpy(x).dtype).replace("torch.", "")


def depth_from_u16(depth, depth_scale=1000.0, depth_trunc=3.0, device=None):
    """a4: Open3D create_from_color_and_depth depth conversion (N/3DM/slam_utils.py:212-220).

    uint16 input -> f32 metres CUDA tensor (`/ depth_scale`, `>= depth_trunc -> 0`); float32 input
    is passed through unchanged (Open3D's integrate takes float depth as is).
    """
    torch = _lib.require_cuda()
    dev = _device(device)
    name = _dtype_name(depth)
    if name == "float32":
        return as_cuda(depth, torch.float32, dev)
    if name != "uint16":
        raise RuntimeError(f"[depth_from_u16] Unsupported image format. (dtype {name})")
    src = as_cuda(depth, torch.uint16, dev)
    out = torch.empty(src.shape, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().bslam_depth_from_u16(_lib.ptr(src), src.numel(), float(depth_scale), float(depth_trunc or 0.0),
                                                    _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


def scale_to_u16(depth_m, scale=256.0, device=None):
    """a1: `(metres * 256).astype(uint16)` -- ZoeDepth's infer_pil(output_type='pil') tail."""
    torch = _lib.require_cuda()
    dev = _device(device)
    src = _aligned(as_cuda(depth_m, torch.float32, dev))
    out = torch.empty(src.shape, dtype=torch.uint16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().bslam_scale_u16(_lib.ptr(src), src.numel(), float(scale), _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


def _pack_rgba(c):
    c = [int(x) & 0xFF for x in c]
    while len(c) < 4:
        c.append(255)
    return c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24)


def colorize_u16(lut_rgba, depth_m=None, depth_u16=None, scale=256.0, invalid_val=None, background=(128, 128, 128, 255),
                 vmin=None, vmax=None, p_lo=2.0, p_hi=85.0, index_table=None, device=None, return_stats=False):
    """K1: fused metric scaling + percentile-normalised colour mapping of a [B,H,W] batch.

    Exactly one of depth_m (f32 metres) / depth_u16.  Returns (rgba [B,H,W,4] u8, u16 [B,H,W]) and,
    with return_stats, the per-image (vmin, vmax) f64 CUDA tensor [B,2].
    """
    torch = _lib.require_cuda()
    dev = _device(device)
    L = _lib.load()
    if (depth_m is None) == (depth_u16 is None):
        raise ValueError("give exactly one of depth_m / depth_u16")
    src = _aligned(as_cuda(depth_m, torch.float32, dev) if depth_m is not None else as_cuda(depth_u16, torch.uint16, dev))
    squeeze = src.dim() == 2
    if squeeze:
        src = src.unsqueeze(0)
    if src.dim() != 3:
        raise RuntimeError(f"[colorize] Unsupported image format. (shape {tuple(src.shape)})")
    B, H, W = src.shape
    lut = as_cuda(np.ascontiguousarray(to_numpy(lut_rgba), dtype=np.uint8).reshape(256, 4), torch.uint8, dev)
    rgba = torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev)
    u16 = torch.empty((B, H, W), dtype=torch.uint16, device=dev) if depth_m is not None else src
    ws = torch.empty(L.bslam_colorize_workspace_bytes(B), dtype=torch.uint8, device=dev)
    stats = torch.empty((B, 2), dtype=torch.float64, device=dev)
    h_over = None
    if vmin is not None or vmax is not None:
        h_over = np.full((B, 2), np.nan, np.float64)
        if vmin is not None:
            h_over[:, 0] = np.asarray(vmin, dtype=np.float64)
        if vmax is not None:
            h_over[:, 1] = np.asarray(vmax, dtype=np.float64)
    has_invalid = invalid_val is not None and float(invalid_val) == int(invalid_val) and 0 <= int(invalid_val) <= 65535
    tab = None
    if index_table is not None:
        tab = as_cuda(np.ascontiguousarray(to_numpy(index_table), dtype=np.uint8).reshape(-1, 65536), torch.uint8, dev)
        if tab.shape[0] == 1 and B > 1:
            tab = tab.expand(B, 65536).contiguous()
    with torch.cuda.device(dev):
        _lib.check(L.bslam_colorize(_lib.ptr(src) if depth_m is not None else None, _lib.ptr(src) if depth_m is None else None,
                                    B, H, W, float(scale), _lib.ptr(u16) if depth_m is not None else None, _lib.ptr(rgba),
                                    _lib.ptr(lut), float(p_lo), float(p_hi), int(has_invalid),
                                    int(invalid_val) if has_invalid else 0, _pack_rgba(background), _lib.ptr(h_over),
                                    _lib.ptr(stats), _lib.ptr(tab), _lib.ptr(ws), _lib.stream_ptr(dev)))
    if squeeze:
        rgba, u16 = rgba[0], u16[0]
    return (rgba, u16, stats) if return_stats else (rgba, u16)


def colorize_f32(lut_rgba, value, invalid_val=None, background=(128, 128, 128, 255), vmin=None, vmax=None, p_lo=2.0, p_hi=85.0,
                 device=None, return_stats=False):
    """K1, float flavour: percentile-normalised colour mapping of a float32 [B,H,W] / [H,W] image in NumPy's
    float32 arithmetic (the reference's `colorize` on a float tensor / array, depth_map_scaling.py:12-45).
    Returns rgba [.., H, W, 4] u8 CUDA (and the per-image (vmin, vmax) f64 CUDA tensor with return_stats)."""
    torch = _lib.require_cuda()
    dev = _device(device)
    L = _lib.load()
    src = as_cuda(value, torch.float32, dev)
    squeeze = src.dim() == 2
    if squeeze:
        src = src.unsqueeze(0)
    if src.dim() != 3:
        raise RuntimeError(f"[colorize] Unsupported image format. (shape {tuple(src.shape)})")
    B, H, W = src.shape
    lut = as_cuda(np.ascontiguousarray(to_numpy(lut_rgba), dtype=np.uint8).reshape(256, 4), torch.uint8, dev)
    rgba = torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev)
    ws = torch.empty(L.bslam_colorize_f32_workspace_bytes(B), dtype=torch.uint8, device=dev)
    stats = torch.empty((B, 2), dtype=torch.float64, device=dev)
    h_over = None
    if vmin is not None or vmax is not None:
        h_over = np.full((B, 2), np.nan, np.float64)
        if vmin is not None:
            h_over[:, 0] = np.asarray(vmin, dtype=np.float64)
        if vmax is not None:
            h_over[:, 1] = np.asarray(vmax, dtype=np.float64)
    with torch.cuda.device(dev):
        _lib.check(L.bslam_colorize_f32(_lib.ptr(src), B, H, W, _lib.ptr(rgba), _lib.ptr(lut), float(p_lo), float(p_hi),
                                        int(invalid_val is not None), float(invalid_val if invalid_val is not None else 0.0),
                                        _pack_rgba(background), _lib.ptr(h_over), _lib.ptr(stats), _lib.ptr(ws), _lib.stream_ptr(dev)))
    if squeeze:
        rgba = rgba[0]
    return (rgba, stats) if return_stats else rgba


def minmax_colormap(depth_u16, lut_bgr=None, device=None):
    """a12: `np.uint8(255*(d-min)/(max-min))` (+ 3-byte LUT) of N/3DM/slam_utils.py:250-264."""
    torch = _lib.require_cuda()
    dev = _device(device)
    L = _lib.load()
    src = _aligned(as_cuda(depth_u16, torch.uint16, dev))
    squeeze = src.dim() == 2
    if squeeze:
        src = src.unsqueeze(0)
    B, H, W = src.shape
    gray = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    rgb = lut = None
    if lut_bgr is not None:
        lut = as_cuda(np.ascontiguousarray(to_numpy(lut_bgr), dtype=np.uint8).reshape(256, 3), torch.uint8, dev)
        rgb = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    ws = torch.empty(L.bslam_colorize_workspace_bytes(B), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.bslam_minmax_u8(_lib.ptr(src), B, H, W, _lib.ptr(gray), _lib.ptr(rgb), _lib.ptr(lut), _lib.ptr(ws), _lib.stream_ptr(dev)))
    if squeeze:
        gray = gray[0]
        rgb = None if rgb is None else rgb[0]
    return gray, rgb


def median_u16(x_u16, invalid_val=None, device=None):
    """exact numpy-median of each image of a [B, ...] u16 batch -> f64 CUDA tensor [B]."""
    torch = _lib.require_cuda()
    dev = _device(device)
    L = _lib.load()
    src = _aligned(as_cuda(x_u16, torch.uint16, dev))
    if src.dim() == 1:
        src = src.unsqueeze(0)
    B = src.shape[0]
    n = src.numel() // B
    out = torch.empty(B, dtype=torch.float64, device=dev)
    ws = torch.empty(L.bslam_colorize_workspace_bytes(B), dtype=torch.uint8, device=dev)
    has_invalid = invalid_val is not None
    with torch.cuda.device(dev):
        _lib.check(L.bslam_median_u16(_lib.ptr(src), B, n, int(has_invalid), int(invalid_val or 0), _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr(dev)))
    return out


def backproject(depth, K, extrinsic=None, color=None, stride=1, valid_only=True, device=None, return_counts=False):
    """K2: pinhole back-projection + rigid transform of a [B,H,W] (or [H,W]) f32 depth batch.

    K = (fx, fy, cx, cy); extrinsic = world->camera 4x4 (or [B,4,4]); points are
    inverse(extrinsic) * [x,y,z,1] like Open3D create_from_depth_image.
    Returns xyz [M,3] f32 (images concatenated, row-major valid pixels) [, rgb [M,3] f32] and with
    return_counts the per-image row counts (i64 CUDA tensor [B]).
    """
    torch = _lib.require_cuda()
    dev = _device(device)
    L = _lib.load()
    d = as_cuda(depth, torch.float32, dev)
    if d.dim() == 2:
        d = d.unsqueeze(0)
    B, H, W = d.shape
    E = np.eye(4)[None].repeat(B, 0) if extrinsic is None else np.asarray(to_numpy(extrinsic), dtype=np.float64).reshape(-1, 4, 4)
    if E.shape[0] == 1 and B > 1:
        E = E.repeat(B, 0)
    M = np.ascontiguousarray(inverse4x4(E)[:, :3, :].reshape(B, 12), dtype=np.float32)   # Eigen's cofactor inverse, like Open3D
    Kf = np.ascontiguousarray(K, dtype=np.float32)
    c8 = None
    if color is not None:
        c8 = as_cuda(color, torch.uint8, dev)
        if c8.numel() != B * H * W * 3:
            raise RuntimeError("[backproject] Unsupported image format. (colour must be HxWx3 uint8)")
    rows = B * ((H + stride - 1) // stride) * ((W + stride - 1) // stride)
    xyz = torch.empty((rows, 3), dtype=torch.float32, device=dev)
    rgb = torch.empty((rows, 3), dtype=torch.float32, device=dev) if c8 is not None else None
    counts = torch.empty(B + 1, dtype=torch.int64, device=dev)
    ws = torch.empty(L.bslam_backproject_workspace_bytes(B, H, W, stride), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.bslam_backproject(_lib.ptr(d), _lib.ptr(c8), B, H, W, int(stride), _lib.ptr(Kf), _lib.ptr(M), int(bool(valid_only)),
                                       _lib.ptr(xyz), _lib.ptr(rgb), rows, _lib.ptr(counts), _lib.ptr(ws), _lib.stream_ptr(dev)))
    if valid_only:
        n = int(counts[B].item())
        xyz = xyz[:n]
        rgb = None if rgb is None else rgb[:n]
    out = (xyz, rgb)
    return out + (counts[:B],) if return_counts else out
