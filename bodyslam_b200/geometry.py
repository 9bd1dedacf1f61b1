"""Duck-typed stand-ins for the Open3D value types that cross the 3DM boundary.

The reference passes Open3D objects through `TSDF.build_3D_map(rgbd, intrinsic, extrinsic)`
(N/3DM/tsdf.py:14-22) and gets Open3D meshes / point clouds back (tsdf.py:39-43).  Open3D is
not a dependency of this package; these classes carry the same attribute names so the
reference's call sites keep working, and every consumer here also accepts the real Open3D
objects (anything `np.asarray` can read).
"""
from __future__ import annotations

import numpy as np


def to_numpy(x):
    """numpy view/copy of a numpy array, torch tensor (CPU/CUDA) or Open3D Image/Vector."""
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def inverse4x4(m):
    """General 4x4 inverse by cofactors, row-major float64 -- the formula Eigen's `Matrix4d::inverse()` evaluates
    (Open3D: `extrinsic.inverse()` in CreateFromDepthImage / ScalableTSDFVolume::Integrate), so that the camera
    pose handed to the kernels has Open3D's rounding rather than LAPACK's.  Accepts [..., 4, 4]; evaluated by the
    library's host helper (bslam_invert4x4), the same routine the unit-activation kernel's poses come from."""
    from . import _lib

    a = np.ascontiguousarray(m, dtype=np.float64)
    out = np.empty_like(a)
    _lib.check(_lib.load().bslam_invert4x4(_lib.ptr(a), _lib.ptr(out), a.size // 16))
    return out


class PinholeCameraIntrinsic:
    """Open3D `camera.PinholeCameraIntrinsic` look-alike (N/3DM/slam_utils.py:48-68)."""

    def __init__(self, width: int = 640, height: int = 480, fx: float = 525.0, fy: float = 525.0,
                 cx: float = 319.5, cy: float = 239.5):
        self.set_intrinsics(width, height, fx, fy, cx, cy)

    def set_intrinsics(self, width, height, fx, fy, cx, cy):
        self.width, self.height = int(width), int(height)
        self.intrinsic_matrix = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)

    def get_focal_length(self):
        return float(self.intrinsic_matrix[0, 0]), float(self.intrinsic_matrix[1, 1])

    def get_principal_point(self):
        return float(self.intrinsic_matrix[0, 2]), float(self.intrinsic_matrix[1, 2])

    def __repr__(self):
        fx, fy = self.get_focal_length()
        cx, cy = self.get_principal_point()
        return f"PinholeCameraIntrinsic({self.width}x{self.height}, fx={fx}, fy={fy}, cx={cx}, cy={cy})"


def intrinsic_params(intrinsic):
    """(width, height, fx, fy, cx, cy) from ours, Open3D's, or a (W, H, 3x3 | fx,fy,cx,cy) tuple."""
    if hasattr(intrinsic, "intrinsic_matrix"):
        K = np.asarray(intrinsic.intrinsic_matrix, dtype=np.float64)
        return int(intrinsic.width), int(intrinsic.height), float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    if isinstance(intrinsic, (tuple, list)) and len(intrinsic) == 3:
        W, H, K = intrinsic
        K = np.asarray(K, dtype=np.float64)
        return int(W), int(H), float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    if isinstance(intrinsic, (tuple, list)) and len(intrinsic) == 6:
        W, H, fx, fy, cx, cy = intrinsic
        return int(W), int(H), float(fx), float(fy), float(cx), float(cy)
    raise TypeError(f"cannot read camera intrinsics from {type(intrinsic).__name__}")


class RGBDImage:
    """`.color` (H,W,3 u8 or H,W f32 intensity) + `.depth` (H,W f32 metres), numpy or torch."""

    def __init__(self, color=None, depth=None):
        self.color = color
        self.depth = depth

    @staticmethod
    def create_from_color_and_depth(color, depth, depth_scale=1000.0, depth_trunc=3.0,
                                    convert_rgb_to_intensity=True, device=None):
        """Open3D semantics (reference call N/3DM/slam_utils.py:212-220).

        u16 depth is converted ON THE GPU (bslam_depth_from_u16); the result stays a CUDA tensor.
        """
        from . import ops

        d = ops.depth_from_u16(depth, depth_scale=depth_scale, depth_trunc=depth_trunc, device=device)
        c = to_numpy(color) if not hasattr(color, "is_cuda") else color
        if hasattr(c, "shape") and tuple(c.shape[:2]) != tuple(d.shape[-2:]):
            raise RuntimeError("[CreateFromColorAndDepth] Unsupported image format.")
        if convert_rgb_to_intensity and getattr(c, "ndim", 0) == 3:
            cn = to_numpy(c).astype(np.float32) / 255.0
            c = (0.2990 * cn[..., 0] + 0.5870 * cn[..., 1] + 0.1140 * cn[..., 2]).astype(np.float32)
        return RGBDImage(c, d)


class _LazyArrays:
    """Attributes are torch CUDA tensors; `.numpy(name)` / np.asarray(getattr(...).cpu()) materialise."""

    _fields = ()

    def numpy(self, name):
        v = getattr(self, name)
        return None if v is None else to_numpy(v)

    def cpu(self):
        out = self.__class__.__new__(self.__class__)
        for k in self._fields:
            v = getattr(self, k)
            setattr(out, k, None if v is None else to_numpy(v))
        return out


class TriangleMesh(_LazyArrays):
    """`.vertices [V,3] f32`, `.triangles [T,3] i32`, `.vertex_colors [V,3] f32|None`.

    `.vertex_keys [V,4] i32` = (x, y, z, axis) of the voxel edge carrying each vertex (the key
    Open3D's extract_triangle_mesh de-duplicates on) -- lets callers canonicalise the order.
    """

    _fields = ("vertices", "triangles", "vertex_colors", "vertex_keys")

    def __init__(self, vertices=None, triangles=None, vertex_colors=None, vertex_keys=None):
        self.vertices, self.triangles = vertices, triangles
        self.vertex_colors, self.vertex_keys = vertex_colors, vertex_keys

    def has_vertex_colors(self):
        return self.vertex_colors is not None

    def canonical(self, dims):
        """Order-independent form (numpy): (edge codes sorted, vertices in that order, triangles as rows of
        ranks rotated to start at their smallest id and sorted) -- equal meshes give equal arrays whatever
        order the vertices / triangles were emitted in (single GPU, z-slabs, the oracle's serial scan)."""
        keys = to_numpy(self.vertex_keys).astype(np.int64)
        verts = to_numpy(self.vertices)
        tri = to_numpy(self.triangles).astype(np.int64)
        nx, ny, nz = (int(d) for d in dims)
        code = ((keys[:, 0] * (ny + 1) + keys[:, 1]) * (nz + 1) + keys[:, 2]) * 4 + keys[:, 3]
        order = np.argsort(code, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        tri = rank[tri] if len(tri) else tri.reshape(0, 3)
        if len(tri):
            k = np.argmin(tri, axis=1)
            tri = np.stack([np.take_along_axis(tri, ((k + i) % 3)[:, None], 1)[:, 0] for i in range(3)], 1)
            tri = tri[np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))]
        return code[order], verts[order], tri

    def canonical_digest(self, dims) -> str:
        """sha256 over the canonical form (edge codes, vertex coordinates bit for bit, triangles)"""
        import hashlib

        code, verts, tri = self.canonical(dims)
        h = hashlib.sha256()
        for a in (code, np.ascontiguousarray(verts, dtype=np.float32), tri):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


class PointCloud(_LazyArrays):
    """`.points`, `.colors`, `.normals` ([P,3] f32 each; colours/normals may be None)."""

    _fields = ("points", "colors", "normals", "point_keys")

    def __init__(self, points=None, colors=None, normals=None, point_keys=None):
        self.points, self.colors, self.normals, self.point_keys = points, colors, normals, point_keys

    def has_colors(self):
        return self.colors is not None

    def has_normals(self):
        return self.normals is not None
