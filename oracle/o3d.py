"""ctypes front-end of ``libo3d_oracle.so`` (TEST INFRASTRUCTURE, see o3d_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libo3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("o3d_oracle.c", "mc_tables.h"))
    ):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        p = C.c_void_p
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_depth_from_u16.argtypes = [p, C.c_size_t, C.c_double, C.c_double, p]
        L.orc_backproject.restype = C.c_int64
        L.orc_backproject.argtypes = [p, p, C.c_int, C.c_int, p, p, C.c_int, C.c_int, p, p]
        L.orc_tsdf_integrate.restype = C.c_int64
        L.orc_tsdf_integrate.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         p, p, p, C.c_int, C.c_int, p, p, C.c_int]
        L.orc_invert4x4.argtypes = [p, p]
        L.orc_set_scalable_schedule.argtypes = [C.c_int]
        L.orc_set_scalable_all_units.argtypes = [C.c_int]
        L.orc_scalable_integrate.restype = C.c_int64
        L.orc_scalable_integrate.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, p, C.c_int, C.c_int, C.c_double, C.c_double,
                                             p, p, C.c_int, C.c_int, p, p, p, C.c_int, p]
        L.orc_extract_mesh.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, p,
                                       p, p, p, C.c_int64, p, p, C.c_int64, p]
        L.orc_extract_points.restype = C.c_int64
        L.orc_extract_points.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, p,
                                         p, p, p, p, C.c_int64]
        L.orc_set_extract_flavour.argtypes = [C.c_double, C.c_double]
        L.orc_vbg_integrate.restype = C.c_int64
        L.orc_vbg_integrate.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, p, C.c_double, C.c_double, p, p, C.c_int, C.c_int, p, p,
                                        C.c_double, C.c_double, p]
        L.orc_count_occupied.restype = C.c_int64
        L.orc_count_occupied.argtypes = [p, C.c_size_t]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def set_scalable_schedule(open3d_like: bool) -> None:
    """threading of `Volume.integrate_scalable`: True = Open3D's own (touched units one after the other, OpenMP
    only over the x loop inside a unit), False (default) = units spread over the threads (faster; same result)"""
    lib().orc_set_scalable_schedule(0 if open3d_like else 1)


def set_extract_flavour(weight_threshold: float = 0.0, pos_half: float = 0.5) -> None:
    """flavour of extract_mesh / extract_points: legacy Open3D (weight != 0, voxel centres) by default; the tensor
    pipeline behind `MAP` uses weight >= 3 and voxel corners (3.0, 0.0).  Global: reset it after use."""
    lib().orc_set_extract_flavour(float(weight_threshold), float(pos_half))


def depth_from_u16(depth_u16, depth_scale=1000.0, depth_trunc=3.0):
    """Appendix A.1 -- reference call site N/3DM/slam_utils.py:212-220."""
    a = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    out = np.empty(a.shape, np.float32)
    lib().orc_depth_from_u16(_ptr(a), a.size, float(depth_scale), float(depth_trunc), _ptr(out))
    return out


def backproject(depth_f32, K, extrinsic=None, rgb=None, stride=1, valid_only=True):
    """Appendix A.2 -- N/3DM/mapping_module.py:37,41 / N/3DM/scaling_system.py:72-77.

    K = (fx, fy, cx, cy); extrinsic = world->camera 4x4 (f64).  Returns (xyz f64 [M,3], rgb f64 [M,3] | None).
    """
    d = np.ascontiguousarray(depth_f32, dtype=np.float32)
    H, W = d.shape
    Kd = np.ascontiguousarray(K, dtype=np.float64)
    E = np.eye(4) if extrinsic is None else np.asarray(extrinsic, dtype=np.float64)
    M = np.zeros(16)
    lib().orc_invert4x4(_ptr(np.ascontiguousarray(E)), _ptr(M))      # Eigen's 4x4 inverse is the cofactor formula
    M = M.reshape(4, 4)
    rows = ((H + stride - 1) // stride) * ((W + stride - 1) // stride)
    xyz = np.empty((rows, 3), np.float64)
    c = None
    col = None
    if rgb is not None:
        c = np.ascontiguousarray(rgb, dtype=np.uint8)
        col = np.empty((rows, 3), np.float64)
    n = lib().orc_backproject(_ptr(d), _ptr(c), W, H, _ptr(Kd), _ptr(M), int(stride), int(bool(valid_only)), _ptr(xyz), _ptr(col))
    return xyz[:n], (None if col is None else col[:n])


class Volume:
    """Dense TSDF box in Open3D order idx=(x*ny+y)*nz+z (UniformTSDFVolume semantics, A.3-A.5)."""

    def __init__(self, resolution, voxel_length, sdf_trunc, origin=(0.0, 0.0, 0.0), gz0=0, with_color=False):
        if np.isscalar(resolution):
            resolution = (int(resolution),) * 3
        self.nx, self.ny, self.nz = (int(r) for r in resolution)
        self.voxel_length = float(voxel_length)
        self.sdf_trunc = float(sdf_trunc)
        self.origin = np.ascontiguousarray(origin, dtype=np.float64)
        self.gz0 = int(gz0)
        n = self.nx * self.ny * self.nz
        self.tsdf = np.zeros(n, np.float32)
        self.weight = np.zeros(n, np.float32)
        self.color = np.zeros(n * 3, np.float32) if with_color else None

    def integrate(self, depth_f32, K, extrinsic, rgb=None, z_restart=8) -> int:
        d = np.ascontiguousarray(depth_f32, dtype=np.float32)
        H, W = d.shape
        Kd = np.ascontiguousarray(K, dtype=np.float64)
        E = np.ascontiguousarray(extrinsic, dtype=np.float64)
        c = None if rgb is None else np.ascontiguousarray(rgb, dtype=np.uint8)
        return lib().orc_tsdf_integrate(_ptr(self.tsdf), _ptr(self.weight), _ptr(self.color if c is not None else None),
                                        self.nx, self.ny, self.nz, self.gz0, self.voxel_length, self.sdf_trunc,
                                        _ptr(self.origin), _ptr(d), _ptr(c), W, H, _ptr(Kd), _ptr(E), int(z_restart))

    def integrate_scalable(self, depth_f32, K, extrinsic, rgb=None, z_restart=0, unit_res=32, stride=8, return_touched=False, all_units=False):
        """ScalableTSDFVolume.integrate (A.3 step 7; what `TSDF()` of N/3DM/tsdf.py:7-12 builds) on this dense
        box, which must consist of whole units aligned to the world unit grid (origin = k * unit_length).
        z_restart = 0 (default): Open3D's literal float32 recurrence inside every unit, from the unit's z = 0."""
        d = np.ascontiguousarray(depth_f32, dtype=np.float32)
        H, W = d.shape
        Kd = np.ascontiguousarray(K, dtype=np.float64)
        E = np.ascontiguousarray(extrinsic, dtype=np.float64)
        M = np.zeros(16)
        lib().orc_invert4x4(_ptr(E), _ptr(M))      # Eigen's 4x4 inverse is the cofactor formula
        c = None if rgb is None else np.ascontiguousarray(rgb, dtype=np.uint8)
        ul = self.voxel_length * unit_res
        u0 = np.rint(self.origin / ul)
        if self.gz0 or np.abs(u0 * ul - self.origin).max() > 1e-9 * max(1.0, np.abs(self.origin).max()) or any(
                n % unit_res for n in (self.nx, self.ny, self.nz)):
            raise ValueError("scalable mode needs a box of whole units on the world unit grid")
        u0 = np.ascontiguousarray(u0, dtype=np.int32)
        touched = np.zeros((self.nx // unit_res) * (self.ny // unit_res) * (self.nz // unit_res), np.uint8)
        lib().orc_set_scalable_all_units(1 if all_units else 0)      # all_units: every unit, no activation test (per-unit arithmetic only)
        n = lib().orc_scalable_integrate(_ptr(self.tsdf), _ptr(self.weight), _ptr(self.color if c is not None else None),
                                         self.nx, self.ny, self.nz, _ptr(u0), int(unit_res), int(stride), self.voxel_length,
                                         self.sdf_trunc, _ptr(d), _ptr(c), W, H, _ptr(Kd), _ptr(E), _ptr(M), int(z_restart), _ptr(touched))
        return (n, touched.reshape(self.nx // unit_res, self.ny // unit_res, self.nz // unit_res)) if return_touched else n

    def integrate_vbg(self, depth_u16, K, pose, rgb=None, depth_scale=1000.0, depth_max=3.0, trunc_voxel_multiplier=8.0, return_touched=False):
        """`MAP.integrate` (N/3DM/tsdf.py:71-83): Open3D tensor VoxelBlockGrid semantics (16^3 blocks, depth-touch
        activation, projective sdf, depth_max) on this dense box; `pose` = camera->world 4x4 (T_frame_to_model);
        voxel_length is the voxel_size, the box origin must sit on the world block grid.  `sdf_trunc` is unused."""
        d = np.ascontiguousarray(depth_u16, dtype=np.uint16)
        H, W = d.shape
        Kd = np.ascontiguousarray(K, dtype=np.float64)
        P = np.ascontiguousarray(pose, dtype=np.float64)
        c = None if rgb is None else np.ascontiguousarray(rgb, dtype=np.uint8)
        bs = self.voxel_length * 16
        b0 = np.rint(self.origin / bs)
        if self.gz0 or np.abs(b0 * bs - self.origin).max() > 1e-9 * max(1.0, np.abs(self.origin).max()):
            raise ValueError("the box origin must sit on the world block grid (multiples of 16 voxels)")
        b0 = np.ascontiguousarray(b0, dtype=np.int32)
        nb = [(n + 15) // 16 for n in (self.nx, self.ny, self.nz)]
        touched = np.zeros(nb[0] * nb[1] * nb[2], np.uint8)
        n = lib().orc_vbg_integrate(_ptr(self.tsdf), _ptr(self.weight), _ptr(self.color if c is not None else None), self.nx, self.ny, self.nz,
                                    _ptr(b0), self.voxel_length, float(trunc_voxel_multiplier), _ptr(d), _ptr(c), W, H, _ptr(Kd), _ptr(P),
                                    float(depth_scale), float(depth_max), _ptr(touched))
        return (n, touched.reshape(nb)) if return_touched else n

    def grid(self, name="tsdf"):
        return getattr(self, name).reshape(self.nx, self.ny, self.nz)

    def occupied(self) -> int:
        return lib().orc_count_occupied(_ptr(self.weight), self.weight.size)

    def extract_mesh(self):
        cap_v, cap_t = 1 << 16, 1 << 17
        while True:
            v = np.empty((cap_v, 3), np.float64)
            key = np.empty((cap_v, 4), np.int32)
            col = np.empty((cap_v, 3), np.float64) if self.color is not None else None
            t = np.empty((cap_t, 3), np.int32)
            tz = np.empty(cap_t, np.int32)
            cnt = np.zeros(2, np.int64)
            lib().orc_extract_mesh(_ptr(self.tsdf), _ptr(self.weight), _ptr(self.color), self.nx, self.ny, self.nz,
                                   self.gz0, self.voxel_length, _ptr(self.origin), _ptr(v), _ptr(key), _ptr(col),
                                   cap_v, _ptr(t), _ptr(tz), cap_t, _ptr(cnt))
            nv, nt = int(cnt[0]), int(cnt[1])
            if nv <= cap_v and nt <= cap_t:
                return {"vertices": v[:nv], "keys": key[:nv], "triangles": t[:nt], "triangle_z": tz[:nt],
                        "colors": None if col is None else col[:nv]}
            cap_v, cap_t = max(cap_v, nv), max(cap_t, nt)

    def extract_points(self, normals=True):
        cap = 1 << 16
        while True:
            p = np.empty((cap, 3), np.float64)
            nrm = np.empty((cap, 3), np.float64) if normals else None
            col = np.empty((cap, 3), np.float64) if self.color is not None else None
            key = np.empty((cap, 4), np.int32)
            n = lib().orc_extract_points(_ptr(self.tsdf), _ptr(self.weight), _ptr(self.color), self.nx, self.ny,
                                         self.nz, self.gz0, self.voxel_length, _ptr(self.origin), _ptr(p), _ptr(nrm),
                                         _ptr(col), _ptr(key), cap)
            if n <= cap:
                return {"points": p[:n], "normals": None if nrm is None else nrm[:n],
                        "colors": None if col is None else col[:n], "keys": key[:n]}
            cap = int(n)
