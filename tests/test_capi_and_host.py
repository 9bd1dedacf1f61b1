"""C-ABI surface and host-side logic that need no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from bodyslam_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "bodyslam_b200.h")).read()
    declared = set(re.findall(r"BSLAM_API\s+[\w\s\*]+?\b(bslam_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert L.bslam_version() >= 100
    assert L.bslam_device_count() >= 0
    assert L.bslam_tsdf_storage_bytes(512, 512, 512, 0) >= 512 ** 3 * 8
    assert L.bslam_tsdf_storage_bytes(60, 52, 44, 1) >= 8 * 7 * 6 * 512 * 20
    assert L.bslam_colorize_workspace_bytes(2) > 2 * 65536 * 5
    assert L.bslam_backproject_workspace_bytes(1, 480, 640, 1) > 0


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bodyslam_b200 import mdem, ops
    from bodyslam_b200.tsdf import TSDF
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        TSDF()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        mdem.colorize(np.zeros((4, 4), np.uint16))
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.backproject(np.zeros((4, 4), np.float32), (1, 1, 0, 0))
    # the C ABI itself reports a CUDA error, not a silent success
    from bodyslam_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    rc = L.bslam_tsdf_create(ctypes.byref(h), 8, 8, 8, 0, 0.01, 0.04, None, 0, 0, None, None)
    assert rc == _lib.E_CUDA and b"cuda" in L.bslam_last_error().lower()


def test_argument_errors_do_not_need_a_device():
    from bodyslam_b200 import _lib
    L = _lib.load()
    h = ctypes.c_void_p()
    assert L.bslam_tsdf_create(ctypes.byref(h), 0, 8, 8, 0, 0.01, 0.04, None, 0, 0, None, None) == _lib.E_ARG
    assert L.bslam_tsdf_create(ctypes.byref(h), 8, 8, 8, 4, 0.01, 0.04, None, 0, 0, None, None) == _lib.E_ARG
    assert b"multiple of 8" in L.bslam_last_error()
    assert L.bslam_scale_u16(None, 10, 256.0, None, None) == _lib.E_ARG
    assert L.bslam_tsdf_integrate(None, None, None, 1, 4, 4, None, None, 8, None, 0, None) == _lib.E_ARG


def test_pose_and_camera_helpers_match_reference_formulas():
    from bodyslam_b200.slam_utils import compute_curr_estimate_global_pose, ensure_so3_v2, get_o3d_intrinsic, pixel_to_3d
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    R = ensure_so3_v2(A)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
    E_prev, T = np.eye(4), np.eye(4)
    E_prev[:3, :3], E_prev[:3, 3] = ensure_so3_v2(rng.normal(size=(3, 3))), [0.1, 0.2, 0.3]
    T[:3, :3], T[:3, 3] = ensure_so3_v2(np.eye(3) + 0.01 * rng.normal(size=(3, 3))), [0.001, 0, 0.002]
    G = compute_curr_estimate_global_pose(E_prev.copy(), T)
    assert np.allclose(G, E_prev @ T, atol=1e-12)
    intr, Kt = get_o3d_intrinsic(600, 480, 383.19, 383.19, 276.47, 124.33)
    assert (intr.width, intr.height) == (600, 480) and intr.get_focal_length() == (383.19, 383.19)
    assert intr.get_principal_point() == (276.47, 124.33) and np.array_equal(Kt, intr.intrinsic_matrix)
    assert np.allclose(pixel_to_3d(300, 200, 0.25, 383.19, 383.19, 276.47, 124.33),
                       [(300 - 276.47) * 0.25 / 383.19, (200 - 124.33) * 0.25 / 383.19, 0.25])


def test_ply_writers_round_trip(tmp_path):
    from bodyslam_b200.geometry import PointCloud, TriangleMesh
    from bodyslam_b200.io import read_ply, write_point_cloud, write_triangle_mesh
    rng = np.random.default_rng(1)
    v = rng.normal(size=(10, 3)).astype(np.float32)
    t = rng.integers(0, 10, size=(7, 3)).astype(np.int32)
    c = rng.uniform(size=(10, 3)).astype(np.float32)
    write_triangle_mesh(str(tmp_path / "m.ply"), TriangleMesh(v, t, c))
    verts, faces = read_ply(str(tmp_path / "m.ply"))
    assert np.allclose(np.stack([verts["x"], verts["y"], verts["z"]], 1), v) and np.array_equal(faces, t)
    assert np.array_equal(verts["red"], np.floor(c[:, 0].astype(np.float64) * 255 + 0.5).astype(np.uint8))
    write_point_cloud(str(tmp_path / "p.ply"), PointCloud(v, None, v))
    verts, faces = read_ply(str(tmp_path / "p.ply"))
    assert faces is None and np.allclose(verts["nz"], v[:, 2])


def test_ply_bytes_follow_open3d_writer_layout(tmp_path):
    """o3d.io.write_triangle_mesh / write_point_cloud (N/3DM/tsdf.py:37,52) for a `.ply` path: rply binary
    little-endian, comment "Created by Open3D", double coordinates (+ double normals), uchar colours =
    round(clamp(c) * 255), faces as uchar count + uint32 indices -- checked byte by byte against a hand-packed file"""
    import struct
    from bodyslam_b200.geometry import PointCloud, TriangleMesh
    from bodyslam_b200.io import write_point_cloud, write_triangle_mesh
    v = np.array([[0.0, 1.5, -2.25], [1e-3, 2.0, 3.0], [4.0, 5.0, 6.0]], np.float32)
    c = np.array([[0.0, 0.5, 1.0], [0.2, 0.999, 1.7], [-0.1, 0.25, 0.75]], np.float32)
    t = np.array([[0, 1, 2], [2, 1, 0]], np.int32)
    write_triangle_mesh(str(tmp_path / "m.ply"), TriangleMesh(v, t, c))
    head = (b"ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\nelement vertex 3\nproperty double x\nproperty double y\n"
            b"property double z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face 2\n"
            b"property list uchar uint vertex_indices\nend_header\n")
    body = b""
    for p, col in zip(v, c):
        body += struct.pack("<3d", *[float(x) for x in p])
        body += struct.pack("<3B", *[int(np.floor(min(1.0, max(0.0, float(x))) * 255.0 + 0.5)) for x in col])
    for tri in t:
        body += struct.pack("<B3I", 3, *[int(i) for i in tri])
    assert open(tmp_path / "m.ply", "rb").read() == head + body
    n = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0]], np.float32)
    write_point_cloud(str(tmp_path / "p.ply"), PointCloud(v, c, n))
    head = (b"ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\nelement vertex 3\nproperty double x\nproperty double y\n"
            b"property double z\nproperty double nx\nproperty double ny\nproperty double nz\nproperty uchar red\nproperty uchar green\n"
            b"property uchar blue\nend_header\n")
    body = b""
    for p, nn, col in zip(v, n, c):
        body += struct.pack("<6d", *[float(x) for x in p], *[float(x) for x in nn])
        body += struct.pack("<3B", *[int(np.floor(min(1.0, max(0.0, float(x))) * 255.0 + 0.5)) for x in col])
    assert open(tmp_path / "p.ply", "rb").read() == head + body


def test_depth_estimator_interface_keeps_reference_surface(tmp_path):
    """signatures / error behaviour of R/src/depth_estimation/interface.py without the network"""
    import inspect
    from PIL import Image
    from bodyslam_b200.mdem import DepthEstimator, MDEMInterface, colorize, process_image, process_images
    assert DepthEstimator.SUPPORTED_MODELS == ['ZoeD_N', 'ZoeD_K', 'ZoeD_NK'] and DepthEstimator.DEFAULT_MODEL == 'ZoeD_NK'
    sig = inspect.signature(colorize)
    assert list(sig.parameters) == ["value", "vmin", "vmax", "cmap", "invalid_val", "invalid_mask", "background_color",
                                    "gamma_corrected", "value_transform"]
    assert sig.parameters["cmap"].default == "gray_r" and sig.parameters["invalid_val"].default == -99
    assert list(inspect.signature(process_image).parameters)[:5] == ["estimator", "input_path", "output_path", "colormap", "invalid_val"]
    assert list(inspect.signature(process_images).parameters)[:4] == ["input_dir", "output_dir", "colormap", "invalid_val"]
    est = DepthEstimator(model=object())
    assert est.model is not None
    p = tmp_path / "in.png"
    Image.new("L", (8, 6)).save(p)
    assert DepthEstimator.load_image(str(p)).mode == "RGB"
    DepthEstimator.save_depth_map(Image.new("I;16", (8, 6)), str(tmp_path / "d.xyz"), extension=".png")
    assert (tmp_path / "d.png").exists()                         # extension REPLACED (interface.py:83-84)
    MDEMInterface.save_depth_map(Image.new("I;16", (8, 6)), str(tmp_path / "e"), ".png")
    assert (tmp_path / "e.png").exists()                         # extension APPENDED (io_utils.py:70)


def test_synthetic_scenes_are_deterministic_and_in_range():
    from bodyslam_b200 import synthetic as S
    for name in ("laparoscopy512", "colonoscopy256", "gastroscopy1024"):
        cfg = S.config(name)
        E = cfg["extrinsics"](cfg["frames"])
        assert E.shape == (cfg["frames"], 4, 4)
        R = E[:, :3, :3]
        assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-9)
        d1, c1 = S.render(cfg["surface"], E[[0, cfg["frames"] // 2]], W=160, H=120, K=tuple(k / 4 for k in cfg["K"]), device="cpu")
        d2, _ = S.render(cfg["surface"], E[[0, cfg["frames"] // 2]], W=160, H=120, K=tuple(k / 4 for k in cfg["K"]), device="cpu")
        d1, d2 = d1.numpy(), d2.numpy()
        assert np.array_equal(d1, d2) and d1.dtype == np.uint16 and c1.shape == (2, 120, 160, 3)
        valid = d1 > 0
        assert 0.9 < valid.mean() < 0.995 and d1[valid].min() >= 5 and d1.max() <= 700


def test_stream_chunks_cover_all_frames_in_order():
    from bodyslam_b200.tsdf import DenseTSDFVolume as D

    for F in (1, 7, 255, 256, 257, 600, 1000, 5000):
        for chunk in (8, 64, 256):
            ch = D.stream_chunks(F, chunk)
            assert ch[0][0] == 0 and ch[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            assert all(0 < f1 - f0 <= chunk for f0, f1 in ch)
    assert D.stream_chunks(1000, 256, ramp=()) == [(0, 250), (250, 500), (500, 750), (750, 1000)]
    assert [b - a for a, b in D.stream_chunks(1000)][:3] == [32, 64, 128]


def test_sharded_ingest_ownership_partitions_every_chunk():
    """ShardedTSDF.ingest_pieces / ingest_share: every frame is fed by exactly one rank, pieces are
    contiguous, in rank order and cover each chunk; chunk sizes divide evenly except possibly the last"""
    from bodyslam_b200.sharding import ShardedTSDF as S
    from bodyslam_b200.tsdf import DenseTSDFVolume as D

    for F in (1, 5, 6, 40, 255, 256, 257, 1000, 5000):
        for N in (2, 3, 4, 8):
            chunks, pieces = S.ingest_pieces(F, N, 256)
            assert chunks[0][0] == 0 and chunks[-1][1] == F and all(0 < b - a <= 256 for a, b in chunks)
            for (f0, f1), pc in zip(chunks, pieces):
                assert len(pc) == N and pc[0][0] == f0 and pc[-1][1] == f1
                assert all(a[1] == b[0] for a, b in zip(pc, pc[1:])) and all(b >= a for a, b in pc)
            assert all((b - a) % N == 0 for a, b in chunks[:-1])
            shares = [S.ingest_share(F, r, N, 256) for r in range(N)]
            assert np.array_equal(np.sort(np.concatenate(shares)), np.arange(F))
            assert all(np.all(np.diff(s) > 0) for s in shares if len(s) > 1)
    assert D.stream_chunks(1000, 256, multiple_of=8)[:3] == [(0, 32), (32, 96), (96, 224)]
    assert S.stream_ramp(1, True) == () and S.stream_ramp(8, True) == (64,) and S.stream_ramp(8, False) == (32, 64, 128)


def test_inverse4x4_is_the_cofactor_formula_the_oracle_uses():
    """camera poses handed to back-projection / unit activation: Eigen-style cofactor inverse (host helper of the C ABI)"""
    import oracle
    from bodyslam_b200.geometry import inverse4x4
    rng = np.random.default_rng(0)
    E = rng.normal(size=(16, 4, 4))
    E[:, 3] = [0, 0, 0, 1]
    inv = inverse4x4(E)
    assert inv.shape == E.shape and np.abs(inv - np.linalg.inv(E)).max() < 1e-9
    for k in range(16):
        M = np.zeros(16)
        oracle.o3d.lib().orc_invert4x4(np.ascontiguousarray(E[k]).ctypes.data, M.ctypes.data)
        assert np.abs(M.reshape(4, 4) - inv[k]).max() <= 4 * np.finfo(np.float64).eps * np.abs(inv[k]).max()
    assert np.array_equal(inverse4x4(np.eye(4)), np.eye(4))


def test_map_and_tsdf_signatures_match_the_reference():
    import inspect
    from bodyslam_b200.tsdf import MAP, TSDF
    p = inspect.signature(MAP.__init__).parameters
    assert list(p)[1:9] == ["width", "height", "intrinsic", "device", "depth_scale", "voxel_size", "block_count", "trunc_voxel_multiplier"]
    assert (p["voxel_size"].default, p["block_count"].default, p["trunc_voxel_multiplier"].default) == (0.0058, 40000, 8.0)
    assert list(inspect.signature(MAP.integrate).parameters)[1:] == ["curr_rgbd", "i", "curr_global_pose"]
    p = inspect.signature(TSDF.__init__).parameters
    assert list(p)[1:3] == ["voxel_length", "sdf_trunc"] and (p["voxel_length"].default, p["sdf_trunc"].default) == (0.001, 0.1)
    for cls in (MAP, TSDF):
        for m in ("extract_pcd", "extract_mesh", "save_pcd", "save_mesh"):
            assert callable(getattr(cls, m))
    assert list(inspect.signature(TSDF.build_3D_map).parameters)[1:] == ["rgbd", "intrinsic", "extrinsic"]
    assert list(inspect.signature(TSDF.build_copy_3D_map).parameters)[1:] == ["rgbd", "intrinsic", "extrinsic"]


def test_stream_chunks_cover_the_trajectory_in_order():
    from bodyslam_b200.tsdf import DenseTSDFVolume
    for F in (1, 5, 255, 256, 257, 1000, 5000):
        for ramp in ((), (64,), (32, 64, 128)):
            for q in (1, 2, 8):
                ch = DenseTSDFVolume.stream_chunks(F, 256, ramp=ramp, multiple_of=q)
                assert ch[0][0] == 0 and ch[-1][1] == F and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
                assert all(0 < f1 - f0 <= 256 + q for f0, f1 in ch)
                assert all((f1 - f0) % q == 0 for f0, f1 in ch[:-1])
