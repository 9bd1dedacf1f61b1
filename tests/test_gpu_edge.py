"""Edge cases of the hot path through the C ABI -- `-m gpu`: empty and ragged inputs, batch-size boundaries, frames that
see nothing, images whose width is not a multiple of the vector width, misaligned views."""
import numpy as np
import pytest
import torch

import oracle
from bodyslam_b200 import mdem, ops
from bodyslam_b200.geometry import PinholeCameraIntrinsic
from bodyslam_b200.tsdf import TSDF, DenseTSDFVolume
from util import small_scene

pytestmark = pytest.mark.gpu


def test_empty_inputs_and_empty_volume(cuda):
    sc = small_scene("laparoscopy512", res=32, frames=2, W=64, H=48, with_color=False)
    vol = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 32, sc["origin"], color=False, device=cuda)
    # nothing integrated yet: extraction returns empty geometry, not an error
    m, p = vol.extract_triangle_mesh(), vol.extract_point_cloud()
    assert m.vertices.shape == (0, 3) and m.triangles.shape == (0, 3) and p.points.shape == (0, 3)
    # F = 0
    vol.integrate_batch(torch.empty((0, 48, 64), device=cuda), None, sc["intrinsic"], np.zeros((0, 4, 4)))
    vol.integrate_u16_batch(torch.empty((0, 48, 64), dtype=torch.uint16, device=cuda), None, sc["intrinsic"], np.zeros((0, 4, 4)))
    assert int((vol.export_dense()[1] != 0).sum()) == 0
    # frames that see nothing: all-invalid depth, and a camera looking away from the box
    depth = ops.depth_from_u16(sc["depth_u16"], 1000.0, 3.0, cuda)
    counts = torch.zeros(2, dtype=torch.int64, device=cuda)
    vol.integrate_batch(torch.zeros_like(depth), None, sc["intrinsic"], sc["E"], update_counts=counts)
    away = sc["E"].copy()
    away[:, :3, 3] += np.array([0.0, 0.0, -50.0])         # the box ends up 50 m behind the camera
    vol.integrate_batch(depth, None, sc["intrinsic"], away, update_counts=counts)
    assert counts.cpu().tolist() == [0, 0] and int((vol.export_dense()[1] != 0).sum()) == 0
    # back-projection of an all-invalid image: zero rows
    xyz, _ = ops.backproject(torch.zeros((1, 48, 64), device=cuda), sc["K"])
    assert xyz.shape == (0, 3)


@pytest.mark.parametrize("W,H", [(37, 23), (66, 50), (8, 8)])
def test_odd_image_sizes_match_oracle(cuda, W, H):
    """widths that are not a multiple of 4 take the scalar paths of the fused conversion / tile-max pass"""
    sc = small_scene("laparoscopy512", res=32, frames=3, W=W, H=H)
    vol = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 32, sc["origin"], color=True, device=cuda)
    counts = torch.zeros(3, dtype=torch.int64, device=cuda)
    vol.integrate_u16_batch(torch.from_numpy(sc["depth_u16"]).to(cuda), torch.from_numpy(sc["color"]).to(cuda), sc["intrinsic"], sc["E"],
                            update_counts=counts)
    V = oracle.o3d.Volume(32, sc["voxel_length"], sc["sdf_trunc"], sc["origin"], with_color=True)
    co = [V.integrate(oracle.o3d.depth_from_u16(sc["depth_u16"][i]), sc["K"], sc["E"][i], rgb=sc["color"][i]) for i in range(3)]
    t, w = (x.cpu().numpy() for x in vol.export_dense())
    assert counts.cpu().tolist() == co and np.array_equal(w, V.grid("weight")) and np.array_equal(t, V.grid("tsdf"))
    # K1 / K2 on the same odd shapes
    lut = mdem.get_cmap_lut("viridis")
    assert np.array_equal(mdem.colorize(sc["depth_u16"][0], cmap="viridis", invalid_val=0), oracle.mdem.colorize(sc["depth_u16"][0], lut, invalid_val=0))
    d = ops.depth_from_u16(sc["depth_u16"][:1], 1000.0, 3.0, cuda)
    xyz, _ = ops.backproject(d, sc["K"], sc["E"][:1])
    ref, _ = oracle.o3d.backproject(d[0].cpu().numpy(), sc["K"], sc["E"][0])
    assert xyz.shape[0] == len(ref) and (len(ref) == 0 or np.abs(xyz.cpu().numpy() - ref).max() < 1e-4)


@pytest.mark.parametrize("F", [255, 256, 257, 513])
def test_batch_size_boundaries(cuda, F):
    """launches hold at most BSLAM_MAX_BATCH = 256 frames: 255 / 256 / 257 / 513 frames give the frame-by-frame result"""
    sc = small_scene("colonoscopy256", res=32, frames=F, W=80, H=60, with_color=False)
    vol = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 32, sc["origin"], color=False, device=cuda)
    counts = torch.zeros(F, dtype=torch.int64, device=cuda)
    vol.integrate_u16_batch(torch.from_numpy(sc["depth_u16"]).to(cuda), None, sc["intrinsic"], sc["E"], update_counts=counts)
    V = oracle.o3d.Volume(32, sc["voxel_length"], sc["sdf_trunc"], sc["origin"])
    co = [V.integrate(oracle.o3d.depth_from_u16(sc["depth_u16"][i]), sc["K"], sc["E"][i]) for i in range(F)]
    t, w = (x.cpu().numpy() for x in vol.export_dense())
    assert counts.cpu().tolist() == co and np.array_equal(w, V.grid("weight")) and np.array_equal(t, V.grid("tsdf"))


def test_misaligned_depth_view_takes_the_scalar_path(cuda):
    """a float32 depth batch whose base is only 4-byte aligned (a view into a larger buffer) is integrated correctly"""
    sc = small_scene("laparoscopy512", res=32, frames=2, W=64, H=48, with_color=False)
    d = ops.depth_from_u16(sc["depth_u16"], 1000.0, 3.0, cuda)
    buf = torch.zeros(d.numel() + 1, dtype=torch.float32, device=cuda)
    buf[1:] = d.reshape(-1)
    view = buf[1:].view(d.shape)
    assert view.data_ptr() % 16 != 0
    a = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 32, sc["origin"], color=False, device=cuda)
    b = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 32, sc["origin"], color=False, device=cuda)
    a._L.bslam_tsdf_integrate  # the C entry is reached directly: as_cuda() would re-pack a non-contiguous tensor, a view like this one is contiguous
    a.integrate_batch(view, None, sc["intrinsic"], sc["E"])
    b.integrate_batch(d, None, sc["intrinsic"], sc["E"])
    for x, y in zip(a.export_dense(), b.export_dense()):
        assert torch.equal(x, y)


def test_colorize_degenerate_images(cuda):
    lut = mdem.get_cmap_lut("viridis")
    one = np.array([[1234]], np.uint16)
    assert np.array_equal(mdem.colorize(one, cmap="viridis", invalid_val=0), oracle.mdem.colorize(one, lut, invalid_val=0))
    onef = np.array([[1.5]], np.float32)
    assert np.array_equal(mdem.colorize(onef, cmap="viridis", invalid_val=0), oracle.mdem.colorize(onef.copy(), lut, invalid_val=0))
    # every pixel invalid: the reference's np.percentile raises on the empty selection; here the image is all background
    inv = np.zeros((4, 5), np.uint16)
    out = mdem.colorize(inv, cmap="viridis", invalid_val=0)
    assert out.shape == (4, 5, 4) and np.all(out == np.array([128, 128, 128, 255], np.uint8))
    outf = mdem.colorize(np.zeros((4, 5), np.float32), cmap="viridis", invalid_val=0)
    assert np.all(outf == np.array([128, 128, 128, 255], np.uint8))


def test_drop_in_tsdf_rejects_mismatched_frames_like_open3d(cuda):
    sc = small_scene("laparoscopy512", res=32, frames=1, W=64, H=48)
    t = TSDF(voxel_length=sc["voxel_length"], sdf_trunc=sc["sdf_trunc"], resolution=32, origin=sc["origin"], device=cuda)
    from bodyslam_b200.geometry import RGBDImage
    d = ops.depth_from_u16(sc["depth_u16"][0], 1000.0, 3.0, cuda)
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        t.build_3D_map(RGBDImage(sc["color"][0], d), PinholeCameraIntrinsic(32, 48, *sc["K"]), sc["E"][0])      # intrinsic size mismatch
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        t.build_3D_map(RGBDImage(None, d), sc["intrinsic"], sc["E"][0])                                          # RGB8 volume without colour
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        t.build_3D_map(RGBDImage(sc["color"][0], sc["depth_u16"][0]), sc["intrinsic"], sc["E"][0])               # u16 depth: Open3D wants float


def test_default_box_recentres_on_the_first_frame(cuda):
    """`TSDF()` with no origin: the reference's SLAM loop starts at the identity pose looking at a surface 0.3 - 0.5 m
    ahead (depth PNG / 1000), which lies outside a box centred on the world origin -- the box follows the first frame
    (with a warning); an explicit origin is left alone"""
    from bodyslam_b200.geometry import RGBDImage
    K = (383.19, 383.19, 276.47, 124.33)             # the reference's 600x480 intrinsics (N/3DM/slam.py:25)
    intr = PinholeCameraIntrinsic(600, 480, *K)
    yy, xx = np.mgrid[0:480, 0:600]
    depth = (0.40 + 0.05 * np.sin(xx / 60.0) * np.cos(yy / 45.0)).astype(np.float32)
    color = np.full((480, 600, 3), 120, np.uint8)
    t = TSDF(device=cuda)                            # the reference's constructor call: TSDF()
    assert np.allclose(t.tsdf.origin, -0.256)
    with pytest.warns(UserWarning, match="re-centred"):
        t.build_3D_map(RGBDImage(color, depth), intr, np.eye(4))
    assert t.tsdf.unit_activation and abs(t.tsdf.origin[2] + 0.256 - 0.40) < 0.05
    st = t.tsdf.clip_stats()
    assert st["outside"] < 0.35 * st["points"]       # the 0.512 m box holds the central part of the 0.6 m wide view
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert t.extract_mesh().triangles.shape[0] > 10000
    # an explicit origin is respected (and the clip warning tells the user what was lost)
    t2 = TSDF(origin=(-0.256,) * 3, device=cuda)
    t2.build_3D_map(RGBDImage(color, depth), intr, np.eye(4))
    assert np.allclose(t2.tsdf.origin, -0.256)
    with pytest.warns(UserWarning, match="outside the volume box"):
        t2.extract_pcd()


def test_image_ops_take_misaligned_views(cuda):
    """a contiguous view into a larger buffer (base only 4- / 2-byte aligned) gives the result of an aligned copy: the
    Python ops re-pack it, and the C entry points refuse the raw pointer instead of faulting in a vector load"""
    import ctypes
    from bodyslam_b200 import _lib
    rng = np.random.default_rng(5)
    m = rng.uniform(0.05, 3.0, size=(3, 48, 64)).astype(np.float32)
    m[rng.uniform(size=m.shape) < 0.05] = 0
    lut = mdem.get_cmap_lut("viridis")
    d = torch.from_numpy(m).to(cuda)
    buf = torch.zeros(d.numel() + 1, dtype=torch.float32, device=cuda)
    buf[1:] = d.reshape(-1)
    view = buf[1:].view(d.shape)
    assert view.data_ptr() % 16 != 0 and view.is_contiguous()
    rgba_a, u16_a = ops.colorize_u16(lut, depth_m=d, invalid_val=0)
    rgba_v, u16_v = ops.colorize_u16(lut, depth_m=view, invalid_val=0)
    assert torch.equal(rgba_a, rgba_v) and torch.equal(u16_a, u16_v)
    assert torch.equal(ops.scale_to_u16(view), ops.scale_to_u16(d))
    ubuf = torch.zeros(u16_a.numel() + 1, dtype=torch.uint16, device=cuda)
    ubuf[1:] = u16_a.reshape(-1)
    uview = ubuf[1:].view(u16_a.shape)
    assert uview.data_ptr() % 8 != 0
    assert torch.equal(ops.colorize_u16(lut, depth_u16=uview, invalid_val=0)[0], rgba_a)
    assert torch.equal(ops.median_u16(uview, invalid_val=0), ops.median_u16(u16_a, invalid_val=0))
    assert torch.equal(ops.minmax_colormap(uview)[0], ops.minmax_colormap(u16_a)[0])
    # the C ABI itself: a misaligned pointer is an argument error, not a CUDA fault
    L = _lib.load()
    ws = torch.empty(L.bslam_colorize_workspace_bytes(3), dtype=torch.uint8, device=cuda)
    out = torch.empty(3, dtype=torch.float64, device=cuda)
    rc = L.bslam_median_u16(_lib.ptr(uview), 3, 48 * 64, 0, 0, _lib.ptr(out), _lib.ptr(ws), _lib.stream_ptr(cuda))
    assert rc != 0 and "aligned" in L.bslam_last_error().decode()
    torch.cuda.synchronize()
