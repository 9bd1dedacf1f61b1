"""K1 / K2 / a4 parity: CUDA colorize, metric scaling, depth conversion and back-projection vs
the NumPy / C oracle and the reference's golden pair -- `-m gpu`.

Bar: u16 depth, LUT indices, RGBA bytes, valid-point counts and order bit-exact; percentiles
equal to numpy's f64 result; coordinates within 1e-4 * metric scale."""
import os

import numpy as np
import pytest
import torch

import oracle
from bodyslam_b200 import mdem, ops
from bodyslam_b200.slam_utils import pixel_to_3d
from util import small_scene

pytestmark = pytest.mark.gpu


def test_colorize_reproduces_reference_golden_pair(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "colorize_golden.npz"))
    out = mdem.colorize(g["depth"], cmap="viridis", invalid_val=0)   # depth_map_scaling.py:71
    assert out.dtype == np.uint8 and out.shape == g["rgba"].shape
    assert np.array_equal(out, g["rgba"])
    _, _, stats = ops.colorize_u16(mdem.get_cmap_lut("viridis"), depth_u16=g["depth"], invalid_val=0, return_stats=True)
    assert stats.cpu().numpy().tolist() == [[float(g["vmin"]), float(g["vmax"])]]
    out_t = mdem.colorize(torch.from_numpy(g["depth"].astype(np.int32)), cmap="viridis", invalid_val=0)
    assert np.array_equal(out_t, g["rgba"])


def random_depth(rng, shape, lo, hi, invalid_frac=0.1, smooth=True):
    H, W = shape
    if smooth:
        yy, xx = np.mgrid[0:H, 0:W]
        d = lo + (hi - lo) * (0.5 + 0.5 * np.sin(xx / rng.uniform(20, 80) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(20, 80)))
    else:
        d = rng.uniform(lo, hi, size=shape)
    d = d.astype(np.uint16)
    d[rng.uniform(size=shape) < invalid_frac] = 0
    return d


@pytest.mark.parametrize("cmap", ["viridis", "gray_r", "jet"])
@pytest.mark.parametrize("case", ["smooth", "wide", "tiny", "constant", "no_invalid"])
def test_colorize_matches_numpy_oracle(cuda, cmap, case):
    rng = np.random.default_rng(hash((cmap, case)) % 2**32)
    if case == "smooth":
        d = random_depth(rng, (480, 640), 300, 520)
    elif case == "wide":
        d = random_depth(rng, (123, 457), 1, 65535, smooth=False)   # exceeds the shared histogram window
    elif case == "tiny":
        d = random_depth(rng, (3, 5), 100, 200, invalid_frac=0.3, smooth=False)
        d[0, 0] = 150
    elif case == "constant":
        d = np.full((64, 64), 777, np.uint16)
        d[::7, ::5] = 0
    else:
        d = random_depth(rng, (240, 320), 5000, 9000, invalid_frac=0.0)
    inv = -99 if case == "no_invalid" else 0
    lut = mdem.get_cmap_lut(cmap)
    ref, idx, vmin, vmax = oracle.mdem.colorize(d, lut, invalid_val=inv, return_index=True)
    out = mdem.colorize(d, cmap=cmap, invalid_val=inv)
    assert np.array_equal(out, ref)
    ref_g = oracle.mdem.colorize(d, lut, invalid_val=inv, gamma_corrected=True, vmin=float(vmin) - 3, vmax=float(vmax) + 10.5)
    out_g = mdem.colorize(d, cmap=cmap, invalid_val=inv, gamma_corrected=True, vmin=float(vmin) - 3, vmax=float(vmax) + 10.5)
    assert np.array_equal(out_g, ref_g)


def test_colorize_value_transform_and_mask(cuda):
    rng = np.random.default_rng(7)
    d = random_depth(rng, (200, 300), 400, 900)
    lut = mdem.get_cmap_lut("viridis")
    tf = lambda x: np.sqrt(np.clip(x, 0, None))
    ref = oracle.mdem.colorize(d, lut, invalid_val=0, value_transform=tf)
    assert np.array_equal(mdem.colorize(d, cmap="viridis", invalid_val=0, value_transform=tf), ref)
    m = rng.uniform(size=d.shape) < 0.2
    ref = oracle.mdem.colorize(d, lut, invalid_mask=m)
    assert np.array_equal(mdem.colorize(d, cmap="viridis", invalid_mask=m), ref)
    with pytest.raises(RuntimeError):
        mdem.colorize(np.linspace(0, 1, 100).reshape(10, 10), cmap="viridis")


def test_fused_scale_colorize_batch(cuda):
    """BASELINE config 3 shape (reduced): [B,H,W] f32 metres -> u16 (x256) + RGBA, per-image percentiles"""
    rng = np.random.default_rng(3)
    B, H, W = 5, 270, 480
    depth = np.stack([rng.uniform(0.3, 2.0 + i, size=(H, W)).astype(np.float32) for i in range(B)])
    depth[:, ::9, ::7] = 0.0
    rgba, u16 = mdem.process_depth_batch(torch.from_numpy(depth), colormap="viridis", invalid_val=0)
    ref_u16 = oracle.mdem.scale_to_u16(depth)
    assert np.array_equal(u16.cpu().numpy(), ref_u16)
    lut = mdem.get_cmap_lut("viridis")
    for b in range(B):
        assert np.array_equal(rgba[b].cpu().numpy(), oracle.mdem.colorize(ref_u16[b], lut, invalid_val=0))
    assert np.array_equal(ops.scale_to_u16(depth).cpu().numpy(), ref_u16)


def test_metric_scaling_fixture_format(cuda, golden_dir):
    """a1 is pinned in FORMAT by the reference fixtures (I;16, metres*256): round-trip them"""
    fx = np.load(os.path.join(golden_dir, "zoedepth_u16_fixtures.npz"))
    for k in ("output_depth_map", "expected_output"):
        u = fx[k]
        metres = (u.astype(np.float32) + 0.5) / 256.0
        assert np.array_equal(ops.scale_to_u16(metres).cpu().numpy(), u)
        pil = mdem.depth_tensor_to_pil(torch.from_numpy(metres))
        assert pil.mode == "I;16" and np.array_equal(np.array(pil), u)


def test_minmax_jet_preview_and_median(cuda):
    import cv2
    rng = np.random.default_rng(11)
    d = random_depth(rng, (480, 600), 20, 450, invalid_frac=0.02)
    norm = oracle.mdem.minmax_u8(d)
    ref = cv2.applyColorMap(norm, cv2.COLORMAP_JET)
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(-1, 1), cv2.COLORMAP_JET).reshape(256, 3)
    gray, bgr = ops.minmax_colormap(d, lut_bgr=lut)
    assert np.array_equal(gray.cpu().numpy(), norm)
    assert np.array_equal(bgr.cpu().numpy(), ref)
    for n in (d.size, d.size - 1):
        x = d.reshape(-1)[:n]
        assert float(ops.median_u16(x)[0].item()) == float(np.median(x))
    gt = rng.uniform(100, 60000, size=5000).astype(np.uint16)
    assert mdem.compute_median_scale_factor(gt, d) == oracle.mdem.compute_median_scale_factor(gt, d)
    gtf = rng.uniform(0.1, 3.0, size=4001)
    assert abs(mdem.compute_median_scale_factor(gtf, gtf * 2.0) - 0.5) < 1e-12


def test_depth_from_u16_exact(cuda):
    rng = np.random.default_rng(5)
    u = rng.integers(0, 65536, size=(3, 97, 131)).astype(np.uint16)
    for scale, trunc in ((1000.0, 3.0), (256.0, 10.0), (1000.0, 0.0)):
        ref = oracle.o3d.depth_from_u16(u, scale, trunc if trunc > 0 else 1e30)
        out = ops.depth_from_u16(u, scale, trunc).cpu().numpy()
        assert np.array_equal(out, ref)


@pytest.mark.parametrize("stride", [1, 8])
def test_backproject_matches_oracle(cuda, stride):
    sc = small_scene("laparoscopy512", res=64, frames=3)
    depth = oracle.o3d.depth_from_u16(sc["depth_u16"])
    xyz, rgb, counts = ops.backproject(depth, sc["K"], sc["E"], color=sc["color"], stride=stride, return_counts=True)
    off = 0
    for i in range(3):
        ref, refc = oracle.o3d.backproject(depth[i], sc["K"], sc["E"][i], rgb=sc["color"][i], stride=stride)
        n = int(counts[i].item())
        assert n == len(ref), "valid point count differs"
        got = xyz[off:off + n].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
        assert np.abs(rgb[off:off + n].cpu().numpy() - refc).max() <= 1e-6
        off += n
    assert off == xyz.shape[0]
    # organised (valid_only=False) output keeps one row per visited pixel, NaN where invalid
    xyz_all, _ = ops.backproject(depth[:1], sc["K"], sc["E"][:1], stride=stride, valid_only=False)
    ref_all, _ = oracle.o3d.backproject(depth[0], sc["K"], sc["E"][0], stride=stride, valid_only=False)
    got = xyz_all.cpu().numpy()
    assert got.shape == ref_all.shape and np.array_equal(np.isnan(got), np.isnan(ref_all))
    # the reference's scalar helper agrees with the dense kernel at identity pose
    p, _ = ops.backproject(depth[0], sc["K"], None, stride=1, valid_only=False)
    u, v = 333, 222
    want = pixel_to_3d(u, v, float(depth[0][v, u]), *sc["K"])
    if depth[0][v, u] > 0:
        assert np.abs(p[v * sc["W"] + u].cpu().numpy() - want).max() < 1e-5


@pytest.mark.parametrize("case", ["metres", "wide", "tiny", "constant", "negative", "two_values", "ties"])
def test_colorize_float32_metres_matches_numpy(cuda, case):
    """`colorize` on a FLOAT32 image (ZoeDepth's metres tensor, depth_map_scaling.py:14-15), numpy and torch, CPU and
    CUDA: NumPy's float32 percentile / normalisation and matplotlib's index rule, reproduced exactly (RGBA bytes
    and vmin / vmax bit for bit) by the two-level radix select."""
    rng = np.random.default_rng(sum(map(ord, case)))
    if case == "metres":
        yy, xx = np.mgrid[0:480, 0:640]
        d = (0.4 + 1.9 * (0.5 + 0.5 * np.sin(xx / 37.0) * np.cos(yy / 53.0)) + rng.normal(0, 1e-3, (480, 640))).astype(np.float32)
        d[rng.uniform(size=d.shape) < 0.05] = 0
    elif case == "wide":
        d = np.exp(rng.uniform(-20, 20, (123, 457))).astype(np.float32)
        d[::9, ::4] = 0
    elif case == "tiny":
        d = rng.uniform(0.1, 3.0, (3, 5)).astype(np.float32)
        d[1, 1] = 0
    elif case == "constant":
        d = np.full((64, 64), 1.2345, np.float32)
        d[::7, ::5] = 0
    elif case == "negative":
        d = rng.normal(0.0, 5.0, (200, 300)).astype(np.float32)
        d[d == 0] = 1e-3
        d[::11, ::3] = 0
    elif case == "two_values":
        d = np.where(rng.uniform(size=(100, 100)) < 0.5, 1.0, 2.5).astype(np.float32)
    else:   # many exact ties around the percentile ranks
        d = np.round(rng.uniform(0.5, 2.0, (240, 320)), 2).astype(np.float32)
        d[::5, ::5] = 0
    lut = mdem.get_cmap_lut("viridis")
    ref, idx, vmin, vmax = oracle.mdem.colorize(d.copy(), lut, invalid_val=0, return_index=True)
    out = mdem.colorize(d.copy(), cmap="viridis", invalid_val=0)
    assert out.dtype == np.uint8 and out.shape == ref.shape
    _, stats = ops.colorize_f32(lut, d, invalid_val=0, return_stats=True)
    assert stats.cpu().numpy().tolist() == [[float(vmin), float(vmax)]], (stats.cpu().numpy().tolist(), vmin, vmax)
    assert np.array_equal(out, ref)
    # torch CPU / CUDA tensors with singleton dims, like a network output [1,1,H,W]
    for t in (torch.from_numpy(d)[None, None], torch.from_numpy(d).to(cuda)[None]):
        assert np.array_equal(mdem.colorize(t, cmap="viridis", invalid_val=0), ref)
    # explicit vmin / vmax, gamma, an explicit invalid mask
    ref2 = oracle.mdem.colorize(d.copy(), lut, vmin=0.5, vmax=2.0, invalid_val=0, gamma_corrected=True)
    assert np.array_equal(mdem.colorize(d.copy(), vmin=0.5, vmax=2.0, cmap="viridis", invalid_val=0, gamma_corrected=True), ref2)
    m = rng.uniform(size=d.shape) < 0.2
    ref3 = oracle.mdem.colorize(d.copy(), lut, invalid_mask=m)
    assert np.array_equal(mdem.colorize(d.copy(), cmap="viridis", invalid_mask=m), ref3)


def test_colorize_float32_batch_is_per_image(cuda):
    rng = np.random.default_rng(3)
    d = rng.uniform(0.2, 4.0, (5, 96, 128)).astype(np.float32) * np.arange(1, 6, dtype=np.float32)[:, None, None]
    d[:, ::6, ::7] = 0
    lut = mdem.get_cmap_lut("jet")
    out = ops.colorize_f32(lut, d, invalid_val=0).cpu().numpy()
    for b in range(5):
        assert np.array_equal(out[b], oracle.mdem.colorize(d[b].copy(), lut, invalid_val=0))
