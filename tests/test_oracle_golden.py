"""Pin the NumPy MDEM oracle against the reference's own artefacts (CPU only)."""
import os

import numpy as np

import oracle


def test_colorize_oracle_reproduces_reference_golden_pair(golden_dir):
    g = np.load(os.path.join(golden_dir, "colorize_golden.npz"))
    L = np.load(os.path.join(golden_dir, "viridis_lut.npz"))
    img, idx, vmin, vmax = oracle.mdem.colorize(g["depth"], L["lut"], invalid_val=0, return_index=True)
    assert (vmin, vmax) == (338.0, 465.0) == (float(g["vmin"]), float(g["vmax"]))
    assert img.dtype == np.uint8 and img.shape == (480, 600, 4)
    assert np.array_equal(img, g["rgba"])
    assert L["pinned"][np.unique(idx)].all()          # the golden only exercises pinned LUT rows
    assert int(L["pinned"].sum()) >= 222
    # under / over colours = first / last LUT row (matplotlib defaults)
    assert tuple(img[g["depth"] < vmin][0]) == (68, 1, 84, 255)
    assert tuple(img[g["depth"] > vmax][0]) == (253, 231, 36, 255)


def test_product_lut_equals_golden_lut(golden_dir):
    L = np.load(os.path.join(golden_dir, "viridis_lut.npz"))
    from bodyslam_b200.mdem import get_cmap_lut
    assert np.array_equal(get_cmap_lut("viridis"), L["lut"])
    assert np.array_equal(get_cmap_lut("viridis_r"), L["lut"][::-1])
    for name in ("gray", "gray_r", "jet", "jet_r"):
        assert np.array_equal(get_cmap_lut(name), oracle.mdem.segment_lut(name))
    g = get_cmap_lut("gray")[:, 0].astype(int)      # trunc(255 * i/255): i or i-1 (float rounding), like matplotlib
    assert np.all((np.arange(256) - g >= 0) & (np.arange(256) - g <= 1)) and g[0] == 0 and g[255] == 255
    assert get_cmap_lut(L["lut"][:, :3]).shape == (256, 4)


def test_metric_scaling_fixtures_are_i16_metres_times_256(golden_dir):
    fx = np.load(os.path.join(golden_dir, "zoedepth_u16_fixtures.npz"))
    for k in ("output_depth_map", "expected_output"):
        u = fx[k]
        assert u.dtype == np.uint16 and u.shape == (480, 600)
        assert 0.5 < u.min() / 256.0 and u.max() / 256.0 < 3.0        # plausible metres
        assert np.array_equal(oracle.mdem.scale_to_u16((u.astype(np.float32) + 0.5) / 256.0), u)
    assert np.array_equal(oracle.mdem.scale_to_u16(np.array([0.0, 1.0, 1.999, 255.99], np.float32)), [0, 256, 511, 65533])


def test_apply_lut_matches_matplotlib_index_rule():
    lut = np.arange(256 * 4, dtype=np.uint32).reshape(256, 4).astype(np.uint8)
    x = np.array([-0.1, 0.0, 0.5, 1.0 - 1e-12, 1.0, 1.5, np.nan, 1 / 256, 255 / 256])
    out, idx = oracle.mdem.apply_lut(x, lut)
    assert idx.tolist() == [0, 0, 128, 255, 255, 255, 0, 1, 255]
    assert tuple(out[6]) == (0, 0, 0, 0)       # "bad" colour, repainted with the background by colorize


def test_minmax_and_median():
    rng = np.random.default_rng(0)
    d = rng.integers(10, 5000, size=(50, 60)).astype(np.uint16)
    n = oracle.mdem.minmax_u8(d)
    assert n.min() == 0 and n.max() == 255 and n.dtype == np.uint8
    assert oracle.mdem.compute_median_scale_factor(d * 2, d) == 2.0


def test_oracle_output_is_pinned(golden_dir):
    """the oracle's own tsdf / weight / colour grids, update counts and mesh sizes on seeded scenes equal the
    committed pins (tests/golden/make_oracle_pins.py) -- every GPU parity test leans on the oracle, so it
    must not drift unnoticed"""
    import importlib.util
    import json
    import os

    spec = importlib.util.spec_from_file_location("make_oracle_pins", os.path.join(golden_dir, "make_oracle_pins.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(golden_dir, "oracle_pins.json")))
    got = mod.compute()
    assert got.keys() == want.keys()
    for k in want:
        assert got[k] == want[k], k
