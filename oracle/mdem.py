"""NumPy restatement of the MDEM depth post-processing -- TEST INFRASTRUCTURE ONLY.

Follows, line by line:
  * ZoeDepth ``infer_pil(..., output_type="pil")`` tail (third party, unpinned torch.hub HEAD;
    call sites R/src/depth_estimation/interface.py:61, N/MDEM/mdem_interface.py:68):
    ``(metres * 256).astype(uint16)``                                     -> ``scale_to_u16``
  * ``colorize`` R/examples/depth_estimation/depth_map_scaling.py:12-45
    (== batch_processing.py:12-45)                                        -> ``colorize``
  * ``compute_median_scale_factor`` N/EVALUATION/MDEM_eval.py:114-127     -> same name
  * ``RGBD._compute_colored_depth`` N/3DM/slam_utils.py:250-264           -> ``minmax_u8``

matplotlib is not installable here; its ``Colormap.__call__(X, bytes=True)`` is restated in
``apply_lut`` (index = trunc(X*N) with X<0 -> under, X==1 -> N-1, X>1 -> over, NaN -> bad) and
the byte LUT is an explicit argument.  PARITY: pinned by the reference's golden pair
tests/golden/colorize_golden.npz (exact) for 222/256 viridis rows; see tests/golden/make_golden.py.
"""
import numpy as np


def scale_to_u16(depth_metres, scale=256):
    d = np.asarray(depth_metres, dtype=np.float32)
    return (d * scale).astype(np.uint16)


def compute_median_scale_factor(ground_truth, predictions):
    return np.median(ground_truth) / np.median(predictions)


def apply_lut(x, lut_u8, bad=(0, 0, 0, 0)):
    """matplotlib ``Colormap.__call__(x, bytes=True)`` for float input x; lut_u8 [N,4]."""
    N = lut_u8.shape[0]
    xa = np.array(x, dtype=np.float64, copy=True)
    mask_bad = np.isnan(xa)
    with np.errstate(invalid="ignore"):
        xa *= N
        xa[xa < 0] = -1
        xa[xa == N] = N - 1
        np.clip(xa, -1, N, out=xa)
    xa[mask_bad] = 0
    xi = xa.astype(int)
    idx = np.where(xi < 0, 0, np.where(xi > N - 1, N - 1, xi))  # under = lut[0], over = lut[N-1]
    out = lut_u8[idx]
    out[mask_bad] = np.asarray(bad, dtype=np.uint8)
    return out, idx


def colorize(value, lut_u8, vmin=None, vmax=None, invalid_val=-99, invalid_mask=None,
             background_color=(128, 128, 128, 255), gamma_corrected=False, value_transform=None,
             return_index=False):
    value = np.asarray(value).squeeze()
    if invalid_mask is None:
        invalid_mask = value == invalid_val
    mask = np.logical_not(invalid_mask)
    vmin = np.percentile(value[mask], 2) if vmin is None else vmin
    vmax = np.percentile(value[mask], 85) if vmax is None else vmax
    if vmin != vmax:
        value = (value - vmin) / (vmax - vmin)
    else:
        value = value * 0.0
    value = np.array(value, dtype=np.float64)
    value[invalid_mask] = np.nan
    if value_transform:
        value = value_transform(value)
    img, idx = apply_lut(value, lut_u8)
    img[invalid_mask] = background_color
    if gamma_corrected:
        img = img / 255
        img = np.power(img, 2.2)
        img = img * 255
        img = img.astype(np.uint8)
    if return_index:
        return img, idx, float(vmin), float(vmax)
    return img


def minmax_u8(depth_u16):
    """``np.uint8(255 * (d - min) / (max - min))`` (slam_utils.py:255-258; cv2.minMaxLoc -> f64)."""
    d = np.asarray(depth_u16)
    mn, mx = float(d.min()), float(d.max())
    return np.uint8(255 * (d - mn) / (mx - mn))


# ---------------------------------------------------------------- LUT builders
def _create_lookup_table(N, data):
    """matplotlib.colors._create_lookup_table for (x, y0, y1) segment data, gamma = 1."""
    adata = np.array(data, dtype=float)
    x, y0, y1 = adata[:, 0], adata[:, 1], adata[:, 2]
    x = x * (N - 1)
    xind = (N - 1) * np.linspace(0, 1, N)
    ind = np.searchsorted(x, xind)[1:-1]
    distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.concatenate([[y1[0]], distance * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]])
    return np.clip(lut, 0.0, 1.0)


_SEGMENTS = {
    "gray": {"red": [(0, 0, 0), (1, 1, 1)], "green": [(0, 0, 0), (1, 1, 1)], "blue": [(0, 0, 0), (1, 1, 1)]},
    "jet": {
        "red": [(0.00, 0, 0), (0.35, 0, 0), (0.66, 1, 1), (0.89, 1, 1), (1.00, 0.5, 0.5)],
        "green": [(0.000, 0, 0), (0.125, 0, 0), (0.375, 1, 1), (0.640, 1, 1), (0.910, 0, 0), (1.000, 0, 0)],
        "blue": [(0.00, 0.5, 0.5), (0.11, 1, 1), (0.34, 1, 1), (0.65, 0, 0), (1.00, 0, 0)],
    },
}


def segment_lut(name, N=256):
    """Byte LUT of a matplotlib LinearSegmentedColormap ('gray', 'jet', and their '_r')."""
    rev = name.endswith("_r")
    seg = _SEGMENTS[name[:-2] if rev else name]
    if rev:  # LinearSegmentedColormap.reversed(): (1-x, y1, y0) for reversed(data)
        seg = {k: [(1.0 - x, y1, y0) for x, y0, y1 in reversed(v)] for k, v in seg.items()}
    lut = np.ones((N, 4), float)
    for c, k in enumerate(("red", "green", "blue")):
        lut[:, c] = _create_lookup_table(N, seg[k])
    return (lut * 255).astype(np.uint8)
