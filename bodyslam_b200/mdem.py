"""MDEM depth post-processing -- drop-in for the reference's depth-estimation surface.

Mirrors, with unchanged signatures:
  * `colorize`                      R/examples/depth_estimation/depth_map_scaling.py:12-45
                                    (== batch_processing.py:12-45)
  * `process_image`, `process_images`   R/examples/depth_estimation/batch_processing.py:47-72
  * `DepthEstimator`                R/src/depth_estimation/interface.py:16-107
  * `MDEMInterface`                 N/MDEM/mdem_interface.py:17-113
  * `compute_median_scale_factor`   N/EVALUATION/MDEM_eval.py:114-127
The ZoeDepth backbone stays the reference's PyTorch forward (torch.hub); everything after the
predicted depth tensor -- x256 metric scaling to uint16, percentile normalisation, colour
mapping -- runs in the K1 CUDA kernels (csrc/bslam_image.cu).  No CPU fallback.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import numpy as np

from . import _lib, ops
from .geometry import to_numpy

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


# ---------------------------------------------------------------- colour maps (byte LUTs)
def _segment_table(N, data):
    # matplotlib.colors._create_lookup_table for (x, y0, y1) rows, gamma 1
    adata = np.array(data, dtype=float)
    x, y0, y1 = adata[:, 0] * (N - 1), adata[:, 1], adata[:, 2]
    xind = (N - 1) * np.linspace(0, 1, N)
    ind = np.searchsorted(x, xind)[1:-1]
    distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.concatenate([[y1[0]], distance * (y0[ind] - y1[ind - 1]) + y1[ind - 1], [y0[-1]]])
    return np.clip(lut, 0.0, 1.0)


_SEGMENTS = {
    "gray": {c: [(0.0, 0.0, 0.0), (1.0, 1.0, 1.0)] for c in ("red", "green", "blue")},
    "jet": {
        "red": [(0.00, 0, 0), (0.35, 0, 0), (0.66, 1, 1), (0.89, 1, 1), (1.00, 0.5, 0.5)],
        "green": [(0.000, 0, 0), (0.125, 0, 0), (0.375, 1, 1), (0.640, 1, 1), (0.910, 0, 0), (1.000, 0, 0)],
        "blue": [(0.00, 0.5, 0.5), (0.11, 1, 1), (0.34, 1, 1), (0.65, 0, 0), (1.00, 0, 0)],
    },
}
_LUT_CACHE = {}


def get_cmap_lut(cmap) -> np.ndarray:
    """[256,4] uint8 table equal to matplotlib's `cmap(np.arange(256)/255.., bytes=True)` rows.

    Accepts a name ('viridis', 'gray', 'jet', and their '_r' reversals), a [256,3|4] array, or a
    matplotlib colormap object (used if matplotlib happens to be installed).
    """
    if isinstance(cmap, str):
        if cmap in _LUT_CACHE:
            return _LUT_CACHE[cmap]
        base, rev = (cmap[:-2], True) if cmap.endswith("_r") else (cmap, False)
        path = os.path.join(_DATA, f"cmap_{base}.npy")
        if os.path.exists(path):
            lut = np.load(path).astype(np.uint8)
            lut = lut[::-1].copy() if rev else lut
        elif base in _SEGMENTS:
            seg = _SEGMENTS[base]
            if rev:
                seg = {k: [(1.0 - x, y1, y0) for x, y0, y1 in reversed(v)] for k, v in seg.items()}
            f = np.ones((256, 4))
            for i, k in enumerate(("red", "green", "blue")):
                f[:, i] = _segment_table(256, seg[k])
            lut = (f * 255).astype(np.uint8)
        else:
            try:  # pragma: no cover - matplotlib is optional
                import matplotlib

                lut = matplotlib.colormaps[cmap](np.arange(256), bytes=True)
            except Exception as e:
                raise ValueError(f"{cmap!r} is not a known colormap (have: viridis, gray, jet and *_r)") from e
        _LUT_CACHE[cmap] = np.ascontiguousarray(lut, dtype=np.uint8)
        return _LUT_CACHE[cmap]
    if callable(cmap) and hasattr(cmap, "N"):
        return np.ascontiguousarray(cmap(np.arange(cmap.N), bytes=True), dtype=np.uint8)
    a = np.asarray(cmap)
    if a.shape == (256, 3):
        a = np.concatenate([a, np.full((256, 1), 255, a.dtype)], 1)
    if a.shape != (256, 4):
        raise ValueError("cmap array must be [256,3] or [256,4] uint8")
    return np.ascontiguousarray(a, dtype=np.uint8)


def _gamma_u8(x):
    # reference: img/255 -> power 2.2 -> *255 -> astype(uint8), all four channels (depth_map_scaling.py:39-43)
    x = np.asarray(x, dtype=np.uint8)
    y = x / 255
    y = np.power(y, 2.2)
    y = y * 255
    return y.astype(np.uint8)


def colorize(value, vmin=None, vmax=None, cmap='gray_r', invalid_val=-99, invalid_mask=None,
             background_color=(128, 128, 128, 255), gamma_corrected=False, value_transform=None):
    """Converts a depth map to a color image.  (reference signature, depth_map_scaling.py:12)

    value: [H,W] (any singleton dims squeezed), numpy or torch (CPU/CUDA): the uint16 image
    `np.array(PIL I;16)` the reference feeds it (any integer-valued depth in [0, 65535]), or a
    float32 image of arbitrary values such as ZoeDepth's metres tensor (evaluated in NumPy's float32
    arithmetic, exactly like the reference does for a float32 array).
    Returns a numpy [H,W,4] uint8 RGBA image like the reference.  Percentiles, normalisation,
    LUT indexing and painting all run on the GPU; the byte LUT (with gamma folded in) and, for
    `value_transform`, the 65536-entry index table are the only host-side preparation.
    """
    return to_numpy(colorize_cuda(value, vmin, vmax, cmap, invalid_val, invalid_mask, background_color,
                                  gamma_corrected, value_transform))


def colorize_cuda(value, vmin=None, vmax=None, cmap='gray_r', invalid_val=-99, invalid_mask=None,
                  background_color=(128, 128, 128, 255), gamma_corrected=False, value_transform=None, device=None, _squeeze=True):
    """`colorize` that leaves the RGBA image on the device ([H,W,4] or [B,H,W,4] uint8 CUDA tensor)."""
    torch = _lib.require_cuda()
    if hasattr(value, "detach"):
        v = value.detach()
        while _squeeze and v.dim() > 2 and 1 in v.shape:
            v = v.squeeze()
        name = str(v.dtype).replace("torch.", "")
    else:
        v = np.asarray(value)
        if _squeeze:
            v = v.squeeze()
        name = str(v.dtype)
    if name not in ("uint16", "uint8", "int16", "int32", "int64", "float32", "float64"):
        raise RuntimeError(f"[colorize] Unsupported image format. (dtype {name})")
    if v.ndim < 2:
        # a single row / a single pixel after the reference's squeeze(): colour it as a one-row image and give the
        # result the squeezed shape + (4,), like `cmapper(value, bytes=True)` does
        shape = tuple(v.shape)
        mask1 = None if invalid_mask is None else to_numpy(invalid_mask).reshape(1, -1)
        out = colorize_cuda(v.reshape(1, -1), vmin, vmax, cmap, invalid_val, mask1, background_color, gamma_corrected, value_transform, device,
                            _squeeze=False)
        return out.reshape(shape + (4,))
    lut = get_cmap_lut(cmap)
    bg = np.asarray(list(background_color) + [255] * (4 - len(background_color)), dtype=np.uint8)
    if gamma_corrected:
        lut, bg = _gamma_u8(lut), _gamma_u8(bg)
    if name == "float32":
        # float metres (a ZoeDepth tensor, depth_map_scaling.py:14-15): NumPy's float32 arithmetic, exact radix select
        if value_transform is not None:
            raise RuntimeError("[colorize] value_transform is only supported for integer-valued (uint16) depth on the CUDA path")
        vt = ops.as_cuda(v, torch.float32, device)
        inv = invalid_val
        if invalid_mask is not None:
            m = ops.as_cuda(invalid_mask, torch.bool, device).reshape(vt.shape)
            inv = float(np.finfo(np.float32).min)
            while bool((vt == inv).any()):
                inv = float(np.nextafter(np.float32(inv), np.float32(0)))
            vt = torch.where(m, torch.full_like(vt, inv), vt)
        return ops.colorize_f32(lut, vt, invalid_val=inv, background=bg, vmin=vmin, vmax=vmax, device=device)
    if name != "uint16":
        vn = to_numpy(v)
        if vn.size and (vn.min() < 0 or vn.max() > 65535 or (name.startswith("float") and np.any(vn != np.floor(vn)))):
            raise RuntimeError("[colorize] float64 / signed input must be integer-valued depth in [0, 65535] on the CUDA path "
                               "(float32 images of any value are supported: pass value.astype(np.float32) / tensor.float())")
        v = vn.astype(np.uint16)
    if invalid_mask is not None:
        # an explicit mask replaces `value == invalid_val` (depth_map_scaling.py:17-18): paint the
        # masked pixels with a sentinel value no valid pixel uses and hand that to the kernel
        vt = ops.as_cuda(v, torch.uint16, device)
        m = ops.as_cuda(invalid_mask, torch.bool, device).reshape(vt.shape)
        used = torch.bincount(vt.reshape(-1).to(torch.int64)[~m.reshape(-1)], minlength=65536)
        free = torch.nonzero(used == 0)
        if free.numel() == 0:
            raise RuntimeError("[colorize] no free uint16 value left to encode invalid_mask")
        invalid_val = int(free[-1].item())
        v = torch.where(m, torch.full_like(vt.to(torch.int32), invalid_val), vt.to(torch.int32)).to(torch.uint16)
    table = None
    if value_transform is not None:
        # value_transform acts on the normalised value, which is a function of the 16-bit depth
        # only: apply it to the 65536 possible values on the host and ship the index table
        _, _, stats = ops.colorize_u16(lut, depth_u16=v, invalid_val=invalid_val, background=bg, vmin=vmin, vmax=vmax,
                                       device=device, return_stats=True)
        st = stats.reshape(-1, 2).cpu().numpy()
        tabs = []
        for lo, hi in st:
            x = np.arange(65536, dtype=np.float64)
            x = (x - lo) / (hi - lo) if lo != hi else x * 0.0
            x = np.array(value_transform(x), dtype=np.float64)
            with np.errstate(invalid="ignore"):
                xa = x * 256
                xa[xa < 0] = -1
                xa[xa == 256] = 255
                xa = np.clip(xa, -1, 256)
            xa[np.isnan(xa)] = 0
            tabs.append(np.clip(xa.astype(int), 0, 255).astype(np.uint8))
        table = np.stack(tabs)
    rgba, _ = ops.colorize_u16(lut, depth_u16=v, invalid_val=invalid_val, background=bg, vmin=vmin, vmax=vmax,
                               index_table=table, device=device)
    return rgba


def compute_median_scale_factor(ground_truth, predictions):
    """median(gt) / median(pred)  (N/EVALUATION/MDEM_eval.py:114-127).

    uint16 inputs go through the GPU histogram median (exact); other dtypes use torch's GPU sort.
    """
    torch = _lib.require_cuda()

    def med(a):
        if ops._dtype_name(a) == "uint16":
            t = ops.as_cuda(a, torch.uint16).reshape(1, -1)
            return float(ops.median_u16(t)[0].item())
        t = ops.as_cuda(a, torch.float64).reshape(-1)
        s, _ = torch.sort(t)
        n = s.numel()
        return float(((s[(n - 1) // 2] + s[n // 2]) / 2).item())

    return med(ground_truth) / med(predictions)


# ---------------------------------------------------------------- depth estimator front-ends
def depth_tensor_to_pil(depth_metres, scale: float = 256.0):
    """ZoeDepth `infer_pil(..., output_type='pil')` tail on the GPU: metres -> I;16 PIL image."""
    from PIL import Image

    u16 = ops.scale_to_u16(depth_metres, scale)
    while u16.dim() > 2:
        u16 = u16[0]
    return Image.fromarray(to_numpy(u16))


class DepthEstimator:
    '''A class to interface with ZOE for monocular depth estimation (interface.py:16)'''

    SUPPORTED_MODELS = ['ZoeD_N', 'ZoeD_K', 'ZoeD_NK']
    DEFAULT_MODEL = 'ZoeD_NK'

    def __init__(self, model_type: str = DEFAULT_MODEL, model=None):
        """`model`: optional pre-built backbone (anything with `.infer(tensor)->metres` or
        `.infer_pil(image, output_type='tensor')`), so offline boxes can inject one."""
        self.model = model if model is not None else self._initialize_model(model_type)

    def _initialize_model(self, model_type: str):
        import torch

        if model_type not in self.SUPPORTED_MODELS:
            warnings.warn(
                f"The model type '{model_type}' is not supported. Using default model '{self.DEFAULT_MODEL}'.")
            model_type = self.DEFAULT_MODEL
        torch.hub.help("intel-isl/MiDaS", "DPT_BEiT_L_384", force_reload=True)
        model = torch.hub.load("isl-org/ZoeDepth", model_type, pretrained=True)
        device = "cuda" if torch.cuda.is_available() else "cpu"
        print(f"[INFO] Model loaded on {device}")
        return model.to(device)

    def infer_depth_tensor(self, path_to_frame: str):
        """metric depth in metres as a float32 tensor, left on the backbone's device."""
        image = self.load_image(path_to_frame)
        try:
            return self.model.infer_pil(image, output_type="tensor")
        except TypeError:
            return self.model.infer_pil(image)

    def infer_depth_map(self, path_to_frame: str):
        """Depth map as a 16-bit PIL Image (metres * 256), interface.py:53-61 -- the scaling runs in K1."""
        return depth_tensor_to_pil(self.infer_depth_tensor(path_to_frame))

    @staticmethod
    def load_image(path: str):
        from PIL import Image

        image = Image.open(path)
        return image.convert('RGB')

    @staticmethod
    def save_depth_map(image, saving_path: str, extension: Optional[str] = None):
        if extension:
            saving_path = os.path.splitext(saving_path)[0] + '.' + extension.lstrip('.')
        image.save(saving_path)

    def debug(self, path_to_frame: str, saving_path: str):
        from PIL import Image

        tests = [
            ("load image", lambda: self.load_image(path_to_frame)),
            ("infer method", lambda: self.infer_depth_map(path_to_frame)),
            ("saving method", lambda: self.save_depth_map(Image.new('RGB', (100, 100)), saving_path)),
        ]
        for test_name, test_func in tests:
            print(f"[DEBUG]: Testing {test_name}...")
            try:
                test_func()
                print(f"[DEBUG]: {test_name} status -> ok")
            except Exception as e:
                print(f"[DEBUG]: OPS :/ -> {e}")


class MDEMInterface:
    """Legacy twin of DepthEstimator (N/MDEM/mdem_interface.py:17)."""

    def __init__(self, model_type: str = "ZoeD_NK", model=None):
        self.zoe = model if model is not None else self._initialize_ZOE(model_type)

    def _initialize_ZOE(self, model_type: str):
        import torch

        torch.hub.help("intel-isl/MiDaS", "DPT_BEiT_L_384", force_reload=True)
        if model_type not in ("ZoeD_N", "ZoeD_K", "ZoeD_NK"):
            warnings.warn(f"The model type selected [{model_type}], does not exist! Using default model [ZoeD_NK]")
        zoe = torch.hub.load("isl-org/ZoeDepth", model_type, pretrained=True)
        DEVICE = "cuda" if torch.cuda.is_available() else "cpu"
        print(f"[INFO] model loaded on {DEVICE}")
        return zoe.to(DEVICE)

    def infer_monocular_depth_map(self, path_to_frame: str):
        from PIL import Image

        image = Image.open(path_to_frame).convert("RGB")
        try:
            depth = self.zoe.infer_pil(image, output_type="tensor")
        except TypeError:
            depth = self.zoe.infer_pil(image)
        return depth_tensor_to_pil(depth)

    @staticmethod
    def save_depth_map(image, saving_path: str, extension: str = None):
        # FrameIO.save_p_img (N/UTILS/io_utils.py:49-74): extension is APPENDED; ValueError is printed
        try:
            if extension is None:
                warnings.warn("No extension has been provided")
                image.save(saving_path)
            else:
                image.save(saving_path + extension)
            return True
        except ValueError as e:
            print(f"Error while saving: {e}")


# ---------------------------------------------------------------- batch_processing.py
def process_image(estimator, input_path, output_path, colormap='viridis', invalid_val=0):
    """Process a single image: estimate depth, colorize, and save. (batch_processing.py:47-58)"""
    from PIL import Image

    depth_map = estimator.infer_depth_map(input_path)
    depth_array = np.array(depth_map)
    colorized_depth = colorize(depth_array, cmap=colormap, invalid_val=invalid_val)
    Image.fromarray(colorized_depth).save(output_path)
    print(f"Processed and saved: {output_path}")


def process_images(input_dir, output_dir, colormap='viridis', invalid_val=0, estimator=None):
    """Process all images in the input directory (batch_processing.py:60-72)."""
    estimator = estimator if estimator is not None else DepthEstimator()
    os.makedirs(output_dir, exist_ok=True)
    for filename in os.listdir(input_dir):
        if filename.lower().endswith(('.png', '.jpg', '.jpeg', '.tiff', '.bmp', '.gif')):
            input_path = os.path.join(input_dir, filename)
            output_path = os.path.join(output_dir, f"depth_{os.path.splitext(filename)[0]}.png")
            process_image(estimator, input_path, output_path, colormap, invalid_val)


def process_depth_batch(depth_metres, colormap='viridis', invalid_val=0, scale: float = 256.0, device=None):
    """The batched shape of BASELINE config 3: [B,H,W] f32 metres -> (rgba [B,H,W,4], u16 [B,H,W]) on the GPU."""
    return ops.colorize_u16(get_cmap_lut(colormap), depth_m=depth_metres, scale=scale, invalid_val=invalid_val, device=device)
