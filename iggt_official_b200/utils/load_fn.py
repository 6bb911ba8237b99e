"""Device-side image pre-processing.

Drop-in for the reference's `iggt.utils.load_fn.load_and_preprocess_images` (iggt/utils/load_fn.py:12-128): same
arguments, modes ("crop" / "pad" / "resize"), errors and result - a float32 [N, 3, H, W] batch in [0, 1] - except that
the batch is produced ON THE GPU.  File decoding and colour-mode conversion stay with Pillow on the host (I/O); the
8-bit bicubic resize, ToTensor, centre crop, white padding and batching - everything the reference does per pixel on
the CPU - run in two CUDA kernels per view that are bit-exact with Pillow's resampler (csrc/preprocess.cu).

`precompute_coeffs` is the host half of that resampler: Pillow's tap tables in double precision, then 22-bit
fixed point (Pillow 12.x src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, bicubic_filter)."""
import numpy as np
import torch

from .. import ops

PRECISION_BITS = 32 - 8 - 2
DEFAULT_TARGET_SIZE = 518            # load_fn.py:56


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def precompute_coeffs(in_size: int, out_size: int):
    """Fixed-point bicubic taps of one pass: (kk int32 [out_size, ksize], bounds int32 [out_size, 2] = (first, count))."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.int64)[None, :]
    w = _bicubic(((x + xmin[:, None]) - center[:, None] + 0.5) * ss)
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1:]                       # sequential sum, the order Pillow accumulates in
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    fixed = np.where(w < 0, np.trunc(-0.5 + w * (1 << PRECISION_BITS)), np.trunc(0.5 + w * (1 << PRECISION_BITS)))
    return fixed.astype(np.int32), np.stack([xmin, xmax], axis=1).astype(np.int32)


def _target_size(width, height, mode, resize_target_size):
    """(new_width, new_height) of load_fn.py:68-79."""
    t = DEFAULT_TARGET_SIZE
    if mode == "pad":
        if width >= height:
            return t, round(height * (t / width) / 14) * 14
        return round(width * (t / height) / 14) * 14, t
    if mode == "resize":
        return tuple(resize_target_size)
    return t, round(height * (t / width) / 14) * 14


_COEFF_CACHE = {}


def _device_coeffs(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    if key not in _COEFF_CACHE:
        kk, bounds = precompute_coeffs(in_size, out_size)
        _COEFF_CACHE[key] = (torch.from_numpy(kk).to(device), torch.from_numpy(bounds).to(device))
    return _COEFF_CACHE[key]


def preprocess_decoded(frames, mode="crop", resize_target_size=None, device="cuda"):
    """frames: list of decoded RGB views, uint8 [H, W, 3] (numpy or torch).  Returns float32 [N, 3, H', W'] on `device`."""
    if not frames:
        raise ValueError("At least 1 image is required")
    if mode not in ["crop", "pad", "resize"]:
        raise ValueError("Mode must be either 'crop', 'pad', or 'resize'")
    if mode == "resize":
        if resize_target_size is None:
            raise ValueError("resize_target_size must be provided as a (width, height) tuple when mode is 'resize'")
        if not (isinstance(resize_target_size, (tuple, list)) and len(resize_target_size) == 2):
            raise ValueError("resize_target_size must be a tuple or list of two integers: (width, height)")
    device = torch.device(device)
    t = DEFAULT_TARGET_SIZE
    plans = []
    for f in frames:
        f = torch.as_tensor(np.ascontiguousarray(f) if isinstance(f, np.ndarray) else f)
        if f.dtype != torch.uint8 or f.dim() != 3 or f.shape[2] != 3:
            raise ValueError("decoded views must be uint8 [H, W, 3]")
        height, width = int(f.shape[0]), int(f.shape[1])
        new_w, new_h = _target_size(width, height, mode, resize_target_size)
        y0, rows, top, left, out_h, out_w = 0, new_h, 0, 0, new_h, new_w
        if mode == "crop" and new_h > t:                     # load_fn.py:86-88
            y0, rows, out_h = (new_h - t) // 2, t, t
        elif mode == "pad":                                  # load_fn.py:89-98 (padding is never negative here)
            top, left = max(t - new_h, 0) // 2, max(t - new_w, 0) // 2
            out_h, out_w = max(t, new_h), max(t, new_w)
        plans.append((f, new_w, new_h, y0, rows, top, left, out_h, out_w))
    shapes = {(p[7], p[8]) for p in plans}
    if len(shapes) > 1:                                      # load_fn.py:104-121
        print(f"Warning: Found images with different shapes after processing: {shapes}")
    max_h, max_w = max(s[0] for s in shapes), max(s[1] for s in shapes)
    padded = any(p[4] != max_h or p[1] != max_w for p in plans)
    batch = (torch.ones if padded else torch.empty)((len(plans), 3, max_h, max_w), dtype=torch.float32, device=device)
    for n, (f, new_w, new_h, y0, rows, top, left, out_h, out_w) in enumerate(plans):
        top += (max_h - out_h) // 2
        left += (max_w - out_w) // 2
        src = f.to(device, non_blocking=True).contiguous()
        kk_h, b_h = _device_coeffs(int(f.shape[1]), new_w, device)
        kk_v, b_v = _device_coeffs(int(f.shape[0]), new_h, device)
        ops.resample_bicubic_u8(src, kk_h, b_h, kk_v, b_v, batch[n, :, top:top + rows, left:left + new_w], oy0=y0)
    return batch


def load_and_preprocess_images(image_path_list, mode="crop", resize_target_size=None, device="cuda"):
    """Reference signature (+ `device`).  Decode with Pillow on the host, everything else on the GPU."""
    from PIL import Image
    if not image_path_list:
        raise ValueError("At least 1 image is required")
    if mode not in ["crop", "pad", "resize"]:
        raise ValueError("Mode must be either 'crop', 'pad', or 'resize'")
    frames = []
    for path in image_path_list:
        img = Image.open(path)
        if img.mode == "RGBA":                               # load_fn.py:62-64: composite onto white
            img = Image.alpha_composite(Image.new("RGBA", img.size, (255, 255, 255, 255)), img)
        frames.append(np.asarray(img.convert("RGB")))
    return preprocess_decoded(frames, mode, resize_target_size, device)
