"""CPU ORACLE of the This&That gesture rasteriser.  *** TEST INFRASTRUCTURE ONLY ***

PARITY PINNED: unlike the denoising oracle, this one is checked against the real thing — `cv2` (the reference's own
dependency, importable in the authoring container) in tests/test_gesture.py, and against golden outputs produced by
executing the reference's own `get_thisthat_sam` (tests/golden/make_gesture_golden.py -> tests/golden/gesture_*.npz).

Restates, in plain numpy, data_loader/video_this_that_dataset.py:28-130 of the reference (duplicated in app.py:282-328):
    :26      blur_kernel = bivariate_Gaussian(99, 10, 10, 0, isotropic=True)   (utils/optical_flow_utils.py:168-219)
    :60-74   255-filled float32 image of the ORIGINAL size, 21 x 21 square around the point — first point [0,0,255],
             later points [0,255,0] (BGR, never converted)
    :77-78   cv2.filter2D(base_img, -1, blur_kernel)      (correlation, anchor at the centre, BORDER_REFLECT_101)
    :85      cv2.resize(..., (width, height), INTER_CUBIC) (a = -0.75, half-pixel centres, replicated border, float32)
    :89-90   optional np.fliplr
    :99      / 255.0
    :106-114 hwc -> chw, written into frame `frame_idx` of a zero [F, 3, H, W] tensor (later points overwrite earlier
             ones on the same frame)
Only tests/ and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def bivariate_gaussian_kernel(kernel_size: int = 99, sigma: float = 10.0) -> np.ndarray:
    """utils/optical_flow_utils.py:168-219 in the isotropic mode: exp(-0.5 x^T Sigma^-1 x) on the integer grid
    [-(K//2), K//2]^2, normalised to sum 1 (float64, like the reference)."""
    ax = np.arange(-kernel_size // 2 + 1.0, kernel_size // 2 + 1.0)
    xx, yy = np.meshgrid(ax, ax)
    kernel = np.exp(-0.5 * (xx * xx + yy * yy) / (sigma * sigma))
    return kernel / np.sum(kernel)


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    """cv2.BORDER_REFLECT_101: ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ... (repeated for kernels wider than the image)."""
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


def filter2d_reflect101(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """cv2.filter2D(img, -1, kernel): correlation with the anchor at the kernel centre, BORDER_REFLECT_101.
    Direct spatial-domain sum in float64 (cv2 switches to a float32 DFT for kernels this large: agreement ~1e-4 on
    the 0..255 scale), rounded to float32 like cv2's output."""
    kh, kw = kernel.shape
    h, w = img.shape[:2]
    rows = _reflect101(np.arange(-(kh // 2), h + kh // 2), h)
    cols = _reflect101(np.arange(-(kw // 2), w + kw // 2), w)
    padded = img.astype(np.float64)[rows][:, cols]
    out = np.zeros(img.shape, dtype=np.float64)
    for i in range(kh):
        for j in range(kw):
            out += kernel[i, j] * padded[i:i + h, j:j + w]
    return out.astype(np.float32)


def _cubic_coeffs(x: np.ndarray) -> np.ndarray:
    """cv2 interpolateCubic, A = -0.75, float32 arithmetic."""
    A = np.float32(-0.75)
    x = x.astype(np.float32)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c3 = np.float32(1.0) - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.float32)


def _cubic_taps(n_src: int, n_dst: int) -> Tuple[np.ndarray, np.ndarray]:
    scale = n_src / n_dst
    f = ((np.arange(n_dst) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = f - s.astype(np.float32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, n_src - 1)  # replicated border
    return idx, _cubic_coeffs(frac)


def resize_cubic(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv2.resize(img, (width, height), interpolation=cv2.INTER_CUBIC) for float32 images: separable, horizontal pass
    first, no antialiasing, no clamping of the overshoot."""
    img = img.astype(np.float32)
    h, w = img.shape[:2]
    xi, xc = _cubic_taps(w, width)
    yi, yc = _cubic_taps(h, height)
    tmp = np.zeros((h, width) + img.shape[2:], dtype=np.float32)
    for k in range(4):
        tmp += img[:, xi[:, k]] * xc[:, k].reshape((1, width) + (1,) * (img.ndim - 2))
    out = np.zeros((height, width) + img.shape[2:], dtype=np.float32)
    for k in range(4):
        out += tmp[yi[:, k]] * yc[:, k].reshape((height, 1) + (1,) * (img.ndim - 2))
    return out


def rasterise(points: Sequence[Tuple[int, int, int]], org_hw: Tuple[int, int], out_hw: Tuple[int, int],
              n_frames: int = 14, dilate: bool = True, flip: bool = False) -> np.ndarray:
    """points: (frame_idx, vertical, horizontal) in data.txt order (the first one is drawn red, the others green).
    Returns the [n_frames, 3, H, W] float32 condition of get_thisthat_sam."""
    org_h, org_w = org_hw
    H, W = out_hw
    cond = np.zeros((n_frames, 3, H, W), dtype=np.float32)
    kernel = bivariate_gaussian_kernel(99, 10.0)
    for idx, (frame_idx, vertical, horizontal) in enumerate(points):
        base = np.full((org_h, org_w, 3), 255.0, dtype=np.float32)
        colour = np.array([0, 0, 255] if idx == 0 else [0, 255, 0], dtype=np.float32)
        r0, r1 = max(vertical - 10, 0), min(vertical + 10, org_h - 1)
        c0, c1 = max(horizontal - 10, 0), min(horizontal + 10, org_w - 1)
        if r0 <= r1 and c0 <= c1:
            base[r0:r1 + 1, c0:c1 + 1] = colour
        if dilate:
            base = filter2d_reflect101(base, kernel)
        base = resize_cubic(base, W, H)
        if flip:
            base = base[:, ::-1]
        cond[frame_idx] = (base / 255.0).transpose(2, 0, 1)
    return cond


def parse_data_txt(lines: List[str]) -> List[Tuple[int, int, int]]:
    """data.txt lines are `frame_idx horizontal vertical` (reference :49-50: int(float(.)) on the coordinates)."""
    pts = []
    for line in lines:
        if not line.strip():
            continue
        frame_idx, horizontal, vertical = line.split(" ")
        pts.append((int(frame_idx), int(float(vertical)), int(float(horizontal))))
    return pts
