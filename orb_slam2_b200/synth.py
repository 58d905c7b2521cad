"""Deterministic synthetic camera frames (no datasets are available offline).

SURVEY.md §8(d): left image = smooth low-frequency background (uniform noise in [70,150] at 1/6
resolution, bicubic-upsampled) + 300 random axis-aligned rectangles (6-40 x 6-30 px, additive
contrast in [-90,90]) + Gaussian sensor noise (sigma 2); right image = the noise-free left image
resampled with a smooth disparity field d(x,y) in [2,80] px plus independent sensor noise, so
that Frame::ComputeStereoMatches (reference src/Frame.cc:466) finds matches for most keypoints.
Everything is a pure function of (seed, stream_id, frame_idx); numpy only.
"""
from __future__ import annotations

import numpy as np

KITTI = (1242, 375)   # Examples/Stereo/KITTI00-02.yaml (shape used by BASELINE configs 2,4)
TUM = (640, 480)      # Examples/RGB-D/TUM1.yaml
EUROC = (752, 480)    # Examples/Stereo/EuRoC.yaml


def _rng(seed: int, stream_id: int, frame_idx: int) -> np.random.Generator:
    return np.random.default_rng(np.random.SeedSequence([int(seed), int(stream_id), int(frame_idx)]))


def _bilinear(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """Sample img (float32 HxW) at float coordinates (broadcastable xs, ys), clamped."""
    h, w = img.shape
    xs = np.clip(xs, 0, w - 1.001)
    ys = np.clip(ys, 0, h - 1.001)
    x0 = np.floor(xs).astype(np.int32)
    y0 = np.floor(ys).astype(np.int32)
    fx = (xs - x0).astype(np.float32)
    fy = (ys - y0).astype(np.float32)
    a = img[y0, x0] * (1 - fx) + img[y0, x0 + 1] * fx
    b = img[y0 + 1, x0] * (1 - fx) + img[y0 + 1, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def _cubic_w(t: np.ndarray, a: float = -0.75) -> np.ndarray:
    t = np.abs(t)
    return np.where(t <= 1, (a + 2) * t ** 3 - (a + 3) * t ** 2 + 1, np.where(t < 2, a * t ** 3 - 5 * a * t ** 2 + 8 * a * t - 4 * a, 0.0))


def _bicubic_upsample(low: np.ndarray, h: int, w: int, factor: float = 6.0) -> np.ndarray:
    """Separable 4-tap bicubic (a=-0.75, the OpenCV INTER_CUBIC kernel) upsampling of a coarse grid (SURVEY §8d:
    "uniform noise in [70,150] at 1/6 resolution, bicubic-upsampled")."""
    lh, lw = low.shape

    def taps(n, ln):
        x = np.arange(n, dtype=np.float64) / factor
        x0 = np.floor(x).astype(np.int64)
        idx = [np.clip(x0 + k, 0, ln - 1) for k in (-1, 0, 1, 2)]
        wt = [_cubic_w(x - (x0 + k)).astype(np.float32) for k in (-1, 0, 1, 2)]
        return idx, wt

    iy, wy = taps(h, lh)
    ix, wx = taps(w, lw)
    low = low.astype(np.float32)
    tmp = sum(low[iy[k], :] * wy[k][:, None] for k in range(4))          # h x lw
    return sum(tmp[:, ix[k]] * wx[k][None, :] for k in range(4)).astype(np.float32)


def clean_left(seed: int, stream_id: int, frame_idx: int, w: int, h: int, n_rect: int = 300) -> np.ndarray:
    """Noise-free left image as float32 (HxW)."""
    rng = _rng(seed, stream_id, frame_idx)
    lw, lh = w // 6 + 3, h // 6 + 3
    low = rng.uniform(70.0, 150.0, (lh, lw)).astype(np.float32)
    img = _bicubic_upsample(low, h, w)
    for _ in range(n_rect):
        rw = int(rng.integers(6, 41))
        rh = int(rng.integers(6, 31))
        x0 = int(rng.integers(-rw // 2, w - rw // 2))
        y0 = int(rng.integers(-rh // 2, h - rh // 2))
        c = float(rng.uniform(-90.0, 90.0))
        img[max(y0, 0):max(y0 + rh, 0), max(x0, 0):max(x0 + rw, 0)] += c
    return img.astype(np.float32)


def _finish(img: np.ndarray, rng: np.random.Generator, sigma: float = 2.0) -> np.ndarray:
    noisy = img + rng.normal(0.0, sigma, img.shape).astype(np.float32)
    return np.clip(np.rint(noisy), 0, 255).astype(np.uint8)


def disparity_field(seed: int, stream_id: int, frame_idx: int, w: int, h: int) -> np.ndarray:
    rng = _rng(seed + 7919, stream_id, frame_idx)
    ph = rng.uniform(0, 2 * np.pi, 3)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    d = 41.0 + 25.0 * np.sin(xs / w * 2.1 * np.pi + ph[0]) * np.cos(ys / h * 1.3 * np.pi + ph[1]) \
        + 14.0 * np.sin((xs + 2 * ys) / (w + 2 * h) * 3.0 * np.pi + ph[2])
    return np.clip(d, 2.0, 80.0).astype(np.float32)


def mono_frame(seed: int, stream_id: int, frame_idx: int, w: int, h: int) -> np.ndarray:
    clean = clean_left(seed, stream_id, frame_idx, w, h)
    return _finish(clean, _rng(seed + 1, stream_id, frame_idx))


def stereo_pair(seed: int, stream_id: int, frame_idx: int, w: int = KITTI[0], h: int = KITTI[1]):
    """Returns (left u8 HxW, right u8 HxW, disparity float32 HxW defined on RIGHT pixels)."""
    clean = clean_left(seed, stream_id, frame_idx, w, h)
    left = _finish(clean, _rng(seed + 1, stream_id, frame_idx))
    d = disparity_field(seed, stream_id, frame_idx, w, h)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    right_clean = _bilinear(clean, xs + d, ys)      # a left point x_L shows up at x_R = x_L - d
    right = _finish(right_clean, _rng(seed + 2, stream_id, frame_idx))
    return left, right, d


def white_noise(seed: int, w: int, h: int) -> np.ndarray:
    """Worst-case corner density (about 10x a natural frame) — stress input for parity tests."""
    return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)
