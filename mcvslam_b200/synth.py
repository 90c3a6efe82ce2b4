"""Seeded synthetic inputs for tests and bench. Not part of parity: oracle and GPU consume the
same bytes. numpy only."""
import numpy as np


def scene(seed, w=640, h=480):
    """Random-rectangle scene + i.i.d. +-6 noise (PCG64)."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128, np.int32)
    for _ in range((w * h) // 600):
        x = int(rng.integers(0, w)); y = int(rng.integers(0, h))
        rw = int(rng.integers(8, 64)); rh = int(rng.integers(8, 64)); g = int(rng.integers(0, 256))
        img[y:min(y + rh + 1, h), x:min(x + rw + 1, w)] = g
    img += rng.integers(-6, 7, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def stereo_right(left, seed, bands=8, dmin=4, dmax=60):
    """Right image: `left` shifted left by a per-row-band integer disparity; vacated columns refilled from a second scene."""
    h, w = left.shape
    fill = scene(seed + 100003, w, h)
    right = np.empty_like(left)
    bh = (h + bands - 1) // bands
    for b in range(bands):
        d = dmin + (dmax - dmin) * b // max(1, bands - 1)
        r0, r1 = b * bh, min(h, (b + 1) * bh)
        right[r0:r1, :w - d] = left[r0:r1, d:]
        right[r0:r1, w - d:] = fill[r0:r1, w - d:]
    return right


def wide(seed, w=640, h=480):
    """Wide camera: centre crop of a 2x larger scene decimated by 2 (fixed recipe; only determinism matters)."""
    big = scene(seed + 200003, 2 * w, 2 * h).astype(np.uint16)
    out = (big[0::2, 0::2] + big[0::2, 1::2] + big[1::2, 0::2] + big[1::2, 1::2] + 2) >> 2
    return out.astype(np.uint8)


def triplet(seed, w=640, h=480):
    """(3, h, w) u8: left, right, wide for one three-camera frame."""
    left = scene(seed, w, h)
    return np.stack([left, stereo_right(left, seed), wide(seed, w, h)])


def descriptors(n, seed, low_entropy=False):
    rng = np.random.default_rng(seed)
    hi = 4 if low_entropy else 256
    return rng.integers(0, hi, (n, 32), dtype=np.uint8)
