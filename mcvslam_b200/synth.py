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


def random_vocabulary(seed, K=10, L=4, weighting=0, norm=1, stop_fraction=0.02, ragged=True):
    """A synthetic DBoW3-shaped vocabulary (the real orbvoc.dbow3 is K=10, L=6, 1 082 073 nodes; same structure): a K-ary tree of
    depth L built like HKmeansStep numbers its nodes (children of a node get consecutive ids), node descriptors = parent's with
    bits flipped (so descents are meaningful and ties between siblings occur), random positive leaf weights with a few zero
    ("stopped") words. Flat arrays: child_off [n+1], child_ids, node_desc [n][32], word_id [n], weight [n] (float64).
    weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; norm: 0 none, 1 L1, 2 L2."""
    rng = np.random.default_rng(seed)
    desc = [rng.integers(0, 256, 32, dtype=np.uint8)]
    children = [[]]
    level = [0]
    frontier = [0]
    for depth in range(1, L + 1):
        nxt = []
        for p in frontier:
            k = K if not ragged or depth == 1 else int(rng.integers(max(2, K - 3), K + 1))
            for _ in range(k):
                i = len(desc)
                d = desc[p].copy()
                nb = int(rng.integers(2, 40))
                bits = rng.choice(256, nb, replace=False)
                np.bitwise_xor.at(d, bits // 8, (1 << (bits % 8)).astype(np.uint8))
                if rng.random() < 0.05 and children[p]:
                    d = desc[children[p][-1]].copy()          # duplicate sibling: exact distance tie
                desc.append(d); children.append([]); level.append(depth); children[p].append(i); nxt.append(i)
        frontier = nxt
    n = len(desc)
    child_off = np.zeros(n + 1, np.int32); ids = []
    for i in range(n):
        child_off[i + 1] = child_off[i] + len(children[i]); ids += children[i]
    word_id = np.full(n, -1, np.int32); weight = np.zeros(n, np.float64)
    leaves = [i for i in range(n) if not children[i]]
    word_id[leaves] = np.arange(len(leaves))
    w = rng.uniform(0.1, 9.0, len(leaves)); w[rng.random(len(leaves)) < stop_fraction] = 0.0
    weight[leaves] = w if weighting in (0, 2) else np.where(w > 0, 1.0, 0.0)
    return dict(child_off=child_off, child_ids=np.array(ids, np.uint32), node_desc=np.stack(desc), word_id=word_id, weight=weight, L=L, K=K,
                weighting=weighting, norm=norm)


def shifted(img, dx, dy, seed, noise=3):
    """`img` translated by the fractional offset (dx, dy) (pure-numpy bilinear blend of four integer shifts, edges wrap) plus
    fresh +-`noise` noise: the second frame of a synthetic optical-flow pair."""
    fx, fy = int(np.floor(dx)), int(np.floor(dy))
    ax, ay = float(dx - fx), float(dy - fy)
    f = img.astype(np.float64)
    s00 = np.roll(f, (fy, fx), (0, 1)); s01 = np.roll(f, (fy, fx + 1), (0, 1))
    s10 = np.roll(f, (fy + 1, fx), (0, 1)); s11 = np.roll(f, (fy + 1, fx + 1), (0, 1))
    out = (1 - ay) * ((1 - ax) * s00 + ax * s01) + ay * ((1 - ax) * s10 + ax * s11)
    rng = np.random.default_rng(seed)
    out = np.rint(out) + (rng.integers(-noise, noise + 1, img.shape) if noise > 0 else 0)
    return np.clip(out, 0, 255).astype(np.uint8)
