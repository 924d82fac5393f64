"""Synthetic input generators for tests and bench (SURVEY.md §8(d), Appendix B `synth`).

Frames are rasterised with cv2 (present in the image); the draw order is part of the spec so that seeds are
reproducible.  Descriptor sets for the matcher benchmarks are random 256-bit words with planted near-duplicates.
"""
import numpy as np


def synth(W, H, seed):
    """8-bit gray frame: mid-gray + W*H/600 random filled rectangles/discs, 3x3 Gaussian, N(0,2^2) noise."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.full((H, W), 128, np.uint8)
    for _ in range(int(W * H / 600)):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        w, h = int(rng.integers(4, 60)), int(rng.integers(4, 60))
        c = int(rng.integers(0, 256))
        if rng.random() < 0.5:
            cv2.rectangle(img, (x, y), (x + w, y + h), c, -1)
        else:
            cv2.circle(img, (x, y), w // 2, c, -1)
    img = cv2.GaussianBlur(img, (3, 3), 0.8)
    return np.clip(img + rng.normal(0, 2.0, (H, W)), 0, 255).astype(np.uint8)


def synth_batch(W, H, seeds):
    return np.stack([synth(W, H, s) for s in seeds])


def synth_descriptors(n, seed, dup_of=None, dup_rate=0.5, max_flip=40):
    """n random 32-byte descriptors; if dup_of is given, a dup_rate fraction are copies of random rows of dup_of
    with 0..max_flip random bit flips (so true matches with Hamming <= max_flip exist)."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if dup_of is not None and len(dup_of):
        k = int(n * dup_rate)
        rows = rng.choice(n, k, replace=False)
        src = rng.integers(0, len(dup_of), k)
        d[rows] = dup_of[src]
        for r in rows:
            nf = int(rng.integers(0, max_flip + 1))
            bits = rng.choice(256, nf, replace=False)
            for b in bits:
                d[r, b >> 3] ^= np.uint8(1 << (b & 7))
    return d
