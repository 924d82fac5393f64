"""Golden vectors for the descriptor distance (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1650-1666) from the real OpenCV:
cv2.norm(a, b, NORM_HAMMING) on descriptor pairs at controlled distances, and cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) distances
of a query set against a database (distances only: tie ORDER is BFMatcher's own business, the reference's loop keeps the first).
Run here with cv2 4.13; the fixture travels, cv2 does not have to."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(2024)


def flip(d, k):
    out = d.copy()
    for b in rng.choice(256, size=k, replace=False):
        out[b >> 3] ^= np.uint8(1 << (b & 7))
    return out


a = rng.integers(0, 256, (600, 32), dtype=np.uint8)
ks = np.concatenate([[0, 1, 2, 49, 50, 51, 99, 100, 101, 255, 256], rng.integers(0, 257, 589)]).astype(np.int32)
b = np.stack([flip(a[i], int(ks[i])) for i in range(600)])
dist = np.array([int(cv2.norm(a[i], b[i], cv2.NORM_HAMMING)) for i in range(600)], np.int32)
assert np.array_equal(dist, ks)

db = rng.integers(0, 256, (500, 32), dtype=np.uint8)
q = np.stack([flip(db[int(rng.integers(0, 500))], int(rng.integers(0, 80))) if i % 3 else rng.integers(0, 256, 32, dtype=np.uint8)
              for i in range(400)])
knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
d1 = np.array([int(m[0].distance) for m in knn], np.int32)
d2 = np.array([int(m[1].distance) for m in knn], np.int32)
i1 = np.array([m[0].trainIdx for m in knn], np.int32)
np.savez_compressed(os.path.join(HERE, "prim_hamming.npz"), a=a, b=b, dist=dist, q=q, db=db, best=d1, second=d2, best_idx=i1)
print("pairs", len(dist), "knn", len(d1), "unique-best", int((d1 < d2).sum()))
