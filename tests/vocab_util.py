"""Synthetic vocabulary trees for the BoW tests (the real ORBvoc.txt is not shipped with the reference checkout)."""
import numpy as np


def make_tree(k, L, seed, ragged=False, interleave=False):
    """Random vocabulary tree in the reference's node order (breadth-first like ORBvoc.txt is not required: any order with
    parent < child works).  Returns parent, desc, weight, is_leaf."""
    rng = np.random.default_rng(seed)
    parent, level = [0], [0]
    frontier = [0]
    for lev in range(1, L + 1):
        nxt = []
        plan = []
        for p in frontier:
            nch = k if not ragged else int(rng.integers(1, k + 1))
            if ragged and lev > 1 and rng.random() < 0.15:
                continue                                       # an early leaf
            plan.append([p, nch])
        if interleave:                                         # siblings are NOT consecutive node ids (generic child-list path)
            while any(c > 0 for _, c in plan):
                for e in plan:
                    if e[1] > 0:
                        e[1] -= 1
                        parent.append(e[0]); level.append(lev); nxt.append(len(parent) - 1)
        else:
            for p, nch in plan:
                for _ in range(nch):
                    parent.append(p); level.append(lev); nxt.append(len(parent) - 1)
        frontier = nxt
    n = len(parent)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    # children close to their parent so that descents are meaningful; plus exact duplicate siblings to exercise "first wins"
    for i in range(1, n):
        if parent[i] != 0:
            d = desc[parent[i]].copy()
            for b in rng.choice(256, 40, replace=False):
                d[b >> 3] ^= np.uint8(1 << (b & 7))
            desc[i] = d
    for i in range(2, n, 17):
        if parent[i] == parent[i - 1]:
            desc[i] = desc[i - 1]
    has_child = np.zeros(n, bool)
    has_child[np.array(parent[1:])] = True
    is_leaf = (~has_child).astype(np.uint8)
    is_leaf[0] = 0
    weight = np.where(is_leaf == 1, rng.uniform(0.0, 9.0, n), 0.0)
    weight[(is_leaf == 1) & (rng.random(n) < 0.05)] = 0.0       # stopped words
    return np.array(parent, np.int32), desc, weight, is_leaf


def write_text(path, k, L, parent, desc, weight, is_leaf, scoring=0, weighting=0, trailing_newline=False):
    """The layout TemplatedVocabulary::saveToTextFile writes (ORBvoc.txt, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1443-1464).
    By default WITHOUT the final newline: the reference's loader (`while(!f.eof()) getline`, :1393) turns the empty last line of a file
    that ends in a newline into one more child of the root whose descriptor is whatever cv::Mat::create left in memory
    (uninitialised) — undefined in the reference; the product skips blank lines."""
    lines = [f"{k} {L} {scoring} {weighting}"]
    for i in range(1, len(parent)):
        lines.append(f"{parent[i]} {int(is_leaf[i])} " + " ".join(str(int(b)) for b in desc[i]) + f" {float(weight[i])!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))
