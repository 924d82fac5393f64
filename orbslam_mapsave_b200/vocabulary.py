"""Python mirror of the ORB vocabulary's BoW assignment (DBoW2 TemplatedVocabulary::transform) over the C ABI."""
import ctypes as C

import numpy as np

from . import capi
from .matcher import FeatureVector


class ORBVocabulary:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_arrays(cls, k, L, parent, descriptors, weights, is_leaf, scoring=0, weighting=0, device=0):
        parent = np.ascontiguousarray(parent, np.int32)
        descriptors = np.ascontiguousarray(descriptors, np.uint8)
        weights = np.ascontiguousarray(weights, np.float64)
        is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        h = C.c_void_p()
        capi.check(capi.lib().orbv_create(C.byref(h), k, L, scoring, weighting, len(parent), capi._p(parent), capi._p(descriptors),
                                          capi._p(weights), capi._p(is_leaf), device))
        return cls(h)

    @classmethod
    def loadFromTextFile(cls, path, device=0):
        h = C.c_void_p()
        capi.check(capi.lib().orbv_load_text(C.byref(h), str(path).encode(), device))
        return cls(h)

    @classmethod
    def loadFromBinaryFile(cls, path, device=0):
        h = C.c_void_p()
        capi.check(capi.lib().orbv_load_binary(C.byref(h), str(path).encode(), device))
        return cls(h)

    def saveToBinaryFile(self, path):
        capi.check(capi.lib().orbv_save_binary(self._h, str(path).encode()))

    def info(self):
        v = [C.c_int() for _ in range(6)]
        capi.check(capi.lib().orbv_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("k", "L", "n_nodes", "n_words", "scoring", "weighting"), (x.value for x in v)))

    def close(self):
        if self._h is not None:
            capi.lib().orbv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def transform_raw(self, desc, levelsup=4):
        """Per descriptor: (word id, weight, node id at level L - levelsup)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        w, nid = np.zeros(n, np.int32), np.zeros(n, np.int32)
        wt = np.zeros(n, np.float64)
        capi.check(capi.lib().orbv_transform(self._h, capi._p(desc), n, levelsup, capi._p(w), capi._p(wt), capi._p(nid)))
        return w, wt, nid

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup): returns (bow: dict word -> value, FeatureVector of the
        features whose word is not stopped), DBoW2 TF-IDF / L1 semantics taken from the vocabulary's header."""
        w, wt, nid = self.transform_raw(desc, levelsup)
        keep = wt > 0
        return w, wt, FeatureVector_from_nodes(nid, keep)


def FeatureVector_from_nodes(node_id, keep):
    idx = np.nonzero(keep)[0]
    nodes = node_id[idx].astype(np.int64)
    order = np.argsort(nodes, kind="stable")
    ids, counts = np.unique(nodes, return_counts=True)
    off = np.concatenate([[0], np.cumsum(counts)])
    return FeatureVector(ids=ids, off=off, feat=idx[order])
