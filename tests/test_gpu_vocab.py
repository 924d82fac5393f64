"""GPU parity of the BoW assignment (DBoW2 TemplatedVocabulary::transform) against the CPU oracle: word ids,
weights, FeatureVector node ids bit-exact; text / binary vocabulary formats of the reference round-trip."""
import numpy as np
import pytest

import orbslam_mapsave_b200 as orb
from orbslam_mapsave_b200.synth import synth_descriptors
from oracle import orb_oracle_py as orc
from vocab_util import make_tree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,ragged,levelsup,interleave", [(10, 3, False, 1, False), (10, 4, False, 4, False), (10, 4, False, 2, True),
                                                            (7, 5, True, 3, False), (20, 2, False, 0, False), (3, 6, True, 4, True),
                                                            (18, 3, True, 1, True)])
def test_transform_matches_oracle(k, L, ragged, levelsup, interleave):
    parent, desc, weight, is_leaf = make_tree(k, L, seed=k * 10 + L, ragged=ragged, interleave=interleave)
    voc = orb.ORBVocabulary.from_arrays(k, L, parent, desc, weight, is_leaf)
    info = voc.info()
    assert info["n_nodes"] == len(parent) and info["n_words"] == int(is_leaf.sum())
    feats = synth_descriptors(3000, 5, dup_of=desc[1:], dup_rate=0.7, max_flip=30)
    w, wt, nid = voc.transform_raw(feats, levelsup)
    ow, owt, onid = orc.voc_transform(parent, desc, weight, is_leaf, L, feats, levelsup)
    assert np.array_equal(w, ow) and np.array_equal(wt, owt) and np.array_equal(nid, onid)
    # the FeatureVector built from it is the matcher's input format: node ids ascending, features ascending inside a node
    _, _, fv = voc.transform(feats, levelsup)
    keep = owt > 0
    ofv = orc.FeatVec(onid[keep])
    assert np.array_equal(fv.ids, ofv.ids) and np.array_equal(fv.off, ofv.off)
    assert np.array_equal(fv.feat, np.nonzero(keep)[0][ofv.feat])


def test_text_and_binary_vocabulary_formats(tmp_path):
    k, L = 10, 3
    parent, desc, weight, is_leaf = make_tree(k, L, seed=99)
    txt = tmp_path / "voc.txt"
    with open(txt, "w") as f:                      # the layout TemplatedVocabulary::saveToTextFile writes (ORBvoc.txt)
        f.write(f"{k} {L} 0 0\n")
        for i in range(1, len(parent)):
            f.write(f"{parent[i]} {int(is_leaf[i])} " + " ".join(str(int(b)) for b in desc[i]) + f" {float(weight[i])!r}\n")
    voc = orb.ORBVocabulary.loadFromTextFile(txt)
    assert voc.info() == dict(k=k, L=L, n_nodes=len(parent), n_words=int(is_leaf.sum()), scoring=0, weighting=0)
    feats = synth_descriptors(1500, 6, dup_of=desc[1:], dup_rate=0.6)
    got = voc.transform_raw(feats, 2)
    want = orc.voc_transform(parent, desc, weight, is_leaf, L, feats, 2)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # binary round trip: weights are stored as float32 there (TemplatedVocabulary.h:1515-1536)
    binp = tmp_path / "voc.bin"
    voc.saveToBinaryFile(binp)
    voc2 = orb.ORBVocabulary.loadFromBinaryFile(binp)
    assert voc2.info() == voc.info()
    got2 = voc2.transform_raw(feats, 2)
    want2 = orc.voc_transform(parent, desc, weight.astype(np.float32).astype(np.float64), is_leaf, L, feats, 2)
    assert all(np.array_equal(a, b) for a, b in zip(got2, want2))
    with pytest.raises(orb.OrbError):
        orb.ORBVocabulary.loadFromBinaryFile(txt)


def test_orb_vocabulary_shape_k10_L6_throughput_sanity():
    """The real ORBvoc (k=10, L=6, ~1.1 M nodes) is not shipped (.MISSING_LARGE_BLOBS); a synthetic tree of that shape checks
    the 35 MB case and levelsup=4 (the value Frame::ComputeBoW uses, src/Frame.cc:513-520)."""
    parent, desc, weight, is_leaf = make_tree(10, 6, seed=3)
    assert len(parent) == 1111111
    voc = orb.ORBVocabulary.from_arrays(10, 6, parent, desc, weight, is_leaf)
    feats = synth_descriptors(2000, 8, dup_of=desc[1:200000], dup_rate=0.8, max_flip=25)
    got = voc.transform_raw(feats, 4)
    want = orc.voc_transform(parent, desc, weight, is_leaf, 6, feats, 4)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert len(np.unique(got[2])) > 20
