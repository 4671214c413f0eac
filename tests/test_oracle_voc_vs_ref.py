"""ORBVocabulary::transform: the restatement (oracle/voc_oracle.cc) against the UNMODIFIED vendored DBoW2 of the
reference (oracle/_ref/libvoc_ref.so) on synthetic vocabularies loaded with the reference's own loadFromTextFile: word
ids, bit-identical double word values, feature vectors."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from vocdata import features_for, make_vocabulary, tree_from, write_text

CASES = [  # k, L, scoring, weighting, ragged
    (10, 3, 0, 0, False),   # the ORBvoc family: TF_IDF weights, L1 scoring
    (10, 4, 0, 0, False),
    (4, 5, 0, 0, True),
    (9, 3, 1, 0, False),    # L2 norm
    (10, 3, 5, 0, False),   # dot product: no normalisation, TF values divided by the number of words
    (6, 3, 5, 1, True),     # TF
    (6, 3, 0, 2, False),    # IDF: addIfNotExist
    (6, 3, 2, 3, True),     # BINARY, chi-square
    (20, 2, 3, 0, False),   # widest branching the loader accepts
    (2, 6, 4, 0, False),
]


def same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.fixture(scope="module")
def vocs():
    made = []

    def get(case):
        k, L, sc, we, rag = case
        path = write_text(make_vocabulary(k, L, sc, we, seed=31 * k + L, ragged=rag))
        v = po.RefVocabulary(path)
        made.append((v, path))
        return v
    yield get
    for v, path in made:
        v.close()
        os.unlink(path)


@pytest.mark.parametrize("case", CASES)
def test_transform_restatement_equals_dbow2(vocs, case):
    k, L, sc, we, rag = case
    voc = make_vocabulary(k, L, sc, we, seed=31 * k + L, ragged=rag)
    v = vocs(case)
    assert (v.k, v.depth, v.scoring, v.weighting) == (k, L, sc, we) and v.n_nodes == len(voc["parent"]) + 1
    tree = v.tree()
    assert np.array_equal(tree["desc"][1:], voc["desc"]) and np.array_equal(tree["weight"][1:], voc["weight"])
    mine = tree_from(voc)  # what the GPU tests hand to eaof_voc_create == what the reference's loader builds
    for key in ("child_start", "child_idx", "word_id", "weight"):
        assert np.array_equal(mine[key], tree[key]), key
    words = 0
    for n, levelsup in [(1000, 4), (1, 4), (0, 4), (37, 0), (500, 1), (500, 2), (300, L), (300, L + 3)]:
        if rag and 0 < L - levelsup:  # a branch may end above the requested level: *nid is then unset in the reference
            depth_ok = levelsup >= L - 2
            if not depth_ok:
                continue
        f = features_for(voc, n, seed=n + levelsup)
        r = v.transform(f, levelsup)
        o = po.o_voc_transform(tree, f, levelsup)
        assert same(r, o), (n, levelsup)
        words += len(r[0])
        if n:
            assert len(r[4]) <= n and np.all(np.diff(r[0].astype(np.int64)) > 0) and np.all(np.diff(r[2].astype(np.int64)) > 0)
    assert words > 50


def test_stopped_words_are_dropped_from_both_vectors(vocs):
    case = CASES[0]
    v = vocs(case)
    tree = v.tree()
    voc = make_vocabulary(*case[:4], seed=31 * case[0] + case[1], ragged=case[4])
    f = features_for(voc, 2000, seed=5)
    wi, wv, ni, ns, fi = v.transform(f)
    assert len(fi) < 2000 and np.all(wv > 0)  # some features fell into weight-0 words
    assert same((wi, wv, ni, ns, fi), po.o_voc_transform(tree, f))
