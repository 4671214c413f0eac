"""Synthetic DBoW2 vocabularies in the text format of TemplatedVocabulary::loadFromTextFile
(/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1349-1437): first line "k L scoring weighting", then one
line per node in creation order: "parent isLeaf d0 .. d31 weight".  The reference's ORBvoc (k=10, L=6) is not in the tree
(.MISSING_LARGE_BLOBS), so tests grow trees of the same shape family: children are noisy copies of their parent, with
fewer flipped bits the deeper the level, so that descents are decided by small distance margins and ties do occur."""
import os
import tempfile

import numpy as np


def make_vocabulary(k, L, scoring=0, weighting=0, seed=0, stop_frac=0.03, ragged=False):
    """Returns dict(k, L, scoring, weighting, parent[], is_leaf[], desc[n,32] u8, weight[] f64) for nodes 1..n (node 0 is
    the root and has no line).  ragged: some inner nodes get fewer than k children and some branches end early."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parent, is_leaf, desc, weight = [], [], [], []
    root = rng.integers(0, 256, 32, dtype=np.uint8)
    frontier = [(0, root, 0)]  # (node id, descriptor, level)
    next_id = 1
    while frontier:
        nid, d, lvl = frontier.pop(0)
        nch = k if not ragged or lvl == 0 else int(rng.integers(max(1, k // 2), k + 1))
        for c in range(nch):
            flips = rng.random(256) < 0.5 / (1.6 ** lvl) / 2
            cd = d ^ np.packbits(flips.astype(np.uint8))
            if c == nch - 1 and c > 0 and rng.random() < 0.2:
                cd = desc[-1].copy()  # twin of the previous sibling: equal distances, the first child must win
            leaf = lvl + 1 == L or (ragged and lvl + 1 >= 2 and rng.random() < 0.1)
            parent.append(nid); is_leaf.append(int(leaf)); desc.append(cd)
            if leaf:
                w = 0.0 if rng.random() < stop_frac else float(rng.uniform(0.1, 12.0))
                weight.append(w)
            else:
                weight.append(0.0)
                frontier.append((next_id, cd, lvl + 1))
            next_id += 1
    return dict(k=k, L=L, scoring=scoring, weighting=weighting, parent=np.asarray(parent, np.int32),
                is_leaf=np.asarray(is_leaf, np.int32), desc=np.asarray(desc, np.uint8), weight=np.asarray(weight, np.float64))


def write_text(voc, path=None):
    """No trailing newline: loadFromTextFile's `while(!f.eof())` would otherwise parse one more (empty) line into a node
    with an uninitialised parent id."""
    if path is None:
        fd, path = tempfile.mkstemp(suffix=".txt", prefix="eaof_voc_")
        os.close(fd)
    lines = [f"{voc['k']} {voc['L']} {voc['scoring']} {voc['weighting']}"]
    for p, l, d, w in zip(voc["parent"], voc["is_leaf"], voc["desc"], voc["weight"]):
        lines.append(f"{p} {l} " + " ".join(str(int(b)) for b in d) + f" {float(w)!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return path


def features_for(voc, n, seed=1, noise=0.04):
    """n query descriptors: noisy copies of random leaves (most), of inner nodes (some) and random ones (a few)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    src = voc["desc"][rng.integers(0, len(voc["desc"]), n)]
    out = src ^ np.packbits((rng.random((n, 256)) < noise).astype(np.uint8), axis=1)
    rnd = rng.random(n) < 0.05
    out[rnd] = rng.integers(0, 256, (int(rnd.sum()), 32), dtype=np.uint8)
    if n > 4:
        out[1] = out[0]  # duplicate features land in the same word: addWeight accumulates
    return out


def tree_from(voc):
    """The arrays eaof_voc_create takes, built directly from the generator's node list (children in node order, which
    is what loadFromTextFile's push_back produces; word ids in leaf order)."""
    n = len(voc["parent"]) + 1
    parent = np.concatenate([[-1], voc["parent"]]).astype(np.int32)
    order = np.argsort(parent[1:], kind="stable") + 1
    child_start = np.searchsorted(parent[order], np.arange(n + 1)).astype(np.int32)
    leaf = np.concatenate([[0], voc["is_leaf"]])
    word_id = np.where(leaf > 0, np.cumsum(leaf) - 1, -1).astype(np.int32)
    return dict(L=voc["L"], child_start=child_start, child_idx=order.astype(np.int32),
                desc=np.concatenate([np.zeros((1, 32), np.uint8), voc["desc"]]),
                weight=np.concatenate([[0.0], voc["weight"]]), word_id=word_id, weighting=voc["weighting"], scoring=voc["scoring"])
