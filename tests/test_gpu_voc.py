"""GPU parity of the bag-of-words conversion (include/eaof_voc.h): word ids, bit-identical double word values and feature
vectors against the oracle restatement (pinned to the unmodified vendored DBoW2 in tests/test_oracle_voc_vs_ref.py), and
the drop-in ORBVocabulary class against the DBoW2 class through the same harness."""
import os

import numpy as np
import pytest

from vocdata import features_for, make_vocabulary, tree_from, write_text

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN_SO = os.path.join(ROOT, "tests", "cpp", "_build", "libvoc_dropin.so")

CASES = [(10, 3, 0, 0, False), (10, 4, 0, 0, False), (4, 5, 0, 0, True), (9, 3, 1, 0, False), (10, 3, 5, 0, False),
         (6, 3, 5, 1, True), (6, 3, 0, 2, False), (6, 3, 2, 3, True), (20, 2, 3, 0, False), (2, 6, 4, 0, False)]


def same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("case", CASES)
def test_transform_matches_oracle(case):
    import eaof
    from oracle import pyoracle as po
    k, L, sc, we, rag = case
    voc = make_vocabulary(k, L, sc, we, seed=31 * k + L, ragged=rag)
    tree = tree_from(voc)
    v = eaof.ORBVocabulary(tree, max_features=2048, max_sets=4)
    words = 0
    for n, levelsup in [(1000, 4), (1, 4), (0, 4), (37, 0), (500, 1), (500, 2), (300, L), (300, L + 3), (2048, 2)]:
        if rag and 0 < L - levelsup and levelsup < L - 2:
            continue  # a branch may end above the requested level: *nid is unset in the reference
        f = features_for(voc, n, seed=n + levelsup)
        g = v.transform(f, levelsup)
        o = po.o_voc_transform(tree, f, levelsup)
        assert same(g, o), (n, levelsup)
        words += len(g[0])
    # several sets in one call
    sets = [features_for(voc, n, seed=70 + n) for n in (700, 0, 1, 333)]
    for g, f in zip(v.transform_sets(sets, min(2, L)), sets):
        assert same(g, po.o_voc_transform(tree, f, min(2, L)))
    assert words > 50
    v.close()


def test_transform_of_extracted_frames_on_the_device(frames640):
    """Extractor batch -> eaof_voc_transform_orb_device, everything device-resident; checked per frame against the host call."""
    import torch
    import eaof
    from oracle import pyoracle as po
    voc = make_vocabulary(10, 4, 0, 0, seed=5)
    tree = tree_from(voc)
    n = 4
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=n)
    res = ex.extract_batch(frames640[:n])
    cap = ex.cap
    v = eaof.ORBVocabulary(tree, max_features=cap, max_sets=n)
    dev = torch.device("cuda:0")
    nw, nn = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev)
    wi, ni, fi = (torch.zeros(n * cap, dtype=torch.int32, device=dev) for _ in range(3))
    wv = torch.zeros(n * cap, dtype=torch.float64, device=dev)
    ns = torch.zeros(n * (cap + 1), dtype=torch.int32, device=dev)
    v.transform_orb_device(ex, n, 4, nw.data_ptr(), wi.data_ptr(), wv.data_ptr(), nn.data_ptr(), ni.data_ptr(), ns.data_ptr(),
                           fi.data_ptr())
    v.sync()
    for f in range(n):
        o = po.o_voc_transform(tree, res[f][1], 4)
        a, b = int(nw[f]), int(nn[f])
        st = ns[f * (cap + 1):f * (cap + 1) + b + 1].cpu().numpy()
        g = (wi[f * cap:f * cap + a].cpu().numpy().astype(np.uint32), wv[f * cap:f * cap + a].cpu().numpy(),
             ni[f * cap:f * cap + b].cpu().numpy().astype(np.uint32), st, fi[f * cap:f * cap + st[-1]].cpu().numpy().astype(np.uint32))
        assert same(g, o), f
        assert a > 300
    v.close()
    ex.close()


def test_bad_trees_are_rejected():
    import eaof
    voc = make_vocabulary(4, 2, seed=1)
    tree = tree_from(voc)
    bad = dict(tree)
    bad["child_idx"] = tree["child_idx"].copy()
    bad["child_idx"][3] = bad["child_idx"][2]
    with pytest.raises(eaof.EaofError):
        eaof.ORBVocabulary(bad)
    v = eaof.ORBVocabulary(tree, max_features=16)
    with pytest.raises(eaof.EaofError):
        v.transform(np.zeros((17, 32), np.uint8))
    v.close()


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[4], CASES[7]])
def test_dropin_vocabulary_class(case):
    """eao-fusion_b200/dropin/ORBVocabulary.h (subclass of the vendored DBoW2 vocabulary, GPU transform through virtual
    dispatch) against the unmodified DBoW2 class: same text file, same features, identical BowVector / FeatureVector."""
    from oracle import pyoracle as po
    if not os.path.exists(DROPIN_SO):
        pytest.fail(f"{DROPIN_SO} missing: run `make` where /root/reference is present")
    k, L, sc, we, rag = case
    voc = make_vocabulary(k, L, sc, we, seed=31 * k + L, ragged=rag)
    path = write_text(voc)
    try:
        r = po.RefVocabulary(path)
        d = po.RefVocabulary(path, L=po.voc_harness_lib(DROPIN_SO))
        assert (d.k, d.depth, d.n_nodes, d.n_words) == (r.k, r.depth, r.n_nodes, r.n_words)
        for n, levelsup in [(1000, min(4, L)), (1, 1), (0, 4), (5000, 2)]:
            if rag and levelsup < L - 2:
                continue
            f = features_for(voc, n, seed=n)
            a, b = r.transform(f, levelsup), d.transform(f, levelsup)
            assert same(a, b), (n, levelsup)
            if n:
                assert r.score(a[:2], b[:2]) == r.score(a[:2], a[:2])
        r.close()
        d.close()
    finally:
        os.unlink(path)


@pytest.mark.parametrize("mode", [0, 1])
def test_search_by_bow_device_resident_end_to_end(frames640, mode):
    """extraction -> ORBVocabulary::transform -> SearchByBoW, nothing leaving the device in between; per pair against the
    oracle's SearchByBoW fed with the oracle's FeatureVectors (TrackReferenceKeyFrame / loop-candidate shape)."""
    import torch
    import eaof
    from oracle import pyoracle as po
    voc = make_vocabulary(10, 4, 0, 0, seed=5)
    tree = tree_from(voc)
    n = 5
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=n)
    res = ex.extract_batch(frames640[:n])
    cap = ex.cap
    v = eaof.ORBVocabulary(tree, max_features=cap, max_sets=n)
    mt = eaof.ORBmatcher(0.7, True, max_features=cap, max_pairs=8)
    dev = torch.device("cuda:0")
    nw, nn = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev)
    wi, ni, fi = (torch.zeros(n * cap, dtype=torch.int32, device=dev) for _ in range(3))
    wv = torch.zeros(n * cap, dtype=torch.float64, device=dev)
    ns = torch.zeros(n * (cap + 1), dtype=torch.int32, device=dev)
    levelsup = 2
    v.transform_orb_device(ex, n, levelsup, nw.data_ptr(), wi.data_ptr(), wv.data_ptr(), nn.data_ptr(), ni.data_ptr(), ns.data_ptr(),
                           fi.data_ptr())
    v.sync()
    pq = np.array([0, 1, 2, 3, 4, 0], np.int32)
    pt = np.array([1, 2, 3, 4, 0, 0], np.int32)
    d_match = torch.full((len(pq), cap), -7, dtype=torch.int32, device=dev)
    d_dist = torch.full((len(pq), cap), -7, dtype=torch.int32, device=dev)
    d_n = torch.zeros(len(pq), dtype=torch.int32, device=dev)
    mt.bow_orb_device(ex, n, mode, pq, pt, nn.data_ptr(), ni.data_ptr(), ns.data_ptr(), fi.data_ptr(), d_match.data_ptr(),
                      d_dist.data_ptr(), d_n.data_ptr())
    mt.sync()
    fvs = [po.o_voc_transform(tree, res[f][1], levelsup) for f in range(n)]
    total = 0
    for p, (q, t) in enumerate(zip(pq, pt)):
        kq, dq = res[q]
        kt, dt = res[t]
        nodes_q = (fvs[q][2].astype(np.int32), fvs[q][3], fvs[q][4].astype(np.int32))
        nodes_t = (fvs[t][2].astype(np.int32), fvs[t][3], fvs[t][4].astype(np.int32))
        on, om, od = po.o_search_by_bow(mode, 0.7, True, dq, kq["angle"], None, nodes_q, dt, kt["angle"], None, nodes_t)
        nout = len(kt) if mode == 0 else len(kq)
        assert int(d_n[p]) == on, (p, int(d_n[p]), on)
        assert np.array_equal(d_match[p, :nout].cpu().numpy(), om) and np.array_equal(d_dist[p, :nout].cpu().numpy(), od), p
        total += on
    assert total > 1000
    mt.close()
    v.close()
    ex.close()


def test_cuda_path_reproduces_the_frozen_reference_answers():
    """CUDA outputs hashed straight against tests/golden/frontend_hashes.json (frozen from the reference's own code), with no
    oracle in between: vocabulary transform and both SearchByBoW overloads."""
    import json
    import eaof
    from golden_frontend import digest
    from matchdata import planted_pair, random_nodes
    with open(os.path.join(ROOT, "tests", "golden", "frontend_hashes.json")) as f:
        gold = json.load(f)
    voc = make_vocabulary(10, 3, 0, 0, seed=313)
    v = eaof.ORBVocabulary(tree_from(voc), max_features=1024)
    assert digest(*v.transform(features_for(voc, 1000, seed=1004), 2)) == gold["voc_transform"]["sha256"]
    v.close()
    for mode in (0, 1):
        q, aq, t, at = planted_pair(600, 600, 16, dup=5)
        nq, nt = eaof.csr_from_nodes(random_nodes(600, 12, 26)), eaof.csr_from_nodes(random_nodes(600, 12, 36))
        rng = np.random.Generator(np.random.PCG64(106))
        vq, vt = (rng.random(600) > 0.1).astype(np.uint8), (rng.random(600) > 0.1).astype(np.uint8)
        m = eaof.ORBmatcher(0.75, True, max_features=1024)
        n, match, _ = m.SearchByBoW(mode, q, aq, vq, nq, t, at, vt, nt)
        assert digest(np.asarray([n], np.int32), match) == gold[f"search_by_bow_mode{mode}"]["sha256"], mode
        m.close()
