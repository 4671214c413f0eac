"""Known-answer cases for everything beside the extractor: one seeded input per reference function, the function's outputs
produced twice — by the reference's own code run here (oracle/_ref binaries, `ref`) and by the oracle restatement
(`oracle`).  tests/golden/make_golden_frontend.py freezes SHA-256 digests of the `ref` outputs in
tests/golden/frontend_hashes.json; tests/test_golden_frontend.py requires the restatement to reproduce them without the
reference binaries (they do not exist on the GPU box unless shipped prebuilt), and the GPU tests compare the CUDA path
with the same restatement."""
import hashlib

import numpy as np

_SF = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
_BOUNDS = (0.0, 640.0, 0.0, 480.0)
_GINV = (np.float32(64) / np.float32(640), np.float32(48) / np.float32(480))
_KW = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF)
_LSF = float(np.log(np.float32(1.2)))
_INV_SIGMA2 = (np.float32(1.0) / (_SF * _SF)).astype(np.float32)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def cases():
    """name -> (ref, oracle): two zero-argument callables returning the tuple of arrays that is hashed."""
    import eaof
    from matchdata import (fuse_scene, init_scene, kf_scene, map_scene, planted_pair, random_nodes, sim3_scene, stereo_pair,
                           tri_inputs, window_scene)
    from oracle import pyoracle as po
    from test_oracle_matcher_vs_ref import (_proj_inputs, fuse_case, fuse_sim3_case, sim3_queries, sim3kf_case)
    from vocdata import features_for, make_vocabulary, write_text
    out = {}
    i32 = lambda v: np.asarray([v], np.int32)

    # --- ORBmatcher: BoW, triangulation, the projection family, initialisation
    for mode in (0, 1):
        q, aq, t, at = planted_pair(600, 600, 16, dup=5)
        nq, nt = eaof.csr_from_nodes(random_nodes(600, 12, 26)), eaof.csr_from_nodes(random_nodes(600, 12, 36))
        rng = np.random.Generator(np.random.PCG64(106))
        vq, vt = (rng.random(600) > 0.1).astype(np.uint8), (rng.random(600) > 0.1).astype(np.uint8)
        args = (mode, 0.75, True, q, aq, vq, nq, t, at, vt, nt)
        out[f"search_by_bow_mode{mode}"] = (lambda a=args: (lambda r: (i32(r[0]), r[1]))(po.r_search_by_bow(*a)),
                                            lambda a=args: (lambda r: (i32(r[0]), r[1]))(po.o_search_by_bow(*a)))
    k1, k2, F12, sf, ls = tri_inputs(51)
    targs = (k1, k2, F12, (320.0, 240.0), sf, ls, False, True)
    out["search_for_triangulation"] = (lambda: (lambda r: (i32(r[0]), r[1]))(po.r_search_for_triangulation(*targs)),
                                       lambda: (lambda r: (i32(r[0]), r[1]))(po.o_search_for_triangulation(*targs)))
    cur, last = _proj_inputs(201, stereo=True, flags=True)
    pkw = dict(bounds=_BOUNDS, grid_inv=_GINV, scale_factors=_SF, mbf=40.0, search_mode=0)
    out["search_by_projection_last"] = (lambda: (lambda r: (i32(r[0]), r[1]))(po.r_search_by_projection(cur, last, 15.0, True, **pkw)),
                                        lambda: (lambda r: (i32(r[0]), r[1]))(po.o_search_by_projection(cur, last, 15.0, True, **pkw)))
    F, mp = window_scene(301, stereo=True, flags=True)
    out["search_by_projection_mappoints"] = (
        lambda: (lambda r: (i32(r[0]), r[1]))(po.r_search_by_projection_mappoints(F, mp, 3.0, 0.8, **_KW)),
        lambda: (lambda r: (i32(r[0]), r[1]))(po.o_search_by_projection_mappoints(F, mp, 3.0, 0.8, **_KW)))
    Fk, kf = kf_scene(401, flags=True)

    def kf_oracle():
        valid, u, v, lvl = po.kf_projection(kf, _LSF, 8)
        kq = dict(valid=valid, u=u, v=v, level=lvl, angle=kf["angle"], desc=kf["desc"])
        r = po.o_search_by_projection_kf(Fk, kq, 15.0, 100, True, **_KW)
        return i32(r[0]), r[1]
    out["search_by_projection_kf"] = (
        lambda: (lambda r: (i32(r[0]), r[1]))(po.r_search_by_projection_kf(Fk, kf, 15.0, 100, True, log_scale_factor=_LSF, **_KW)),
        kf_oracle)
    F1, F2, prev = init_scene(501)
    ikw = dict(bounds=_BOUNDS, grid_inv=_GINV)
    out["search_for_initialization"] = (
        lambda: (lambda r: (i32(r[0]), r[1], r[2]))(po.r_search_for_initialization(F1, F2, prev, 100, 0.9, True, **ikw)),
        lambda: (lambda r: (i32(r[0]), r[1], r[2]))(po.o_search_for_initialization(F1, F2, prev, 100, 0.9, True, **ikw)))

    # --- map-side matchers
    KFa, ptsa, qa = sim3kf_case(601, True)
    out["search_by_projection_sim3kf"] = (
        lambda: (lambda r: (i32(r[0]), r[1]))(po.r_search_by_projection_sim3kf(KFa, ptsa, 10, log_scale_factor=_LSF, scw_scale=2.0, **_KW)),
        lambda: (lambda r: (i32(r[0]), np.where(r[1] >= 0, r[1], KFa["matched_q"]).astype(np.int32)))(
            po.o_search_by_projection_sim3kf(KFa, qa, 10, **_KW)))
    KFb, ptsb, qb = fuse_case(701, True)

    def fuse_oracle():
        _, mq, _ = po.o_window_best(1, KFb, qb, 3.0, 50, inv_level_sigma2=_INV_SIGMA2, **_KW)
        r = po.o_fuse_apply(mq, ptsb["state"], ptsb["qid"], ptsb["obs"], KFb["slot_state"], KFb["slot_obs"])
        return (i32(r[0]),) + tuple(r[1:])
    out["fuse"] = (lambda: (lambda r: (i32(r[0]),) + tuple(r[1:]))(
        po.r_fuse(KFb, ptsb, 3.0, inv_level_sigma2=_INV_SIGMA2, log_scale_factor=_LSF, mbf=20.0, **_KW)), fuse_oracle)
    KFc, ptsc, qc = fuse_sim3_case(801)

    def fuse_sim3_oracle():
        _, mq, _ = po.o_window_best(0, KFc, qc, 4.0, 50, **_KW)
        r = po.o_fuse_sim3_apply(mq, KFc["slot_state"], KFc["slot_query"])
        return (i32(r[0]),) + tuple(r[1:])
    out["fuse_sim3"] = (lambda: (lambda r: (i32(r[0]),) + tuple(r[1:]))(
        po.r_fuse_sim3(KFc, ptsc, 4.0, log_scale_factor=_LSF, scw_scale=4.0, **_KW)), fuse_sim3_oracle)
    K1, K2, pre12 = sim3_scene(901, flags=True)

    def sim3_oracle():
        q1, q2 = sim3_queries(K1, K2, pre12)
        _, m1, _ = po.o_window_best(0, K2, q1, 7.5, 100, **_KW)
        _, m2, _ = po.o_window_best(0, K1, q2, 7.5, 100, **_KW)
        n, m = po.o_sim3_agreement(m1, m2)
        return i32(n), np.where(m >= 0, m, pre12).astype(np.int32)
    out["search_by_sim3"] = (lambda: (lambda r: (i32(r[0]), r[1]))(po.r_search_by_sim3(K1, K2, pre12, 7.5, log_scale_factor=_LSF, **_KW)),
                             sim3_oracle)

    # --- MapPoint::ComputeDistinctiveDescriptors
    rng = np.random.Generator(np.random.PCG64(78))
    base = rng.integers(0, 256, size=(1, 32), dtype=np.uint8)
    dd = base ^ np.packbits((rng.random((33, 256)) < rng.uniform(0.0, 0.3, (33, 1))).astype(np.uint8), axis=1)
    dd[32] = dd[1]
    out["distinctive_descriptor"] = (lambda: (i32(po.r_distinctive_descriptor(dd)[0]),),
                                     lambda: (i32(po.o_distinctive_descriptor(dd)[0]),))

    # --- ORBVocabulary::transform
    voc = make_vocabulary(10, 3, 0, 0, seed=313)
    feats = features_for(voc, 1000, seed=1004)

    def voc_ref():
        import os
        path = write_text(voc)
        try:
            v = po.RefVocabulary(path)
            r = v.transform(feats, 2)
            v.close()
        finally:
            os.unlink(path)
        return r

    def voc_oracle():
        from vocdata import tree_from
        return po.o_voc_transform(tree_from(voc), feats, 2)
    out["voc_transform"] = (voc_ref, voc_oracle)

    # --- Frame::ComputeStereoMatches / ComputeStereoFromRGBD on a synthetic rectified pair
    left, right = stereo_pair(0, disparity=14, half_pixel=True)

    def stereo_inputs():
        exl, exr = po.RefExtractor(1000), po.RefExtractor(1000)
        kL, dL = exl.extract(left, keep_pyramid=True)
        kR, dR = exr.extract(right, keep_pyramid=True)
        t = exl.tables()
        return (kL, dL, kR, dR, [exl.level(l, with_border=True) for l in range(8)], [exr.level(l, with_border=True) for l in range(8)],
                t["scale"], t["inv_scale"])

    def stereo_inputs_oracle():
        kL, dL, pL, _, _ = po.o_extract(left, dumps=True)
        kR, dR, pR, _, _ = po.o_extract(right, dumps=True)
        t = po.o_tables(1000)
        return kL, dL, kR, dR, pL, pR, t["scale"], t["inv_scale"]
    out["stereo_matches"] = (lambda: po.r_stereo_matches(*stereo_inputs(), 0.1, 40.0),
                             lambda: po.o_stereo_matches(*stereo_inputs_oracle(), 0.1, 40.0)[:2])
    depth = np.random.Generator(np.random.PCG64(8)).uniform(-0.5, 8.0, (480, 640)).astype(np.float32)

    out["stereo_from_rgbd"] = (lambda: po.r_stereo_from_rgbd(po.RefExtractor(1000).extract(left)[0], depth, 40.0),
                               lambda: po.o_stereo_from_rgbd(po.o_extract(left)[0], depth, 40.0))
    return out
