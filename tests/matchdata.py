"""Synthetic matcher inputs shared by the CPU and GPU matcher tests (SURVEY.md §8(d) 'Matching inputs')."""
import numpy as np


def planted_pair(n_q, n_t, seed, flip=0.08, frac=0.7, dup=0):
    """Descriptors T = Q with each bit flipped w.p. `flip` for `frac` of the rows (shuffled), random for the rest;
    angles T = Q + N(0, 5 deg).  dup > 0 appends exact duplicates of some rows (forced ties)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q = rng.integers(0, 256, size=(n_q, 32), dtype=np.uint8)
    aq = rng.uniform(0, 360, n_q).astype(np.float32)
    t = rng.integers(0, 256, size=(n_t, 32), dtype=np.uint8)
    at = rng.uniform(0, 360, n_t).astype(np.float32)
    m = min(int(frac * min(n_q, n_t)), n_q, n_t)
    src = rng.permutation(n_q)[:m]
    dst = rng.permutation(n_t)[:m]
    noise = (rng.random((m, 256)) < flip).astype(np.uint8)
    t[dst] = q[src] ^ np.packbits(noise, axis=1)
    at[dst] = np.mod(aq[src] + rng.normal(0, 5, m), 360).astype(np.float32)
    if dup > 0 and n_t > 2 * dup:
        d = rng.permutation(n_t)[:2 * dup]
        t[d[dup:]] = t[d[:dup]]
    return q, aq, t, at


def random_nodes(n, n_nodes, seed, unassigned=0.05):
    rng = np.random.Generator(np.random.PCG64(seed))
    node = rng.integers(0, n_nodes, n) * 3 + 7  # sparse ascending ids
    node[rng.random(n) < unassigned] = -1
    return node


def frame_features(kps, desc):
    return dict(x=kps["x"].copy(), y=kps["y"].copy(), octave=kps["octave"].copy(), angle=kps["angle"].copy(), desc=desc.copy())


def projected_last(kps, desc, dx, dy):
    return dict(u=(kps["x"] + np.float32(dx)).astype(np.float32), v=(kps["y"] + np.float32(dy)).astype(np.float32),
                octave=kps["octave"].copy(), angle=kps["angle"].copy(), desc=desc.copy())


def _csr(node):
    import eaof
    return eaof.csr_from_nodes(node)


def tri_inputs(seed, n1=500, n2=600, nn=8):
    """Two keyframes whose planted pairs share a vocabulary node and lie on each other's epipolar line
    (F12 = [e_x]_x for a pure x-translation: the line of (x1, y1) is y = y1), plus ties and distractors."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d1 = rng.integers(0, 256, size=(n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, size=(n2, 32), dtype=np.uint8)
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    a2 = rng.uniform(0, 360, n2).astype(np.float32)
    x1, y1 = rng.uniform(0, 640, n1).astype(np.float32), rng.uniform(0, 480, n1).astype(np.float32)
    x2, y2 = rng.uniform(0, 640, n2).astype(np.float32), rng.uniform(0, 480, n2).astype(np.float32)
    node1, node2 = random_nodes(n1, nn, seed + 1), random_nodes(n2, nn, seed + 2)
    m = int(0.7 * min(n1, n2))
    src, dst = rng.permutation(n1)[:m], rng.permutation(n2)[:m]
    d2[dst] = d1[src] ^ np.packbits((rng.random((m, 256)) < 0.06).astype(np.uint8), axis=1)
    a2[dst] = np.mod(a1[src] + rng.normal(0, 5, m), 360).astype(np.float32)
    y2[dst] = (y1[src] + rng.normal(0, 1.5, m)).astype(np.float32)
    node2[dst] = node1[src]
    # exact duplicates of some planted targets inside the same node: equal distance, the LATER candidate must win (:738)
    ntw = min(20, m)
    extra = rng.permutation(n2)[:ntw]
    twin = dst[:ntw]
    d2[extra] = d2[twin]; y2[extra] = y2[twin]; node2[extra] = node2[twin]; a2[extra] = a2[twin]
    k1 = dict(desc=d1, angle=a1, x=x1, y=y1, free=(rng.random(n1) > 0.2).astype(np.uint8),
              stereo=(rng.random(n1) > 0.5).astype(np.uint8), nodes=_csr(node1))
    k2 = dict(desc=d2, angle=a2, x=x2, y=y2, octave=rng.integers(0, 8, n2).astype(np.int32),
              free=(rng.random(n2) > 0.2).astype(np.uint8), stereo=(rng.random(n2) > 0.5).astype(np.uint8), nodes=_csr(node2))
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32) + rng.normal(0, 1e-5, (3, 3)).astype(np.float32)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    return k1, k2, F12, sf, (sf * sf).astype(np.float32)


def window_scene(seed, n=900, stereo=False, flags=False):
    """A frame of n features and n projected map points, half of which land next to a feature whose descriptor is a
    noisy copy of theirs, with ties, crowded windows and (optionally) stereo / taken / validity flags."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q, aq, t, at = planted_pair(n, n, seed, flip=0.06, frac=0.8, dup=8)
    x = rng.uniform(-5, 645, n).astype(np.float32)
    y = rng.uniform(-5, 485, n).astype(np.float32)
    octv = rng.integers(0, 8, n).astype(np.int32)
    F = dict(x=x, y=y, octave=octv, angle=at, desc=t)
    perm = rng.permutation(n)
    mp = dict(x=(x[perm] + rng.normal(0, 2, n)).astype(np.float32), y=(y[perm] + rng.normal(0, 2, n)).astype(np.float32),
              level=np.clip(octv[perm] + rng.integers(0, 2, n), 0, 7).astype(np.int32), angle=aq, desc=q,
              cos=rng.choice(np.array([0.9, 0.999, 0.9985], np.float32), n))
    half = perm[: n // 2]
    F["desc"][half] = q[: n // 2] ^ np.packbits((rng.random((n // 2, 256)) < 0.05).astype(np.uint8), axis=1)
    # near-ties on the same level: a second feature right next to some planted ones with an almost equal descriptor
    twins = half[:60]
    extra = perm[n // 2: n // 2 + 60]
    F["x"][extra] = F["x"][twins] + 1; F["y"][extra] = F["y"][twins]; F["octave"][extra] = F["octave"][twins]
    F["desc"][extra] = F["desc"][twins] ^ np.packbits((rng.random((60, 256)) < 0.01).astype(np.uint8), axis=1)
    if stereo:
        F["uright"] = np.where(rng.random(n) > 0.3, x - 20 + rng.normal(0, 4, n), -1).astype(np.float32)
        mp["xr"] = (mp["x"] - 20 + rng.normal(0, 3, n)).astype(np.float32)
    if flags:
        F["taken"] = (rng.random(n) < 0.1).astype(np.uint8)
        mp["in_view"] = (rng.random(n) > 0.1).astype(np.uint8)
        mp["bad"] = (rng.random(n) < 0.05).astype(np.uint8)
        mp["obs"] = (rng.random(n) > 0.2).astype(np.uint8)
    return F, mp


def kf_scene(seed, n=900, flags=False):
    """Cur frame + keyframe map points for SearchByProjection(Cur,KF): world points (u, v, 1) seen from an identity
    pose, distance gates and predicted levels spread over the pyramid."""
    F, mp = window_scene(seed, n)
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    d = np.sqrt(mp["x"].astype(np.float64) ** 2 + mp["y"].astype(np.float64) ** 2 + 1.0)
    lvl = mp["level"]
    # mfMaxDistance so that ceil(log(max/d)/log(1.2)) == lvl with margin: max = d * 1.2^(lvl-0.5)
    max_dist = (d * 1.2 ** (lvl - 0.5)).astype(np.float32)
    min_dist = (max_dist / sf[7]).astype(np.float32)
    state = np.full(n, 3, np.uint8)
    if flags:
        state = rng.choice(np.array([0, 1, 2, 3], np.uint8), n, p=[0.05, 0.05, 0.1, 0.8])
        F["taken"] = (rng.random(n) < 0.1).astype(np.uint8)
        far = rng.random(n) < 0.05
        max_dist[far] = (d[far] * 0.5).astype(np.float32)  # outside the scale pyramid -> skipped by the distance gate
    kf = dict(state=state, wx=mp["x"], wy=mp["y"], wz=np.ones(n, np.float32), max_dist=max_dist, min_dist=min_dist,
              angle=mp["angle"], desc=mp["desc"])
    return F, kf


def init_scene(seed, n=1200):
    """Two monocular frames for SearchForInitialization: F2 = F1 moved by a few pixels, descriptors noisy copies, with
    clusters of look-alike F2 features so that later F1 features steal earlier matches (:443-469)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a1 = rng.uniform(0, 360, n).astype(np.float32)
    x1, y1 = rng.uniform(0, 640, n).astype(np.float32), rng.uniform(0, 480, n).astype(np.float32)
    o1 = rng.choice(np.arange(3), n, p=[0.7, 0.2, 0.1]).astype(np.int32)
    d2 = d1 ^ np.packbits((rng.random((n, 256)) < 0.05).astype(np.uint8), axis=1)
    a2 = np.mod(a1 + rng.normal(0, 4, n), 360).astype(np.float32)
    x2, y2 = (x1 + rng.normal(0, 3, n)).astype(np.float32), (y1 + rng.normal(0, 3, n)).astype(np.float32)
    o2 = o1.copy()
    # look-alike groups: several F1 features close together share (almost) one descriptor
    for g in range(40):
        idx = rng.permutation(n)[:4]
        x1[idx] = x1[idx[0]] + rng.uniform(-4, 4, 4); y1[idx] = y1[idx[0]] + rng.uniform(-4, 4, 4)
        x2[idx] = x1[idx] + 1; y2[idx] = y1[idx]
        o1[idx] = 0; o2[idx] = 0
        d1[idx] = d1[idx[0]] ^ np.packbits((rng.random((4, 256)) < 0.01).astype(np.uint8), axis=1)
        d2[idx] = d1[idx[0]] ^ np.packbits((rng.random((4, 256)) < 0.02).astype(np.uint8), axis=1)
    perm = rng.permutation(n)  # F2 in a different order
    F1 = dict(octave=o1, angle=a1, desc=d1)
    F2 = dict(x=x2[perm], y=y2[perm], octave=o2[perm], angle=a2[perm], desc=d2[perm])
    prev = np.stack([x1, y1], 1).astype(np.float32)
    return F1, F2, prev


def _dist_range(x, y, lvl, scale=1.0):
    """mfMaxDistance / mfMinDistance so that MapPoint::PredictScale gives lvl for the point (x, y, 1)*scale seen from
    the origin (margin half a level), as kf_scene does."""
    sf7 = np.float32(1.2) ** 7
    d = scale * np.sqrt(np.asarray(x, np.float64) ** 2 + np.asarray(y, np.float64) ** 2 + 1.0)
    max_dist = (d * 1.2 ** (np.asarray(lvl) - 0.5)).astype(np.float32)
    return max_dist, (max_dist / sf7).astype(np.float32)


def map_scene(seed, n=900, flags=False, stereo=False, noise=2.0):
    """A keyframe of n features and n candidate map points given as world points (u, v, +-1) seen from an identity
    pose, for the map-side matchers (SearchByProjection(KF,Scw), Fuse x2): distance ranges that put the predicted
    level on or one above the paired feature's octave, viewing normals (some turned away), points behind the camera,
    outside the image and outside their scale range."""
    F, mp = window_scene(seed, n, stereo=stereo)
    rng = np.random.Generator(np.random.PCG64(seed + 11))
    if noise != 2.0:  # re-draw the projection noise around the paired features
        mp["x"] = (mp["x"] + rng.normal(0, noise, n)).astype(np.float32)
        mp["y"] = (mp["y"] + rng.normal(0, noise, n)).astype(np.float32)
    wx, wy, wz = mp["x"].copy(), mp["y"].copy(), np.ones(n, np.float32)
    max_dist, min_dist = _dist_range(wx, wy, mp["level"])
    normal = np.stack([wx, wy, wz], 1).astype(np.float32)
    bad = np.zeros(n, np.uint8)
    if flags:
        wz[rng.random(n) < 0.05] = -1.0
        away = rng.random(n) < 0.08
        normal[away] *= -1.0
        far = rng.random(n) < 0.05
        max_dist[far] = (max_dist[far] * 0.2).astype(np.float32)
        bad = (rng.random(n) < 0.05).astype(np.uint8)
    KF = dict(x=F["x"], y=F["y"], octave=F["octave"], desc=F["desc"])
    if stereo:
        KF["uright"] = F["uright"]
    pts = dict(wx=wx, wy=wy, wz=wz, max_dist=max_dist, min_dist=min_dist, normal=normal, desc=mp["desc"], bad=bad)
    return KF, pts


def fuse_scene(seed, n=900, flags=True, stereo=False):
    """map_scene + the keyframe's own map points (slot_state / slot_obs) and candidate states for
    Fuse(KeyFrame*, vpMapPoints, th): NULL / bad / already-in-keyframe candidates, the same MapPoint* listed twice,
    several candidates landing on one feature."""
    KF, pts = map_scene(seed, n, flags=flags, stereo=stereo, noise=1.0)
    rng = np.random.Generator(np.random.PCG64(seed + 13))
    KF["slot_state"] = rng.choice(np.array([0, 1, 2], np.uint8), n, p=[0.5, 0.42, 0.08])
    KF["slot_obs"] = rng.integers(1, 6, n).astype(np.int32)
    state = np.where(pts["bad"] > 0, 1, 3).astype(np.uint8)
    if flags:
        state[rng.random(n) < 0.05] = 0
        state[rng.random(n) < 0.05] = 2
    qid = np.arange(n, dtype=np.int32)
    dup = rng.permutation(np.arange(n // 2, n))[:40]           # later entries repeating an earlier MapPoint*
    src = rng.integers(0, n // 2, 40)
    for d, s in zip(dup, src):
        qid[d] = s
        for k in ("wx", "wy", "wz", "max_dist", "min_dist", "normal", "desc"):
            pts[k][d] = pts[k][s]
        state[d] = state[s]
    # crowds: a few candidates copied next to another candidate so that they compete for one feature
    crowd = rng.permutation(n // 2)[:40]
    for c in crowd:
        t = (c + 1) % (n // 2)
        if qid[t] != t or qid[c] != c or np.any(qid[n // 2:] == t):
            continue
        for k in ("wx", "wy", "wz", "max_dist", "min_dist", "normal"):
            pts[k][t] = pts[k][c]
        pts["desc"][t] = pts["desc"][c] ^ np.packbits((rng.random(256) < 0.01).astype(np.uint8))
    pts["state"], pts["qid"], pts["obs"] = state, qid, rng.integers(1, 6, n).astype(np.int32)
    return KF, pts


def sim3_scene(seed, n=800, flags=True):
    """Two keyframes for SearchBySim3: KF2's features are KF1's moved by a few pixels (in another order) with noisy
    descriptors; the map point held by a feature projects next to the paired feature of the other keyframe, so that
    most pairs agree in both directions, some only in one, some on neither."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x1, y1 = rng.uniform(5, 635, n).astype(np.float32), rng.uniform(5, 475, n).astype(np.float32)
    o1 = rng.integers(0, 7, n).astype(np.int32)
    d1 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    perm = rng.permutation(n)               # KF2 feature j pairs with KF1 feature perm[j]
    x2 = (x1[perm] + rng.normal(0, 2, n)).astype(np.float32)
    y2 = (y1[perm] + rng.normal(0, 2, n)).astype(np.float32)
    o2 = o1[perm].copy()
    flip = lambda d, p: d ^ np.packbits((rng.random((len(d), 256)) < p).astype(np.uint8), axis=1)
    d2 = flip(d1[perm], 0.06)
    inv = np.argsort(perm)                  # KF1 feature i pairs with KF2 feature inv[i]
    def points(xo, yo, lvl_of_pair, dk):
        wx = (xo + rng.normal(0, 1.5, n)).astype(np.float32)
        wy = (yo + rng.normal(0, 1.5, n)).astype(np.float32)
        lvl = np.clip(lvl_of_pair + rng.integers(0, 2, n), 0, 7)
        mx, mn = _dist_range(wx, wy, lvl)
        return dict(wx=wx, wy=wy, wz=np.ones(n, np.float32), max_dist=mx, min_dist=mn, pdesc=flip(dk, 0.03))
    K1 = dict(x=x1, y=y1, octave=o1, desc=d1, **points(x2[inv], y2[inv], o2[inv], d1))
    K2 = dict(x=x2, y=y2, octave=o2, desc=d2, **points(x1[perm], y1[perm], o1[perm], d2))
    K1["state"] = np.full(n, 3, np.uint8)
    K2["state"] = np.full(n, 3, np.uint8)
    pre12 = np.full(n, -1, np.int32)
    if flags:
        K1["state"] = rng.choice(np.array([0, 1, 3], np.uint8), n, p=[0.1, 0.05, 0.85])
        K2["state"] = rng.choice(np.array([0, 1, 3], np.uint8), n, p=[0.1, 0.05, 0.85])
        pre = rng.permutation(n)[:60]
        ok = K2["state"][inv[pre]] > 0
        pre12[pre[ok]] = inv[pre[ok]]
        # one-sided pairs: move the KF2 point of some pairs far away
        lost = rng.permutation(n)[:80]
        K2["wx"][lost] = rng.uniform(5, 635, 80).astype(np.float32)
        K2["max_dist"][lost], K2["min_dist"][lost] = _dist_range(K2["wx"][lost], K2["wy"][lost], np.clip(o1[perm][lost], 0, 7))
    return K1, K2, pre12


def stereo_pair(seed=0, width=640, height=480, disparity=14, half_pixel=True, tex=None):
    """A rectified pair cut from one texture: the right image is the left one shifted by `disparity` px (a fronto-parallel
    plane), optionally blended with the next shift so that the SAD parabola lands between pixels, plus a little noise."""
    from eaof import synth
    rng = np.random.Generator(np.random.PCG64(seed))
    if tex is None:
        tex = synth.base_texture(width, height, seed=40 + seed)
    ox, oy = synth.frame_offset(3 * seed)
    left = tex[oy:oy + height, ox:ox + width].copy()
    a = tex[oy:oy + height, ox + disparity:ox + disparity + width].astype(np.int32)
    if half_pixel:
        b = tex[oy:oy + height, ox + disparity + 1:ox + disparity + 1 + width].astype(np.int32)
        a = (a + b + 1) >> 1
    right = np.clip(a + rng.integers(-2, 3, a.shape), 0, 255).astype(np.uint8)
    return left, right
