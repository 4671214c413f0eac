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
