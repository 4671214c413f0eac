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
