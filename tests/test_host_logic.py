"""Host-side logic that needs no GPU: C-ABI surface, sharding plans (incl. a world_size-2 gloo run), matcher oracle
against independent pure-Python restatements."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from matchdata import planted_pair, random_nodes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in ("eaof_orb.h", "eaof_match.h", "eaof_voc.h", "eaof_sweep.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(eaof_[a-z0-9_]+)\s*\(", src))
    return names


def test_cabi_library_exports_every_declared_symbol():
    import eaof
    if not os.path.exists(eaof.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(eaof.LIB_PATH)  # loads without a GPU; no compute calls here
    decl = _declared_symbols()
    assert len(decl) >= 25
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    lib.eaof_abi_version.restype = ctypes.c_int
    assert lib.eaof_abi_version() == 1


def test_product_fails_loudly_without_a_gpu():
    import eaof
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(eaof.EaofError, match="no CUDA device"):
        eaof.ORBextractor(width=640, height=480)
    with pytest.raises(eaof.EaofError, match="no CUDA device"):
        eaof.ORBmatcher(0.9)
    from eaof import sweep
    with pytest.raises(eaof.EaofError, match="no CUDA device"):
        sweep.Sweep(0, 1, 0)
    with pytest.raises(eaof.EaofError, match="no CUDA device"):
        eaof.FrameRing(4, 640, 480)


def test_frame_ring_argument_errors():
    """include/eaof_orb.h eaof_ring_*: shape and null-pointer errors come back as EAOF_ERR_ARG before any CUDA call."""
    import eaof
    L = eaof.lib()
    h = ctypes.c_void_p()
    for slots, w, hh, ch in ((0, 640, 480, 1), (4, 0, 480, 1), (4, 640, 480, 2), (4, 640, -1, 3)):
        assert L.eaof_ring_create(slots, w, hh, ch, ctypes.byref(h)) == -1
        assert b"bad ring shape" in L.eaof_last_error() and not h.value
    assert L.eaof_ring_create(4, 640, 480, 1, None) == -1
    assert L.eaof_ring_pending(None) == -1
    assert L.eaof_ring_release(None, 0) == -1
    assert L.eaof_ring_commit(None, 0.0) == -1
    n = ctypes.c_int(7)
    assert L.eaof_orb_extract_ring(None, None, 1, 0, 0, None, None, 0, None, None, ctypes.byref(n)) == -1
    L.eaof_ring_destroy(None)  # a no-op, like eaof_orb_destroy


def test_sweep_cabi_argument_errors_and_nccl_binding():
    """include/eaof_sweep.h without a GPU: argument errors come back as EAOF_ERR_ARG with a reason, NCCL is bound at run
    time (the id call works wherever a libnccl.so.2 exists and fails with EAOF_ERR_NCCL where it does not)."""
    import eaof
    L = eaof.lib()
    L.eaof_sweep_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    h = ctypes.c_void_p()
    assert L.eaof_sweep_create(None, 2, 2, 0, ctypes.byref(h)) == -1 and b"rank 2 outside a world of 2" in L.eaof_last_error()
    assert L.eaof_sweep_create(None, 0, 2, 0, ctypes.byref(h)) == -1 and b"unique id" in L.eaof_last_error()
    buf = (ctypes.c_uint8 * 128)()
    rc = L.eaof_sweep_unique_id(buf)
    assert rc in (0, -5), L.eaof_last_error()
    if rc == 0:
        assert any(buf)
        v = ctypes.c_int()
        assert L.eaof_sweep_nccl_version(ctypes.byref(v)) == 0 and v.value >= 20000
    else:
        assert b"libnccl" in L.eaof_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eao-fusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                bad = re.findall(r'#\s*include\s*[<"][^>"]*oracle|^\s*(?:from|import)\s+oracle|pyoracle|dlopen\([^)]*oracle', txt, flags=re.M)
                assert not bad, f"{f} pulls in oracle/: {bad}"


def test_shard_plans_cover_everything_once():
    from eaof import shard
    for n, world in ((1000, 1), (1000, 8), (10, 4), (7, 8), (100000, 8)):
        blocks = [shard.frame_block(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        pairs = [p for r in range(world) for p in shard.consecutive_pairs(*blocks[r], n)]
        assert pairs == [(f - 1, f) for f in range(1, n)]
        for r in range(world):
            hb = shard.halo_block(*blocks[r])
            assert all(hb[0] <= a and b < hb[1] for a, b in shard.consecutive_pairs(*blocks[r], n))
        sw = sorted(i for r in range(world) for i in shard.sweep_pairs(n, r, world))
        assert sw == list(range(n))


def test_sharding_world_size_2_gloo(tmp_path):
    """Two processes over gloo: each extracts its frame block with the CPU oracle (stand-in for the GPU on this box),
    the all-gathered keypoint counts must equal the single-process result -> frame sharding is deterministic."""
    script = tmp_path / "w2.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "eao-fusion_b200"))
import numpy as np, torch, torch.distributed as dist
from eaof import shard, synth
from oracle import pyoracle as po
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
tex = synth.base_texture(320, 240, seed=3)
frames = synth.make_frames(5, 320, 240, tex=tex)
b, e = shard.frame_block(len(frames), rank, world)
mine = torch.zeros(len(frames), dtype=torch.int64)
for f in range(b, e):
    k, d = po.o_extract(frames[f], nfeatures=300)
    mine[f] = len(k) * 1000003 + int(d.sum()) % 1000003
dist.all_reduce(mine)
if rank == 0:
    full = [len(k) * 1000003 + int(d.sum()) % 1000003 for k, d in (po.o_extract(fr, nfeatures=300) for fr in frames)]
    assert mine.tolist() == full, (mine.tolist(), full)
    print("W2 OK")
dist.destroy_process_group()
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert "W2 OK" in out.stdout, out.stdout + out.stderr


def test_sweep_block_gather_world_size_2_gloo(tmp_path):
    """The one exchange step of the path (all-gather of descriptor blocks for the cross-frame sweep) over gloo with CPU
    tensors: gathered blocks land at global_block(frame), padding blocks are empty, and the round-robin pair shares of
    the two ranks cover the global pair list exactly once."""
    script = tmp_path / "sw2.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "eao-fusion_b200"))
import numpy as np, torch, torch.distributed as dist
from eaof import shard, sweep
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_frames, stride = 7, 5                      # ragged: rank 0 owns 4 frames, rank 1 owns 3 (+1 padding block)
b, e = shard.frame_block(n_frames, rank, world)
def block(f):
    rng = np.random.Generator(np.random.PCG64(f))
    return rng.integers(0, 256, (stride, 32), dtype=np.uint8), rng.uniform(0, 360, stride).astype(np.float32), 1 + f % stride
loc = [block(f) for f in range(b, e)]
desc = torch.from_numpy(np.stack([x[0] for x in loc])); ang = torch.from_numpy(np.stack([x[1] for x in loc]))
cnt = torch.tensor([x[2] for x in loc], dtype=torch.int32)
gd, ga, gc = sweep.gather_blocks(desc, ang, cnt, n_frames, dist)
per = sweep.padded_blocks(n_frames, world)
assert gd.shape == (world * per, stride, 32) and gc.shape == (world * per,)
for f in range(n_frames):
    g = sweep.global_block(f, n_frames, world)
    d, a, c = block(f)
    assert np.array_equal(gd[g].numpy(), d) and np.array_equal(ga[g].numpy(), a) and int(gc[g]) == c, f
assert int(gc[world * per - 1]) == 0          # padding block of the last rank
pairs = np.array([(i, j) for i in range(n_frames) for j in range(n_frames) if i != j])
sel, pq, pt = sweep.my_pairs(pairs, n_frames, rank, world)
assert all(pq[k] == sweep.global_block(pairs[i, 0], n_frames, world) for k, i in enumerate(sel))
mine = torch.zeros(len(pairs), dtype=torch.int64); mine[torch.from_numpy(sel)] = 1
dist.all_reduce(mine)
assert mine.tolist() == [1] * len(pairs)
if rank == 0: print("SW2 OK")
dist.destroy_process_group()
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert "SW2 OK" in out.stdout, out.stdout + out.stderr


# ---------------------------------------------------------------------------------------------------------------
# matcher oracle vs. independent pure-Python restatements of the same reference loops

def _ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


def _three_maxima(sizes):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if m2 < np.float32(0.1) * np.float32(m1):
        i2 = i3 = -1
    elif m3 < np.float32(0.1) * np.float32(m1):
        i3 = -1
    return i1, i2, i3


def _py_bow(mode, ratio, ori, q, aq, vq, nodes_q, t, at, vt, nodes_t):
    (iq, sq, xq), (it, st, xt) = nodes_q, nodes_t
    nout = len(t) if mode == 0 else len(q)
    match = np.full(nout, -1, np.int32)
    dist = np.full(nout, -1, np.int32)
    matched = np.zeros(len(t), bool)
    hist = [[] for _ in range(30)]
    n = 0
    tmap = {int(k): j for j, k in enumerate(it)}
    for a, nid in enumerate(iq):
        if int(nid) not in tmap:
            continue
        b = tmap[int(nid)]
        for qi in xq[sq[a]:sq[a + 1]]:
            if vq is not None and not vq[qi]:
                continue
            b1, b2, bi = 256, 256, -1
            for ti in xt[st[b]:st[b + 1]]:
                if matched[ti] or (mode == 1 and vt is not None and not vt[ti]):
                    continue
                d = _ham(q[qi], t[ti])
                if d < b1:
                    b2, b1, bi = b1, d, ti
                elif d < b2:
                    b2 = d
            ok = b1 <= 50 if mode == 0 else b1 < 50
            if ok and np.float32(b1) < np.float32(ratio) * np.float32(b2):
                matched[bi] = True
                o = bi if mode == 0 else qi
                match[o] = qi if mode == 0 else bi
                dist[o] = b1
                if ori:
                    rot = np.float32(aq[qi]) - np.float32(at[bi])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360))
                    v = np.float32(rot * np.float32(1.0 / 30))
                    bn = int(np.floor(abs(v) + np.float32(0.5)) * np.sign(v))  # C round(): half away from zero
                    hist[0 if bn == 30 else bn].append(o)
                n += 1
    if ori:
        keep = _three_maxima([len(h) for h in hist])
        for i, h in enumerate(hist):
            if i not in keep:
                for o in h:
                    match[o] = -1
                    dist[o] = -1
                    n -= 1
    return n, match, dist


def test_matcher_oracle_primitives():
    from oracle import pyoracle as po
    cv2 = pytest.importorskip("cv2")
    rng = np.random.Generator(np.random.PCG64(2))
    a = rng.integers(0, 256, size=(300, 32), dtype=np.uint8)
    b = rng.integers(0, 256, size=(300, 32), dtype=np.uint8)
    a[:3] = 0
    b[:3] = 255
    d = po.o_hamming(a, b)
    assert np.array_equal(d, np.unpackbits(a ^ b, axis=1).sum(1))
    assert all(d[i] == int(cv2.norm(a[i], b[i], cv2.NORM_HAMMING)) for i in range(300))
    assert d[0] == 256
    for _ in range(200):
        sizes = rng.integers(0, 12, 30) * (rng.random(30) < 0.5)
        assert po.o_three_maxima(sizes) == _three_maxima([int(s) for s in sizes])
    assert po.o_three_maxima([0] * 30) == (-1, -1, -1)


@pytest.mark.parametrize("mode", [0, 1])
def test_matcher_oracle_bow_vs_python(mode):
    import eaof
    from oracle import pyoracle as po
    for seed, (nq, nt, ratio, ori) in enumerate([(120, 150, 0.9, True), (80, 80, 0.6, True), (60, 200, 0.75, False),
                                                 (1, 5, 0.9, True), (40, 40, 0.9, True)]):
        q, aq, t, at = planted_pair(nq, nt, 70 + seed, dup=3 if nt > 30 else 0)
        if seed == 4:  # heavy ties
            t[:] = t[0]
            q[:] = t[0]
        nodes_q = eaof.csr_from_nodes(random_nodes(nq, 3, seed) if seed % 2 else np.zeros(nq, int))
        nodes_t = eaof.csr_from_nodes(random_nodes(nt, 3, seed + 9) if seed % 2 else np.zeros(nt, int))
        rng = np.random.Generator(np.random.PCG64(seed))
        vq = (rng.random(nq) > 0.1).astype(np.uint8)
        vt = (rng.random(nt) > 0.1).astype(np.uint8)
        got = po.o_search_by_bow(mode, ratio, ori, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        exp = _py_bow(mode, ratio, ori, q, aq, vq, nodes_q, t, at, vt, nodes_t)
        assert got[0] == exp[0] and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])


def test_matcher_oracle_grid_vs_python():
    from oracle import pyoracle as po
    rng = np.random.Generator(np.random.PCG64(4))
    x = rng.uniform(-5, 645, 800).astype(np.float32)
    y = rng.uniform(-5, 485, 800).astype(np.float32)
    iw, ih = np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)
    cs, ci = po.o_build_grid(x, y, 0.0, 0.0, iw, ih)
    cells = [[] for _ in range(64 * 48)]
    for i in range(800):
        vx, vy = np.float32(x[i] * iw), np.float32(y[i] * ih)
        px = int(np.floor(abs(vx) + np.float32(0.5)) * np.sign(vx))
        py = int(np.floor(abs(vy) + np.float32(0.5)) * np.sign(vy))
        if 0 <= px < 64 and 0 <= py < 48:
            cells[px * 48 + py].append(i)
    flat = [i for c in cells for i in c]
    assert list(ci) == flat
    assert list(cs[:-1]) == list(np.cumsum([0] + [len(c) for c in cells])[:-1])


def test_dropin_class_loads_and_refuses_to_run_without_cuda():
    """The drop-in ORB_SLAM2::ORBextractor answers its getters before any frame (host restatement of
    src/ORBextractor.cc:415-433) and throws — never falls back to a CPU path — when no CUDA device exists."""
    import numpy as np
    import torch
    from dropin import DropinExtractor
    from oracle import pyoracle as po
    d = DropinExtractor(1000, 1.2, 8, 20, 7)
    t, r = d.tables(), po.o_tables(1000, 1.2, 8)
    assert np.array_equal(t["scale"], r["scale"]) and np.array_equal(t["inv_sigma2"], r["inv_sigma2"])
    n, rows, _, _ = d.extract(None, pre_n=3)  # empty image: silent return, outputs untouched (:1046-1047)
    assert (n, rows) == (3, 3)
    if not torch.cuda.is_available():
        import ctypes as C
        img = np.zeros((240, 320), np.uint8)
        rc = d.L.dropin_extract(d.h, img.ctypes.data, 320, 240, 320, None, None, 0, 0, None)
        assert rc == -2
    d.close()


def test_matcher_dropin_loads_and_refuses_to_run_without_cuda():
    """tests/cpp/_build/libmatch_dropin.so = the drop-in ORBmatcher.cc behind the reference harness: on a box without a
    GPU every search must throw (no CPU fallback) — the harness lets the C++ exception terminate a child process."""
    so = os.path.join(ROOT, "tests", "cpp", "_build", "libmatch_dropin.so")
    if not os.path.exists(so):
        pytest.skip("drop-in matcher harness not built (needs /root/reference headers)")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = f"""
import sys
sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'eao-fusion_b200')!r}, {os.path.join(ROOT, 'tests')!r}]
import numpy as np
from oracle import pyoracle as po
from matchdata import planted_pair
import eaof
D = po.match_harness_lib({so!r})
assert D.mref_th_low() == 50
q, aq, t, at = planted_pair(50, 60, 1)
n0 = eaof.csr_from_nodes(np.zeros(50, int)); n1 = eaof.csr_from_nodes(np.zeros(60, int))
po.r_search_by_bow(0, 0.9, True, q, aq, None, n0, t, at, None, n1, L=D)
print("RETURNED")
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert "RETURNED" not in out.stdout and out.returncode != 0
    assert "no CUDA device" in out.stderr or "eaof_matcher_create" in out.stderr, out.stderr[-2000:]


def test_workload_sequences_shard_exactly():
    """eaof/workload.py: global frame t is the same whichever rank builds it (scene t // 1000 with its own texture), scene 0
    is the configs[1] sequence, and a rank's block + halo is a plain slice of the whole."""
    from eaof import shard, synth, workload
    seq = workload.Sequence(96, 80)
    whole = seq.frames(0, 2 * workload.SEQ_LEN)
    assert np.array_equal(whole[:7], synth.make_frames(7, 96, 80, tex=synth.base_texture(96, 80, seed=1235)))
    assert not np.array_equal(whole[0], whole[workload.SEQ_LEN])  # a new scene
    for world in (2, 3):
        for r in range(world):
            b, e = shard.frame_block(2 * workload.SEQ_LEN, r, world)
            hb, he = shard.halo_block(b, e)
            assert np.array_equal(workload.Sequence(96, 80).frames(hb, he), whole[hb:he])
    d = [workload.frame_digest(np.zeros(3, "<f4"), np.zeros((3, 32), np.uint8)), workload.pair_digest(np.arange(4))]
    assert workload.combine(d) == workload.combine(list(d)) and len(workload.combine(d)) == 64


def test_dropin_selects_the_blur_arithmetic_of_its_opencv():
    """dropin/ORBextractor.cc probes the cv::GaussianBlur it is compiled against (here: the shim, switchable between the three
    known arithmetics) and must select the matching EAOF_BLUR_* mode at construction — no GPU involved."""
    import time
    import dropin
    if not os.path.exists(dropin.SO):
        import __graft_entry__ as g
        g.build()
    L = dropin.lib()
    L.dropin_blur_mode.argtypes = [ctypes.c_void_p]
    old = os.environ.pop("EAOF_BLUR_MODE", None)
    try:
        for mode in (0, 1, 2):
            L.dropin_shim_set_blur_mode(mode)
            t0 = time.perf_counter()
            h = L.dropin_create(1000, 1.2, 8, 20, 7)
            dt = time.perf_counter() - t0
            assert h and L.dropin_blur_mode(h) == mode
            assert dt < 2.0, f"the probe took {dt:.2f} s"
            L.dropin_destroy(h)
        os.environ["EAOF_BLUR_MODE"] = "1"   # the override wins over the probe
        L.dropin_shim_set_blur_mode(0)
        h = L.dropin_create(1000, 1.2, 8, 20, 7)
        assert L.dropin_blur_mode(h) == 1
        L.dropin_destroy(h)
    finally:
        os.environ.pop("EAOF_BLUR_MODE", None)
        if old is not None:
            os.environ["EAOF_BLUR_MODE"] = old
        L.dropin_shim_set_blur_mode(0)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): stdout is ONE JSON line with the contract's keys —
    anything a library prints on descriptor 1 after argument parsing lands on stderr (bench.guard_stdout)."""
    import json
    code = ("import os, sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; "
            "import bench; bench.guard_stdout(); os.write(1, b'noise from a library\\n'); bench.main()")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:500]
    assert "noise from a library" in r.stderr
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
