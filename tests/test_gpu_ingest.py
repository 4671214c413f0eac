"""Pinned frame ring (include/eaof_orb.h eaof_ring_*): what the camera callback
(ros_test/src/message_flow.cc:250-254) writes into the ring must come out of eaof_orb_extract_ring exactly as if the same
frames had been handed to the extractor directly — across the wrap-around of the slot array, with a producer thread running
beside the consumer, for gray and colour frames."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    return len(a) == len(b) and all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b))


def test_ring_batches_equal_direct_extraction_with_a_producer_thread(frames640):
    import eaof
    n = len(frames640)
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=3)
    direct = []
    for i in range(0, n, 3):
        direct += ex.extract_batch(frames640[i:i + 3])
    ring = eaof.FrameRing(4, 640, 480)
    dropped = []

    def camera():
        for i in range(n):
            while True:
                try:
                    ring.push(frames640[i], timestamp=100.0 + i)
                    break
                except eaof.RingFull:
                    dropped.append(i)  # the callback would drop the frame; the test waits instead
                    time.sleep(0.001)

    t = threading.Thread(target=camera)
    t.start()
    got, stamps, sizes = [], [], []
    deadline = time.time() + 60
    while len(got) < n and time.time() < deadline:
        res, ts = ring.extract(ex)
        if not res:
            time.sleep(0.0005)
            continue
        for k in range(len(res)):  # the frames are still in the ring: peek sees the pixels the callback wrote
            img, tk = ring.peek(k)
            assert tk == ts[k] and np.array_equal(img, frames640[len(got) + k])
        ring.release(len(res))
        got += res
        stamps += list(ts)
        sizes.append(len(res))
    t.join()
    assert len(got) == n and stamps == [100.0 + i for i in range(n)]
    assert max(sizes) <= 3 and ring.pending() == 0
    assert _same(got, direct)
    ring.close()
    ex.close()


def test_ring_full_wraparound_and_release(frames640):
    import eaof
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=4)
    ring = eaof.FrameRing(3, 640, 480)
    for i in range(3):
        ring.push(frames640[i], i)
    with pytest.raises(eaof.RingFull):
        ring.push(frames640[3], 3)
    assert ring.pending() == 3
    with pytest.raises(eaof.EaofError):
        ring.peek(3)
    with pytest.raises(eaof.EaofError):
        ring.release(4)
    res, ts = ring.extract(ex, max_frames=2)
    assert len(res) == 2 and list(ts) == [0.0, 1.0]
    ring.release(2)
    ring.push(frames640[3], 3)  # slot 0 again
    ring.push(frames640[4], 4)  # slot 1
    res2, ts2 = ring.extract(ex)  # frame 2 sits in the last slot: the run stops at the end of the slot array
    assert len(res2) == 1 and list(ts2) == [2.0]
    ring.release(1)
    res3, ts3 = ring.extract(ex)
    assert len(res3) == 2 and list(ts3) == [3.0, 4.0]
    ring.release(2)
    res4, _ = ring.extract(ex)
    assert res4 == [] and ring.pending() == 0
    assert _same(res + res2 + res3, ex.extract_batch(frames640[:4]) + ex.extract_batch(frames640[4:5]))
    # a ring of another frame size is refused
    small = eaof.FrameRing(2, 320, 240)
    small.push(np.zeros((240, 320), np.uint8))
    with pytest.raises(eaof.EaofError, match="extractor built for"):
        small.extract(ex)
    small.close()
    ring.close()
    ex.close()


@pytest.mark.parametrize("color", [0, 3])
def test_colour_ring(frames640, color):
    import eaof
    from test_gpu_extract import _color_frames
    col = _color_frames(frames640[:3], color)
    ex = eaof.ORBextractor(1000, 1.2, 8, 20, 7, width=640, height=480, max_batch=3)
    want = ex.extract_batch_color(col, color, eaof.GRAY_CV331)
    ring = eaof.FrameRing(4, 640, 480, channels=col.shape[-1])
    for f in range(3):
        ring.push(col[f], f)
    with pytest.raises(eaof.EaofError, match="colour order"):
        ring.extract(ex, color=2 if color == 0 else 0)
    res, ts = ring.extract(ex, color=color, gray_mode=eaof.GRAY_CV331)
    assert list(ts) == [0.0, 1.0, 2.0] and _same(res, want)
    ring.close()
    ex.close()
