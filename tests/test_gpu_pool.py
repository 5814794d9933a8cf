"""StreamPool (bgsb_pool_*): N camera streams over GPU groups with per-GPU worker threads and pinned rings, against
per-stream oracle chains.  Two groups on ONE device exercise the multi-group logic on a single-GPU box; a real second
device is used when there is one."""
import numpy as np
import pytest

from tracking_b200 import capi

pytestmark = pytest.mark.gpu


def oracle_chain(oracle, fg, chain):
    clean = fg
    for op, it in chain:
        clean = oracle.morph(clean, op, it)
    return (clean,) + tuple(oracle.ccl8(clean, True))


def run_pool(oracle, clips, algo, devices, nstreams, ring, nframes=18, want_masks=True):
    from tracking_b200.streams import StreamPool
    clip = clips["video_clip"]
    h, w = clip.shape[1:3]
    chain = (("erode", 1), ("dilate", 1))
    pool = StreamPool(algo, nstreams, w, h, devices=devices, ring=ring, morph=chain)
    os_ = [oracle.ALGOS[algo]() for _ in range(nstreams)]
    frame_of = lambda s, t: clip[(t * (1 + s % 3) + 5 * s) % len(clip)]          # a different walk per stream

    def check(t):
        slot = t % ring
        valid = pool.wait(slot)
        for s in range(nstreams):
            ofg, _ = os_[s].process(frame_of(s, t))
            assert valid == (ofg is not None), (s, t)
            if ofg is None:
                continue
            eclean, en, elab, est, eext = oracle_chain(oracle, ofg, chain)
            if want_masks:
                assert np.array_equal(pool.mask(s, slot), eclean), (s, t)
            else:
                assert pool.mask(s, slot) is None
            comps = pool.components(s, slot)
            assert len(comps) == en, (s, t)
            for c, st_, e in zip(comps, est, eext):
                assert (c["x"], c["y"], c["x"] + c["w"] - 1, c["y"] + c["h"] - 1, c["area"], c["first_index"]) == \
                    tuple(int(v) for v in st_)
                assert c["external"] == int(e)

    # keep ring - 1 frame sets in flight: the upload of set t overlaps the work on set t-1
    for t in range(nframes):
        if t >= ring:
            check(t - ring)
        slot = t % ring
        for s in range(nstreams):
            pool.frame_buffer(s, slot)[...] = frame_of(s, t)
        pool.submit(slot, want_masks)
    for t in range(max(0, nframes - ring), nframes):
        check(t)
    assert sorted({pool.device_of(s) for s in range(nstreams)}) == sorted(set(devices[:nstreams]))
    pool.close()


@pytest.mark.parametrize("algo", [capi.ALGO_MOG2, capi.ALGO_FRAME_DIFFERENCE, capi.ALGO_WEIGHTED_MOVING_VARIANCE])
@pytest.mark.parametrize("ring", [1, 3])
def test_pool_one_gpu_group(oracle, clips, algo, ring):
    run_pool(oracle, clips, algo, [0], nstreams=5, ring=ring)


def test_pool_two_groups_on_one_device(oracle, clips):
    """devices = [0, 0]: two worker threads, two stream groups (3 + 2 streams), one GPU."""
    run_pool(oracle, clips, capi.ALGO_MOG2, [0, 0], nstreams=5, ring=2)
    run_pool(oracle, clips, capi.ALGO_MOG2, [0, 0, 0, 0], nstreams=3, ring=2, want_masks=False)     # more GPUs than streams


def test_pool_two_devices(oracle, clips):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box (two groups on one device are covered above)")
    run_pool(oracle, clips, capi.ALGO_MOG2, [0, 1], nstreams=6, ring=3)


def test_pool_argument_errors():
    from tracking_b200.streams import StreamPool
    with pytest.raises(capi.BgsbError):
        StreamPool(capi.ALGO_MOG2, 0, 64, 64, devices=[0])
    with pytest.raises(capi.BgsbError):
        StreamPool(capi.ALGO_MOG2, 2, 64, 64, devices=[0], ring=99)
    with pytest.raises(capi.BgsbError):
        StreamPool(4, 2, 64, 64, devices=[0])                   # not a hot-path plugin id
    p = StreamPool(capi.ALGO_MOG2, 2, 64, 48, devices=[0], ring=2)
    with pytest.raises(capi.BgsbError):
        p.wait(1)                                               # nothing submitted for that slot
    with pytest.raises(IndexError):
        p.frame_buffer(2, 0)
    p.close()
