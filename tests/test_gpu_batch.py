"""GPU suite: full-size batches (BASELINE.json configs[2] shape) checked through size-independent properties, plus
the error behaviour of the batch entry points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_1080p_batch_round_trip_and_position_independence(orc):
    """64 distinct 1080p frames in one device-resident batch:
       * every JPEG entropy-decodes (independent baseline decoder in the oracle) to exactly the quantised levels the
         FDCT kernel produced -> encode -> decode round trip through Huffman tables, scan, stuffing, header;
       * the same picture at batch positions 0, 31 and 63 gives identical bytes (frames never interact);
       * a sample of frames is compared with the oracle in full."""
    import torch

    import h2j_b200

    w, h, n = 1920, 1080, 64
    base = [orc.pack_i420(*orc.synth_planes(w, h, "textured", seed=100 + s, amp=15 + 9 * s)) for s in range(6)]
    order = [0, 1, 2, 3, 4, 5] * 10 + [1, 2, 3, 0]
    order[31] = 0
    order[63] = 0
    frames = np.stack([base[i] for i in order])
    d = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    nblk = ((w + 15) // 16) * ((h + 15) // 16) * 6
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=n, n_slots=1) as e:
        e.submit_device(0, d.data_ptr(), frames.shape[1], n, w, h)
        res = e.collect(0)
        assert res.status == [0] * n
        assert res.jpegs[0] == res.jpegs[31] == res.jpegs[63]
        by_src = {}
        for i, src in enumerate(order):
            by_src.setdefault(src, res.jpegs[i])
            assert res.jpegs[i] == by_src[src], f"frame {i} differs from another copy of picture {src}"
        for i in (0, 7, 62):
            levels, info = orc.decode_coefs(res.jpegs[i])
            got = e.coefficients(0, i, nblk)
            assert (levels == got).all()
            fi = e.frame_info(0, i)
            assert info[2] == fi.stuffed_ff and (fi.scan_bits + 7) // 8 + fi.stuffed_ff == info[3]
    for src in (0, 3, 5):
        y, u, v = h2j_b200.split_planes(base[src], w, h)
        want, _, _ = orc.oracle_encode(np.ascontiguousarray(y), np.ascontiguousarray(u), np.ascontiguousarray(v))
        assert by_src[src] == want


def test_slots_overlap_and_reuse(orc):
    import h2j_b200

    w, h = 320, 192
    sets = [[orc.synth_planes(w, h, "textured", seed=10 * k + s, amp=20 + 5 * s) for s in range(3)] for k in range(4)]
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=3, n_slots=2) as e:
        got = {}
        for k in range(4):  # two slots, four batches: every slot is reused while the other is in flight
            slot = k % 2
            if k >= 2:
                got[k - 2] = e.collect(slot)
            frames = np.stack([orc.pack_i420(*p) for p in sets[k]])
            e.submit_host(slot, frames.ctypes.data, frames.shape[1], 3, w, h)
            e.wait(slot)  # frames is a temporary: the copy must be done before it goes away
        got[2] = e.collect(0)
        got[3] = e.collect(1)
    for k in range(4):
        for (y, u, v), j in zip(sets[k], got[k].jpegs):
            want, _, _ = orc.oracle_encode(y, u, v)
            assert j == want


def test_error_behaviour(orc):
    import h2j_b200

    y, u, v = orc.synth_planes(64, 64, "noise", seed=1, amp=100)
    frames = orc.pack_i420(y, u, v)[None, :].copy()
    with h2j_b200.Encoder(max_width=64, max_height=64, max_batch=1, n_slots=1, max_jpeg_bytes=1024) as e:
        # geometry outside the configured maximum
        with pytest.raises(h2j_b200.H2JError) as ei:
            e.submit_host(0, frames.ctypes.data, frames.shape[1], 1, 128, 64)
        assert ei.value.status == h2j_b200.ERR_INVALID_ARG or ei.value.status == h2j_b200.ERR_UNSUPPORTED
        # batch larger than configured
        with pytest.raises(h2j_b200.H2JError) as ei:
            e.submit_host(0, frames.ctypes.data, frames.shape[1], 2, 64, 64)
        assert ei.value.status == h2j_b200.ERR_INVALID_ARG
        # nothing to collect
        with pytest.raises(h2j_b200.H2JError) as ei:
            e.collect(0)
        assert ei.value.status == h2j_b200.ERR_BUSY
        # a JPEG that does not fit max_jpeg_bytes is reported, not truncated silently
        e.submit_host(0, frames.ctypes.data, frames.shape[1], 1, 64, 64)
        with pytest.raises(h2j_b200.H2JError) as ei:
            e.submit_host(0, frames.ctypes.data, frames.shape[1], 1, 64, 64)  # slot busy
        assert ei.value.status == h2j_b200.ERR_BUSY
        with pytest.raises(h2j_b200.H2JError) as ei:
            e.collect(0)
        assert ei.value.status == h2j_b200.ERR_OUTPUT_TOO_SMALL
        # the encoder stays usable afterwards
    with h2j_b200.Encoder(max_width=64, max_height=64, max_batch=1, n_slots=1) as e:
        assert e.yuv2jpeg(y, u, v) == orc.oracle_encode(y, u, v)[0]


def test_collect_with_a_short_buffer_can_be_repeated(orc):
    """h2j_collect with a caller buffer shorter than the batch: H2J_ERR_BUFFER_TOO_SMALL, the sizes are reported, nothing is
    lost -- the same slot collects into a buffer of the reported size -- and a frame that outgrew max_jpeg_bytes comes back as
    a zero-length entry with its status while its neighbours are delivered (strict=False)."""
    import h2j_b200

    w, h = 320, 192
    planes = [orc.synth_planes(w, h, "textured", seed=70 + s, amp=20 + 10 * s) for s in range(3)]
    frames = np.stack([orc.pack_i420(*p) for p in planes])
    want = [orc.oracle_encode(*p)[0] for p in planes]
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=3, n_slots=1) as e:
        e.submit_host(0, frames.ctypes.data, frames.shape[1], 3, w, h)
        small = np.empty(1000, np.uint8)
        with pytest.raises(h2j_b200.H2JError) as ei:
            e.collect_into(0, small.ctypes.data, small.size)
        assert ei.value.status == h2j_b200.ERR_BUFFER_TOO_SMALL
        sizes = np.diff(ei.value.offsets)
        assert [int(x) for x in sizes] == [len(j) for j in want]
        big = np.empty(int(ei.value.offsets[-1]), np.uint8)
        offs, st = e.collect_into(0, big.ctypes.data, big.size)
        assert list(st) == [0, 0, 0]
        for i in range(3):
            assert big[int(offs[i]): int(offs[i + 1])].tobytes() == want[i]
        with pytest.raises(h2j_b200.H2JError) as ei:  # now the slot is free
            e.collect_into(0, big.ctypes.data, big.size)
        assert ei.value.status == h2j_b200.ERR_BUSY
    # one frame of three does not fit max_jpeg_bytes
    noisy = orc.synth_planes(w, h, "noise", seed=5, amp=120)
    frames2 = np.stack([orc.pack_i420(*planes[0]), orc.pack_i420(*noisy), orc.pack_i420(*planes[2])])
    cap = max(len(want[0]), len(want[2])) + 64
    assert len(orc.oracle_encode(*noisy)[0]) > cap
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=3, n_slots=1, max_jpeg_bytes=cap) as e:
        e.submit_host(0, frames2.ctypes.data, frames2.shape[1], 3, w, h)
        res = e.collect(0, strict=False)
        assert res.status[0] == 0 and res.status[2] == 0 and res.status[1] == h2j_b200.ERR_OUTPUT_TOO_SMALL
        assert res.jpegs[0] == want[0] and res.jpegs[2] == want[2] and res.jpegs[1] == b""
