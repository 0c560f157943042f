"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs.  Bar: identical quantised coefficients, identical tables, identical JPEG bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # (w, h, kind, amp, seed)
    (16, 16, "textured", 40, 1),
    (2, 2, "noise", 100, 2),
    (17, 17, "noise", 60, 3),
    (33, 47, "blocks", 30, 4),
    (64, 64, "binary", 0, 5),
    (131, 77, "textured", 80, 6),
    (322, 242, "textured", 40, 7),
    (641, 479, "noise", 20, 8),
    (1280, 720, "textured", 30, 9),
    (1920, 1080, "textured", 40, 10),
    (1918, 1078, "textured", 40, 11),
    (1920, 1080, "const", 0, 12),
    (1920, 1080, "ff", 0, 13),
]


@pytest.fixture(scope="module")
def enc():
    import h2j_b200

    e = h2j_b200.Encoder(max_width=1920, max_height=1088, max_batch=4, n_slots=2, max_jpeg_bytes=6 * 1024 * 1024)
    yield e
    e.close()


def _compare(enc, orc, y, u, v, **kw):
    h, w = y.shape
    frames = orc.pack_i420(y, u, v)[None, :].copy()
    enc.submit_host(0, frames.ctypes.data, frames.shape[1], 1, w, h)
    res = enc.collect(0)
    want, dbg, coefs = orc.oracle_encode(y, u, v, want_coefs=True, **kw)
    info = enc.frame_info(0, 0)
    assert info.mb_var_sum == dbg.mb_var_sum
    assert info.qscale == dbg.qscale
    assert bytes(info.intra_matrix) == bytes(dbg.intra_matrix)
    got_coefs = enc.coefficients(0, 0, coefs.shape[0])
    bad = np.nonzero((got_coefs != coefs).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} blocks differ, first {bad[:5]}: {got_coefs[bad[0]]} vs {coefs[bad[0]]}"
    for t in range(4):
        assert list(info.hist[t]) == list(dbg.hist[t]), f"histogram {t}"
        assert info.nvals[t] == dbg.nvals[t]
        assert bytes(info.bits[t]) == bytes(dbg.bits[t]), f"BITS {t}"
        assert bytes(info.vals[t])[: info.nvals[t]] == bytes(dbg.vals[t])[: dbg.nvals[t]], f"HUFFVAL {t}"
    assert info.header_bytes == dbg.header_bytes
    assert info.scan_bits == dbg.scan_bits
    assert res.status == [0]
    got = res.jpegs[0]
    assert len(got) == len(want)
    assert got == want
    return got


@pytest.mark.parametrize("w,h,kind,amp,seed", CASES)
def test_frame_matches_oracle(enc, orc, w, h, kind, amp, seed):
    y, u, v = orc.synth_planes(w, h, kind, seed=seed, amp=amp)
    _compare(enc, orc, y, u, v)


def test_single_frame_entry_point_with_strides(enc, orc):
    y, u, v = orc.synth_planes(322, 242, "textured", seed=21)
    # AVFrame-like padded linesizes
    yp = np.zeros((242, 384), np.uint8); yp[:, :322] = y
    up = np.zeros((121, 192), np.uint8); up[:, :161] = u
    vp = np.zeros((121, 192), np.uint8); vp[:, :161] = v
    got = enc.yuv2jpeg(yp[:, :322], up[:, :161], vp[:, :161])
    want, _, _ = orc.oracle_encode(y, u, v)
    assert got == want


def test_batch_of_different_frames(enc, orc):
    w, h = 640, 368
    planes = [orc.synth_planes(w, h, k, seed=s, amp=a) for k, s, a in [("textured", 1, 20), ("noise", 2, 90), ("const", 3, 0), ("blocks", 4, 10)]]
    frames = np.stack([orc.pack_i420(*p) for p in planes])
    res = enc.encode_batch(frames, w, h, slot=1)
    assert res.status == [0, 0, 0, 0]
    for (y, u, v), got in zip(planes, res.jpegs):
        want, _, _ = orc.oracle_encode(y, u, v)
        assert got == want


def test_dense_tiles_take_the_windowed_path(orc):
    """Noise at qscale 1 needs far more than 8 Kibit per 32-block unit, so K4 runs its windowed emission."""
    import h2j_b200

    w, h = 272, 208
    y, u, v = orc.synth_planes(w, h, "noise", seed=31, amp=127)
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=2, n_slots=1, fixed_qscale=1, max_jpeg_bytes=8 * 1024 * 1024) as e:
        got = _compare(e, orc, y, u, v, fixed_qscale=1)
        info = e.frame_info(0, 0)
    n_tiles = -(-info.mcu_w * info.mcu_h * 6 // 192)
    assert info.scan_bits / n_tiles > 6 * 8192, "the case no longer exercises the windowed path (8 Kibit per 32-block unit)"
    assert len(got) > 0


def test_fixed_qscale_range(orc):
    import h2j_b200

    y, u, v = orc.synth_planes(160, 96, "textured", seed=5, amp=60)
    for q in (1, 3, 31):
        with h2j_b200.Encoder(max_width=160, max_height=96, max_batch=1, n_slots=1, fixed_qscale=q) as e:
            _compare(e, orc, y, u, v, fixed_qscale=q)


def test_blocks_larger_than_their_slot_are_emitted_directly(orc):
    """K4 encodes every block into a private 256-bit slot; a block that needs more is emitted straight into the
    warp's window instead.  Flat frame with scattered noisy blocks at qscale 1: those blocks need ~600 bits while
    their 32-block unit stays far below the 8 Kibit window, so both merge paths run side by side in one warp."""
    import h2j_b200

    w, h = 256, 128
    rng = np.random.default_rng(77)
    y = np.full((h, w), 120, np.uint8)
    u = np.full((h // 2, w // 2), 128, np.uint8)
    v = np.full((h // 2, w // 2), 128, np.uint8)
    for (by, bx) in [(0, 0), (1, 5), (3, 7), (8, 30), (15, 31), (9, 9), (9, 10)]:
        y[by * 8: by * 8 + 8, bx * 8: bx * 8 + 8] = rng.integers(0, 256, (8, 8))
    u[8:16, 40:48] = rng.integers(0, 256, (8, 8))
    v[56:64, 120:128] = rng.integers(0, 256, (8, 8))
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1, fixed_qscale=1) as e:
        got = _compare(e, orc, y, u, v, fixed_qscale=1)
        info = e.frame_info(0, 0)
    levels, _ = orc.decode_coefs(got)
    nbits_max = 0
    assert info.scan_bits < 8192 * (info.mcu_w * info.mcu_h * 6 // 32), "units must stay below the window"
    assert (np.count_nonzero(levels, axis=1) > 40).sum() >= 7, "expected blocks with far more than 256 bits"


@pytest.mark.parametrize("w,h", [(65500, 16), (16, 65500), (4096, 18), (7680, 4320), (2, 4098)])
def test_extreme_geometries(orc, w, h):
    """The JPEG limit (65500), thin strips in both directions (tiles wrap an MCU row every MCU / never), and an 8K frame
    (129,600 MCUs: grid, prefix and staging arithmetic well past the 1080p sizes)."""
    import h2j_b200

    y, u, v = orc.synth_planes(w, h, "textured", seed=w + h, amp=45)
    cap = max(2 * 1024 * 1024, w * h)
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1, max_jpeg_bytes=cap) as e:
        got = e.yuv2jpeg(y, u, v)
        info = e.frame_info(0, 0)
    want, dbg, _ = orc.oracle_encode(y, u, v)
    assert info.qscale == dbg.qscale and info.scan_bits == dbg.scan_bits
    assert got == want


def test_randomised_geometries_and_contents(orc):
    """Seeded sweep: 48 random sizes (tiles that wrap MCU rows at every phase, partial last tiles/units, single-MCU
    frames), random content and qscale, three frames per submit.  Byte-identical JPEGs and identical histograms."""
    import h2j_b200

    rng = np.random.default_rng(20260)
    kinds = ["textured", "noise", "blocks", "binary", "const", "ff"]
    with h2j_b200.Encoder(max_width=700, max_height=520, max_batch=3, n_slots=2, max_jpeg_bytes=4 << 20) as e_auto:
        encs = {0: e_auto}
        for case in range(48):
            w = int(rng.integers(2, 700)); h = int(rng.integers(2, 520))
            if case % 6 == 0:
                w = int(rng.choice([16, 17, 31, 32, 33, 255, 256, 257, 272])); h = int(rng.choice([16, 17, 31, 33, 48]))
            fq = int(rng.choice([0, 0, 0, 1, 2, 7, 31]))
            if fq not in encs:
                encs[fq] = h2j_b200.Encoder(max_width=700, max_height=520, max_batch=3, n_slots=2, fixed_qscale=fq, max_jpeg_bytes=4 << 20)
            e = encs[fq]
            planes = [orc.synth_planes(w, h, str(rng.choice(kinds)), seed=int(rng.integers(1 << 30)), amp=int(rng.integers(0, 128))) for _ in range(3)]
            frames = np.stack([orc.pack_i420(*p) for p in planes])
            slot = case & 1
            res = e.encode_batch(frames, w, h, slot=slot)
            assert res.status == [0, 0, 0], (case, w, h, fq)
            for i, (y, u, v) in enumerate(planes):
                want, dbg, _ = orc.oracle_encode(y, u, v, fixed_qscale=fq)
                info = e.frame_info(slot, i)
                assert info.qscale == dbg.qscale, (case, w, h, fq, i)
                for t in range(4):
                    assert list(info.hist[t]) == list(dbg.hist[t]), (case, w, h, fq, i, t)
                assert res.jpegs[i] == want, (case, w, h, fq, i)
        for e in encs.values():
            e.close()


def test_rate_control_beyond_the_lookup_table(orc):
    """A 4K frame of 0/255 noise drives predict_size() past the 65,536-entry qscale table (n ~ 80,000): the clamp must be
    exact, i.e. the rate control has saturated (qscale 25) before the table ends.  Also the largest scan in the suite."""
    import h2j_b200

    w, h = 3840, 2160
    y, u, v = orc.synth_planes(w, h, "binary", seed=99, amp=0)
    want, dbg, _ = orc.oracle_encode(y, u, v)
    assert dbg.qscale == 25 and 826.0 * (dbg.mb_var_sum ** 0.5) / 236.0 > 65536
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1, max_jpeg_bytes=len(want) + 65536) as e:
        got = e.yuv2jpeg(y, u, v)
        info = e.frame_info(0, 0)
    assert info.qscale == 25 and info.mb_var_sum == dbg.mb_var_sum
    assert got == want


def test_single_frame_from_pinned_planes_skips_the_staging_copy(enc, orc):
    """h2j_encode_frame with planes in page-locked memory (strided, as an AVFrame from a pinned pool would be): uploaded from
    where they are; same bytes as from pageable planes."""
    import h2j_b200

    w, h = 1918, 1078  # odd width: pitched device layout, strided 2-D copies
    y, u, v = orc.synth_planes(w, h, "textured", seed=31, amp=33)
    ys, cs = 1920 + 64, 960 + 32
    buf = h2j_b200.PinnedBuffer(ys * h + 2 * cs * ((h + 1) // 2))
    a = buf.array
    py = a[: ys * h].reshape(h, ys)[:, :w]
    pu = a[ys * h: ys * h + cs * ((h + 1) // 2)].reshape((h + 1) // 2, cs)[:, : (w + 1) // 2]
    pv = a[ys * h + cs * ((h + 1) // 2):].reshape((h + 1) // 2, cs)[:, : (w + 1) // 2]
    py[:], pu[:], pv[:] = y, u, v
    want = orc.oracle_encode(y, u, v)[0]
    assert enc.yuv2jpeg(py, pu, pv) == want
    assert enc.yuv2jpeg(y, u, v) == want
    del py, pu, pv, a
    buf.free()
