"""GPU suite: kernel (1) of the north star -- yuv420p limited -> yuvj420p full range with MCU edge padding -- on
its own (h2j_convert_pad) and fused in front of the encoder (range_mode = LIMITED_TO_FULL), against the oracle's
libswscale restatement (pinned to libswscale 5.8.100 by tests/test_oracle_vs_reference.py and
tests/golden/swscale_lut.npz)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_range(orc, y, u, v):
    lib = orc.oracle()
    oy, ou, ov = np.zeros_like(y), np.zeros_like(u), np.zeros_like(v)
    lib.orc_range_luma(y.ctypes.data, y.strides[0], oy.ctypes.data, oy.strides[0], y.shape[1], y.shape[0])
    lib.orc_range_chroma(u.ctypes.data, u.strides[0], ou.ctypes.data, ou.strides[0], u.shape[1], u.shape[0])
    lib.orc_range_chroma(v.ctypes.data, v.strides[0], ov.ctypes.data, ov.strides[0], v.shape[1], v.shape[0])
    return oy, ou, ov


def _pad(p, pw, ph, uw, uh):
    """What the encoder sees: the first uw x uh samples, edges replicated out to pw x ph."""
    q = p[:uh, :uw]
    return np.pad(q, ((0, ph - uh), (0, pw - uw)), mode="edge")


# 720x576 and 272x208: w = 16 (mod 32), so the chroma pitch w/2 = 8 (mod 16) and every odd chroma row starts 8 bytes off
# a 16-byte boundary (the 128-bit load path must be chosen per row, not per frame)
@pytest.mark.parametrize("w,h", [(64, 48), (322, 242), (1918, 1078), (1920, 1080), (33, 17), (720, 576), (272, 208), (848, 480)])
def test_convert_pad_matches_swscale_and_edge_rule(orc, w, h):
    import h2j_b200

    rng = np.random.default_rng(w * 31 + h)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = rng.integers(0, 256, (h, w)).astype(np.uint8)
    u = rng.integers(0, 256, (ch, cw)).astype(np.uint8)
    v = rng.integers(0, 256, (ch, cw)).astype(np.uint8)
    mw, mh = (w + 15) // 16, (h + 15) // 16
    with h2j_b200.Encoder(max_width=1920, max_height=1088, max_batch=1, n_slots=1) as e:
        for mode in (h2j_b200.RANGE_PASSTHROUGH, h2j_b200.RANGE_LIMITED_TO_FULL):
            gy, gu, gv = e.convert_pad(y, u, v, mode)
            ry, ru, rv = (y, u, v) if mode == 0 else _oracle_range(orc, y, u, v)
            assert (gy == _pad(ry, mw * 16, mh * 16, w, h)).all()
            # chroma: the encoder reads w>>1 x h>>1 samples (mpegvideo_enc.c load_input_picture)
            assert (gu == _pad(ru, mw * 8, mh * 8, w >> 1, h >> 1)).all()
            assert (gv == _pad(rv, mw * 8, mh * 8, w >> 1, h >> 1)).all()


def test_all_256_values_follow_the_committed_swscale_table():
    import h2j_b200

    lut = np.load(os.path.join(G, "swscale_lut.npz"))["lut"]
    y = np.tile(np.arange(256, dtype=np.uint8), (16, 1))
    u = np.tile(np.arange(0, 256, 2, dtype=np.uint8), (8, 1))
    u2 = np.tile(np.arange(1, 256, 2, dtype=np.uint8), (8, 1))
    with h2j_b200.Encoder(max_width=256, max_height=16, max_batch=1, n_slots=1) as e:
        gy, gu, gv = e.convert_pad(y, u, u2, h2j_b200.RANGE_LIMITED_TO_FULL)
    assert (gy[0, :256] == lut[0]).all()
    assert (gu[0, :128] == lut[1][0::2]).all() and (gv[0, :128] == lut[1][1::2]).all()


@pytest.mark.parametrize("w,h,kind", [(322, 242, "textured"), (1920, 1080, "textured"), (131, 77, "noise")])
def test_fused_range_conversion_encode(orc, w, h, kind):
    import h2j_b200

    y, u, v = orc.synth_planes(w, h, kind, seed=77, amp=70)
    with h2j_b200.Encoder(max_width=1920, max_height=1088, max_batch=1, n_slots=1, range_mode=h2j_b200.RANGE_LIMITED_TO_FULL) as e:
        got = e.yuv2jpeg(y, u, v)
        info = e.frame_info(0, 0)
    want, dbg, _ = orc.oracle_encode(y, u, v, range_mode=1)
    assert info.mb_var_sum == dbg.mb_var_sum and info.qscale == dbg.qscale
    assert got == want
    # and the two-step form (convert, then encode as-is) is the same stream
    oy, ou, ov = _oracle_range(orc, y, u, v)
    two_step, _, _ = orc.oracle_encode(oy, ou, ov)
    assert got == two_step
