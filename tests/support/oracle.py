"""ctypes access to the CPU oracle (oracle/libh2j_oracle.so) and, when it has been built, to the compiled
reference (oracle/_ref/libh2j_ref.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ORACLE_SO = os.path.join(ROOT, "oracle", "libh2j_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libh2j_ref.so")
NOPTS = -(1 << 63)

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63])
MPEG1_INTRA = np.array([8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26,
                        27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58, 26, 27, 29, 34, 38, 46,
                        56, 69, 27, 29, 35, 38, 46, 56, 69, 83], np.uint16)


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("pts", C.c_int64), ("fixed_qscale", C.c_int), ("range_mode", C.c_int),
                ("comment", C.c_char_p), ("chroma_format", C.c_int)]


CHROMA_420, CHROMA_422, CHROMA_444 = 0, 1, 2
PIX_FMT_OF = {CHROMA_420: 12, CHROMA_422: 13, CHROMA_444: 14}  # AV_PIX_FMT_YUVJ420P / YUVJ422P / YUVJ444P


def chroma_shape(w, h, chroma_format):
    """(rows, columns) of a chroma plane as the decoder hands it over: ceil(h / 2^vshift) x ceil(w / 2^hshift)"""
    hs = 0 if chroma_format == CHROMA_444 else 1
    vs = 1 if chroma_format == CHROMA_420 else 0
    return (h + (1 << vs) - 1) >> vs, (w + (1 << hs) - 1) >> hs


def blocks_of(w, h, chroma_format):
    """number of 8x8 blocks the encoder codes"""
    mh = (h + 15) // 16
    if chroma_format == CHROMA_444:
        return ((w + 7) // 8) * mh * 6
    return ((w + 15) // 16) * mh * (6 if chroma_format == CHROMA_420 else 8)


class Debug(C.Structure):
    _fields_ = [("qscale", C.c_int), ("lambda_", C.c_int), ("mb_var_sum", C.c_int64), ("mcu_w", C.c_int), ("mcu_h", C.c_int),
                ("intra_matrix", C.c_uint8 * 64), ("qmat16", C.c_uint16 * 64), ("bias16", C.c_uint16 * 64),
                ("hist", (C.c_uint32 * 256) * 4), ("bits", (C.c_uint8 * 17) * 4), ("vals", (C.c_uint8 * 256) * 4),
                ("nvals", C.c_int * 4), ("scan_bits", C.c_int64), ("header_bytes", C.c_int), ("coefs", C.c_void_p)]


_oracle = None
_ref = None


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle"], check=True)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        vp, ci = C.c_void_p, C.c_int
        lib.orc_encode_frame.restype = C.c_long
        lib.orc_encode_frame.argtypes = [vp, ci, vp, ci, vp, ci, C.POINTER(Params), vp, C.c_long, C.POINTER(Debug)]
        lib.orc_mb_var_sum.restype = C.c_int64
        lib.orc_mb_var_sum.argtypes = [vp, ci, ci, ci]
        lib.orc_rate_control_qscale.argtypes = [C.c_int64, C.c_int64, vp]
        lib.orc_fdct_sse2.argtypes = [vp]
        lib.orc_fdct_islow.argtypes = [vp]
        lib.orc_build_matrices.argtypes = [ci, vp, vp, vp]
        lib.orc_quantize.argtypes = [vp, vp, vp, vp]
        lib.orc_huffman_table.argtypes = [vp, vp, vp, vp]
        lib.orc_range_luma.argtypes = [vp, ci, vp, ci, ci, ci]
        lib.orc_range_chroma.argtypes = [vp, ci, vp, ci, ci, ci]
        lib.orc_encode_batch_mt.argtypes = [vp, C.c_long, ci, C.POINTER(Params), vp, C.c_long, vp, ci]
        _oracle = lib
    return _oracle


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def reference() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        vp, ci = C.c_void_p, C.c_int
        lib.ref_yuv2jpeg.restype = C.c_long
        lib.ref_yuv2jpeg.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, C.c_int64, ci, vp, C.c_long]
        lib.ref_yuv2jpeg_file.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, C.c_int64, ci, C.c_char_p, ci]
        lib.ref_h265_to_jpeg.argtypes = [C.c_char_p, C.c_char_p, ci]
        lib.ref_decode_first_frame.argtypes = [C.c_char_p, vp, vp, vp, C.c_long, vp]
        lib.ref_fdct.argtypes = [vp, ci]
        lib.ref_sws_limited_to_full.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
        lib.ref_version.restype = C.c_char_p
        lib.ref_mjpeg_encode_fmt.restype = C.c_long
        lib.ref_mjpeg_encode_fmt.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci, ci, vp, C.c_long]
        _ref = lib
    return _ref


def oracle_encode(y, u, v, fixed_qscale=0, range_mode=0, comment=None, want_coefs=False, pts=NOPTS, chroma_format=0):
    """Returns (jpeg bytes, Debug, coefs or None)."""
    h, w = y.shape
    lib = oracle()
    p = Params(w, h, pts, fixed_qscale, range_mode, comment, chroma_format)
    d = Debug()
    coefs = None
    if want_coefs:
        nblk = blocks_of(w, h, chroma_format)
        coefs = np.zeros((nblk, 64), np.int16)
        d.coefs = coefs.ctypes.data
    cap = w * h * 4 + 65536
    out = np.empty(cap, np.uint8)
    n = lib.orc_encode_frame(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], C.byref(p),
                             out.ctypes.data, cap, C.byref(d))
    assert n > 0, n
    return out[:n].tobytes(), d, coefs


def reference_encode(y, u, v, pts=NOPTS, pix_fmt=0) -> bytes:
    h, w = y.shape
    cap = 2 * 1024 * 1024 + 4096
    out = np.empty(cap, np.uint8)
    n = reference().ref_yuv2jpeg(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], w, h, pts,
                                 pix_fmt, out.ctypes.data, cap)
    assert n > 0, n
    return out[:n].tobytes()


def reference_encode_fmt(y, u, v, chroma_format) -> bytes:
    """libavcodec's mjpeg encoder opened as the reference opens it, at another chroma format (oracle/ref_harness.cpp)."""
    h, w = y.shape
    cap = 8 * 1024 * 1024
    out = np.empty(cap, np.uint8)
    n = reference().ref_mjpeg_encode_fmt(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], w, h,
                                         PIX_FMT_OF[chroma_format], out.ctypes.data, cap)
    assert n > 0, n
    return out[:n].tobytes()


def synth_planes_fmt(w, h, chroma_format, kind="textured", seed=0, amp=40):
    """synth_planes with chroma planes of the given format's size"""
    y, _, _ = synth_planes(w, h, kind, seed=seed, amp=amp)
    ch, cw = chroma_shape(w, h, chroma_format)
    cy, _, _ = synth_planes(cw, ch, kind, seed=seed + 1000, amp=max(1, amp // 2))
    cz, _, _ = synth_planes(cw, ch, kind, seed=seed + 2000, amp=max(1, amp // 2))
    return y, cy, cz


def synth_planes(w, h, kind="textured", seed=0, amp=40):
    """Deterministic synthetic 4:2:0 planes.  'textured' is the bench workload: smooth gradients +
    band-limited texture + a little noise, so that the DCT spectrum looks like camera footage."""
    rng = np.random.default_rng(seed)
    cw, ch = (w + 1) // 2, (h + 1) // 2

    def plane(hh, ww, a):
        yy, xx = np.mgrid[0:hh, 0:ww].astype(np.float32)
        if kind == "textured":
            p = 128 + 50 * np.sin(xx / 97.0 + seed) * np.cos(yy / 61.0 - seed) + a * np.sin(xx / 3.1 + yy / 4.3 + seed) * np.sin(yy / 2.3)
            p = p + rng.normal(0, a / 6.0, (hh, ww))
        elif kind == "noise":
            p = rng.integers(128 - a, 128 + a + 1, (hh, ww))
        elif kind == "binary":
            p = rng.integers(0, 2, (hh, ww)) * 255
        elif kind == "const":
            p = np.full((hh, ww), int(rng.integers(0, 256)))
        elif kind == "blocks":
            p = np.kron(rng.integers(0, 256, ((hh + 7) // 8, (ww + 7) // 8)), np.ones((8, 8)))[:hh, :ww] + rng.integers(-a, a + 1, (hh, ww))
        elif kind == "ff":  # maximises 0xFF bytes in the scan: strong texture
            p = 128 + 127 * np.sign(np.sin(xx * 1.7) * np.sin(yy * 1.3))
        else:
            raise ValueError(kind)
        return np.clip(p, 0, 255).astype(np.uint8)

    return plane(h, w, amp), plane(ch, cw, amp // 2), plane(ch, cw, amp // 2)


def pack_i420(y, u, v) -> np.ndarray:
    return np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)])


def golden_planes(w, h, seed, amp):
    """Integer-only deterministic frame generator for the committed golden vectors (no floating point, so the
    planes are bit-identical on every platform): random 8x8-block means + random walk texture, clipped."""
    rng = np.random.default_rng(seed)
    cw, ch = (w + 1) // 2, (h + 1) // 2

    def plane(hh, ww, a):
        blocks = rng.integers(32, 224, ((hh + 7) // 8, (ww + 7) // 8), dtype=np.int64)
        base = np.kron(blocks, np.ones((8, 8), dtype=np.int64))[:hh, :ww]
        walk = np.cumsum(rng.integers(-a, a + 1, (hh, ww), dtype=np.int64), axis=1) // 4
        tex = rng.integers(-a, a + 1, (hh, ww), dtype=np.int64)
        return np.clip(base + walk + tex, 0, 255).astype(np.uint8)

    return plane(h, w, amp), plane(ch, cw, max(1, amp // 2)), plane(ch, cw, max(1, amp // 2))


def golden_planes_fmt(w, h, seed, amp, chroma_format):
    """golden_planes() with chroma planes of the given format's size (integer-only, platform independent)."""
    rng = np.random.default_rng(seed + 7919 * (chroma_format + 1))
    ch, cw = chroma_shape(w, h, chroma_format)

    def plane(hh, ww, a):
        blocks = rng.integers(32, 224, ((hh + 7) // 8, (ww + 7) // 8), dtype=np.int64)
        base = np.kron(blocks, np.ones((8, 8), dtype=np.int64))[:hh, :ww]
        walk = np.cumsum(rng.integers(-a, a + 1, (hh, ww), dtype=np.int64), axis=1) // 4
        tex = rng.integers(-a, a + 1, (hh, ww), dtype=np.int64)
        return np.clip(base + walk + tex, 0, 255).astype(np.uint8)

    return plane(h, w, amp), plane(ch, cw, max(1, amp // 2)), plane(ch, cw, max(1, amp // 2))


def decode_coefs(jpeg: bytes):
    """Independent baseline entropy decode (oracle/mjpeg_oracle.c orc_jpeg_decode_coefs).
    Returns (levels[n_blocks, 64] zigzag, info = [w, h, n_ff00, scan_bytes])."""
    lib = oracle()
    lib.orc_jpeg_decode_coefs.restype = C.c_long
    lib.orc_jpeg_decode_coefs.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
    jb = np.frombuffer(jpeg, np.uint8)
    # geometry from SOF0
    i = jpeg.find(b"\xff\xc0")
    h = (jpeg[i + 5] << 8) | jpeg[i + 6]
    w = (jpeg[i + 7] << 8) | jpeg[i + 8]
    hv = [(jpeg[i + 11 + 3 * c] >> 4, jpeg[i + 11 + 3 * c] & 15) for c in range(3)]
    per_mcu = hv[0][0] * hv[0][1] + 2 * hv[1][0] * hv[1][1]
    nblk = ((w + 8 * hv[0][0] - 1) // (8 * hv[0][0])) * ((h + 15) // 16) * per_mcu
    out = np.zeros((nblk, 64), np.int16)
    info = np.zeros(4, np.int64)
    n = lib.orc_jpeg_decode_coefs(jb.ctypes.data, len(jpeg), out.ctypes.data, nblk, info.ctypes.data)
    assert n == nblk, f"entropy decode failed: {n}"
    return out, info
