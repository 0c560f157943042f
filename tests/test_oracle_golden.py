"""CPU suite: the oracle against the committed golden vectors (all generated FROM THE REFERENCE by
tests/golden/make_golden.py — see that file for provenance)."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_golden_frames_byte_exact(orc):
    d = json.load(open(os.path.join(G, "golden_frames.json")))
    assert d["reference"] == "Lavc58.117.101"
    for fr in d["frames"]:
        if fr["w"] * fr["h"] > 1920 * 1088:
            continue  # the 4K vector is covered by the GPU suite; keeps the CPU suite short
        y, u, v = orc.golden_planes(fr["w"], fr["h"], fr["seed"], fr["amp"])
        assert sha(y.tobytes() + u.tobytes() + v.tobytes()) == fr["planes_sha256"], "frame generator drifted"
        j, dbg, _ = orc.oracle_encode(y, u, v)
        assert len(j) == fr["size"], fr
        assert sha(j) == fr["sha256"], fr
        assert j[27] == fr["dqt1"]


def test_reference_fixture_crops_byte_exact(orc):
    z = np.load(os.path.join(G, "ref_img_crops.npz"))
    for key in ("img01_h264", "img01_h265"):
        y, u, v = z[key + "_y"], z[key + "_u"], z[key + "_v"]
        j, _, _ = orc.oracle_encode(np.ascontiguousarray(y), np.ascontiguousarray(u), np.ascontiguousarray(v))
        assert j == z[key + "_jpeg"].tobytes(), key


def test_fdct_vectors(orc):
    z = np.load(os.path.join(G, "fdct_vectors.npz"))
    blocks, want = z["blocks"], z["fdct"]
    got = np.ascontiguousarray(blocks.copy())
    lib = orc.oracle()
    for i in range(len(got)):
        lib.orc_fdct_sse2(got[i].ctypes.data)
    assert (got == want).all()
    # bounds used by the CUDA code: nothing near int16 saturation for 8-bit samples
    assert np.abs(want[:, 1:]).max() <= 8160 and want[:, 0].max() <= 16320 and want[:, 0].min() >= 0


def test_ratecontrol_pairs(orc):
    d = json.load(open(os.path.join(G, "ratecontrol.json")))
    lib = orc.oracle()
    qs = set()
    for var, q in d["pairs"]:
        assert lib.orc_rate_control_qscale(int(var), 0, None) == q, (var, q)
        qs.add(q)
    assert qs == set(range(2, 26))  # every value the first-frame rate control can produce (lambda clips at 2926 -> 25)


def test_pts_does_not_matter(orc):
    lib = orc.oracle()
    for var in (1000, 123456, 9876543):
        q0 = lib.orc_rate_control_qscale(var, orc.NOPTS, None)
        for pts in (0, 1, 25, 90000, -7):
            assert lib.orc_rate_control_qscale(var, pts, None) == q0


def test_swscale_lut(orc):
    lut = np.load(os.path.join(G, "swscale_lut.npz"))["lut"]
    src = np.arange(256, dtype=np.uint8).reshape(16, 16).copy()
    dst = np.zeros_like(src)
    lib = orc.oracle()
    lib.orc_range_luma(src.ctypes.data, 16, dst.ctypes.data, 16, 16, 16)
    assert (dst.reshape(-1) == lut[0]).all()
    lib.orc_range_chroma(src.ctypes.data, 16, dst.ctypes.data, 16, 16, 16)
    assert (dst.reshape(-1) == lut[1]).all()


def test_entropy_round_trip_and_structure(orc):
    """encode -> independent entropy decode -> same levels; stuffing count and markers are consistent."""
    for (w, h, kind, seed) in [(322, 242, "textured", 1), (640, 368, "ff", 2), (17, 17, "noise", 3), (2, 2, "const", 4)]:
        y, u, v = orc.synth_planes(w, h, kind, seed=seed)
        j, dbg, coefs = orc.oracle_encode(y, u, v, want_coefs=True)
        got, info = orc.decode_coefs(j)
        assert (got == coefs).all()
        assert info[0] == w and info[1] == h
        assert j[:2] == b"\xff\xd8" and j[-2:] == b"\xff\xd9"
        scan = j[dbg.header_bytes:-2]
        assert scan.count(b"\xff") == scan.count(b"\xff\x00") == info[2]
        assert len(scan) == (dbg.scan_bits + 7) // 8 + info[2]


def test_fixed_qscale_matrix(orc):
    lib = orc.oracle()
    for q in (1, 2, 8, 31):
        im = np.zeros(64, np.uint8); q16 = np.zeros(64, np.uint16); b16 = np.zeros(64, np.uint16)
        lib.orc_build_matrices(q, im.ctypes.data, q16.ctypes.data, b16.ctypes.data)
        assert im[0] == 8
        want = np.minimum((orc.MPEG1_INTRA.astype(int) * q) >> 3, 255)
        assert (im[1:] == want[1:]).all()
        assert (q16 == (131072 // (16 * im.astype(int)))).all()


def test_oracle_matches_the_422_and_444_golden_frames(orc):
    """Row f3: the same libavcodec encoder at its other MCU geometries (yuvj422p: 16x16 MCUs of 4 Y + 2 Cb + 2 Cr; yuvj444p:
    8x16 MCUs of 2 Y + 2 Cb + 2 Cr).  The committed digests come from the libavcodec the reference vendors, opened the way
    the reference opens it (tests/golden/make_golden.py formats)."""
    import hashlib
    import json

    d = json.load(open(os.path.join(G, "golden_frames_fmt.json")))
    assert d["reference"] == "Lavc58.117.101" and len(d["frames"]) >= 20
    for fr in d["frames"]:
        y, u, v = orc.golden_planes_fmt(fr["w"], fr["h"], fr["seed"], fr["amp"], fr["chroma_format"])
        assert hashlib.sha256(y.tobytes() + u.tobytes() + v.tobytes()).hexdigest() == fr["planes_sha256"], "frame generator drifted"
        j, dbg, coefs = orc.oracle_encode(y, u, v, chroma_format=fr["chroma_format"], want_coefs=True)
        assert len(j) == fr["size"] and hashlib.sha256(j).hexdigest() == fr["sha256"], fr
        # and the oracle's own baseline decoder gets the quantised levels back out of the stream
        levels, info = orc.decode_coefs(j)
        assert levels.shape == coefs.shape and (levels == coefs).all()
