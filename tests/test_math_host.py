"""CPU suite: the integer code the kernels execute (csrc/h2j_math.cuh, compiled here for the host) against the
oracle: FDCT, quantiser, matrix set-up for every qscale, lambda->qscale."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "libmathshim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "support", "math_shim.cpp")], check=True)
    lib = C.CDLL(so)
    lib.shim_fdct.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.shim_quant.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.shim_matrix.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _blocks():
    rng = np.random.default_rng(0)
    blocks = [rng.integers(0, 256, (3000, 64))]
    ext = [np.full(64, 255), np.zeros(64, np.int64)]
    for pat in range(256):
        col = np.array([(255 if (pat >> k) & 1 else 0) for k in range(8)])
        ext.append(np.tile(col[:, None], (1, 8)).reshape(64)); ext.append(np.tile(col[None, :], (8, 1)).reshape(64))
    return np.ascontiguousarray(np.concatenate(blocks + [np.array(ext)]).astype(np.int16))


def test_fdct_equals_oracle(orc, shim):
    blocks = _blocks()
    got = np.zeros_like(blocks)
    shim.shim_fdct(blocks.ctypes.data, got.ctypes.data, len(blocks))
    want = blocks.copy()
    lib = orc.oracle()
    for i in range(len(want)):
        lib.orc_fdct_sse2(want[i].ctypes.data)
    assert (got == want).all()


def test_quantiser_and_matrices_equal_oracle(orc, shim):
    blocks = _blocks()[::5]
    f = blocks.copy()
    lib = orc.oracle()
    for i in range(len(f)):
        lib.orc_fdct_sse2(f[i].ctypes.data)
    mp = orc.MPEG1_INTRA.copy()
    for qs in range(1, 32):
        im = np.zeros(64, np.uint8); q16 = np.zeros(64, np.uint16); b16 = np.zeros(64, np.uint16)
        lib.orc_build_matrices(qs, im.ctypes.data, q16.ctypes.data, b16.ctypes.data)
        dqt = np.zeros(64, np.uint8); pk = np.zeros(64, np.uint32)
        shim.shim_matrix(qs, mp.ctypes.data, dqt.ctypes.data, pk.ctypes.data)
        assert (dqt == im).all()
        assert ((pk & 0xffff) == q16).all() and ((pk >> 16) == q16.astype(np.uint32) * b16).all()
        out = np.zeros_like(f)
        shim.shim_quant(f.ctypes.data, out.ctypes.data, len(f), qs, mp.ctypes.data)
        oz = np.zeros(64, np.int16)
        for i in range(0, len(f), 7):
            lib.orc_quantize(f[i].ctypes.data, oz.ctypes.data, q16.ctypes.data, b16.ctypes.data)
            assert (oz == out[i][orc.ZIGZAG]).all(), (qs, i)


def test_lambda_to_qscale(shim):
    for lam in range(1, 4000):
        q = (lam * 139 + 8192) >> 14
        assert shim.shim_lambda_to_qscale(lam) == min(max(q, 2), 31)
