"""CPU suite: ff_fdct_sse2 works in saturating 16-bit lanes (paddsw/psubsw/packssdw); csrc/h2j_math.cuh uses plain
32-bit adds.  The two agree iff no intermediate saturates for 8-bit samples.  The oracle keeps the saturating lane
semantics, so equality on the extreme patterns below (every separable 0/255 pattern, AND/XOR/OR combined, plus
their complements and sparse impulses) is the check DESIGN.md §2 refers to."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shimb") / "libmathshim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "support", "math_shim.cpp")], check=True)
    lib = C.CDLL(so)
    lib.shim_fdct.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.shim_quant.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.shim_matrix.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _extreme_blocks():
    bits = ((np.arange(256)[:, None] >> np.arange(8)[None, :]) & 1).astype(np.int16)  # [pattern, k]
    out = []
    step = 5  # 52 x 52 pattern pairs per operator: the butterflies are symmetric, a stride keeps the suite short
    rows, cols = bits[::step], bits[::step]
    r = rows[:, None, :, None]
    c = cols[None, :, None, :]
    for op in (np.bitwise_and, np.bitwise_xor, np.bitwise_or):
        m = op(np.broadcast_to(r, (len(rows), len(cols), 8, 8)), np.broadcast_to(c, (len(rows), len(cols), 8, 8)))
        out.append((m * 255).reshape(-1, 64))
        out.append(((1 - m) * 255).reshape(-1, 64))
    eye = np.eye(64, dtype=np.int16)
    out.append(eye * 255)
    out.append((1 - eye) * 255)
    return np.ascontiguousarray(np.concatenate(out).astype(np.int16))


def test_no_saturation_on_extreme_patterns(orc, shim):
    blocks = _extreme_blocks()
    got = np.zeros_like(blocks)
    shim.shim_fdct(blocks.ctypes.data, got.ctypes.data, len(blocks))
    want = blocks.copy()
    lib = orc.oracle()
    f = lib.orc_fdct_sse2
    base = want.ctypes.data
    for i in range(len(want)):
        f(base + i * 128)
    assert (got == want).all()
    # head-room actually observed: far from the int16 limits the SSE2 code would clamp at
    assert int(np.abs(want).max()) <= 16320


def test_dc_is_the_plain_sample_sum(orc):
    """K2 predicts the DC across tile borders from pixel sums alone (csrc/h2j_k_fdct.cuh)."""
    rng = np.random.default_rng(11)
    blocks = np.ascontiguousarray(rng.integers(0, 256, (2000, 64)).astype(np.int16))
    sums = blocks.astype(np.int64).sum(axis=1)
    lib = orc.oracle()
    for i in range(len(blocks)):
        lib.orc_fdct_sse2(blocks[i].ctypes.data)
    assert (blocks[:, 0] == sums).all()
