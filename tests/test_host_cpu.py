"""CPU suite: the host mirror's batch machinery (host/Encoder.cpp) without a GPU.  With no CUDA device every encode fails
LOUDLY (there is no CPU path), but the scope itself -- begin / end, several caller threads, the counters -- must run to completion and report
every picture as refused or failed instead of hanging or inventing output."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "h264-h265-to-jpeg_b200", "lib", "libh2j_host.so")


def _lib():
    lib = C.CDLL(HOST_SO)
    lib.h2j_host_yuv2jpeg_file.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    lib.h2j_host_batch_end.argtypes = [C.POINTER(C.c_int)]
    return lib


def _frame(w, h, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (h, w)).astype(np.uint8), rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2)).astype(np.uint8),
            rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2)).astype(np.uint8))


def test_host_paths_fail_loudly_and_terminate_without_a_gpu(tmp_path, capfd):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_dropin_gpu.py covers the real thing")
    lib = _lib()
    y, u, v = _frame(64, 48, 1)
    out = str(tmp_path / "a.jpeg").encode()
    # synchronous call: false, no file
    assert lib.h2j_host_yuv2jpeg_file(y.ctypes.data, 64, u.ctypes.data, 32, v.ctypes.data, 32, 64, 48, 0, out) == 0
    assert not os.path.exists(out)
    # batch scope from three threads, two sizes: every picture is accepted into the queue, none can be encoded
    assert lib.h2j_host_batch_end(None) == -1
    assert lib.h2j_host_batch_begin(4) == 0
    assert lib.h2j_host_batch_begin(4) == -1
    accepted = []

    def caller(t):
        for i in range(7):
            w, h = (64, 48) if (i + t) % 3 else (80, 32)
            yy, uu, vv = _frame(w, h, 10 * t + i)
            p = str(tmp_path / f"t{t}_{i}.jpeg").encode()
            accepted.append(lib.h2j_host_yuv2jpeg_file(yy.ctypes.data, w, uu.ctypes.data, (w + 1) // 2, vv.ctypes.data, (w + 1) // 2, w, h, 0, p))

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
        assert not t.is_alive()
    failed = C.c_int(-1)
    written = lib.h2j_host_batch_end(C.byref(failed))
    # (without a CUDA driver not even the pinned staging can be had: the pictures are refused at the door; with a driver
    # but no usable device they are queued and fail in the worker -- either way nothing is written and nothing is lost count of)
    assert len(accepted) == 21 and written == 0 and failed.value == sum(accepted)
    assert not list(tmp_path.glob("*.jpeg"))
    log = capfd.readouterr().out
    assert "h2j_create failed" in log and ("pinned allocation" in log or sum(accepted) == 21)
    assert lib.h2j_host_devices_in_use() == 0


def test_stream_copy_is_a_memcpy():
    """h2j_stream_copy (non-temporal stores where the CPU has AVX2) against numpy on every head / tail alignment."""
    import h2j_b200

    lib = h2j_b200.load_library()
    lib.h2j_stream_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.h2j_stream_copy.restype = None
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, 1 << 20, dtype=np.uint8)
    for n in (0, 1, 31, 32, 33, 127, 128, 129, 4095, 4096, 4097, 70001, 262144, 1000003):
        for so in (0, 1, 7, 13):
            for do in (0, 3, 16, 29):
                if so + n > src.size:
                    continue
                dst = np.full(n + 64, 0xEE, np.uint8)
                lib.h2j_stream_copy(dst.ctypes.data + do, src.ctypes.data + so, n)
                assert (dst[do: do + n] == src[so: so + n]).all(), (n, so, do)
                assert (dst[:do] == 0xEE).all() and (dst[do + n:] == 0xEE).all(), (n, so, do)


def test_dropin_library_exports_the_reference_surface():
    """The drop-in (the reference's Decoder.cpp + JNI bridge over this repo's Encoder) loads without a GPU and exports what
    the reference's library exports -- IDecoder::getInstance, the JNI entry -- plus the batch scope."""
    so = os.path.join(ROOT, "h264-h265-to-jpeg_b200", "lib", "libH265ToJpeg_b200.so")
    if not os.path.exists(so):
        pytest.skip("drop-in library not built (needs /root/reference at build time)")
    lib = C.CDLL(so)
    for name in ("_ZN8IDecoder11getInstanceEv", "Java_com_autonavi_socol_occtiltedserver_service_H265DecodeService_decode",
                 "h2j_host_batch_begin", "h2j_host_batch_end", "h2j_host_configure", "h2j_host_devices_in_use", "dropin_h265_to_jpeg",
                 "dropin_loop", "_ZN7Encoder8yuv2JpegEP7AVFrame"):
        assert hasattr(lib, name), name
    # the reference's own Encoder machinery is NOT in it (no second encoder hiding behind the class)
    assert not hasattr(lib, "_ZN7Encoder19constructOutputDataEv")
