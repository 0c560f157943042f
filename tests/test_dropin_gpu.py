"""GPU suite: the boundary.  (1) host/Encoder.cpp -- the C++ mirror of the reference's Encoder class -- through its
C shim; (2) the DROP-IN: the reference's own src/Decoder.cpp + JNI bridge compiled against this repo's Encoder
(host/Makefile target `dropin`, with this repo's Encoder.h force-included in place of the reference's), driven through the unchanged IDecoder::H265ToJpeg(in, out) on the reference's own
test/img fixtures (copied next to the .so at build time), compared byte for byte with what the unmodified
reference (oracle/_ref/libh2j_ref.so) writes for the same file."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "h264-h265-to-jpeg_b200", "lib", "libh2j_host.so")
DROPIN_SO = os.path.join(ROOT, "h264-h265-to-jpeg_b200", "lib", "libH265ToJpeg_b200.so")
FIXTURES = os.path.join(ROOT, "oracle", "_ref", "fixtures")  # the reference's test/img pictures (test infrastructure)
G = os.path.join(ROOT, "tests", "golden")


def test_host_encoder_class_writes_the_oracle_bytes(orc, tmp_path):
    assert os.path.exists(HOST_SO), "libh2j_host.so missing: run __graft_entry__.build()"
    lib = C.CDLL(HOST_SO)
    lib.h2j_host_yuv2jpeg_file.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    for (w, h, fmt) in ((322, 242, 0), (1920, 1080, 12), (2562, 1442, 0)):  # the last one makes the shared encoder grow
        y, u, v = orc.synth_planes(w, h, "textured", seed=w, amp=35)
        out = str(tmp_path / f"o_{w}.jpeg")
        assert lib.h2j_host_yuv2jpeg_file(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], w, h, fmt,
                                          out.encode()) == 1
        want, _, _ = orc.oracle_encode(y, u, v)
        assert open(out, "rb").read() == want
    # error behaviour of the reference class: false, no file, on a bad path / bad format
    y, u, v = orc.synth_planes(64, 64, "textured", seed=1)
    assert lib.h2j_host_yuv2jpeg_file(y.ctypes.data, 64, u.ctypes.data, 32, v.ctypes.data, 32, 64, 64, 0, b"") == 0
    assert lib.h2j_host_yuv2jpeg_file(y.ctypes.data, 64, u.ctypes.data, 32, v.ctypes.data, 32, 64, 64, 2, str(tmp_path / "x.jpeg").encode()) == 0
    assert not os.path.exists(tmp_path / "x.jpeg")


@pytest.mark.skipif(not os.path.exists(DROPIN_SO), reason="drop-in library not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name", ["img01.h264", "img01.h265"])
def test_dropin_library_on_the_reference_fixtures(orc, name, tmp_path):
    lib = C.CDLL(DROPIN_SO)
    lib.dropin_h265_to_jpeg.argtypes = [C.c_char_p, C.c_char_p]
    src = os.path.join(FIXTURES, name)
    out = str(tmp_path / (name + ".jpeg"))
    assert lib.dropin_h265_to_jpeg(src.encode(), out.encode()) == 1
    got = open(out, "rb").read()
    # (a) against the committed golden digest of the reference's full-frame JPEG
    z = np.load(os.path.join(G, "ref_img_crops.npz"))
    key = name.replace(".", "_")
    assert hashlib.sha256(got).hexdigest() == z[key + "_full_sha256"].tobytes().decode()
    assert len(got) == int(z[key + "_full_dims"][2])
    # (b) against the unmodified reference run live on the same file, when it travelled to this box
    if orc.have_reference():
        ref_out = str(tmp_path / (name + ".ref.jpeg"))
        assert orc.reference().ref_h265_to_jpeg(src.encode(), ref_out.encode(), 1) == 1
        assert got == open(ref_out, "rb").read()


def test_host_batch_scope_mixed_sizes(orc, tmp_path):
    """h2j_host_batch_begin/end: yuv2Jpeg only queues, the GPU encodes runs of equal-sized pictures as batches, the files
    appear at the latest when the scope ends, and every file is byte-identical to the oracle."""
    lib = C.CDLL(HOST_SO)
    lib.h2j_host_yuv2jpeg_file.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    lib.h2j_host_batch_end.argtypes = [C.POINTER(C.c_int)]
    assert lib.h2j_host_batch_end(None) == -1  # no scope open
    assert lib.h2j_host_batch_begin(4) == 0
    assert lib.h2j_host_batch_begin(4) == -1   # already open
    sizes = [(322, 242), (322, 242), (322, 242), (64, 64), (64, 64), (640, 368), (322, 242)]  # a bigger one forces a pool re-cut
    want, outs = [], []
    for i, (w, h) in enumerate(sizes):
        y, u, v = orc.synth_planes(w, h, "textured", seed=50 + i, amp=25 + 5 * i)
        out = str(tmp_path / f"b_{i}.jpeg")
        assert lib.h2j_host_yuv2jpeg_file(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], w, h, 0,
                                          out.encode()) == 1
        want.append(orc.oracle_encode(y, u, v)[0])
        outs.append(out)
    assert not os.path.exists(outs[-1])  # the tail of the queue is still waiting
    failed = C.c_int(-1)
    assert lib.h2j_host_batch_end(C.byref(failed)) == len(sizes) and failed.value == 0
    for out, w_ in zip(outs, want):
        assert open(out, "rb").read() == w_
    # outside a scope the call is synchronous again
    y, u, v = orc.synth_planes(64, 64, "noise", seed=9, amp=30)
    out = str(tmp_path / "sync.jpeg")
    assert lib.h2j_host_yuv2jpeg_file(y.ctypes.data, 64, u.ctypes.data, 32, v.ctypes.data, 32, 64, 64, 0, out.encode()) == 1
    assert open(out, "rb").read() == orc.oracle_encode(y, u, v)[0]


@pytest.mark.skipif(not os.path.exists(DROPIN_SO), reason="drop-in library not built (needs /root/reference at build time)")
def test_dropin_batch_scope_on_the_reference_fixtures(tmp_path):
    """The reference's own Decoder::H265ToJpeg in a loop, inside a batch scope: H.264 and H.265 inputs decoded by
    libavcodec, queued by this repo's Encoder, encoded on the GPU as batches."""
    lib = C.CDLL(DROPIN_SO)
    lib.dropin_h265_to_jpeg.argtypes = [C.c_char_p, C.c_char_p]
    lib.h2j_host_batch_end.argtypes = [C.POINTER(C.c_int)]
    z = np.load(os.path.join(G, "ref_img_crops.npz"))
    names = ["img01.h264", "img01.h265", "img01.h265", "img01.h264", "img01.h264"]
    assert lib.h2j_host_batch_begin(3) == 0
    outs = []
    for i, name in enumerate(names):
        out = str(tmp_path / f"{i}_{name}.jpeg")
        assert lib.dropin_h265_to_jpeg(os.path.join(FIXTURES, name).encode(), out.encode()) == 1
        outs.append(out)
    failed = C.c_int(-1)
    assert lib.h2j_host_batch_end(C.byref(failed)) == len(names) and failed.value == 0
    for out, name in zip(outs, names):
        key = name.replace(".", "_")
        assert hashlib.sha256(open(out, "rb").read()).hexdigest() == z[key + "_full_sha256"].tobytes().decode()


def test_hub_uses_a_second_device_when_the_first_is_busy(orc, tmp_path):
    """The host mirror spreads over the box by itself (h2j_host_configure(-1, ...), the default): inside a batch scope fed by
    several caller threads the batches pile up while the first GPU is still coming up, a second device context is started,
    and every file -- whichever GPU made it -- is the oracle's.  Needs two GPUs."""
    import threading

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU: the hub has nothing to spread over")
    lib = C.CDLL(HOST_SO)
    lib.h2j_host_yuv2jpeg_file.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
    lib.h2j_host_batch_end.argtypes = [C.POINTER(C.c_int)]
    lib.h2j_host_configure.argtypes = [C.c_int, C.c_int]
    lib.h2j_host_configure.restype = None
    lib.h2j_host_configure(-1, 0)
    w, h, per_thread, n_threads = 640, 368, 12, 4
    planes = [orc.synth_planes(w, h, "textured", seed=900 + i, amp=20 + 3 * i) for i in range(6)]
    want = [orc.oracle_encode(*p)[0] for p in planes]
    assert lib.h2j_host_batch_begin(2) == 0
    ok = []

    def caller(t):
        for i in range(per_thread):
            y, u, v = planes[(t + i) % len(planes)]
            out = str(tmp_path / f"t{t}_{i}.jpeg").encode()
            ok.append(lib.h2j_host_yuv2jpeg_file(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], w, h, 0, out))

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
        assert not t.is_alive()
    failed = C.c_int(-1)
    assert lib.h2j_host_batch_end(C.byref(failed)) == per_thread * n_threads and failed.value == 0 and all(ok)
    assert lib.h2j_host_devices_in_use() >= 2
    for t in range(n_threads):
        for i in range(per_thread):
            assert open(tmp_path / f"t{t}_{i}.jpeg", "rb").read() == want[(t + i) % len(planes)]
    # the synchronous path from several threads at once: every call finds a device
    outs = []

    def sync_caller(t):
        y, u, v = planes[t % len(planes)]
        out = str(tmp_path / f"s{t}.jpeg")
        for _ in range(5):
            outs.append((t, lib.h2j_host_yuv2jpeg_file(y.ctypes.data, y.strides[0], u.ctypes.data, u.strides[0], v.ctypes.data, v.strides[0], w, h, 0, out.encode())))

    threads = [threading.Thread(target=sync_caller, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert len(outs) == 20 and all(r == 1 for _, r in outs)
    for t in range(4):
        assert open(tmp_path / f"s{t}.jpeg", "rb").read() == want[t % len(planes)]
    lib.h2j_host_configure(0, 0)  # the other tests of this process expect device 0
