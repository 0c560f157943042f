#!/usr/bin/env python
"""A handful of small encodes for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h264-h265-to-jpeg_b200")]
import h2j_b200
from tests.support import oracle as orc

ok = True
cases = [(322, 242, "textured", 40, 0, 0), (131, 77, "noise", 80, 0, 0), (640, 368, "textured", 30, 0, 1), (272, 208, "noise", 127, 1, 0), (1920, 1080, "textured", 40, 0, 0)]
for (w, h, kind, amp, fq, rm) in cases:
    y, u, v = orc.synth_planes(w, h, kind, seed=w, amp=amp)
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=3, n_slots=2, fixed_qscale=fq, range_mode=rm, max_jpeg_bytes=8 << 20) as e:
        frames = np.stack([orc.pack_i420(y, u, v)] * 3)
        res = e.encode_batch(frames, w, h, slot=1)
        one = e.yuv2jpeg(y, u, v)
        e.convert_pad(y, u, v, 1)
    want = orc.oracle_encode(y, u, v, fixed_qscale=fq, range_mode=rm)[0]
    good = all(j == want for j in res.jpegs) and one == want
    print(w, h, kind, "ok" if good else "MISMATCH", len(want))
    ok = ok and good
# the other chroma formats (single-component roles; 4:2:2's four-unit tiles), several tiles per CTA on a partial last tile
for fmt in (1, 2):
    for (w, h, kind, amp) in ((322, 242, "textured", 40), (131, 77, "noise", 80), (641, 479, "textured", 25)):
        y, u, v = orc.synth_planes_fmt(w, h, fmt, kind, seed=w + fmt, amp=amp)
        with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=3, n_slots=1, chroma_format=fmt, max_jpeg_bytes=8 << 20) as e:
            e.set_knob("fdct_tiles_per_cta", 3)
            res = e.encode_batch(np.stack([orc.pack_i420(y, u, v)] * 3), w, h)
            one = e.yuv2jpeg(y, u, v)
        want = orc.oracle_encode(y, u, v, chroma_format=fmt)[0]
        good = all(j == want for j in res.jpegs) and one == want
        print("fmt", fmt, w, h, kind, "ok" if good else "MISMATCH", len(want))
        ok = ok and good
# NV12 device input and a frame whose scan is dense with 0xFF bytes (stuffing) / has units larger than a window
import torch
w, h = 322, 242
y, u, v = orc.synth_planes(w, h, "ff", seed=5)
cw, ch = (w + 1) // 2, (h + 1) // 2
pitch = 336
nv = np.zeros(pitch * (h + ch), np.uint8)
nv[: pitch * h].reshape(h, pitch)[:, :w] = y
uvp = nv[pitch * h:].reshape(ch, pitch)
uvp[:, 0:2 * cw:2] = u
uvp[:, 1:2 * cw:2] = v
d = torch.from_numpy(nv).cuda()
with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1, max_jpeg_bytes=8 << 20) as e:
    e.submit_device_nv12(0, d.data_ptr(), nv.size, pitch, pitch * h, 1, w, h)
    got = e.collect(0).jpegs[0]
good = got == orc.oracle_encode(y, u, v)[0]
print("nv12 ff", "ok" if good else "MISMATCH")
ok = ok and good
sys.exit(0 if ok else 1)
