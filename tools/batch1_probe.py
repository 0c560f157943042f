"""Per-kernel times of ONE picture (batch of one, device resident, per-kernel CUDA-event brackets): where the yuv2Jpeg call
shape spends its kernel time.  usage (GPU box): python tools/batch1_probe.py [w h]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h264-h265-to-jpeg_b200")]
import numpy as np, torch
import h2j_b200
from tests.support import oracle as orc

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
y, u, v = orc.synth_planes(w, h, "textured", seed=1, amp=40)
frame = torch.from_numpy(orc.pack_i420(y, u, v)).cuda()
fb = frame.numel()
with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1, profile=True) as e:
    acc, n = {}, 0
    for it in range(60):
        e.submit_device(0, frame.data_ptr(), fb, 1, w, h)
        e.collect_device(0)
        if it >= 10:
            n += 1
            for name, ms in e.kernel_ms(0):
                acc[name] = acc.get(name, 0.0) + ms
    per = {k: round(1e3 * v_ / n, 2) for k, v_ in acc.items()}
    e.set_profile(False)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it in range(200):
        e.submit_device(0, frame.data_ptr(), fb, 1, w, h)
        e.collect_device(0)
    ev1.record(); torch.cuda.synchronize()
print(json.dumps({"size": [w, h], "kernel_us": per, "sum_us": round(sum(per.values()), 2), "us_per_picture_unbracketed": round(1e3 * ev0.elapsed_time(ev1) / 200, 2)}))
