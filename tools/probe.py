#!/usr/bin/env python
"""Side measurements for DESIGN.md (not the bench contract): single-frame latency through h2j_encode_frame, device-resident
throughput against batch size, and 4K / odd-size throughput.  Run on the GPU box: python tools/probe.py"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h264-h265-to-jpeg_b200")]
import torch
import h2j_b200
from bench import make_frames_torch

dev = torch.device("cuda", 0)
res = {}

def throughput(w, h, n, reps=5):
    d, fb, stride = make_frames_torch(min(n, 64), w, h, dev)
    if n > d.shape[0]:
        d = d.repeat((n + d.shape[0] - 1) // d.shape[0], 1)[:n].contiguous()
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=n, n_slots=1) as e:
        for _ in range(2):
            e.submit_device(0, d.data_ptr(), stride, n, w, h); e.collect_device(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            e.submit_device(0, d.data_ptr(), stride, n, w, h); e.collect_device(0)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return n / dt, dt * 1e3

for n in (1, 4, 16, 64, 256, 512, 1024):
    fps, ms = throughput(1920, 1080, n)
    res[f"1080p_batch{n}"] = {"frames_per_s": round(fps), "ms_per_batch": round(ms, 3)}
for (w, h, n) in ((3840, 2160, 64), (1918, 1078, 256), (1280, 720, 256)):
    fps, ms = throughput(w, h, n)
    res[f"{w}x{h}_batch{n}"] = {"frames_per_s": round(fps), "mpixel_per_s": round(fps * w * h / 1e6), "ms_per_batch": round(ms, 3)}

# single frame, host planes in -> JPEG bytes out (the Encoder::yuv2Jpeg drop-in call)
from tests.support import oracle as orc
for (w, h) in ((1920, 1080), (3840, 2160)):
    y, u, v = orc.synth_planes(w, h, "textured", seed=1, amp=40)
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1) as e:
        for _ in range(3):
            e.yuv2jpeg(y, u, v)
        t0 = time.perf_counter()
        for _ in range(20):
            j = e.yuv2jpeg(y, u, v)
        dt = (time.perf_counter() - t0) / 20
    res[f"single_frame_{w}x{h}"] = {"ms_host_to_host": round(dt * 1e3, 3), "jpeg_bytes": len(j)}
print(json.dumps(res, indent=1))
