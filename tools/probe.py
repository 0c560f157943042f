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

# NV12 device input (NVDEC layout: pitch 2048, chroma plane behind 1088 luma rows), 256 x 1080p
def nv12_throughput(w, h, n, pitch, rows, reps=5):
    d, fb, stride = make_frames_torch(min(n, 32), w, h, dev)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    k = d.shape[0]
    nv = torch.zeros((k, pitch * (rows + ch)), dtype=torch.uint8, device=dev)
    nv[:, : pitch * h].view(k, h, pitch)[:, :, :w] = d[:, : w * h].view(k, h, w)
    uvv = nv[:, pitch * rows :].view(k, ch, pitch)
    uvv[:, :, 0 : 2 * cw : 2] = d[:, w * h : w * h + cw * ch].view(k, ch, cw)
    uvv[:, :, 1 : 2 * cw : 2] = d[:, w * h + cw * ch : w * h + 2 * cw * ch].view(k, ch, cw)
    nv = nv.repeat((n + k - 1) // k, 1)[:n].contiguous()
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=n, n_slots=1) as e:
        for _ in range(2):
            e.submit_device_nv12(0, nv.data_ptr(), nv.shape[1], pitch, pitch * rows, n, w, h); e.collect_device(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            e.submit_device_nv12(0, nv.data_ptr(), nv.shape[1], pitch, pitch * rows, n, w, h); e.collect_device(0)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return n / dt, dt * 1e3

fps, ms = nv12_throughput(1920, 1080, 256, 2048, 1088)
res["nv12_1080p_batch256"] = {"frames_per_s": round(fps), "ms_per_batch": round(ms, 3)}
print(json.dumps(res, indent=1))
