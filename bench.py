#!/usr/bin/env python
"""bench.py — JPEG encodes/s of the B200 YUV->JPEG path on synthetic textured 1080p frames.

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's ffmpeg-mjpeg CPU path on the host cores

A step is one pass of the hot path over one batch of `--frames` distinct frames (default 2048 x 1080p = 6.4 GB
of input, larger than the 126 MB L2, so nothing is served from cache between steps).
  value : frames/s, whole job, inputs already resident in HBM (h2j_submit_device), CUDA-event timed.
  e2e   : frames/s through the public C ABI with HOST buffers: pinned I420 frames in (H2D every step), JPEG
          bytes out to pinned host memory (D2H every step), wall-clock between device synchronisations.
  roofline : the dominant kernel's algorithmic bytes / its CUDA-event duration, against MEASURED_PEAKS.json.
  cpu_baseline : the compiled reference (oracle/_ref, the reference's own Encoder.cpp + vendored libavcodec)
          timed on this box's host cores on a bounded sample of the same frames.
Frames shard across ranks with no collective (images are independent): scaling is "weak".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "h264-h265-to-jpeg_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "jpeg_encodes_per_sec_1080p"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=2048, help="frames per step per GPU")
    ap.add_argument("--sub-batch", type=int, default=0, help="frames per submitted batch of the device-resident pass (0 = --frames)")
    ap.add_argument("--slots", type=int, default=1, help="batches in flight (streams) of the device-resident pass; 1 keeps the per-kernel "
                                                         "CUDA-event brackets free of other streams' kernels")
    ap.add_argument("--e2e-sub-batch", type=int, default=64, help="frames per submitted batch of the host-to-host pass")
    ap.add_argument("--e2e-slots", type=int, default=4, help="batches in flight of the host-to-host pass (overlaps H2D, kernels, D2H)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU baseline sample (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall time the CPU baseline sample should take")
    ap.add_argument("--max-jpeg-bytes", type=int, default=0, help="per-frame output capacity (0 = 2 MiB, the reference's HEAP_SIZE)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-every", type=int, default=5, help="per-kernel CUDA-event brackets on every K-th timed step (1 = all)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="skip the informational two-stream pass")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# synthetic workload: textured frames (smooth gradients + band-limited texture + mild noise), made on the GPU
# ------------------------------------------------------------------------------------------------------
def make_frames_torch(n, w, h, device, seed0=0):
    import torch

    cw, ch = (w + 1) // 2, (h + 1) // 2
    fb = w * h + 2 * cw * ch
    stride = (fb + 255) // 256 * 256
    out = torch.zeros((n, stride), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)

    def plane(hh, ww, amp, seed):
        yy = torch.arange(hh, device=device, dtype=torch.float32)[:, None]
        xx = torch.arange(ww, device=device, dtype=torch.float32)[None, :]
        s = float(seed % 97)
        p = 128 + 50 * torch.sin(xx / 97.0 + s) * torch.cos(yy / 61.0 - s) + amp * torch.sin(xx / 3.1 + yy / 4.3 + s) * torch.sin(yy / 2.3 + 0.37 * s)
        g.manual_seed(seed)
        p = p + torch.randn((hh, ww), device=device, generator=g) * (amp / 6.0)
        return p.clamp(0, 255).to(torch.uint8)

    for i in range(n):
        seed = seed0 + i
        amp = 20 + (seed * 7) % 40  # varies the rate-control outcome from frame to frame
        out[i, : w * h] = plane(h, w, amp, 3 * seed).reshape(-1)
        out[i, w * h: w * h + cw * ch] = plane(ch, cw, amp // 2, 3 * seed + 1).reshape(-1)
        out[i, w * h + cw * ch: fb] = plane(ch, cw, amp // 2, 3 * seed + 2).reshape(-1)
    return out, fb, stride


def make_frames_numpy(n, w, h, seed0=0):
    from tests.support import oracle as orc

    base = [orc.pack_i420(*orc.synth_planes(w, h, "textured", seed=seed0 + s, amp=20 + (s * 7) % 40)) for s in range(min(n, 8))]
    return np.stack([base[i % len(base)] for i in range(n)])


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML is polled from a thread every couple of milliseconds
    (the timed region of the default run lasts ~30 ms: `nvidia-smi -lms` does not even start up in that time, which is
    what an 8-GPU run showed); `nvidia-smi` stays as the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.h = None
        self.p = None
        self.f = None
        self.thread = None
        self.samples = []
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(gpu_index)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self._stop:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            import threading

            self._stop = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "source": "nvml"}
            nv = self.nv
            names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                     ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
            reasons = sorted({n for _, r in self.samples for n, bit in names if r & bit})
            return {"sm_mhz": float(np.median([m for m, _ in self.samples])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(self.samples), "source": "nvml, polled every 2 ms during the timed region"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------------
# host memory placement: pinned buffers are allocated (and first touched) on the GPU's own NUMA node
# ------------------------------------------------------------------------------------------------------
class NumaLocal:
    """Context manager: while active, the process runs on the CPUs of the NUMA node the GPU hangs off, so that the
    page-locked buffers allocated inside land in that node's memory (one PCIe root away from the GPU instead of across
    the socket interconnect -- matters when eight ranks pull 50 GB/s each).  The previous affinity is restored on exit:
    the CPU baseline and everything else keep all the cores.  Does nothing when the topology cannot be read."""

    def __init__(self, device_index):
        self.node = None
        self.saved = None
        try:
            import torch

            pr = torch.cuda.get_device_properties(device_index)
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            if node < 0:
                return
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            self.cpus = cpus & os.sched_getaffinity(0)
            if self.cpus:
                self.node = node
        except Exception:
            self.node = None

    def __enter__(self):
        if self.node is not None:
            self.saved = os.sched_getaffinity(0)
            os.sched_setaffinity(0, self.cpus)
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False

# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------------
def reference_lib():
    from tests.support import oracle as orc

    if not orc.have_reference():
        return None
    lib = orc.reference()
    lib.ref_yuv2jpeg_batch_mt.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return lib


def cpu_encode_sample(frames: np.ndarray, w, h, n, threads):
    """Encode n frames (cycling over `frames`) on `threads` host threads; returns (seconds, kind)."""
    from tests.support import oracle as orc

    reps = (n + frames.shape[0] - 1) // frames.shape[0]
    sample = np.ascontiguousarray(np.concatenate([frames] * reps)[:n])
    sizes = np.zeros(n, np.int64)
    ref = reference_lib()
    if ref is not None:
        t0 = time.perf_counter()
        ok = ref.ref_yuv2jpeg_batch_mt(sample.ctypes.data, sample.shape[1], n, w, h, sizes.ctypes.data, threads)
        dt = time.perf_counter() - t0
        assert ok == n, f"reference encoded {ok}/{n} frames"
        return dt, "reference", sizes
    lib = orc.oracle()
    p = orc.Params(w, h, orc.NOPTS, 0, 0, None)
    cap = 4 * 1024 * 1024
    out = np.empty(cap * min(n, 64), np.uint8)
    t0 = time.perf_counter()
    done = 0
    while done < n:
        m = min(64, n - done)
        lib.orc_encode_batch_mt(sample[done:].ctypes.data, sample.shape[1], m, C.byref(p), out.ctypes.data, cap, sizes[done:].ctypes.data, threads)
        done += m
    return time.perf_counter() - t0, "port", sizes


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    w, h = a.width, a.height
    cores = os.cpu_count() or 1
    frames = make_frames_numpy(8, w, h)
    # bounded sample per step: sized from a calibration pass so that warmup + steps take about a minute in all
    n0 = max(2 * cores, 16)
    dt0, _, _ = cpu_encode_sample(frames, w, h, n0, cores)
    per_step_s = min(6.0, 60.0 / max(1, a.steps + a.warmup))
    n = a.cpu_frames or max(n0, int(per_step_s * n0 / dt0))
    for _ in range(a.warmup):
        cpu_encode_sample(frames, w, h, n, cores)
    times = []
    kind = "port"
    for _ in range(a.steps):
        dt, kind, _ = cpu_encode_sample(frames, w, h, n, cores)
        times.append(dt)
    total = sum(times)
    value = n * a.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16",
        "data": "synthetic", "mpixel_per_s": value * w * h / 1e6,
        "config": {"workload": f"{n} synthetic textured {w}x{h} yuv420p frames per step through the reference's Encoder::yuv2Jpeg "
                               f"(ffmpeg mjpeg, libavcodec 58.117.101) on {cores} host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{n} frames x {a.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    import h2j_b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the h2j_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w, h = a.width, a.height
    F, SB, NS = a.frames, a.sub_batch or a.frames, a.slots
    assert F % SB == 0 and F % a.e2e_sub_batch == 0
    nsub = F // SB
    lo, hi = h2j_b200.shard_range(world * F, rank, world)  # the job is world*F distinct frames, sharded per image
    assert hi - lo == F
    d_frames, fb, stride = make_frames_torch(F, w, h, dev, seed0=lo)
    torch.cuda.synchronize()

    enc = h2j_b200.Encoder(max_width=w, max_height=h, max_batch=SB, n_slots=NS, device=local_rank, profile=True,
                           max_jpeg_bytes=a.max_jpeg_bytes)
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    for s_i, st in enumerate(streams):
        enc.set_stream(s_i, st.cuda_stream)
    main = torch.cuda.current_stream(dev)

    kernel_ms = {}
    kernel_calls = {}

    def harvest(slot):
        for name, ms in enc.kernel_ms(slot):
            kernel_ms[name] = kernel_ms.get(name, 0.0) + ms
            kernel_calls[name] = kernel_calls.get(name, 0) + 1

    sizes_total = [0]

    def device_step(record=False):
        inflight = []
        for i in range(nsub):
            slot = i % NS
            if len(inflight) == NS:
                s0 = inflight.pop(0)
                _, _, sizes, st = enc.collect_device(s0)
                if record:
                    harvest(s0)
                    sizes_total[0] += int(sizes.sum())
            enc.submit_device(slot, d_frames.data_ptr() + i * SB * stride, stride, SB, w, h)
            inflight.append(slot)
        for s0 in inflight:
            _, _, sizes, st = enc.collect_device(s0)
            if record:
                harvest(s0)
                sizes_total[0] += int(sizes.sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity spot check against the oracle (outside every timed region) ---------------------------
    from tests.support import oracle as orc

    parity = None
    try:
        host2 = d_frames[:2].cpu().numpy()
        enc.submit_device(0, d_frames.data_ptr(), stride, 2, w, h)
        res = enc.collect(0)
        ok = True
        for i in range(2):
            y, u, v = h2j_b200.split_planes(host2[i], w, h)
            want, _, _ = orc.oracle_encode(np.ascontiguousarray(y), np.ascontiguousarray(u), np.ascontiguousarray(v))
            ok = ok and (res.jpegs[i] == want)
        parity = bool(ok)
    except Exception as ex:  # the bench still reports; the tests are the gate
        parity = f"not checked: {ex}"

    # ---- value: device-resident -------------------------------------------------------------------
    for _ in range(a.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = enc.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main)
    for st in streams:
        st.wait_stream(main)
    t0 = time.perf_counter()
    for step_i in range(a.steps):
        # per-kernel event brackets on a sample of the timed steps: a bracket leaves the GPU idle for a few microseconds at
        # every kernel boundary (eight per step, 2.5 % of a step), which is instrumentation, not the encoder
        enc.set_profile(step_i % max(1, a.profile_every) == 0)
        device_step(record=True)
    for st in streams:
        main.wait_stream(st)
    ev1.record(main)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    launches = enc.kernel_launches - launches0
    dev_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([dev_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * F * a.steps / (dev_ms_max / 1000.0)
    if world > 1:  # every rank's own time and clock, so that a slow GPU (power cap, clocks) shows in the line
        g = torch.zeros(world, 2, device=dev)
        g[rank, 0] = dev_ms
        g[rank, 1] = clocks.get("sm_mhz") or 0.0
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        flags = torch.zeros(world, device=dev)
        flags[rank] = 1.0 if clocks.get("reasons") and clocks["reasons"] != ["no samples"] else 0.0
        dist.all_reduce(flags, op=dist.ReduceOp.SUM)
        clocks["per_rank"] = {"ms_timed_region": [round(float(x), 3) for x in g[:, 0].tolist()], "sm_mhz": [float(x) for x in g[:, 1].tolist()],
                              "ranks_with_throttle_reasons": [i for i, x in enumerate(flags.tolist()) if x > 0]}

    # ---- informational: the same job with two half-batches in flight on two streams ------------------
    # (the HBM-bound K1 and the latency-bound K3 of one half run under the issue-bound kernels of the other; not the
    # headline `value`, whose single stream keeps the per-kernel brackets clean)
    overlap = None
    if not a.no_overlap and NS == 1 and nsub == 1 and F % 2 == 0 and F >= 64:
        try:
            H = F // 2
            enc_o = h2j_b200.Encoder(max_width=w, max_height=h, max_batch=H, n_slots=2, device=local_rank, max_jpeg_bytes=a.max_jpeg_bytes)
            ost = [torch.cuda.Stream(device=dev) for _ in range(2)]
            for s_i, st in enumerate(ost):
                enc_o.set_stream(s_i, st.cuda_stream)

            def overlap_step():
                for s_i in range(2):
                    enc_o.submit_device(s_i, d_frames.data_ptr() + s_i * H * stride, stride, H, w, h)
                for s_i in range(2):
                    enc_o.collect_device(s_i)

            for _ in range(a.warmup):
                overlap_step()
            barrier()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record(main)
            for st in ost:
                st.wait_stream(main)
            for _ in range(a.steps):
                overlap_step()
            for st in ost:
                main.wait_stream(st)
            o1.record(main)
            torch.cuda.synchronize()
            barrier()
            to = torch.tensor([o0.elapsed_time(o1)], device=dev)
            if world > 1:
                dist.all_reduce(to, op=dist.ReduceOp.MAX)
            overlap = {"value": world * F * a.steps / (float(to.item()) / 1000.0), "unit": UNIT, "slots": 2, "sub_batch": H,
                       "note": "informational: two half-batches in flight on two streams, device-resident, CUDA events"}
            enc_o.close()
        except Exception as ex:
            overlap = {"error": str(ex)}

    # ---- e2e: pinned host in, pinned host out ------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        ESB, ENS = a.e2e_sub_batch, a.e2e_slots
        ensub = F // ESB
        enc.close()
        out_cap = ESB * 2 * 1024 * 1024
        numa = NumaLocal(local_rank)
        with numa:
            enc2 = h2j_b200.Encoder(max_width=w, max_height=h, max_batch=ESB, n_slots=ENS, device=local_rank)
            h_in = h2j_b200.PinnedBuffer(F * stride)
            h_in.array[:] = d_frames.reshape(-1).cpu().numpy()
            h_out = [h2j_b200.PinnedBuffer(out_cap) for _ in range(ENS)]
            for b in h_out:
                b.array[::4096] = 0
        d2h = [0]
        launches_e2e0 = enc2.kernel_launches

        def e2e_step():
            inflight = []
            for i in range(ensub):
                slot = i % ENS
                if len(inflight) == ENS:
                    s0 = inflight.pop(0)
                    offs, st = enc2.collect_into(s0, h_out[s0].ptr, out_cap)
                    d2h[0] += int(offs[-1])
                enc2.submit_host(slot, h_in.ptr + i * ESB * stride, stride, ESB, w, h)
                inflight.append(slot)
            for s0 in inflight:
                offs, st = enc2.collect_into(s0, h_out[s0].ptr, out_cap)
                d2h[0] += int(offs[-1])

        for _ in range(a.warmup):
            e2e_step()
        barrier()
        d2h[0] = 0
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_max = float(t.item())
        e2e = {"value": world * F * a.steps / dt_max, "unit": UNIT, "h2d_bytes_per_step": world * F * fb,
               "d2h_bytes_per_step": world * d2h[0] // a.steps, "timing": "wall clock between device synchronisations, max over ranks",
               "sub_batch": ESB, "slots": ENS, "h2d_gbs": world * F * fb * a.steps / dt_max / 1e9, "pinned_numa_node": numa.node,
               "note": "host-pinned I420 in, packed JPEG bytes out to pinned host memory, every step; bound by the H2D copy "
                       "(3.11 MB of pixels per frame over PCIe)"}
        enc2.close()
        h_in.free()
        for b in h_out:
            b.free()

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    mcu = ((w + 15) // 16) * ((h + 15) // 16)
    nblk = mcu * 6
    avg_jpeg = sizes_total[0] / max(1, F * a.steps)
    alg_bytes = {  # per frame; DESIGN.md section 5
        "mbvar_kernel": w * h,
        "fdct_quant_kernel": fb + nblk * 136,
        "entropy_walk_kernel": nblk * 136 + avg_jpeg,
        "scan_place_kernel": 2 * avg_jpeg,
        "stuff_kernel": 2 * avg_jpeg,
        "huffman_kernel": nblk * 2,
    }
    per_kernel = {}
    for name, ms in kernel_ms.items():
        calls = kernel_calls[name]
        avg_ms = ms / calls
        entry = {"avg_ms": avg_ms, "launches": calls}
        if name in alg_bytes:
            entry["gbs"] = alg_bytes[name] * SB / (avg_ms * 1e-3) / 1e9
        per_kernel[name] = entry
    dom = max((k for k in per_kernel if k in ("fdct_quant_kernel", "entropy_walk_kernel", "mbvar_kernel", "stuff_kernel", "scan_place_kernel")),
              key=lambda k: per_kernel[k]["avg_ms"])
    traffic, traffic_src = None, None
    try:  # DRAM bytes per frame of the dominant kernel from the committed ncu --set full capture (tools/ncu_traffic.py)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj["kernels"][dom]["dram_bytes_per_frame"] * SB
        traffic_src = f"{tj['source']} ({tj['frames_per_launch']}-frame launch, scaled to {SB})"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": per_kernel[dom]["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes[dom] * SB,
            "note": f"per-kernel CUDA events on the launching stream inside the timed region, on every {max(1, a.profile_every)}-th step "
                    "(one slot: no other stream's kernels inside a bracket)" if NS == 1 else "per-kernel CUDA events on the launching stream; several slots in flight, so a "
                    "bracket can include a neighbour stream's kernels (lower bound on the kernel's own rate)"}
    for k in ("fdct_quant_kernel", "mbvar_kernel", "entropy_walk_kernel"):
        if k in per_kernel and "gbs" in per_kernel[k]:
            roof[k + "_frac"] = per_kernel[k]["gbs"] / peak
    if roof.get("mbvar_kernel_frac", 0) > 1.0:
        roof["mbvar_kernel_note"] = ("read-only kernel: it streams faster than the peak, which was measured with a copy (reads and writes "
                                     "share the bus and pay the read/write turnarounds)")

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        host8 = d_frames[:8, :fb].cpu().numpy()
        n0 = max(4 * cores, 32)
        dt0, _, _ = cpu_encode_sample(host8, w, h, n0, cores)  # warm + calibrate
        n = a.cpu_frames or max(n0, int(a.cpu_seconds * n0 / dt0))
        dt, kind, _ = cpu_encode_sample(host8, w, h, n, cores)
        cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n} frames (8 of the bench frames, cycled) on {cores} host threads, {dt:.1f} s of wall time"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16", "data": "synthetic",
            "mpixel_per_s": value * w * h / 1e6,
            "config": {"workload": f"{F} distinct synthetic textured {w}x{h} yuv420p frames per GPU per step (configs[2] shape), "
                                   f"sub-batches of {SB} on {NS} streams, bit-exact ffmpeg-mjpeg output",
                       "frames_per_step_per_gpu": F, "sub_batch": SB, "slots": NS,
                       "l2_policy": f"inputs larger than L2: {F * fb / 1e6:.0f} MB of frames + {F * nblk * 128 / 1e6:.0f} MB of coefficients per step",
                       "avg_jpeg_bytes": avg_jpeg, "parity_spot_check_vs_oracle": parity},
            "clocks": clocks, "e2e": e2e, "two_stream": overlap, "gpu_launches": launches, "wall_ms_per_step": 1000 * wall / a.steps,
            "roofline": roof, "kernels": per_kernel, "cpu_baseline": cpu,
        }
        emit_line(line)
    enc.close()  # (idempotent: already closed when the host-to-host pass ran)
    if world > 1:
        dist.destroy_process_group()


_json_out = None


def emit_line(line):
    print(json.dumps(line), file=_json_out or sys.stdout, flush=True)


def main():
    global _json_out
    a = parse_args()
    # stdout carries exactly ONE line, the JSON: whatever libraries write to file descriptor 1 (NCCL prints its version
    # there) goes to stderr instead
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
