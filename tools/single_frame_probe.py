import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h264-h265-to-jpeg_b200")]
import h2j_b200
from tests.support import oracle as orc
res = {}
for (w, h) in ((1920, 1080), (3840, 2160), (1280, 720), (1918, 1078)):
    y, u, v = orc.synth_planes(w, h, "textured", seed=1, amp=40)
    with h2j_b200.Encoder(max_width=w, max_height=h, max_batch=1, n_slots=1) as e:
        for _ in range(5):
            e.yuv2jpeg(y, u, v)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter(); j = e.yuv2jpeg(y, u, v); ts.append(time.perf_counter() - t0)
    ts.sort()
    res[f"single_frame_{w}x{h}"] = {"ms_median": round(ts[len(ts)//2] * 1e3, 3), "ms_min": round(ts[0] * 1e3, 3), "jpeg_bytes": len(j)}
print(json.dumps(res))
