"""GPU suite: the CUDA path against the COMMITTED golden vectors, i.e. against what the reference itself produced
(tests/golden/make_golden.py ran the reference's own Encoder::yuv2Jpeg) -- no oracle in between."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def enc4k():
    import h2j_b200

    e = h2j_b200.Encoder(max_width=3840, max_height=2160, max_batch=2, n_slots=1)
    yield e
    e.close()


def test_golden_frames_byte_exact_including_4k(enc4k, orc):
    d = json.load(open(os.path.join(G, "golden_frames.json")))
    assert d["reference"] == "Lavc58.117.101"
    seen_4k = False
    for fr in d["frames"]:
        y, u, v = orc.golden_planes(fr["w"], fr["h"], fr["seed"], fr["amp"])
        assert sha(y.tobytes() + u.tobytes() + v.tobytes()) == fr["planes_sha256"], "frame generator drifted"
        j = enc4k.yuv2jpeg(y, u, v)
        assert len(j) == fr["size"], fr
        assert sha(j) == fr["sha256"], fr
        seen_4k = seen_4k or (fr["w"], fr["h"]) == (3840, 2160)
    assert seen_4k


def test_reference_fixture_crops_byte_exact(enc4k):
    z = np.load(os.path.join(G, "ref_img_crops.npz"))
    for key in ("img01_h264", "img01_h265"):
        y, u, v = (np.ascontiguousarray(z[key + s]) for s in ("_y", "_u", "_v"))
        assert enc4k.yuv2jpeg(y, u, v) == z[key + "_jpeg"].tobytes(), key


def test_4k_batch_matches_oracle(enc4k, orc):
    """configs[3]: 3840x2160 and the odd-sized 1918x1078 through the batch entry points."""
    for (w, h) in ((3840, 2160), (1918, 1078)):
        planes = [orc.synth_planes(w, h, "textured", seed=s, amp=a) for s, a in ((41, 25), (42, 60))]
        frames = np.stack([orc.pack_i420(*p) for p in planes])
        res = enc4k.encode_batch(frames, w, h)
        assert res.status == [0, 0]
        for (y, u, v), got in zip(planes, res.jpegs):
            want, _, _ = orc.oracle_encode(y, u, v)
            assert got == want
