"""tools/make_golden.py -- writes tests/golden/encode_golden.json from the CPU oracle (run in the authoring container).

The reference (kunpengcompute/media) holds no golden vectors for this path and its arithmetic lives in the
absent libopenh264, so these fixtures pin OUR specification of the path: every stream listed here was also decoded with
FFmpeg's independent h264 decoder and matched the oracle's reconstruction bit for bit at generation time (asserted below).
GPU tests compare the CUDA path with these hashes without needing the oracle at run time."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import orc_py
from media_b200.synth import Content
import avdec

CASES = [
    dict(name="qcif_A_qp26", w=176, h=144, kind="A", frames=5, qp=26, slices=1, sr=16),
    dict(name="tiny_A_qp26", w=64, h=48, kind="A", frames=3, qp=26, slices=1, sr=16, keep_stream=True),
    dict(name="qcif_D_qp30_2sl", w=176, h=144, kind="D", frames=3, qp=30, slices=2, sr=16),
    dict(name="odd_322x182_qp10_sr64", w=322, h=182, kind="A", frames=3, qp=10, slices=2, sr=64),
    dict(name="screen_320x180_3sl_sr32", w=320, h=180, kind="B", frames=4, qp=26, slices=3, sr=32),
    dict(name="static_640x360_4sl", w=640, h=360, kind="C", frames=3, qp=26, slices=4, sr=16),
    dict(name="qp51_96x80", w=96, h=80, kind="D", frames=3, qp=51, slices=1, sr=16),
    dict(name="qp0_48x48", w=48, h=48, kind="D", frames=2, qp=0, slices=1, sr=16),
    dict(name="one_mb", w=16, h=16, kind="A", frames=3, qp=26, slices=1, sr=16),
    dict(name="p720_A_qp26", w=1280, h=720, kind="A", frames=3, qp=26, slices=1, sr=16),
    # CABAC back end (profile 1 = Main, 2 = High; the wrapper's iEntropyCodingModeFlag = 1, VideoEncoderOpenH264.cpp:291)
    dict(name="main_qcif_A_qp26", w=176, h=144, kind="A", frames=5, qp=26, slices=1, sr=16, profile=1),
    dict(name="main_tiny_A_qp26", w=64, h=48, kind="A", frames=3, qp=26, slices=1, sr=16, profile=1, keep_stream=True),
    dict(name="high_screen_320x180_3sl_sr32", w=320, h=180, kind="B", frames=4, qp=26, slices=3, sr=32, profile=2),
    dict(name="main_noise_qp12_96x80", w=96, h=80, kind="D", frames=3, qp=12, slices=2, sr=16, profile=1),
    dict(name="main_one_mb", w=16, h=16, kind="A", frames=3, qp=26, slices=1, sr=16, profile=1),
    dict(name="high_p720_A_qp30", w=1280, h=720, kind="A", frames=3, qp=30, slices=1, sr=16, profile=2),
]

out = []
for c in CASES:
    e = orc_py.Encoder(c["w"], c["h"], num_slices=c["slices"], search_range=c["sr"], profile=c.get("profile", 0))
    content = Content(c["kind"], c["w"], c["h"])
    aus, recs = [], []
    for t in range(c["frames"]):
        aus.append(e.encode(content.frame(t), t == 0, c["qp"])); recs.append(e.recon())
    dec = avdec.decode_stream(aus)
    assert len(dec) == len(recs) and all(np.array_equal(a, b) for a, b in zip(dec, recs)), c["name"]
    rec = dict(c)
    rec["au_sha256"] = [hashlib.sha256(a).hexdigest() for a in aus]
    rec["au_bytes"] = [len(a) for a in aus]
    rec["recon_sha256"] = [hashlib.sha256(r.tobytes()).hexdigest() for r in recs]
    if c.get("keep_stream"):
        rec["stream_hex"] = [a.hex() for a in aus]
    out.append(rec)
    print(c["name"], rec["au_bytes"])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "encode_golden.json"), "w"), indent=1)
