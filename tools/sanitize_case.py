"""tools/sanitize_case.py -- the two workloads run under compute-sanitizer (tools/sanitize.sh): the smoke configuration (176x144,
CAVLC, IDR + 2 P pictures) and one multi-slice CABAC batch (3 sessions of 320x192, Main and High, 3 slices, IDR + 2 P pictures).
Streams are checked against the oracle so a sanitizer-clean run is also a correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from media_b200 import enc          # noqa: E402
from media_b200.synth import Content  # noqa: E402
from oracle import orc_py           # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("smoke", "all"):
    w, h, qp = 176, 144, 26
    g = enc.Session(w, h, const_qp=qp, gop=1000, device=0); o = orc_py.Encoder(w, h); c = Content("A", w, h)
    for t in range(3):
        f = c.frame(t)
        assert g.encode(f)[0] == o.encode(f, t == 0, qp), f"smoke frame {t}"
    g.close()
    print("smoke case ok")
if which in ("cabac", "all"):
    w, h, qp, n = 320, 192, 30, 3
    prof = [1, 2, 1]
    ss = [enc.Session(w, h, const_qp=qp, gop=1000, device=0, profile=prof[i], num_slices=3) for i in range(n)]
    os_ = [orc_py.Encoder(w, h, num_slices=3, profile=prof[i]) for i in range(n)]
    cs = [Content("A" if i != 1 else "B", w, h, seed=40 + i) for i in range(n)]
    b = enc.Batch(0, ss)
    for t in range(3):
        fr = [cs[i].frame(t) for i in range(n)]
        out, _ = b.encode(fr)
        for i in range(n):
            assert out[i] == os_[i].encode(fr[i], t == 0, qp), f"cabac session {i} frame {t}"
    b.close()
    for s in ss:
        s.close()
    print("cabac batch case ok")
