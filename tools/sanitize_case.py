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
# "checked": everything, plus the worst cases for the slots / rings / lists, through libb200enc_checked.so (B200ENC_LIB); prints the count of
# device-side bound-check failures (tests/test_gpu_parity.py::test_checked_build_reports_no_bound_violation)
checked = which == "checked"
if checked:
    which = "all"
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
if checked:
    import ctypes as C
    for (w, h, kind, qp, slices, profile) in ((96, 80, "D", 0, 1, 0), (96, 80, "D", 0, 2, 1), (176, 144, "D", 4, 3, 2), (640, 368, "A", 20, 2, 2), (1280, 720, "B", 30, 0, 1)):
        g = enc.Session(w, h, const_qp=qp, num_slices=slices, gop=3, device=0, profile=profile)
        ps, ks = g.slice_counts()       # the engine's automatic rules (key pictures of CABAC sessions take more slices)
        o = orc_py.Encoder(w, h, num_slices=ps, key_slices=ks, profile=profile)
        c = Content(kind, w, h)
        for t in range(4):
            f = c.frame(t)
            assert g.encode(f)[0] == o.encode(f, t % 3 == 0, qp), (w, h, kind, qp, t)
        g.close()
    L = enc.lib()
    L.b200k_check_failures.restype = C.c_int; L.b200k_check_failures.argtypes = [C.c_int, C.POINTER(C.c_int)]
    site = C.c_int()
    n = L.b200k_check_failures(0, C.byref(site))
    print(f"check failures {n} first site {site.value}" if n >= 0 else "not a checked build")
