"""tools/dbk_timing.py -- per-phase cycle counts of the deblocking row loop (a -DDBK_TIMING build of the library, selected with B200ENC_LIB):
one 1080p session, a few P pictures; row 0 of session 0 prints its averages from the device."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from media_b200 import enc
from media_b200.synth import Content
w, h = 1920, 1080
s = enc.Session(w, h, const_qp=int(sys.argv[1]) if len(sys.argv) > 1 else 34, gop=1000, device=0)
c = Content("A", w, h)
for t in range(4):
    s.encode(c.frame(t))
    print("frame", t, "kernel ms", round(s.kernel_ms(), 3), flush=True)
s.close()
