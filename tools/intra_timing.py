"""tools/intra_timing.py -- phase cycle counts of the intra macroblock (needs a -DINTRA_TIMING build via B200ENC_LIB)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from media_b200 import enc
from media_b200.synth import Content
L = enc.lib()
w, h = 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 16
s = enc.Session(w, h, const_qp=30, gop=1000, device=0)
c = Content("A", w, h)
out = (C.c_longlong * 8)()
L.b200k_intra_timing(out, 1)
s.encode(c.frame(0))
L.b200k_intra_timing(out, 1)
n = max(1, out[7])
print(f"IDR, {n} MBs: cycles/MB  neighbours+params {out[0] // n}  I16 decision {out[1] // n}  I4x4 trial {out[2] // n}  coding {out[3] // n}   | per 4x4 block of the trial: edge filter {out[4] // n // 16}  mode evaluation {out[5] // n // 16}  transform+recon {out[6] // n // 16}")
s.close()
