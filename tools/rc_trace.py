"""tools/rc_trace.py -- rate-control trace of one CBR session (run on a GPU box): QP, bytes and running bitrate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from media_b200 import enc
from media_b200.synth import Content, psnr
w, h, kind, br, n = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
s = enc.Session(w, h, fps=30, bitrate=br, gop=300, const_qp=-1, device=0)
c = Content(kind, w, h)
sizes, qps = [], []
for t in range(n):
    f = c.frame(t)
    bs, info = s.encode(f); sizes.append(len(bs)); qps.append(info.qp)
rec = s.recon()
print("qp", qps)
print("KB", [round(x / 1024, 1) for x in sizes])
for a in range(0, n, 30):
    print(f"frames {a}-{a+29}: {sum(sizes[a:a+30]) * 8 / 1e6:.2f} Mbit/s")
print("overall", sum(sizes) * 8 * 30 / n / 1e6, "Mbit/s; last-frame Y-PSNR", round(psnr(f[:w*h], rec[:w*h]), 2))
