"""tools/wave_one_row.py -- one session of 1920x16 (a single macroblock row, no wavefront waits): the workload for an ncu
source-level profile of the per-macroblock latency of k_intra_wave / k_deblock_wave."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from media_b200 import enc
from media_b200.synth import Content
w, h = 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 16
s = enc.Session(w, h, const_qp=30, gop=1000, device=0)
c = Content("A", w, h)
for t in range(4):
    s.encode(c.frame(t))
s.close()
print("ok")
