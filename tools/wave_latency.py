"""tools/wave_latency.py -- per-MB latency of the wavefront kernels: one session of width 1920 and 1..N macroblock rows,
per-kernel CUDA-event times of the IDR and of a P frame (run on a GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from media_b200 import enc
from media_b200.synth import Content

for h in (16, 32, 48, 64, 128, 256):
    w = 1920
    s = enc.Session(w, h, const_qp=30, gop=1000, device=0)
    b = enc.Batch(0, [s]); b.set_profiling(True)
    c = Content("A", w, h)
    for t in range(4):
        b.encode([c.frame(t)])
        kt = dict(b.kernel_times())
        if t in (0, 3):
            print(f"rows={h // 16} frame {t}: intra_wave {kt.get('k_intra_wave', 0):.3f} ms  deblock_wave {kt.get('k_deblock_wave', 0):.3f} ms  "
                  f"-> per MB-step of the critical path ({120 + 2 * (h // 16 - 1)} steps): intra {kt.get('k_intra_wave', 0) * 1e3 / (120 + 2 * (h // 16 - 1)):.2f} us, "
                  f"deblock {kt.get('k_deblock_wave', 0) * 1e3 / (120 + 2 * (h // 16 - 1)):.2f} us")
    s.close(); b.close()
