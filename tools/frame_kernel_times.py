"""tools/frame_kernel_times.py -- per-kernel device times of ONE session's IDR and P frames (batch of one, profiling on):
what a frame's latency is made of, per profile / slice count."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from media_b200 import enc
from media_b200.synth import Content
w, h = 1920, 1080
kind = sys.argv[1] if len(sys.argv) > 1 else "A"
c = Content(kind, w, h)
for profile, slices, qp in ((0, 1, -1), (1, 1, -1), (1, 4, -1), (1, 8, -1)):
    s = enc.Session(w, h, bitrate=4_000_000, gop=4, num_slices=slices, device=0, profile=profile, const_qp=qp)
    b = enc.Batch(0, [s]); b.set_profiling(1)
    for t in range(6):
        f = c.frame(t)
        bs, infos = b.encode([f])
        kt = b.kernel_times()
        if t in (1, 4):
            top = sorted(kt, key=lambda x: -x[1])[:6]
            print(f"profile {profile} slices {slices} frame {t} type {infos[0].frame_type} qp {infos[0].qp} bytes {len(bs[0])} total {b.kernel_ms():.2f} ms: " + ", ".join(f"{n} {m:.2f}" for n, m in top))
    b.close(); s.close()
