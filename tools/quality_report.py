"""tools/quality_report.py -- rate / distortion table of the encoder specification (CPU oracle, bit-exact with the CUDA path) on the
synthetic contents, with the mode-decision tools switched off one at a time (oracle-only flags). Writes markdown to stdout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import orc_py
from media_b200.synth import Content, psnr

W, H, FRAMES = 640, 368, 10
print(f"| content | QP | tools | IDR KB | P-frame kbit/frame | Y-PSNR dB | MB types of the last frame: P16x16 / I16x16 / I4x4 / skip / P8x8 |")
print("|---|---|---|---|---|---|---|")
for kind in "AB":
    for qp in (24, 30, 36):
        for name, kw in (("all", {}), ("no Intra_4x4", dict(no_i4x4=1)), ("no P_8x8", dict(no_p8x8=1))):
            e = orc_py.Encoder(W, H, **kw); c = Content(kind, W, H)
            bits, ps = [], []
            for t in range(FRAMES):
                f = c.frame(t); au = e.encode(f, t == 0, qp)
                bits.append(len(au) * 8); ps.append(psnr(f[:W * H], e.recon()[:W * H]))
            types = np.bincount(e.mb_info()["mb_type"], minlength=5).tolist()
            print(f"| {kind} | {qp} | {name} | {bits[0] / 8192:.1f} | {np.mean(bits[1:]) / 1000:.1f} | {np.mean(ps):.2f} | {' / '.join(map(str, types))} |")
