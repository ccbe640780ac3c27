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

# High profile: the 8x8 transform on inter macroblocks against the same stream without it (transform_8x8_mode_flag = 0)
print()
print("| content | QP | High profile | P-frame kbit/frame | Y-PSNR dB (P frames) | inter MBs with transform_size_8x8_flag (last frame) |")
print("|---|---|---|---|---|---|")
for kind in "AB":
    for qp in (24, 30, 36, 42):
        for name, kw in (("4x4 only", dict(no_t8x8=1)), ("4x4 / 8x8", {})):
            e = orc_py.Encoder(W, H, profile=2, **kw); c = Content(kind, W, H)
            bits, ps = [], []
            for t in range(FRAMES):
                f = c.frame(t); au = e.encode(f, t == 0, qp)
                bits.append(len(au) * 8); ps.append(psnr(f[:W * H], e.recon()[:W * H]))
            mi = e.mb_info(); inter = (mi["mb_type"] == 0) | (mi["mb_type"] == 4)
            print(f"| {kind} | {qp} | {name} | {np.mean(bits[1:]) / 1000:.1f} | {np.mean(ps[1:]):.2f} | {int(((mi['i16_mode'] >> 2) & 1).sum())} of {int(inter.sum())} |")

# High profile groundwork (oracle only so far): Intra_8x8 on key frames
print()
print("| content | QP | High profile key frame | IDR KB | Y-PSNR dB (IDR) | MBs Intra_16x16 / Intra_4x4 / Intra_8x8 |")
print("|---|---|---|---|---|---|")
for kind in "AB":
    for qp in (24, 30, 36, 42):
        for name, kw in (("I16x16 / I4x4", {}), ("+ Intra_8x8 (oracle only)", dict(intra8x8=1))):
            e = orc_py.Encoder(W, H, profile=2, **kw); c = Content(kind, W, H)
            f = c.frame(0); au = e.encode(f, True, qp); ty = e.mb_info()["mb_type"]
            print(f"| {kind} | {qp} | {name} | {len(au) / 1024:.1f} | {psnr(f[:W * H], e.recon()[:W * H]):.2f} | {int((ty == 1).sum())} / {int((ty == 2).sum())} / {int((ty == 5).sum())} |")
