import os, sys
sys.path.insert(0, '/tmp/orcx')
import numpy as np
from oracle import orc_py
from media_b200.synth import Content, psnr
import avdec_x as avdec
W, H, FR = 640, 368, 8
mode = sys.argv[1] if len(sys.argv) > 1 else 'rd'
def run(kind, qp, part, check=False):
    if part: os.environ['ORC_PART'] = '1'
    else: os.environ.pop('ORC_PART', None)
    e = orc_py.Encoder(W, H); c = Content(kind, W, H)
    bits, ps, aus, recs = [], [], [], []
    for t in range(FR):
        f = c.frame(t); au = e.encode(f, t == 0, qp); aus.append(au)
        bits.append(len(au) * 8); ps.append(psnr(f[:W * H], e.recon()[:W * H])); recs.append(e.recon().copy())
    ty = np.bincount(e.mb_info()["mb_type"], minlength=8).tolist()
    if check:
        dec = avdec.decode_stream(aus)
        assert len(dec) == FR, len(dec)
        for t in range(FR):
            assert np.array_equal(np.asarray(dec[t]).ravel()[:recs[t].size], recs[t].ravel()), f"decode mismatch frame {t}"
    return np.mean(bits[1:]) / 1000, np.mean(ps[1:]), ty
if mode == 'check':
    for kind in "ABE":
        for qp in (26, 38):
            r = run(kind, qp, 1, True); print(kind, qp, r, "decoder ok")
else:
    for kind in "ABE":
        for qp in (26, 32, 38):
            a = run(kind, qp, 0); b = run(kind, qp, 1)
            print(f"| {kind} | {qp} | {a[0]:.2f} kbit, {a[1]:.2f} dB, {a[2][4]} P_8x8 | {b[0]:.2f} kbit, {b[1]:.2f} dB, {b[2][6]} 16x8 / {b[2][7]} 8x16 / {b[2][4]} P_8x8 | {100*(b[0]/a[0]-1):+.2f} % bits, {b[1]-a[1]:+.3f} dB |")
