"""tools/rc_sim.py -- closed-loop simulation of the rate control (media_b200/csrc/rate_control.h) on the CPU: the picture sizes come
from the oracle encoder (the CUDA path produces the same bytes for the same QP), the QPs from the very object the sessions use
(b200k_rc_* hooks of libb200enc.so, or a host-only build given by RC_LIB). Prints the QP / size trace, the achieved bitrate per
second and overall, the VBV peak and the number of pictures coded twice.
usage: rc_sim.py <w> <h> <content A|B|C|D> <bitrate> <frames> [gop] [idr_every_forced]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from media_b200.synth import Content, psnr  # noqa: E402
from oracle import orc_py  # noqa: E402


def rc_lib():
    p = os.environ.get("RC_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "media_b200", "csrc", "libb200enc.so")
    L = C.CDLL(p)
    L.b200k_rc_create.restype = C.c_void_p; L.b200k_rc_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.b200k_rc_pick.restype = C.c_int; L.b200k_rc_pick.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b200k_rc_retry_qp.restype = C.c_int; L.b200k_rc_retry_qp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.b200k_rc_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.b200k_rc_vbv.restype = C.c_double; L.b200k_rc_vbv.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.b200k_rc_destroy.argtypes = [C.c_void_p]
    return L


def simulate(w, h, kind, bitrate, frames, gop=300, fps=30, max_bitrate=0, min_qp=0, max_qp=51, force_at=(), verbose=False, content=None):
    L = rc_lib()
    rc = L.b200k_rc_create(bitrate, max_bitrate, fps, min_qp, max_qp, w, h)
    o = orc_py.Encoder(w, h)
    c = content or Content(kind, w, h)
    sizes, qps, types, retries, vbv_peak, since = [], [], [], 0, 0.0, 0
    size = C.c_double()
    for t in range(frames):
        f = c.frame(t)
        idr = t == 0 or since >= gop or t in force_at
        budget, cap = C.c_double(), C.c_double()
        qp = L.b200k_rc_pick(rc, 1 if idr else 0, C.byref(budget), C.byref(cap))
        bs = o.encode(f, idr, qp, trial=True)
        kind_coded = 1 if o.last_was_idr() else 0
        q2 = L.b200k_rc_retry_qp(rc, 1 if idr else 0, kind_coded, len(bs) * 8.0)
        if q2 >= 0:      # second attempt (engine.cu encode_impl): coarser QP, and a promoted picture is planned as the IDR it turned into
            retries += 1; qp = q2; idr = idr or bool(kind_coded)
        bs = o.encode(f, idr, qp)
        kind_coded = 1 if o.last_was_idr() else 0
        L.b200k_rc_update(rc, kind_coded, qp, len(bs) * 8.0)
        since = 1 if kind_coded else since + 1
        sizes.append(len(bs)); qps.append(qp); types.append(kind_coded)
        vbv_peak = max(vbv_peak, L.b200k_rc_vbv(rc, C.byref(size)))
    rec = o.recon()
    out = dict(rate=sum(sizes) * 8 * fps / frames, sizes=sizes, qps=qps, types=types, retries=retries, vbv_peak=vbv_peak, vbv_size=size.value,
               psnr_last=psnr(f[:w * h], rec[:w * h]))
    L.b200k_rc_destroy(rc)
    return out


if __name__ == "__main__":
    w, h, kind, br, n = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    gop = int(sys.argv[6]) if len(sys.argv) > 6 else 300
    r = simulate(w, h, kind, br, n, gop=gop)
    print("qp", r["qps"])
    print("KB", [round(x / 1024, 1) for x in r["sizes"]])
    for a in range(0, n, 30):
        print(f"frames {a}-{a + 29}: {sum(r['sizes'][a:a + 30]) * 8 / 1e6:.3f} Mbit/s")
    print(f"overall {r['rate'] / 1e6:.3f} Mbit/s ({100 * (r['rate'] / br - 1):+.1f} %), retries {r['retries']}, VBV peak {r['vbv_peak'] / r['vbv_size']:.2f} of the bucket, last-frame Y-PSNR {r['psnr_last']:.2f}")
