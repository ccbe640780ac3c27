"""tests/ref_adapter_client.py -- drives oracle/_ref/libVideoCodecRef.so: the reference's VideoEncoderOpenH264 + CreateVideoEncoder compiled
UNMODIFIED from /root/reference (oracle/ref_adapter.mk), exactly as the cloud-phone caller does (properties, Create -> Init -> Start ->
EncodeOneFrame x N with a key-frame request through the property -> Stop -> Destroy). The wrapper dlopen()s "libopenh264.so"
(VideoEncoderOpenH264.cpp:46,203); this script pre-loads the library given on the command line under that SONAME -- the repo's ABI
look-alike media_b200/shim/libopenh264.so (then the GPU encodes) or cisco's real library (then it is the CPU baseline).
usage: ref_adapter_client.py <libopenh264.so> <in.i420> <w> <h> <frames> <bitrate> <gop> <profile> <force_at> <out.h264> <out.sizes>"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib, src, w, h, n, br, gop, profile, force_at, out, sizes = sys.argv[1:12]
w, h, n, force_at = int(w), int(h), int(n), int(force_at)
C.CDLL(os.path.abspath(lib), mode=C.RTLD_GLOBAL)
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libVideoCodecRef.so"))
L.vc_create.argtypes = [C.POINTER(C.c_void_p)]
for f in ("vc_init", "vc_start", "vc_stop", "vc_destroy"):
    getattr(L, f).argtypes = [C.c_void_p]
L.vc_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
L.vc_prop_set.argtypes = [C.c_char_p, C.c_char_p]
for k, v in (("ro.vmi.demo.video.encode.format", "0"), ("ro.sys.vmi.cloudphone", "video"), ("ro.hardware.width", w), ("ro.hardware.height", h),
             ("ro.hardware.fps", 30), ("persist.vmi.video.encode.bitrate", br), ("persist.vmi.video.encode.gopsize", gop),
             ("persist.vmi.video.encode.profile", profile), ("persist.vmi.video.encode.param_adjusting", "0"), ("persist.vmi.video.encode.keyframe", "0")):
    L.vc_prop_set(k.encode(), str(v).encode())
e = C.c_void_p()
assert L.vc_create(C.byref(e)) == 0, "CreateVideoEncoder"
assert L.vc_init(e) == 0, "InitEncoder"
assert L.vc_start(e) == 0, "StartEncoder"
fb = w * h * 3 // 2
data = open(src, "rb").read()
stream, lens = b"", []
p, m = C.c_void_p(), C.c_uint32()
for t in range(n):
    if t == force_at:
        L.vc_prop_set(b"persist.vmi.video.encode.keyframe", b"1")
    assert L.vc_encode(e, data[t * fb:(t + 1) * fb], fb, C.byref(p), C.byref(m)) == 0, f"EncodeOneFrame {t}"
    au = C.string_at(p.value, m.value); stream += au; lens.append(len(au))
assert L.vc_stop(e) == 0 and L.vc_destroy(e) == 0
open(out, "wb").write(stream)
open(sizes, "w").write(" ".join(str(x) for x in lens))
