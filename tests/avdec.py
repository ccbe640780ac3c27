"""Independent H.264 decoder for the round-trip tests: FFmpeg's native `h264` decoder from the libavcodec
bundled in the opencv wheel, driven through ctypes. It pins every normative stage of the
encode path: decode(stream) must equal the encoder's own reconstruction bit for bit."""
import ctypes as C
import glob
import os
import numpy as np

_state = {}


def _load():
    if _state:
        return _state
    import cv2  # noqa: F401  (resolves the wheel's bundled dependencies first)
    d = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
    avu = C.CDLL(glob.glob(os.path.join(d, "libavutil-*"))[0], C.RTLD_GLOBAL)
    avc = C.CDLL(glob.glob(os.path.join(d, "libavcodec-*"))[0], C.RTLD_GLOBAL)
    avc.avcodec_version.restype = C.c_uint
    ver = avc.avcodec_version()
    assert (ver >> 16, (ver >> 8) & 255) == (62, 11), "struct offsets below are for libavcodec 62.11"
    vp = C.c_void_p
    avc.avcodec_find_decoder.restype = vp; avc.avcodec_find_decoder.argtypes = [C.c_int]
    avc.avcodec_alloc_context3.restype = vp; avc.avcodec_alloc_context3.argtypes = [vp]
    avc.avcodec_open2.restype = C.c_int; avc.avcodec_open2.argtypes = [vp, vp, vp]
    avc.av_packet_alloc.restype = vp
    avc.av_packet_free.argtypes = [vp]
    avc.avcodec_send_packet.restype = C.c_int; avc.avcodec_send_packet.argtypes = [vp, vp]
    avc.avcodec_receive_frame.restype = C.c_int; avc.avcodec_receive_frame.argtypes = [vp, vp]
    avc.avcodec_free_context.argtypes = [vp]
    avu.av_frame_alloc.restype = vp
    avu.av_frame_free.argtypes = [vp]
    avu.av_frame_unref.argtypes = [vp]
    _state.update(avu=avu, avc=avc)
    return _state


def available():
    try:
        _load()
        return True
    except Exception:
        return False


class H264Decoder:
    def __init__(self):
        s = _load(); self.avc, self.avu = s["avc"], s["avu"]
        dec = self.avc.avcodec_find_decoder(27)
        assert dec, "h264 decoder missing"
        self.ctx = self.avc.avcodec_alloc_context3(dec)
        assert self.avc.avcodec_open2(self.ctx, dec, None) == 0
        self.pkt = self.avc.av_packet_alloc()
        self.frm = self.avu.av_frame_alloc()

    def _receive(self):
        out = []
        while self.avc.avcodec_receive_frame(self.ctx, self.frm) == 0:
            f = self.frm
            w = C.c_int.from_address(f + 104).value; h = C.c_int.from_address(f + 108).value
            fmt = C.c_int.from_address(f + 116).value
            assert fmt in (0, 12), f"unexpected pixel format {fmt}"   # yuv420p / yuvj420p
            planes = []
            for i in range(3):
                ptr = C.c_void_p.from_address(f + 8 * i).value
                ls = C.c_int.from_address(f + 64 + 4 * i).value
                pw, ph = (w, h) if i == 0 else (w // 2, h // 2)
                buf = (C.c_uint8 * (ls * ph)).from_address(ptr)
                planes.append(np.frombuffer(buf, np.uint8).reshape(ph, ls)[:, :pw].copy().ravel())
            out.append(np.concatenate(planes))
            self.avu.av_frame_unref(self.frm)
        return out

    def decode(self, au: bytes):
        """Feed one access unit; returns the list of I420 frames that came out."""
        buf = C.create_string_buffer(au + b"\0" * 64, len(au) + 64)
        C.c_void_p.from_address(self.pkt + 24).value = C.addressof(buf)
        C.c_int.from_address(self.pkt + 32).value = len(au)
        r = self.avc.avcodec_send_packet(self.ctx, self.pkt)
        assert r == 0, f"avcodec_send_packet -> {r}"
        return self._receive()

    def flush(self):
        self.avc.avcodec_send_packet(self.ctx, None)
        return self._receive()


def decode_stream(aus):
    d = H264Decoder(); frames = []
    for au in aus:
        frames += d.decode(au)
    frames += d.flush()
    return frames
