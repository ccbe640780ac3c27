"""oracle/orc_py.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/orc.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The oracle restates the work behind ISVCEncoder::EncodeFrame
(/root/reference/video_codec/VideoEncoderOpenH264.cpp:344); see orc.h for what is and is not pinned.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liborc.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


class OrcConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("num_slices", C.c_int),
                ("search_range", C.c_int), ("level_idc", C.c_int), ("fps", C.c_int), ("no_i4x4", C.c_int), ("no_p8x8", C.c_int), ("no_scene_change", C.c_int), ("profile", C.c_int), ("no_t8x8", C.c_int),
                ("background_detection", C.c_int), ("complexity_set", C.c_int), ("complexity", C.c_int), ("intra8x8", C.c_int), ("key_slices", C.c_int)]


MBINFO_DTYPE = np.dtype([("mb_type", "u1"), ("i16_mode", "u1"), ("chroma_mode", "u1"), ("cbp", "u1"),
                         ("mv", "<i2", (2,)), ("i4_mode", "u1", (16,)), ("nnz", "u1", (24,))])
MBCOEF_DTYPE = np.dtype([("luma", "<i2", (16, 16)), ("luma_dc", "<i2", (16,)),
                         ("chroma_dc", "<i2", (2, 4)), ("chroma_ac", "<i2", (2, 4, 16))])
MBSIDE_DTYPE = np.dtype([("mvd", "<i2", (4, 2)), ("dc_cbf", "u1"), ("pad", "u1", (3,))])     # mvd aliases i4_syn[16] for Intra_4x4 MBs
assert MBINFO_DTYPE.itemsize == 48 and MBCOEF_DTYPE.itemsize == 816 and MBSIDE_DTYPE.itemsize == 20

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, u8p, i16p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int32)
        L.orc_create.restype = vp; L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [vp]
        L.orc_encode.restype = C.c_int; L.orc_encode.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int]
        L.orc_encode_trial.restype = C.c_int; L.orc_encode_trial.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int]
        L.orc_get_recon.argtypes = [vp, vp]
        L.orc_last_frame_was_idr.restype = C.c_int; L.orc_last_frame_was_idr.argtypes = [vp]
        L.orc_mb_info.restype = vp; L.orc_mb_info.argtypes = [vp]
        L.orc_mb_coef.restype = vp; L.orc_mb_coef.argtypes = [vp]
        L.orc_mb_count.restype = C.c_int; L.orc_mb_count.argtypes = [vp]
        L.orc_plane.restype = vp; L.orc_plane.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_me_level.restype = vp; L.orc_me_level.argtypes = [vp, C.c_int]
        L.orc_inter_cost.restype = vp; L.orc_inter_cost.argtypes = [vp]
        L.orc_dbg_qpel.restype = C.c_int; L.orc_dbg_qpel.argtypes = [vp, C.c_int, C.c_int]
        L.orc_dbg_build_halfpel.argtypes = [vp]
        L.orc_write_sps.restype = C.c_int; L.orc_write_sps.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_write_pps.restype = C.c_int; L.orc_write_pps.argtypes = [vp, C.c_int, C.c_int]
        L.orc_mb_side.restype = vp; L.orc_mb_side.argtypes = [vp]
        L.orc_slice_bins.restype = C.c_int; L.orc_slice_bins.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.orc_cabac_code_bins.restype = C.c_int; L.orc_cabac_code_bins.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_level_for.restype = C.c_int; L.orc_level_for.argtypes = [C.c_int] * 3
        L.orc_sad.restype = C.c_int; L.orc_sad.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
        L.orc_satd4x4.restype = C.c_int; L.orc_satd4x4.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_satd16x16.restype = C.c_int; L.orc_satd16x16.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_dct4x4.argtypes = [vp, vp]; L.orc_idct4x4.argtypes = [vp, vp]
        L.orc_quant4x4.restype = C.c_int; L.orc_quant4x4.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.orc_dequant4x4.argtypes = [vp, vp, C.c_int, C.c_int]
        L.orc_rgba_to_i420.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_nv12_to_i420.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_downsample2.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_interp_luma.restype = C.c_int; L.orc_interp_luma.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_interp_chroma.restype = C.c_int; L.orc_interp_chroma.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_deblock_frame.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_pred_i8.restype = C.c_int; L.orc_pred_i8.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.orc_escape_rbsp.restype = C.c_int; L.orc_escape_rbsp.argtypes = [vp, C.c_int, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Encoder:
    """One oracle session: encode(i420, idr, qp) -> Annex-B bytes; stage dumps as numpy arrays."""

    def __init__(self, width, height, num_slices=1, search_range=16, fps=30, level_idc=0, no_i4x4=0, no_p8x8=0, scene_change=1, profile=0, no_t8x8=0, intra8x8=1,
                 background_detection=0, complexity=None, key_slices=0):
        self.L = lib()
        self.cfg = OrcConfig(width, height, num_slices, search_range, level_idc, fps, no_i4x4, no_p8x8, 0 if scene_change else 1, profile, no_t8x8,
                             background_detection, 0 if complexity is None else 1, complexity or 0, intra8x8, key_slices)
        self.h = self.L.orc_create(C.byref(self.cfg))
        self.width, self.height = width, height
        self.mbw, self.mbh = (width + 15) // 16, (height + 15) // 16
        self._out = np.zeros(self.mbw * self.mbh * 1024 + 65536, np.uint8)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def encode(self, i420, idr, qp, trial=False):
        """trial=True codes the picture without advancing the stream state (a rate-control attempt that is thrown away)"""
        i420 = np.ascontiguousarray(i420, np.uint8).ravel()
        assert i420.size >= self.width * self.height * 3 // 2
        n = (self.L.orc_encode_trial if trial else self.L.orc_encode)(self.h, _p(i420), 1 if idr else 0, int(qp), _p(self._out), self._out.size)
        if n < 0:
            raise RuntimeError("orc_encode failed")
        return self._out[:n].tobytes()

    def last_was_idr(self):
        return bool(self.L.orc_last_frame_was_idr(self.h))

    def recon(self):
        out = np.zeros(self.width * self.height * 3 // 2, np.uint8)
        self.L.orc_get_recon(self.h, _p(out))
        return out

    def mb_info(self):
        n = self.L.orc_mb_count(self.h)
        buf = (C.c_uint8 * (n * 48)).from_address(self.L.orc_mb_info(self.h))
        return np.frombuffer(buf, MBINFO_DTYPE).copy()

    def mb_coef(self):
        n = self.L.orc_mb_count(self.h)
        buf = (C.c_uint8 * (n * 816)).from_address(self.L.orc_mb_coef(self.h))
        return np.frombuffer(buf, MBCOEF_DTYPE).copy()

    def mb_side(self):
        n = self.L.orc_mb_count(self.h)
        buf = (C.c_uint8 * (n * 20)).from_address(self.L.orc_mb_side(self.h))
        return np.frombuffer(buf, MBSIDE_DTYPE).copy()

    def slice_bins(self, s):
        p = C.c_void_p()
        n = self.L.orc_slice_bins(self.h, s, C.byref(p))
        return np.frombuffer((C.c_uint16 * n).from_address(p.value), np.uint16).copy() if n else np.zeros(0, np.uint16)

    def plane(self, which, comp):
        st, w, h = C.c_int(), C.c_int(), C.c_int()
        p = self.L.orc_plane(self.h, which, comp, C.byref(st), C.byref(w), C.byref(h))
        buf = (C.c_uint8 * (st.value * h.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h.value, st.value)[:, :w.value].copy()

    def me_level(self, level):
        n = self.L.orc_mb_count(self.h)
        buf = (C.c_int16 * (n * 2)).from_address(self.L.orc_me_level(self.h, level))
        return np.frombuffer(buf, np.int16).reshape(n, 2).copy()

    def inter_cost(self):
        n = self.L.orc_mb_count(self.h)
        buf = (C.c_int32 * n).from_address(self.L.orc_inter_cost(self.h))
        return np.frombuffer(buf, np.int32).copy()
