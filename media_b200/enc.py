"""media_b200/enc.py -- ctypes host binding of libb200enc.so (include/b200enc.h).

Mirrors the call flow of the reference wrapper (video_codec/VideoEncoderOpenH264.cpp: InitEncoder :131,
EncodeOneFrame :304, ForceKeyFrame :406, DestroyEncoder :373). There is no CPU fallback: importing works anywhere,
but creating a session without the CUDA library or without a GPU raises."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ENC_LIB") or os.path.join(_HERE, "csrc", "libb200enc.so")   # override only for A/B experiments

FMT_I420, FMT_NV12, FMT_RGBA = 0, 1, 2
STAGE = dict(mbinfo=0, mbcoef=1, me2=2, me1=3, me0=4, inter_cost=5, src=6, rec_pre=7, rec=8, mbside=9, bin_count=10, bin_off=11, bins=12)
PROFILE_BASELINE, PROFILE_MAIN, PROFILE_HIGH = 0, 1, 2
MB_BIN_SLOT = 3136

MBINFO_DTYPE = np.dtype([("mb_type", "u1"), ("i16_mode", "u1"), ("chroma_mode", "u1"), ("cbp", "u1"),
                         ("mv", "<i2", (2,)), ("i4_mode", "u1", (16,)), ("nnz", "u1", (24,))])
MBCOEF_DTYPE = np.dtype([("luma", "<i2", (16, 16)), ("luma_dc", "<i2", (16,)),
                         ("chroma_dc", "<i2", (2, 4)), ("chroma_ac", "<i2", (2, 4, 16))])
MBSIDE_DTYPE = np.dtype([("mvd", "<i2", (4, 2)), ("dc_cbf", "u1"), ("pad", "u1", (3,))])     # mvd aliases i4_syn[16] for Intra_4x4 MBs


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "fps", "bitrate", "gop", "const_qp", "num_slices",
                                       "search_range", "input_format", "device", "level_idc", "debug", "scene_change", "auto_batch", "profile",
                                       "max_bitrate", "min_qp", "max_qp", "background_detection", "complexity", "key_slices")]


class FrameInfo(C.Structure):
    _fields_ = [("frame_type", C.c_int), ("qp", C.c_int), ("size_bytes", C.c_uint32), ("frame_index", C.c_uint32)]


class B200EncError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libb200enc.so; raises if it was not built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200EncError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.b200enc_default_config.argtypes = [C.POINTER(Config)]
        L.b200enc_create.restype = C.c_int; L.b200enc_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.b200enc_destroy.argtypes = [vp]
        L.b200enc_encode.restype = C.c_int
        L.b200enc_encode.argtypes = [vp, vp, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(FrameInfo)]
        L.b200enc_force_idr.restype = C.c_int; L.b200enc_force_idr.argtypes = [vp]
        L.b200enc_get_parameter_sets.restype = C.c_int; L.b200enc_get_parameter_sets.argtypes = [vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]
        L.b200enc_device_of.restype = C.c_int; L.b200enc_device_of.argtypes = [vp]
        L.b200enc_frame_bytes.restype = C.c_size_t; L.b200enc_frame_bytes.argtypes = [vp]
        L.b200enc_last_cuda_error.restype = C.c_int
        L.b200enc_strerror.restype = C.c_char_p; L.b200enc_strerror.argtypes = [C.c_int]
        L.b200enc_device_count.restype = C.c_int
        L.b200enc_scheduler_stats.restype = C.c_int; L.b200enc_scheduler_stats.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.b200enc_batch_create.restype = C.c_int; L.b200enc_batch_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.b200enc_batch_destroy.argtypes = [vp]
        L.b200enc_batch_encode.restype = C.c_int
        L.b200enc_batch_encode.argtypes = [vp, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(FrameInfo)]
        L.b200enc_batch_last_status.restype = C.c_int; L.b200enc_batch_last_status.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
        L.b200enc_rc_retries.restype = C.c_uint32; L.b200enc_rc_retries.argtypes = [vp]
        L.b200enc_batch_last_kernel_ms.restype = C.c_float; L.b200enc_batch_last_kernel_ms.argtypes = [vp]
        L.b200enc_last_kernel_ms.restype = C.c_float; L.b200enc_last_kernel_ms.argtypes = [vp]
        L.b200enc_batch_last_launches.restype = C.c_int; L.b200enc_batch_last_launches.argtypes = [vp]
        L.b200enc_batch_set_profiling.restype = C.c_int; L.b200enc_batch_set_profiling.argtypes = [vp, C.c_int]
        L.b200enc_batch_kernel_times.restype = C.c_int
        L.b200enc_batch_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
        L.b200enc_host_alloc.restype = vp; L.b200enc_host_alloc.argtypes = [C.c_size_t]
        L.b200enc_host_free.argtypes = [vp]
        L.b200enc_dev_alloc.restype = vp; L.b200enc_dev_alloc.argtypes = [C.c_int, C.c_size_t]
        L.b200enc_dev_upload.restype = C.c_int; L.b200enc_dev_upload.argtypes = [C.c_int, vp, vp, C.c_size_t]
        L.b200enc_dev_free.argtypes = [C.c_int, vp]
        L.b200enc_get_stage.restype = C.c_int; L.b200enc_get_stage.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.b200enc_get_recon.restype = C.c_int; L.b200enc_get_recon.argtypes = [vp, vp, C.c_size_t]
        L.b200enc_slice_counts.restype = C.c_int; L.b200enc_slice_counts.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200k_convert_to_i420.restype = C.c_int
        L.b200k_convert_to_i420.argtypes = [C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200k_downsample2.restype = C.c_int; L.b200k_downsample2.argtypes = [C.c_int, vp, C.c_int, C.c_int, vp]
        L.b200k_sad16x16.restype = C.c_int; L.b200k_sad16x16.argtypes = [C.c_int, vp, vp, C.c_int, C.c_int, vp, vp]
        L.b200k_satd16x16.restype = C.c_int; L.b200k_satd16x16.argtypes = [C.c_int, vp, vp, C.c_int, C.c_int, vp, vp]
        L.b200k_transform_block.restype = C.c_int; L.b200k_transform_block.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.b200k_transform_block8.restype = C.c_int; L.b200k_transform_block8.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.b200k_deblock.restype = C.c_int; L.b200k_deblock.argtypes = [C.c_int, vp, C.c_int, C.c_int, vp, C.c_int]
        L.b200k_cabac_code.restype = C.c_int; L.b200k_cabac_code.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
        L.b200k_vabsdiff4_peak.restype = C.c_int; L.b200k_vabsdiff4_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def check(rc, what="b200enc"):
    if rc != 0:
        L = lib()
        raise B200EncError(f"{what}: {L.b200enc_strerror(rc).decode()} (code {rc}, cuda error {L.b200enc_last_cuda_error()})")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Session:
    def __init__(self, width, height, fps=30, bitrate=4_000_000, gop=30, const_qp=-1, num_slices=1, search_range=16,
                 input_format=FMT_I420, device=-1, level_idc=0, debug=0, auto_batch=0, scene_change=1, profile=PROFILE_BASELINE,
                 max_bitrate=0, min_qp=0, max_qp=51, background_detection=0, complexity=2, key_slices=0):
        L = lib()
        self.cfg = Config(width, height, fps, bitrate, gop, const_qp, num_slices, search_range, input_format, device, level_idc, debug, scene_change, auto_batch, profile,
                          max_bitrate, min_qp, max_qp, background_detection, complexity, key_slices)
        self.h = C.c_void_p()
        check(L.b200enc_create(C.byref(self.cfg), C.byref(self.h)), "b200enc_create")
        self.width, self.height = width, height
        self.mbw, self.mbh = (width + 15) // 16, (height + 15) // 16
        self.frame_bytes = L.b200enc_frame_bytes(self.h)
        self.device = L.b200enc_device_of(self.h)

    def slice_counts(self):
        """(slices of P pictures, slices of key pictures) after the automatic rules"""
        a, b = C.c_int(), C.c_int()
        check(lib().b200enc_slice_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().b200enc_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def force_idr(self):
        check(lib().b200enc_force_idr(self.h))

    def encode(self, frame):
        """frame: numpy uint8 array in HOST memory. Returns (bytes, FrameInfo)."""
        frame = np.ascontiguousarray(frame, np.uint8).ravel()
        bs, n, info = C.c_void_p(), C.c_uint32(), FrameInfo()
        check(lib().b200enc_encode(self.h, _p(frame), frame.size, C.byref(bs), C.byref(n), C.byref(info)), "b200enc_encode")
        return C.string_at(bs.value, n.value), info

    def kernel_ms(self):
        return lib().b200enc_last_kernel_ms(self.h)

    def recon(self):
        out = np.zeros(self.width * self.height * 3 // 2, np.uint8)
        check(lib().b200enc_get_recon(self.h, _p(out), out.size))
        return out

    def stage(self, name):
        n = self.mbw * self.mbh
        ny = self.mbw * self.mbh * 256
        sizes = dict(mbinfo=n * 48, mbcoef=n * 816, me2=n * 4, me1=n * 4, me0=n * 4, inter_cost=n * 4,
                     src=ny * 3 // 2, rec_pre=ny * 3 // 2, rec=ny * 3 // 2, mbside=n * 20, bin_count=n * 4, bin_off=n * 4, bins=n * MB_BIN_SLOT * 2)
        buf = np.zeros(sizes[name], np.uint8)
        wr = C.c_size_t()
        check(lib().b200enc_get_stage(self.h, STAGE[name], _p(buf), buf.size, C.byref(wr)), f"get_stage({name})")
        if name == "mbinfo":
            return buf.view(MBINFO_DTYPE)
        if name == "mbcoef":
            return buf.view(MBCOEF_DTYPE)
        if name in ("me2", "me1", "me0"):
            return buf.view(np.int16).reshape(n, 2)
        if name == "inter_cost":
            return buf.view(np.int32)
        if name == "mbside":
            return buf.view(MBSIDE_DTYPE)
        if name in ("bin_count", "bin_off"):
            return buf.view(np.uint32)
        if name == "bins":
            return buf.view(np.uint16)
        return buf


class Batch:
    """N sessions of one GPU stepping one frame per call (one chain of kernel launches)."""

    def __init__(self, device, sessions):
        L = lib()
        self.sessions = list(sessions)
        self.n = len(self.sessions)
        self.h = C.c_void_p()
        check(L.b200enc_batch_create(device, self.n, C.byref(self.h)), "b200enc_batch_create")
        self._sess = (C.c_void_p * self.n)(*[s.h for s in self.sessions])
        self._frames = (C.c_void_p * self.n)()
        self._bs = (C.c_void_p * self.n)()
        self._sizes = (C.c_uint32 * self.n)()
        self._infos = (FrameInfo * self.n)()

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().b200enc_batch_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode_ptrs(self, ptrs, device_input):
        """ptrs: sequence of integer addresses (host or device). Returns list of sizes; bitstreams stay in the sessions' buffers."""
        for i, p in enumerate(ptrs):
            self._frames[i] = p
        check(lib().b200enc_batch_encode(self.h, self._sess, self.n, self._frames, int(device_input), self._bs, self._sizes, self._infos),
              "b200enc_batch_encode")
        return self._sizes

    def encode(self, frames):
        frames = [np.ascontiguousarray(f, np.uint8).ravel() for f in frames]
        self.encode_ptrs([f.ctypes.data for f in frames], 0)
        return [C.string_at(self._bs[i], self._sizes[i]) for i in range(self.n)], [self._infos[i] for i in range(self.n)]

    def bitstream(self, i):
        return C.string_at(self._bs[i], self._sizes[i])

    def kernel_ms(self):
        return lib().b200enc_batch_last_kernel_ms(self.h)

    def launches(self):
        return lib().b200enc_batch_last_launches(self.h)

    def set_profiling(self, on):
        check(lib().b200enc_batch_set_profiling(self.h, int(on)))

    def kernel_times(self):
        names = (C.c_char_p * 32)(); ms = (C.c_float * 32)()
        k = lib().b200enc_batch_kernel_times(self.h, names, ms, 32)
        return [(names[i].decode(), ms[i]) for i in range(k)]
