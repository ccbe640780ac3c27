"""media_b200: a B200-native H.264 Baseline encoder behind kunpengcompute/media's VideoCodecApi encoder surface.

The product is the C-ABI library media_b200/csrc -> libb200enc.so (include/b200enc.h) and the C++ sibling of
VideoEncoderOpenH264 (libVideoCodec.so, include/VideoCodecApi.h). This package holds the thin ctypes host
binding used by bench.py and tests; it fails loudly when the CUDA library is missing (no CPU fallback)."""
