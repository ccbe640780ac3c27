// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY: C entry points around the reference's C++ VideoEncoder vtable
// (/root/reference/video_codec/VideoCodecApi.h:22-96) so that pytest / bench.py can drive oracle/_ref/libVideoCodecRef.so
// -- the reference adapter compiled unmodified by oracle/ref_adapter.mk -- through ctypes.
#include "VideoCodecApi.h"
#include <sys/system_properties.h>
extern "C" {
uint32_t vc_create(void **enc) { VideoEncoder *e = nullptr; uint32_t rc = CreateVideoEncoder(&e); *enc = e; return rc; }
uint32_t vc_destroy(void *enc) { return DestroyVideoEncoder(static_cast<VideoEncoder *>(enc)); }
uint32_t vc_init(void *enc) { return static_cast<VideoEncoder *>(enc)->InitEncoder(); }
uint32_t vc_start(void *enc) { return static_cast<VideoEncoder *>(enc)->StartEncoder(); }
uint32_t vc_encode(void *enc, const uint8_t *in, uint32_t size, uint8_t **out, uint32_t *out_size) { return static_cast<VideoEncoder *>(enc)->EncodeOneFrame(in, size, out, out_size); }
uint32_t vc_stop(void *enc) { return static_cast<VideoEncoder *>(enc)->StopEncoder(); }
void vc_destroy_encoder(void *enc) { static_cast<VideoEncoder *>(enc)->DestroyEncoder(); }
int vc_prop_set(const char *k, const char *v) { return __system_property_set(k, v); }
int vc_prop_get(const char *k, char *v) { return __system_property_get(k, v); }
}
