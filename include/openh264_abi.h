/*
 * include/openh264_abi.h -- binary layout of the part of the openh264 encoder API that kunpengcompute/media uses,
 * declared from scratch as offset-checked plain structs (x86-64 / AArch64 LP64).
 *
 * The reference's VideoEncoderOpenH264 dlopen()s "libopenh264.so", resolves WelsCreateSVCEncoder / WelsDestroySVCEncoder
 * (video_codec/VideoEncoderOpenH264.cpp:197-226) and talks to the encoder through the ISVCEncoder vtable
 * (vendor/openh264/codec_api.h:269-339) with SEncParamExt / SSourcePicture / SFrameBSInfo
 * (vendor/openh264/codec_app_def.h:540-593, 621-648, 653-660). media_b200/shim/libopenh264.so implements exactly that
 * ABI on top of libb200enc.so, so the UNMODIFIED reference wrapper drives the GPU encoder.
 * tests/test_host_cpu.py compiles a layout check against the reference's vendored headers where they are available.
 */
#ifndef OPENH264_ABI_H
#define OPENH264_ABI_H
#include <stddef.h>
#include <stdint.h>

namespace oh264 {

enum { kVideoFormatI420 = 23 };
enum { kFrameInvalid = 0, kFrameIDR = 1, kFrameI = 2, kFrameP = 3, kFrameSkip = 4 };
enum { kLayerNonVcl = 0, kLayerVcl = 1 };
enum { kRcQuality = 0, kRcBitrate = 1, kRcBufferBased = 2, kRcTimestamp = 3, kRcOff = -1 };
enum { kSliceSingle = 0, kSliceFixedNum = 1 };
enum { kOptDataFormat = 0, kOptIdrInterval = 1, kOptParamBase = 2, kOptParamExt = 3, kOptFrameRate = 4, kOptBitrate = 5, kOptMaxBitrate = 6, kOptRcMode = 8 };
enum { kMaxLayers = 128, kMaxNalsPerLayer = 128 };

struct SliceArgument { uint32_t mode, num; uint32_t mb_num[35]; uint32_t size_constraint; };            /* 152 bytes */
struct SpatialLayer {                                                                                     /* 200 bytes */
    int32_t width, height; float frame_rate; int32_t bitrate, max_bitrate; int32_t profile_idc, level_idc, dlayer_qp;
    SliceArgument slice; uint8_t vui_and_rest[16];
};
struct EncParamBase { int32_t usage, width, height, target_bitrate, rc_mode; float max_frame_rate; };    /* 24 bytes */
struct EncParamExt {                                                                                      /* 916 bytes */
    int32_t usage, width, height, target_bitrate, rc_mode; float max_frame_rate; int32_t temporal_layers, spatial_layers;
    SpatialLayer layers[4];
    int32_t complexity; uint32_t intra_period; int32_t num_ref, sps_pps_id_strategy;
    uint8_t prefix_nal, ssei, simulcast_avc, pad0; int32_t padding_flag, entropy_mode;
    uint8_t frame_skip, pad1[3]; int32_t max_bitrate, max_qp, min_qp; uint32_t max_nal_size;
    uint8_t ltr, pad2[3]; int32_t ltr_ref_num; uint32_t ltr_mark_period;
    uint16_t multiple_thread_idc; uint8_t load_balancing, pad3;
    int32_t loop_filter_disable_idc, loop_filter_alpha, loop_filter_beta;
    uint8_t denoise, background_detection, adaptive_quant, frame_cropping, scene_change_detect, lossless_link, pad4[2];
};
struct SourcePicture { int32_t color_format; int32_t stride[4]; uint8_t *data[4]; int32_t width, height; long long timestamp; };   /* 72 bytes */
struct LayerBSInfo {                                                                                      /* 40 bytes */
    uint8_t temporal_id, spatial_id, quality_id, pad0; int32_t frame_type; uint8_t layer_type, pad1[3];
    int32_t sub_seq_id, nal_count; int32_t pad2; int32_t *nal_length; uint8_t *bs_buf;
};
struct FrameBSInfo { int32_t layer_num; int32_t pad0; LayerBSInfo layers[kMaxLayers]; int32_t frame_type, frame_size; long long timestamp; };   /* 5144 bytes */
struct BitrateInfo { int32_t layer, bitrate; };

static_assert(sizeof(SliceArgument) == 152 && sizeof(SpatialLayer) == 200 && sizeof(EncParamBase) == 24, "openh264 ABI layout");
static_assert(sizeof(EncParamExt) == 916 && offsetof(EncParamExt, layers) == 32 && offsetof(EncParamExt, complexity) == 832, "openh264 ABI layout");
static_assert(offsetof(EncParamExt, intra_period) == 836 && offsetof(EncParamExt, entropy_mode) == 856 && offsetof(EncParamExt, max_bitrate) == 864, "openh264 ABI layout");
static_assert(offsetof(EncParamExt, max_nal_size) == 876 && offsetof(EncParamExt, multiple_thread_idc) == 892 && offsetof(EncParamExt, loop_filter_disable_idc) == 896, "openh264 ABI layout");
static_assert(offsetof(EncParamExt, scene_change_detect) == 912 && offsetof(SpatialLayer, dlayer_qp) == 28 && offsetof(SpatialLayer, slice) == 32, "openh264 ABI layout");
static_assert(sizeof(SourcePicture) == 72 && offsetof(SourcePicture, data) == 24 && offsetof(SourcePicture, width) == 56, "openh264 ABI layout");
static_assert(sizeof(LayerBSInfo) == 40 && offsetof(LayerBSInfo, frame_type) == 4 && offsetof(LayerBSInfo, layer_type) == 8 && offsetof(LayerBSInfo, nal_count) == 16, "openh264 ABI layout");
static_assert(offsetof(LayerBSInfo, nal_length) == 24 && offsetof(LayerBSInfo, bs_buf) == 32, "openh264 ABI layout");
static_assert(sizeof(FrameBSInfo) == 5144 && offsetof(FrameBSInfo, layers) == 8 && offsetof(FrameBSInfo, frame_type) == 5128 && offsetof(FrameBSInfo, frame_size) == 5132, "openh264 ABI layout");

/* virtual function order of ISVCEncoder (vendor/openh264/codec_api.h:276-338); the destructor comes last */
class SvcEncoder {
public:
    virtual int Initialize(const EncParamBase *p) = 0;
    virtual int InitializeExt(const EncParamExt *p) = 0;
    virtual int GetDefaultParams(EncParamExt *p) = 0;
    virtual int Uninitialize() = 0;
    virtual int EncodeFrame(const SourcePicture *src, FrameBSInfo *out) = 0;
    virtual int EncodeParameterSets(FrameBSInfo *out) = 0;
    virtual int ForceIntraFrame(bool idr, int layer_id = -1) = 0;
    virtual int SetOption(int option, void *value) = 0;
    virtual int GetOption(int option, void *value) = 0;
    virtual ~SvcEncoder() {}
};

} // namespace oh264

extern "C" {
int WelsCreateSVCEncoder(oh264::SvcEncoder **enc);       /* vendor/openh264/codec_api.h:545 */
void WelsDestroySVCEncoder(oh264::SvcEncoder *enc);      /* vendor/openh264/codec_api.h:552 */
}
#endif
