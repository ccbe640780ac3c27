// media_b200/shim/openh264_shim.cpp -> libopenh264.so: the openh264 encoder ABI (include/openh264_abi.h) on top of libb200enc.so.
//
// With this library on the loader path the reference's VideoEncoderOpenH264 works unmodified: its dlopen("libopenh264.so") /
// dlsym at video_codec/VideoEncoderOpenH264.cpp:197-226 resolve here, InitializeExt receives the SEncParamExt it fills at
// :228-296, EncodeFrame the SSourcePicture of :354-365, and the SFrameBSInfo it reads at :349-350 is laid out the way openh264
// does it: all NALs contiguous from sLayerInfo[0].pBsBuf, parameter sets in a NON_VIDEO_CODING_LAYER in front of an IDR.
#include "openh264_abi.h"
#include "b200enc.h"
#include <cstring>
#include <new>
#include <vector>

namespace {
using namespace oh264;

class B200SvcEncoder : public SvcEncoder {
public:
    ~B200SvcEncoder() override { Uninitialize(); }

    int GetDefaultParams(EncParamExt *p) override
    {
        if (!p) return 1;
        memset(p, 0, sizeof *p);
        p->usage = 0; p->rc_mode = kRcQuality; p->max_frame_rate = 30.f; p->temporal_layers = 1; p->spatial_layers = 1;
        p->complexity = 1; p->num_ref = -1; p->entropy_mode = 0; p->max_qp = 51; p->min_qp = 0; p->multiple_thread_idc = 1;
        p->frame_skip = 1; p->background_detection = 1; p->scene_change_detect = 1; p->frame_cropping = 1; p->ltr_mark_period = 30;
        for (auto &l : p->layers) { l.profile_idc = 0; l.level_idc = 0; l.dlayer_qp = 26; l.slice.mode = kSliceSingle; l.slice.num = 1; l.slice.size_constraint = 1500; }
        return 0;
    }
    int Initialize(const EncParamBase *b) override
    {
        if (!b) return 1;
        EncParamExt p; GetDefaultParams(&p);
        p.usage = b->usage; p.width = b->width; p.height = b->height; p.target_bitrate = b->target_bitrate; p.rc_mode = b->rc_mode; p.max_frame_rate = b->max_frame_rate;
        p.layers[0].width = b->width; p.layers[0].height = b->height; p.layers[0].frame_rate = b->max_frame_rate; p.layers[0].bitrate = b->target_bitrate;
        return InitializeExt(&p);
    }
    int InitializeExt(const EncParamExt *p) override
    {
        if (!p || p->spatial_layers != 1 || p->temporal_layers != 1) return 1;      // simulcast / temporal scalability are not part of this path
        Uninitialize();
        b200enc_default_config(&cfg_);
        const SpatialLayer &l = p->layers[0];
        cfg_.width = l.width > 0 ? l.width : p->width; cfg_.height = l.height > 0 ? l.height : p->height;
        const float fps = l.frame_rate > 0 ? l.frame_rate : p->max_frame_rate;
        cfg_.fps = fps >= 1.f ? (int)(fps + 0.5f) : 30;
        cfg_.bitrate = l.bitrate > 0 ? l.bitrate : p->target_bitrate;
        cfg_.gop = p->intra_period ? (int)p->intra_period : 1 << 30;               // 0 = only the first frame is intra
        cfg_.const_qp = p->rc_mode == kRcOff ? (l.dlayer_qp >= 0 && l.dlayer_qp <= 51 ? l.dlayer_qp : 26) : -1;
        if (cfg_.const_qp < 0 && cfg_.bitrate <= 0) cfg_.bitrate = 1000000;
        cfg_.num_slices = l.slice.mode == kSliceFixedNum && l.slice.num > 0 ? (int)l.slice.num : 1;       // SM_SINGLE_SLICE (the wrapper, :247)
        cfg_.scene_change = p->scene_change_detect ? 1 : 0;                       // bEnableSceneChangeDetect (the wrapper sets it, :283)
        cfg_.background_detection = p->background_detection ? 1 : 0;              // bEnableBackgroundDetection (:282)
        cfg_.complexity = p->complexity >= 0 && p->complexity <= 2 ? p->complexity : 2;       // iComplexityMode (:289)
        // rate-control bounds: iMaxBitrate (the wrapper: = target, :239-240; 0 = UNSPECIFIED_BIT_RATE = no tighter than the target), iMinQp / iMaxQp (:230)
        cfg_.max_bitrate = l.max_bitrate > 0 ? l.max_bitrate : p->max_bitrate > 0 ? p->max_bitrate : 0;
        cfg_.min_qp = p->min_qp >= 0 && p->min_qp <= 51 ? p->min_qp : 0;
        cfg_.max_qp = p->max_qp > 0 && p->max_qp <= 51 && p->max_qp >= cfg_.min_qp ? p->max_qp : 51;
        cfg_.auto_batch = 1;
        // entropy_mode = 1 (CABAC, requested by the wrapper at :291) takes effect for profile main (77) / high (100); Baseline has no CABAC
        cfg_.profile = !p->entropy_mode ? 0 : l.profile_idc == 100 ? 2 : l.profile_idc == 77 ? 1 : 0;
        if (cfg_.profile && l.slice.mode != kSliceFixedNum) cfg_.num_slices = 0;    // CABAC: the engine's automatic slice count (the coder is serial per slice)
        return create();
    }
    int Uninitialize() override { if (s_) { b200enc_destroy(s_); s_ = nullptr; } return 0; }

    int EncodeFrame(const SourcePicture *src, FrameBSInfo *out) override
    {
        if (!s_ || !src || !out) return 1;
        if (src->color_format != kVideoFormatI420 || src->width != cfg_.width || src->height != cfg_.height || !src->data[0]) return 1;
        if (dirty_) { if (create() != 0) return 1; }
        const int w = cfg_.width, h = cfg_.height;
        const uint8_t *frame = src->data[0];
        const bool packed = src->stride[0] == w && src->stride[1] == w / 2 && src->stride[2] == w / 2 &&
                            src->data[1] == src->data[0] + (size_t)w * h && src->data[2] == src->data[1] + (size_t)(w / 2) * (h / 2);
        if (!packed) {      // gather strided planes into one tightly packed frame
            staging_.resize((size_t)w * h * 3 / 2);
            uint8_t *d = staging_.data();
            for (int c = 0; c < 3; c++) {
                const int pw = c ? w / 2 : w, ph = c ? h / 2 : h;
                if (!src->data[c]) return 1;
                for (int y = 0; y < ph; y++, d += pw) memcpy(d, src->data[c] + (size_t)y * src->stride[c], pw);
            }
            frame = staging_.data();
        }
        const uint8_t *bs = nullptr; uint32_t size = 0; b200enc_frame_info info;
        if (b200enc_encode(s_, frame, (uint32_t)((size_t)w * h * 3 / 2), &bs, &size, &info) != B200ENC_OK) return 1;
        fill_info(out, bs, size, info.frame_type == B200ENC_FRAME_IDR ? kFrameIDR : kFrameP);
        out->timestamp = src->timestamp;
        return 0;
    }
    int EncodeParameterSets(FrameBSInfo *out) override
    {
        if (!s_ || !out) return 1;
        ps_.resize(256); uint32_t n = 0;
        if (b200enc_get_parameter_sets(s_, ps_.data(), (uint32_t)ps_.size(), &n) != B200ENC_OK) return 1;
        fill_info(out, ps_.data(), n, kFrameInvalid);
        return 0;
    }
    int ForceIntraFrame(bool idr, int) override { if (!idr) return 1; return s_ && b200enc_force_idr(s_) == B200ENC_OK ? 0 : 1; }
    int SetOption(int opt, void *v) override
    {
        if (!v) return 1;
        switch (opt) {
        case kOptDataFormat: return *static_cast<int *>(v) == kVideoFormatI420 ? 0 : 1;
        case kOptIdrInterval: { int g = *static_cast<int *>(v); cfg_.gop = g > 0 ? g : 1 << 30; dirty_ = true; return 0; }
        case kOptFrameRate: { float f = *static_cast<float *>(v); if (f < 1.f) return 1; cfg_.fps = (int)(f + 0.5f); dirty_ = true; return 0; }
        case kOptBitrate: case kOptMaxBitrate: { int b = static_cast<BitrateInfo *>(v)->bitrate; if (b <= 0) return 1; (opt == kOptBitrate ? cfg_.bitrate : cfg_.max_bitrate) = b; dirty_ = true; return 0; }
        case kOptRcMode: { int m = *static_cast<int *>(v); if (m == kRcOff && cfg_.const_qp < 0) cfg_.const_qp = 26; if (m != kRcOff) cfg_.const_qp = -1; dirty_ = true; return 0; }
        case kOptParamExt: return InitializeExt(static_cast<EncParamExt *>(v));
        case kOptParamBase: return Initialize(static_cast<EncParamBase *>(v));
        default: return 0;       // options without a counterpart here (trace, LTR, complexity, ...) are accepted and ignored
        }
    }
    int GetOption(int opt, void *v) override
    {
        if (!v) return 1;
        switch (opt) {
        case kOptDataFormat: *static_cast<int *>(v) = kVideoFormatI420; return 0;
        case kOptIdrInterval: *static_cast<int *>(v) = cfg_.gop; return 0;
        case kOptFrameRate: *static_cast<float *>(v) = (float)cfg_.fps; return 0;
        case kOptBitrate: static_cast<BitrateInfo *>(v)->bitrate = cfg_.bitrate; return 0;
        case kOptMaxBitrate: static_cast<BitrateInfo *>(v)->bitrate = cfg_.max_bitrate > 0 ? cfg_.max_bitrate : cfg_.bitrate; return 0;
        case kOptRcMode: *static_cast<int *>(v) = cfg_.const_qp >= 0 ? kRcOff : kRcBitrate; return 0;
        default: return 1;
        }
    }

private:
    int create()
    {
        if (s_) { b200enc_destroy(s_); s_ = nullptr; }
        dirty_ = false;
        return b200enc_create(&cfg_, &s_) == B200ENC_OK ? 0 : 1;
    }
    // split the Annex-B access unit at its 4-byte start codes and describe it the way openh264 does
    void fill_info(FrameBSInfo *out, const uint8_t *bs, uint32_t size, int frame_type)
    {
        memset(out, 0, sizeof *out);
        nal_len_.clear();
        std::vector<int> kinds;
        std::vector<uint32_t> pos;
        for (uint32_t i = 0; i + 4 < size;) {
            if (bs[i] == 0 && bs[i + 1] == 0 && bs[i + 2] == 0 && bs[i + 3] == 1) { pos.push_back(i); i += 4; } else i++;
        }
        for (size_t k = 0; k < pos.size(); k++) {
            nal_len_.push_back((int)((k + 1 < pos.size() ? pos[k + 1] : size) - pos[k]));
            kinds.push_back(bs[pos[k] + 4] & 31);
        }
        int layer = 0; size_t k = 0; uint32_t off = 0;
        while (k < kinds.size() && layer < kMaxLayers) {
            const bool vcl = kinds[k] == 1 || kinds[k] == 5;
            LayerBSInfo &l = out->layers[layer++];
            l.frame_type = vcl ? frame_type : kFrameInvalid; l.layer_type = vcl ? kLayerVcl : kLayerNonVcl;
            l.nal_length = nal_len_.data() + k; l.bs_buf = const_cast<uint8_t *>(bs) + off;
            while (k < kinds.size() && ((kinds[k] == 1 || kinds[k] == 5) == vcl) && l.nal_count < kMaxNalsPerLayer) { off += (uint32_t)nal_len_[k]; l.nal_count++; k++; }
        }
        out->layer_num = layer; out->frame_type = frame_type; out->frame_size = (int)size;
    }

    b200enc_config cfg_{};
    b200enc_session *s_ = nullptr;
    bool dirty_ = false;
    std::vector<uint8_t> staging_, ps_;
    std::vector<int> nal_len_;
};
} // namespace

extern "C" {
int WelsCreateSVCEncoder(oh264::SvcEncoder **enc)
{
    if (!enc) return 1;
    *enc = new (std::nothrow) B200SvcEncoder();
    return *enc ? 0 : 1;
}
void WelsDestroySVCEncoder(oh264::SvcEncoder *enc) { delete enc; }
}
