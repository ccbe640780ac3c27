// Factory of libVideoCodec.so for this repository: the reference's CreateVideoEncoder (video_codec/VideoCodecApi.cpp:21-44)
// with one added selector value. Selectors 0..2 (openh264, Netint H.264/H.265) belong to the reference tree's own siblings and
// are not built here; INTEGRATION.md shows the one-case patch that adds value 3 to the reference factory.
#define LOG_TAG "VideoCodecApi"
#include "VideoCodecApi.h"
#include <new>
#include "MediaLog.h"
#include "Property.h"
#include "VideoEncoderB200.h"

namespace {
enum EncoderType : uint32_t { ENCODER_TYPE_OPENH264 = 0, ENCODER_TYPE_NETINTH264 = 1, ENCODER_TYPE_NETINTH265 = 2, ENCODER_TYPE_B200H264 = 3 };
}

EncoderRetCode CreateVideoEncoder(VideoEncoder **encoder)
{
    if (encoder == nullptr) return VIDEO_ENCODER_CREATE_FAIL;
    const uint32_t encType = (uint32_t)GetIntEncParam("ro.vmi.demo.video.encode.format");
    INFO("create video encoder: encoder type %u", encType);
    switch (encType) {
        case ENCODER_TYPE_B200H264:
            *encoder = new (std::nothrow) VideoEncoderB200();
            break;
        case ENCODER_TYPE_OPENH264:
        case ENCODER_TYPE_NETINTH264:
        case ENCODER_TYPE_NETINTH265:
            ERR("create video encoder failed: encoder type %u is provided by the reference tree, not by this build", encType);
            return VIDEO_ENCODER_CREATE_FAIL;
        default:
            ERR("create video encoder failed: unknown encoder type %u", encType);
            return VIDEO_ENCODER_CREATE_FAIL;
    }
    if (*encoder == nullptr) { ERR("create video encoder failed: encoder type %u", encType); return VIDEO_ENCODER_CREATE_FAIL; }
    return VIDEO_ENCODER_SUCCESS;
}

EncoderRetCode DestroyVideoEncoder(VideoEncoder *encoder)
{
    if (encoder == nullptr) { WARN("input encoder is null"); return VIDEO_ENCODER_SUCCESS; }
    delete encoder;
    return VIDEO_ENCODER_SUCCESS;
}
