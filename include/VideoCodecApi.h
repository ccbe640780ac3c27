/*
 * include/VideoCodecApi.h -- the encoder surface of kunpengcompute/media, re-declared for this repository
 * (the GPU box receives only this tree, so the reference header cannot be included from /root/reference).
 *
 * Binary contract kept identical to video_codec/VideoCodecApi.h:8-20 (EncoderRetCode values), :22-78 (virtual
 * function order of VideoEncoder: dtor, InitEncoder, StartEncoder, EncodeOneFrame, StopEncoder, DestroyEncoder,
 * ResetEncoder) and :80-96 (the two extern "C" factory symbols), so a caller compiled against the reference
 * header can dlopen this libVideoCodec.so unchanged.
 */
#ifndef VIDEO_CODEC_API_H
#define VIDEO_CODEC_API_H
#include <cstdint>

enum EncoderRetCode : uint32_t {
    VIDEO_ENCODER_SUCCESS                = 0x00,
    VIDEO_ENCODER_CREATE_FAIL            = 0x01,
    VIDEO_ENCODER_INIT_FAIL              = 0x02,
    VIDEO_ENCODER_START_FAIL             = 0x03,
    VIDEO_ENCODER_ENCODE_FAIL            = 0x04,
    VIDEO_ENCODER_STOP_FAIL              = 0x05,
    VIDEO_ENCODER_DESTROY_FAIL           = 0x06,
    VIDEO_ENCODER_REGISTER_FAIL          = 0x07,
    VIDEO_ENCODER_RESET_FAIL             = 0x08,
    VIDEO_ENCODER_FORCE_KEY_FRAME_FAIL   = 0x09,
    VIDEO_ENCODER_SET_ENCODE_PARAMS_FAIL = 0x0A
};

class VideoEncoder {
public:
    VideoEncoder() = default;
    virtual ~VideoEncoder() = default;
    virtual EncoderRetCode InitEncoder() = 0;
    virtual EncoderRetCode StartEncoder() = 0;
    /* one tightly packed frame in (I420 unless the b200 input-format property says otherwise), one Annex-B access
     * unit out; *outputData stays owned by the encoder and valid until the next call on the same object */
    virtual EncoderRetCode EncodeOneFrame(const uint8_t *inputData, uint32_t inputSize,
        uint8_t **outputData, uint32_t *outputSize) = 0;
    virtual EncoderRetCode StopEncoder() = 0;
    virtual void DestroyEncoder() = 0;
    virtual EncoderRetCode ResetEncoder() = 0;
};

extern "C" {
EncoderRetCode CreateVideoEncoder(VideoEncoder** encoder);
EncoderRetCode DestroyVideoEncoder(VideoEncoder* encoder);
}

#endif  // VIDEO_CODEC_API_H
