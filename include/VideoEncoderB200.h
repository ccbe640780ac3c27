/*
 * include/VideoEncoderB200.h -- B200 sibling of VideoEncoderOpenH264 behind the VideoCodecApi encoder surface.
 *
 * Same property keys, validation ranges, defaults, return codes and reset / force-key-frame behaviour as
 * video_codec/VideoEncoderOpenH264.{h,cpp} (:62-122 parameters, :131-157 init, :304-352 per-frame flow,
 * :388-429 reset / key frame / parameter change); the codec underneath is libb200enc.so instead of the
 * dlopen'ed libopenh264.so (:197-226). Selected by ro.vmi.demo.video.encode.format = 3.
 */
#ifndef VIDEO_ENCODER_B200_H
#define VIDEO_ENCODER_B200_H
#include <atomic>
#include <string>
#include "VideoCodecApi.h"
#include "b200enc.h"

class VideoEncoderB200 : public VideoEncoder {
public:
    VideoEncoderB200();
    ~VideoEncoderB200() override;
    EncoderRetCode InitEncoder() override;
    EncoderRetCode StartEncoder() override;
    EncoderRetCode EncodeOneFrame(const uint8_t *inputData, uint32_t inputSize, uint8_t **outputData, uint32_t *outputSize) override;
    EncoderRetCode StopEncoder() override;
    void DestroyEncoder() override;
    EncoderRetCode ResetEncoder() override;
    EncoderRetCode ForceKeyFrame();
    EncoderRetCode SetEncodeParams();
    bool EncodeParamsChange();

private:
    struct EncodeParams {
        uint32_t framerate = 30, bitrate = 5000000, gopsize = 30;
        std::string profile = "baseline";
        uint32_t width = 720, height = 1280;
    };
    bool GetRoEncParam();
    bool GetPersistEncParam();
    bool VerifyEncodeRoParams(int32_t width, int32_t height, int32_t framerate);
    bool VerifyEncodeParams(std::string &bitrate, std::string &gopsize, std::string &profile);
    void Release();

    EncodeParams m_encParams, m_tmpEncParams;
    b200enc_session *m_session = nullptr;
    std::atomic<bool> m_resetFlag{ false };
    uint32_t m_frameSize = 0;
};
#endif
