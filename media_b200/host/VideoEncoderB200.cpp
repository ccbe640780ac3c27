// B200 sibling of VideoEncoderOpenH264 (see include/VideoEncoderB200.h for the reference lines mirrored here).
#define LOG_TAG "VideoEncoderB200"
#include "VideoEncoderB200.h"
#include <cstring>
#include "MediaLog.h"
#include "Property.h"

namespace {
constexpr int32_t WH_MIN = 16, WH_MAX = 4096, FRAMERATE_MIN = 30, FRAMERATE_MAX = 60;
constexpr int32_t BITRATE_MIN = 1000000, BITRATE_MAX = 10000000, GOPSIZE_MIN = 30, GOPSIZE_MAX = 3000;
// extension keys of this sibling (all optional; unset keeps the reference behaviour: I420 in, 1 slice, CBR)
const char *KEY_INPUT_FORMAT = "persist.vmi.b200.encode.input_format";   // "i420" (default) | "nv12" | "rgba"
const char *KEY_DEVICE = "persist.vmi.b200.encode.device";               // CUDA ordinal; unset: least-loaded GPU
const char *KEY_SLICES = "persist.vmi.b200.encode.slices";
const char *KEY_SEARCH_RANGE = "persist.vmi.b200.encode.search_range";
const char *KEY_CONST_QP = "persist.vmi.b200.encode.const_qp";           // test hook: fixed QP instead of rate control
const char *KEY_AUTO_BATCH = "persist.vmi.b200.encode.auto_batch";       // "0" disables the per-GPU session scheduler (default on)
}

VideoEncoderB200::VideoEncoderB200() { INFO("VideoEncoderB200 constructor"); }
VideoEncoderB200::~VideoEncoderB200() { Release(); INFO("VideoEncoderB200 destructor"); }

bool VideoEncoderB200::GetRoEncParam()
{
    int32_t width = 0, height = 0, framerate = 0;
    const std::string mode = GetStrEncParam("ro.sys.vmi.cloudphone");
    if (mode == "video") {
        width = GetIntEncParam("ro.hardware.width"); height = GetIntEncParam("ro.hardware.height"); framerate = GetIntEncParam("ro.hardware.fps");
    } else if (mode == "instruction") {
        width = GetIntEncParam("persist.vmi.demo.video.encode.width"); height = GetIntEncParam("persist.vmi.demo.video.encode.height");
        framerate = GetIntEncParam("persist.vmi.demo.video.encode.framerate");
    } else {
        ERR("Invalid property value[%s] for property[ro.sys.vmi.cloudphone], get property failed!", mode.c_str());
        return false;
    }
    if (!VerifyEncodeRoParams(width, height, framerate)) { ERR("encoder params is not supported"); return false; }
    m_tmpEncParams.width = width; m_tmpEncParams.height = height; m_tmpEncParams.framerate = framerate;
    return true;
}

bool VideoEncoderB200::GetPersistEncParam()
{
    std::string bitrate, gopsize, profile;
    const std::string mode = GetStrEncParam("ro.sys.vmi.cloudphone");
    if (mode == "video") {
        bitrate = GetStrEncParam("persist.vmi.video.encode.bitrate"); gopsize = GetStrEncParam("persist.vmi.video.encode.gopsize");
        profile = GetStrEncParam("persist.vmi.video.encode.profile");
    } else if (mode == "instruction") {
        bitrate = GetStrEncParam("persist.vmi.demo.video.encode.bitrate"); gopsize = GetStrEncParam("persist.vmi.demo.video.encode.gopsize");
        profile = GetStrEncParam("persist.vmi.demo.video.encode.profile");
    } else {
        ERR("Invalid property value[%s] for property[ro.sys.vmi.cloudphone], get property failed!", mode.c_str());
        return false;
    }
    if (!VerifyEncodeParams(bitrate, gopsize, profile)) {
        // keep the last good values and write them back, as the reference does (:111-115)
        SetEncParam("persist.vmi.video.encode.bitrate", std::to_string(m_encParams.bitrate).c_str());
        SetEncParam("persist.vmi.video.encode.gopsize", std::to_string(m_encParams.gopsize).c_str());
        SetEncParam("persist.vmi.video.encode.profile", m_encParams.profile.c_str());
    } else {
        m_tmpEncParams.bitrate = StrToInt(bitrate); m_tmpEncParams.gopsize = StrToInt(gopsize); m_tmpEncParams.profile = profile;
    }
    return true;
}

bool VideoEncoderB200::VerifyEncodeRoParams(int32_t width, int32_t height, int32_t framerate)
{
    bool ok = true;
    if (width > WH_MAX || height > WH_MAX || width < WH_MIN || height < WH_MIN) {
        ERR("Invalid property value[%dx%d] for property[width,height], get property failed!", width, height); ok = false;
    }
    if (framerate != FRAMERATE_MIN && framerate != FRAMERATE_MAX) {
        ERR("Invalid property value[%d] for property[framerate], get property failed!", framerate); ok = false;
    }
    return ok;
}

bool VideoEncoderB200::VerifyEncodeParams(std::string &bitrate, std::string &gopsize, std::string &profile)
{
    bool ok = true;
    if (StrToInt(bitrate) < BITRATE_MIN || StrToInt(bitrate) > BITRATE_MAX) {
        WARN("Invalid property value[%s] for property[bitrate], use last correct encode bitrate[%u]", bitrate.c_str(), m_encParams.bitrate); ok = false;
    }
    if (StrToInt(gopsize) < GOPSIZE_MIN || StrToInt(gopsize) > GOPSIZE_MAX) {
        WARN("Invalid property value[%s] for property[gopsize], use last correct encode gopsize[%u]", gopsize.c_str(), m_encParams.gopsize); ok = false;
    }
    if (profile != "baseline" && profile != "main" && profile != "high") {
        WARN("Invalid property value[%s] for property[profile], use last correct encode profile[%s]", profile.c_str(), m_encParams.profile.c_str()); ok = false;
    }
    return ok;
}

bool VideoEncoderB200::EncodeParamsChange()
{
    return m_tmpEncParams.bitrate != m_encParams.bitrate || m_tmpEncParams.gopsize != m_encParams.gopsize || m_tmpEncParams.profile != m_encParams.profile ||
           m_tmpEncParams.width != m_encParams.width || m_tmpEncParams.height != m_encParams.height || m_tmpEncParams.framerate != m_encParams.framerate;
}

EncoderRetCode VideoEncoderB200::InitEncoder()
{
    if (!GetRoEncParam() || !GetPersistEncParam()) { ERR("init encoder failed: GetEncParam failed"); return VIDEO_ENCODER_INIT_FAIL; }
    m_encParams = m_tmpEncParams;
    b200enc_config cfg;
    b200enc_default_config(&cfg);
    cfg.width = (int)m_encParams.width; cfg.height = (int)m_encParams.height; cfg.fps = (int)m_encParams.framerate;
    cfg.bitrate = (int)m_encParams.bitrate; cfg.gop = (int)m_encParams.gopsize;
    // the policy the reference wrapper fixes in InitParams (VideoEncoderOpenH264.cpp:239-240,282-283,289): max bitrate = target bitrate,
    // background detection and scene-change detection on, HIGH_COMPLEXITY; QP bounds stay at the encoder's defaults (:230)
    cfg.max_bitrate = cfg.bitrate; cfg.background_detection = 1; cfg.scene_change = 1; cfg.complexity = 2;
    // profile property -> uiProfileIdc as in the reference (VideoEncoderOpenH264.cpp:248-253); the wrapper asks for CABAC (:291), which the
    // Baseline profile does not have, so baseline stays CAVLC and main / high are coded with CABAC
    cfg.profile = m_encParams.profile == "high" ? 2 : m_encParams.profile == "main" ? 1 : 0;
    const std::string fmt = GetStrEncParam(KEY_INPUT_FORMAT);
    cfg.input_format = fmt == "rgba" ? B200ENC_FMT_RGBA : fmt == "nv12" ? B200ENC_FMT_NV12 : B200ENC_FMT_I420;
    const int32_t dev = GetIntEncParam(KEY_DEVICE), slices = GetIntEncParam(KEY_SLICES), range = GetIntEncParam(KEY_SEARCH_RANGE), cqp = GetIntEncParam(KEY_CONST_QP);
    cfg.device = dev >= 0 ? dev : -1;
    cfg.auto_batch = GetStrEncParam(KEY_AUTO_BATCH) == "0" ? 0 : 1;
    if (slices > 0) cfg.num_slices = slices;
    if (range > 0) cfg.search_range = range;
    if (GetStrEncParam(KEY_CONST_QP) != "" && cqp >= 0 && cqp <= 51) cfg.const_qp = cqp;
    if ((cfg.width & 1) || (cfg.height & 1)) { ERR("init encoder failed: odd frame size %dx%d", cfg.width, cfg.height); return VIDEO_ENCODER_INIT_FAIL; }
    const int rc = b200enc_create(&cfg, &m_session);
    if (rc != B200ENC_OK) { ERR("init encoder failed: create encoder failed, rc = %d (%s)", rc, b200enc_strerror(rc)); m_session = nullptr; return VIDEO_ENCODER_INIT_FAIL; }
    m_frameSize = (uint32_t)b200enc_frame_bytes(m_session);
    INFO("init encoder success");
    return VIDEO_ENCODER_SUCCESS;
}

EncoderRetCode VideoEncoderB200::StartEncoder() { INFO("start encoder success"); return VIDEO_ENCODER_SUCCESS; }

EncoderRetCode VideoEncoderB200::EncodeOneFrame(const uint8_t *inputData, uint32_t inputSize, uint8_t **outputData, uint32_t *outputSize)
{
    if (m_session == nullptr || inputData == nullptr || outputData == nullptr || outputSize == nullptr) { ERR("encode failed: encoder not initialised or null argument"); return VIDEO_ENCODER_ENCODE_FAIL; }
    if (inputSize < m_frameSize) { ERR("input size error: input size(%u) < frame size(%u)", inputSize, m_frameSize); return VIDEO_ENCODER_ENCODE_FAIL; }
    const std::string isParamChange = GetStrEncParam("persist.vmi.video.encode.param_adjusting");
    if (isParamChange == "1") {
        if (!GetPersistEncParam()) { ERR("init encoder failed: GetEncParam failed"); return VIDEO_ENCODER_INIT_FAIL; }   // same code as the reference (:316)
        SetEncodeParams();
        SetEncParam("persist.vmi.video.encode.param_adjusting", "0");
    } else if (isParamChange != "0") {
        WARN("Invalid property value[%s] for encode param adjusting", isParamChange.c_str());
        SetEncParam("persist.vmi.video.encode.param_adjusting", "0");
    }
    if (m_resetFlag) {
        if (ResetEncoder() != VIDEO_ENCODER_SUCCESS) { ERR("reset encoder failed while encoding"); return VIDEO_ENCODER_ENCODE_FAIL; }
        m_resetFlag = false;
    }
    const std::string isKeyframeChange = GetStrEncParam("persist.vmi.video.encode.keyframe");
    if (isKeyframeChange == "1") {
        INFO("Encoder set key frame");
        ForceKeyFrame();
        SetEncParam("persist.vmi.video.encode.keyframe", "0");
    } else if (isKeyframeChange != "0") {
        WARN("Invalid property value[%s] for property[keyFrame], set to [0]", isKeyframeChange.c_str());
        SetEncParam("persist.vmi.video.encode.keyframe", "0");
    }
    const uint8_t *bs = nullptr; uint32_t size = 0;
    const int rc = b200enc_encode(m_session, inputData, inputSize, &bs, &size, nullptr);
    if (rc != B200ENC_OK) { ERR("encoder encode frame failed, rc = %d (%s)", rc, b200enc_strerror(rc)); return VIDEO_ENCODER_ENCODE_FAIL; }
    *outputData = const_cast<uint8_t *>(bs);
    *outputSize = size;
    return VIDEO_ENCODER_SUCCESS;
}

EncoderRetCode VideoEncoderB200::StopEncoder() { INFO("stop encoder success"); return VIDEO_ENCODER_SUCCESS; }
void VideoEncoderB200::DestroyEncoder() { Release(); INFO("destroy encoder success"); }
void VideoEncoderB200::Release() { if (m_session != nullptr) { b200enc_destroy(m_session); m_session = nullptr; } }

EncoderRetCode VideoEncoderB200::ResetEncoder()
{
    INFO("resetting encoder");
    DestroyEncoder();
    EncoderRetCode ret = InitEncoder();
    if (ret != VIDEO_ENCODER_SUCCESS) { ERR("init encoder failed %#x while resetting", ret); return VIDEO_ENCODER_RESET_FAIL; }
    ret = StartEncoder();
    if (ret != VIDEO_ENCODER_SUCCESS) { ERR("start encoder failed %#x while resetting", ret); return VIDEO_ENCODER_RESET_FAIL; }
    INFO("reset encoder success");
    return VIDEO_ENCODER_SUCCESS;
}

EncoderRetCode VideoEncoderB200::ForceKeyFrame()
{
    if (m_session == nullptr || b200enc_force_idr(m_session) != B200ENC_OK) { ERR("encoder force intra frame failed"); return VIDEO_ENCODER_FORCE_KEY_FRAME_FAIL; }
    INFO("force key frame success");
    return VIDEO_ENCODER_SUCCESS;
}

EncoderRetCode VideoEncoderB200::SetEncodeParams()
{
    if (EncodeParamsChange()) {
        m_encParams = m_tmpEncParams; m_resetFlag = true;
        INFO("Handle encoder config change: [bitrate, gopsize, profile] = [%u,%u,%s]", m_encParams.bitrate, m_encParams.gopsize, m_encParams.profile.c_str());
    } else {
        INFO("Using encoder config: [bitrate, gopsize, profile] = [%u,%u,%s]", m_encParams.bitrate, m_encParams.gopsize, m_encParams.profile.c_str());
    }
    return VIDEO_ENCODER_SUCCESS;
}
