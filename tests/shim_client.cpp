// Test client of media_b200/shim/libopenh264.so: does what the reference wrapper does with openh264
// (video_codec/VideoEncoderOpenH264.cpp:197-296 load + configure, :344-350 per frame, :406-415 key frame), through the vtable.
// usage: shim_client <libopenh264.so> <in.i420> <w> <h> <frames> <bitrate> <gop> <force_idr_at> <out.h264> <out.info>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>
#include "openh264_abi.h"
using namespace oh264;
int main(int argc, char **argv)
{
    if (argc < 11) return 2;
    void *h = dlopen(argv[1], RTLD_LAZY);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
    auto create = (int (*)(SvcEncoder **))dlsym(h, "WelsCreateSVCEncoder");
    auto destroy = (void (*)(SvcEncoder *))dlsym(h, "WelsDestroySVCEncoder");
    if (!create || !destroy) return 4;
    const int w = atoi(argv[3]), hgt = atoi(argv[4]), n = atoi(argv[5]), bitrate = atoi(argv[6]), gop = atoi(argv[7]), force_at = atoi(argv[8]);
    SvcEncoder *e = nullptr;
    if (create(&e) != 0) return 5;
    EncParamExt p;
    if (e->GetDefaultParams(&p) != 0) return 6;
    // the wrapper's policy (VideoEncoderOpenH264.cpp:236-255, 270-296)
    p.width = w; p.height = hgt; p.target_bitrate = bitrate; p.max_bitrate = bitrate; p.max_frame_rate = 30.f; p.intra_period = gop;
    p.layers[0].width = w; p.layers[0].height = hgt; p.layers[0].frame_rate = 30.f; p.layers[0].bitrate = bitrate;
    p.layers[0].slice.mode = kSliceSingle; p.layers[0].profile_idc = argc > 11 ? atoi(argv[11]) : 66; p.layers[0].level_idc = 32;
    p.usage = 0; p.rc_mode = kRcBitrate; p.frame_skip = 0; p.temporal_layers = 1; p.spatial_layers = 1; p.sps_pps_id_strategy = 0;
    p.background_detection = 1; p.scene_change_detect = 1; p.complexity = 2; p.num_ref = 1; p.entropy_mode = 1; p.max_nal_size = 0;
    p.multiple_thread_idc = 1; p.loop_filter_disable_idc = 0;
    if (e->InitializeExt(&p) != 0) { fprintf(stderr, "InitializeExt failed\n"); return 7; }
    int fmt = kVideoFormatI420;
    if (e->SetOption(kOptDataFormat, &fmt) != 0) return 8;
    FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[9], "wb"), *fn = fopen(argv[10], "w");
    if (!fi || !fo || !fn) return 9;
    const size_t fb = (size_t)w * hgt * 3 / 2;
    std::vector<uint8_t> frame(fb);
    static FrameBSInfo info;
    for (int t = 0; t < n; t++) {
        if (fread(frame.data(), 1, fb, fi) != fb) return 10;
        if (t == force_at && e->ForceIntraFrame(true) != 0) return 11;
        SourcePicture src; memset(&src, 0, sizeof src);
        src.width = w; src.height = hgt; src.color_format = kVideoFormatI420;
        src.stride[0] = w; src.stride[1] = w / 2; src.stride[2] = w / 2;
        src.data[0] = frame.data(); src.data[1] = src.data[0] + (size_t)w * hgt; src.data[2] = src.data[1] + ((size_t)w * hgt >> 2);
        if (e->EncodeFrame(&src, &info) != 0) { fprintf(stderr, "EncodeFrame failed at %d\n", t); return 12; }
        // what the wrapper hands to its caller: everything from layer 0's buffer, iFrameSizeInBytes long (:349-350)
        fwrite(info.layers[0].bs_buf, 1, (size_t)info.frame_size, fo);
        int nal_sum = 0, nals = 0;
        for (int l = 0; l < info.layer_num; l++) for (int k = 0; k < info.layers[l].nal_count; k++) { nal_sum += info.layers[l].nal_length[k]; nals++; }
        fprintf(fn, "%d %d %d %d %d %d %d\n", t, info.frame_type, info.frame_size, info.layer_num, nals, nal_sum, (int)info.layers[0].layer_type);
    }
    static FrameBSInfo ps;
    if (e->EncodeParameterSets(&ps) != 0) return 13;
    fprintf(fn, "ps %d %d %d\n", ps.frame_size, ps.layer_num, ps.layers[0].nal_count);
    fclose(fo); fclose(fn); fclose(fi);
    e->Uninitialize();
    destroy(e);
    return 0;
}
