"""CPU tests of the host side: the C ABI exports what include/b200enc.h declares, the libVideoCodec.so factory and the
VideoEncoderB200 property handling behave like the reference wrapper (video_codec/VideoCodecApi.cpp:21-55,
VideoEncoderOpenH264.cpp:62-195), and the multi-GPU sharding helper works over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    return True


def test_c_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "b200enc.h")).read()
    names = set(re.findall(r"\b(b200enc_\w+|b200k_\w+)\s*\(", hdr))
    names -= {"b200enc_config", "b200enc_frame_info", "b200enc_session", "b200enc_batch"}
    assert len(names) >= 30
    L = C.CDLL(os.path.join(ROOT, "media_b200", "csrc", "libb200enc.so"))
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/b200enc.h but not exported"


def test_macroblock_index_arithmetic_is_exact_for_every_accepted_width(built):
    """Every warp-per-MB kernel turns its macroblock index into (mx, my) with one multiply-high by floor(2^32 / mbw) and one correction
    (h264_dev.cuh: mb_xy) instead of a division; the host build of the same function must agree with / and % for every picture width the
    encoder accepts (16..4096 samples = 1..256 macroblocks) and every index of a 256-row picture."""
    L = C.CDLL(os.path.join(ROOT, "media_b200", "csrc", "libb200enc.so"))
    L.b200k_mb_xy_mismatches.argtypes = [C.c_int, C.c_int]; L.b200k_mb_xy_mismatches.restype = C.c_int
    for mbw in range(1, 257):
        assert L.b200k_mb_xy_mismatches(mbw, mbw * 256 + 3) == 0, mbw
    assert L.b200k_mb_xy_mismatches(120, 1 << 24) == 0 and L.b200k_mb_xy_mismatches(1, 1 << 20) == 0
    assert L.b200k_mb_xy_mismatches(0, 10) == -1


def test_codec_library_exports_the_reference_factory(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "media_b200", "host", "libVideoCodec.so")], text=True)
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert {"CreateVideoEncoder", "DestroyVideoEncoder"} <= syms      # reference: video_codec/VideoCodecApi.h:80-96


def _vc():
    L = C.CDLL(os.path.join(ROOT, "media_b200", "host", "libVideoCodec.so"))
    L.vc_create.argtypes = [C.POINTER(C.c_void_p)]; L.vc_create.restype = C.c_uint32
    for f in ("vc_init", "vc_start", "vc_stop", "vc_reset", "vc_destroy"):
        getattr(L, f).argtypes = [C.c_void_p]; getattr(L, f).restype = C.c_uint32
    L.vc_prop_set.argtypes = [C.c_char_p, C.c_char_p]
    L.vc_prop_get.argtypes = [C.c_char_p, C.c_char_p]
    return L


def test_factory_selector_and_error_codes(built):
    L = _vc()
    e = C.c_void_p()
    L.vc_prop_set(b"ro.vmi.demo.video.encode.format", b"")
    assert L.vc_create(C.byref(e)) == 1            # unset selector -> -1 -> CREATE_FAIL (VideoCodecApi.cpp:23,36-38)
    L.vc_prop_set(b"ro.vmi.demo.video.encode.format", b"7")
    assert L.vc_create(C.byref(e)) == 1
    L.vc_prop_set(b"ro.vmi.demo.video.encode.format", b"3")
    assert L.vc_create(C.byref(e)) == 0 and e.value
    assert L.vc_destroy(None) == 0                 # null encoder: warning + SUCCESS (VideoCodecApi.cpp:48-51)
    # invalid phone mode / geometry / fps -> INIT_FAIL without touching the GPU (VideoEncoderOpenH264.cpp:76-79,159-171)
    L.vc_prop_set(b"ro.sys.vmi.cloudphone", b"bogus")
    assert L.vc_init(e) == 2
    L.vc_prop_set(b"ro.sys.vmi.cloudphone", b"video")
    for k, v in ((b"ro.hardware.width", b"8"), (b"ro.hardware.height", b"720"), (b"ro.hardware.fps", b"30")):
        L.vc_prop_set(k, v)
    assert L.vc_init(e) == 2
    L.vc_prop_set(b"ro.hardware.width", b"1280"); L.vc_prop_set(b"ro.hardware.fps", b"25")
    assert L.vc_init(e) == 2
    # invalid persist params are replaced by the last good ones and written back (VideoEncoderOpenH264.cpp:111-115)
    L.vc_prop_set(b"ro.hardware.fps", b"30")
    L.vc_prop_set(b"persist.vmi.video.encode.bitrate", b"12"); L.vc_prop_set(b"persist.vmi.video.encode.gopsize", b"30")
    L.vc_prop_set(b"persist.vmi.video.encode.profile", b"baseline")
    L.vc_init(e)    # SUCCESS on a GPU box, INIT_FAIL (no device) here; either way the write-back happened before the device is touched
    buf = C.create_string_buffer(92); L.vc_prop_get(b"persist.vmi.video.encode.bitrate", buf)
    assert buf.value == b"5000000"                 # default of the reference (VideoEncoderOpenH264.h:19-21)
    assert L.vc_stop(e) == 0 and L.vc_destroy(e) == 0


def test_no_cpu_fallback_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this checks the CPU-only failure mode")
    from media_b200 import enc
    with pytest.raises(enc.B200EncError, match="no usable CUDA device"):
        enc.Session(64, 48, const_qp=26)


def test_session_sharding_over_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, json\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch, torch.distributed as dist\n"
        "from media_b200 import shard\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "mine = shard.sessions_of_rank(11, w, r)\n"
        "t = shard.max_over_ranks(1.0 + r)\n"
        "tot = shard.sum_over_ranks(len(mine))\n"
        f"open(os.path.join({str(tmp_path)!r}, 'r%d.json' % r), 'w').write(json.dumps({{'rank': r, 'mine': mine, 't': t, 'tot': tot}}))\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                                   "--master-port", "29631", str(script)], text=True, env=env, timeout=300, stderr=subprocess.DEVNULL)
    import json
    rows = [json.load(open(tmp_path / f"r{r}.json")) for r in range(2)]
    assert len(rows) == 2
    ids = sorted(sum((r["mine"] for r in rows), []))
    assert ids == list(range(11)) and all(r["t"] == 2.0 and r["tot"] == 11 for r in rows)
    assert abs(len(rows[0]["mine"]) - len(rows[1]["mine"])) <= 1


REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vendor", "openh264")), reason="reference tree only exists in the authoring container")


@needs_reference
def test_openh264_abi_layout_matches_the_vendored_headers(tmp_path):
    """include/openh264_abi.h is written from scratch; here it is checked field by field against the reference's vendored
    openh264 headers (vendor/openh264/codec_api.h, codec_app_def.h)"""
    src = tmp_path / "layout.cpp"
    pairs = [("SEncParamExt", "oh264::EncParamExt", [("iPicWidth", "width"), ("iTargetBitrate", "target_bitrate"), ("iRCMode", "rc_mode"), ("fMaxFrameRate", "max_frame_rate"),
              ("iSpatialLayerNum", "spatial_layers"), ("sSpatialLayers", "layers"), ("uiIntraPeriod", "intra_period"), ("iNumRefFrame", "num_ref"),
              ("iEntropyCodingModeFlag", "entropy_mode"), ("bEnableFrameSkip", "frame_skip"), ("iMaxBitrate", "max_bitrate"), ("iMaxQp", "max_qp"), ("iMinQp", "min_qp"),
              ("uiMaxNalSize", "max_nal_size"), ("iMultipleThreadIdc", "multiple_thread_idc"), ("iLoopFilterDisableIdc", "loop_filter_disable_idc"),
              ("bEnableBackgroundDetection", "background_detection"), ("bEnableSceneChangeDetect", "scene_change_detect"), ("iComplexityMode", "complexity"), ("eSpsPpsIdStrategy", "sps_pps_id_strategy")]),
             ("SSpatialLayerConfig", "oh264::SpatialLayer", [("iVideoWidth", "width"), ("fFrameRate", "frame_rate"), ("iSpatialBitrate", "bitrate"), ("iMaxSpatialBitrate", "max_bitrate"),
              ("uiProfileIdc", "profile_idc"), ("uiLevelIdc", "level_idc"), ("iDLayerQp", "dlayer_qp"), ("sSliceArgument", "slice")]),
             ("SSliceArgument", "oh264::SliceArgument", [("uiSliceMode", "mode"), ("uiSliceNum", "num"), ("uiSliceSizeConstraint", "size_constraint")]),
             ("SSourcePicture", "oh264::SourcePicture", [("iColorFormat", "color_format"), ("iStride", "stride"), ("pData", "data"), ("iPicWidth", "width"), ("iPicHeight", "height"), ("uiTimeStamp", "timestamp")]),
             ("SLayerBSInfo", "oh264::LayerBSInfo", [("eFrameType", "frame_type"), ("uiLayerType", "layer_type"), ("iNalCount", "nal_count"), ("pNalLengthInByte", "nal_length"), ("pBsBuf", "bs_buf")]),
             ("SFrameBSInfo", "oh264::FrameBSInfo", [("iLayerNum", "layer_num"), ("sLayerInfo", "layers"), ("eFrameType", "frame_type"), ("iFrameSizeInBytes", "frame_size"), ("uiTimeStamp", "timestamp")]),
             ("SEncParamBase", "oh264::EncParamBase", [("iPicWidth", "width"), ("iRCMode", "rc_mode"), ("fMaxFrameRate", "max_frame_rate")]),
             ("SBitrateInfo", "oh264::BitrateInfo", [("iBitrate", "bitrate")])]
    body = ['#include <cstddef>', '#include "codec_api.h"', '#define WelsCreateSVCEncoder WelsCreateSVCEncoder_abi', '#define WelsDestroySVCEncoder WelsDestroySVCEncoder_abi', '#include "openh264_abi.h"']
    for a, b, fields in pairs:
        body.append(f'static_assert(sizeof({a}) == sizeof({b}), "{a} size");')
        for fa, fb in fields:
            body.append(f'static_assert(offsetof({a}, {fa}) == offsetof({b}, {fb}), "{a}.{fa}");')
    body += ['static_assert((int)videoFormatI420 == oh264::kVideoFormatI420 && (int)videoFrameTypeIDR == oh264::kFrameIDR && (int)videoFrameTypeP == oh264::kFrameP, "enums");',
             'static_assert((int)RC_OFF_MODE == oh264::kRcOff && (int)RC_BITRATE_MODE == oh264::kRcBitrate && (int)SM_FIXEDSLCNUM_SLICE == oh264::kSliceFixedNum, "enums");',
             'static_assert((int)ENCODER_OPTION_DATAFORMAT == oh264::kOptDataFormat && (int)ENCODER_OPTION_BITRATE == oh264::kOptBitrate && (int)ENCODER_OPTION_RC_MODE == oh264::kOptRcMode && (int)ENCODER_OPTION_FRAME_RATE == oh264::kOptFrameRate, "enums");',
             'static_assert((int)NON_VIDEO_CODING_LAYER == oh264::kLayerNonVcl && MAX_LAYER_NUM_OF_FRAME == oh264::kMaxLayers, "enums");', 'int main() { return 0; }']
    src.write_text("\n".join(body))
    subprocess.check_call(["g++", "-std=c++14", "-I", os.path.join(REF, "vendor", "openh264"), "-I", os.path.join(ROOT, "include"), str(src), "-o", str(tmp_path / "layout")])


@needs_reference
def test_unmodified_reference_wrapper_loads_the_shim(built, tmp_path):
    """the reference's own VideoCodecApi.cpp + VideoEncoderOpenH264.cpp, compiled untouched from /root/reference, dlopen
    "libopenh264.so" = our shim, create the encoder through its vtable and read the defaults; without a GPU InitializeExt
    then fails and the wrapper reports VIDEO_ENCODER_INIT_FAIL (on a B200 the same flow encodes: see the GPU shim test)"""
    host = os.path.join(ROOT, "media_b200", "host")
    srcs = [os.path.join(REF, p) for p in ("video_codec/VideoCodecApi.cpp", "video_codec/VideoEncoderOpenH264.cpp", "video_codec/VideoEncoderNetint.cpp",
                                           "common/log/MediaLog.cpp", "common/log/MediaLogManager.cpp", "common/prop/Property.cpp")]
    inc = [os.path.join(host, "shim")] + [os.path.join(REF, p) for p in ("video_codec", "common/log", "common/prop", "vendor/openh264", "vendor/netint")]
    drv = tmp_path / "drv.cpp"
    drv.write_text('#include "VideoCodecApi.h"\n#include <sys/system_properties.h>\n#include <cstdio>\n'
                   'int main() { const char *kv[][2] = {{"ro.vmi.demo.video.encode.format","0"},{"ro.sys.vmi.cloudphone","video"},{"ro.hardware.width","1280"},'
                   '{"ro.hardware.height","720"},{"ro.hardware.fps","30"},{"persist.vmi.video.encode.bitrate","4000000"},{"persist.vmi.video.encode.gopsize","30"},'
                   '{"persist.vmi.video.encode.profile","baseline"}};\n for (auto &p : kv) __system_property_set(p[0], p[1]);\n'
                   ' VideoEncoder *e = nullptr; unsigned c = CreateVideoEncoder(&e); unsigned i = e ? e->InitEncoder() : 99;\n'
                   ' printf("create=%u init=%u\\n", c, i); if (e) DestroyVideoEncoder(e); return 0; }\n')
    exe = tmp_path / "drv"
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-w"] + [f"-I{i}" for i in inc] + srcs + [os.path.join(host, "PropertyStore.cpp"), str(drv), "-ldl", "-o", str(exe)])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "media_b200", "shim") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([str(exe)], env=env, capture_output=True, text=True, timeout=120)
    assert "create=0" in out.stdout, out.stdout + out.stderr
    import torch
    assert ("init=0" if torch.cuda.is_available() else "init=2") in out.stdout, out.stdout + out.stderr
    assert "load openh264 shared lib failed" not in (out.stdout + out.stderr)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` times the CPU restatement of the path on the host cores (the arm the driver runs beside ours) and prints one
    JSON line with our arm's metric / unit / config; no GPU is needed for it"""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "1080p H.264 encode frames/s per GPU" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "1920x1080" in cb["sample"]
