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
