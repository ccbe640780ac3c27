"""Rate control (media_b200/csrc/rate_control.h): the wrapper asks for RC_BITRATE_MODE with iMaxBitrate = iTargetBitrate and the default
iMinQp / iMaxQp (reference video_codec/VideoEncoderOpenH264.cpp:230,239-240,274). The control law is host logic, so it is tested on
the CPU in closed loop with the oracle encoder (which produces the bytes the CUDA path produces for the same QP), and on the GPU the
session's QP / size trace must equal that simulation picture by picture."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def sim():
    import __graft_entry__ as ge
    ge.build()
    import rc_sim
    return rc_sim


def _bucket_ok(sizes, bitrate, fps, max_bitrate=None):
    """one-second leaky bucket drained at max_bitrate: peak level as a fraction of its size"""
    drain = (max_bitrate or bitrate) / fps; level = peak = 0.0
    for s in sizes:
        level = max(0.0, level + 8 * s - drain); peak = max(peak, level)
    return peak / (max_bitrate or bitrate)


def test_cbr_closed_loop_hits_the_target_within_5_percent(sim):
    w, h, br, n = 320, 192, 300_000, 300
    r = sim.simulate(w, h, "A", br, n, gop=300)
    assert abs(r["rate"] / br - 1) < 0.05, r["rate"]
    per_second = [sum(r["sizes"][a:a + 30]) * 8 for a in range(30, n, 30)]
    assert all(abs(x / br - 1) < 0.10 for x in per_second), per_second      # every second after the first within 10 %
    assert _bucket_ok(r["sizes"], br, 30) < 1.0 and r["vbv_peak"] < r["vbv_size"]     # the one-second bucket never overflows
    assert max(abs(a - b) for a, b in zip(r["qps"][2:], r["qps"][3:])) <= 3              # P pictures move at most 3 QP steps
    assert min(r["qps"]) >= 10 and max(r["qps"]) <= 51


def test_cbr_with_periodic_key_frames_and_screen_content(sim):
    w, h, br, n = 320, 192, 300_000, 150
    r = sim.simulate(w, h, "B", br, n, gop=30)
    assert sum(r["types"]) == 5 and abs(r["rate"] / br - 1) < 0.05, (r["rate"], r["types"])
    assert _bucket_ok(r["sizes"], br, 30) < 1.0
    idr = [s for s, t in zip(r["sizes"], r["types"]) if t][1:]
    assert max(idr) * 8 <= 8 * br / 30 * 1.01, idr          # key frames stay under the hard cap of 8 picture budgets


def test_qp_bounds_from_the_abi_are_honoured(sim):
    w, h = 320, 192
    r = sim.simulate(w, h, "A", 300_000, 60, min_qp=30, max_qp=33)
    assert min(r["qps"]) >= 30 and max(r["qps"]) <= 33
    r = sim.simulate(w, h, "D", 100_000, 20, max_qp=40)          # noise far above the budget: pinned to the upper bound, never beyond
    assert max(r["qps"]) == 40 and r["qps"][-1] == 40


def test_oversized_pictures_are_coded_again(sim):
    """a first key frame several times the cap (no statistics yet) must trigger exactly the second attempt, with a coarser QP"""
    L = sim.rc_lib()
    import ctypes as C
    rc = L.b200k_rc_create(1e6, 0, 30, 0, 51, 640, 368)
    b, c = C.c_double(), C.c_double()
    q = L.b200k_rc_pick(rc, 1, C.byref(b), C.byref(c))
    T = 1e6 / 30
    assert abs(b.value - 5 * T) < 1 and abs(c.value - 8 * T) < 1
    assert L.b200k_rc_retry_qp(rc, 1, 1, 7.9 * T) == -1
    q2 = L.b200k_rc_retry_qp(rc, 1, 1, 20 * T)
    assert q + 2 <= q2 <= min(51, q + 12)
    L.b200k_rc_update(rc, 1, q2, 6 * T)
    assert L.b200k_rc_pick(rc, 0, None, None) == q2 - 3           # first P picture after the first key frame
    L.b200k_rc_destroy(rc)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,gop,frames", [("A", 300, 300), ("B", 30, 150)])
def test_gpu_session_follows_the_simulated_trace(sim, kind, gop, frames):
    """CBR on the GPU: same control law + same bytes per QP => the session's (QP, size, type) trace equals the CPU closed-loop
    simulation with the oracle; the achieved bitrate is within 5 % and the one-second bucket never overflows"""
    from media_b200 import enc
    from media_b200.synth import Content
    w, h, br = 640, 368, 1_000_000
    r = sim.simulate(w, h, kind, br, frames, gop=gop)
    s = enc.Session(w, h, fps=30, bitrate=br, gop=gop, const_qp=-1, device=0)
    c = Content(kind, w, h)
    sizes, qps, types = [], [], []
    for t in range(frames):
        bs, info = s.encode(c.frame(t)); sizes.append(len(bs)); qps.append(info.qp); types.append(info.frame_type)
    assert qps == r["qps"] and sizes == r["sizes"] and types == r["types"]
    assert enc.lib().b200enc_rc_retries(s.h) == r["retries"]
    rate = sum(sizes) * 8 * 30 / frames
    assert abs(rate / br - 1) < 0.05, rate
    assert _bucket_ok(sizes, br, 30) < 1.0
    s.close()
