"""CPU tests of the oracle (oracle/): it is pinned by (a) FFmpeg's independent H.264 decoder reproducing its
reconstruction bit for bit, (b) the CAVLC / deblock / QP tables matching the copies inside that decoder's binary,
(c) the committed golden hashes. The reference itself holds no test vectors for this path."""
import glob
import hashlib
import json
import os
import re

import numpy as np
import pytest

import avdec
from media_b200.synth import Content

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "encode_golden.json")))
needs_decoder = pytest.mark.skipif(not avdec.available(), reason="libavcodec from the opencv wheel not loadable")


def encode_case(orc, c, frames=None):
    e = orc.Encoder(c["w"], c["h"], num_slices=c["slices"], search_range=c["sr"], profile=c.get("profile", 0))
    content = Content(c["kind"], c["w"], c["h"])
    aus, recs = [], []
    for t in range(frames or c["frames"]):
        aus.append(e.encode(content.frame(t), t == 0, c["qp"])); recs.append(e.recon())
    return aus, recs


@pytest.mark.parametrize("case", [c for c in GOLDEN if c["w"] * c["h"] <= 640 * 360], ids=lambda c: c["name"])
def test_oracle_matches_golden(orc, case):
    aus, recs = encode_case(orc, case)
    assert [hashlib.sha256(a).hexdigest() for a in aus] == case["au_sha256"]
    assert [hashlib.sha256(r.tobytes()).hexdigest() for r in recs] == case["recon_sha256"]
    if "stream_hex" in case:
        assert [a.hex() for a in aus] == case["stream_hex"]


@needs_decoder
@pytest.mark.parametrize("case", [c for c in GOLDEN if c["w"] * c["h"] <= 640 * 360], ids=lambda c: c["name"])
def test_oracle_stream_decodes_to_its_reconstruction(orc, case):
    aus, recs = encode_case(orc, case)
    dec = avdec.decode_stream(aus)
    assert len(dec) == len(recs)
    for t, (d, r) in enumerate(zip(dec, recs)):
        assert np.array_equal(d, r), f"frame {t}: decoder output differs from the oracle reconstruction"


@needs_decoder
def test_golden_stream_decodes_without_the_oracle():
    cases = [c for c in GOLDEN if "stream_hex" in c]
    assert {c.get("profile", 0) for c in cases} == {0, 1}          # a CAVLC and a CABAC stream are committed byte for byte
    for case in cases:
        dec = avdec.decode_stream([bytes.fromhex(h) for h in case["stream_hex"]])
        assert [hashlib.sha256(d.tobytes()).hexdigest() for d in dec] == case["recon_sha256"], case["name"]


@needs_decoder
def test_forced_idr_and_slices_mid_stream(orc):
    w, h = 128, 96
    e = orc.Encoder(w, h, num_slices=3, search_range=32)
    c = Content("A", w, h)
    aus, recs = [], []
    for t in range(6):
        aus.append(e.encode(c.frame(t), t in (0, 3), 28 + t)); recs.append(e.recon())   # QP changes per frame, IDR at 3
    assert aus[3][4] == 0x67 and aus[1][4] == 0x61
    dec = avdec.decode_stream(aus)
    assert all(np.array_equal(d, r) for d, r in zip(dec, recs)) and len(dec) == 6


def test_tables_match_the_independent_decoder():
    """Tables 9-5, 9-7..9-10, 8-15, 8-16 and the zig-zag scan are present byte for byte inside libavcodec."""
    if not avdec.available():
        pytest.skip("no libavcodec")
    import cv2
    d = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
    blob = open(glob.glob(os.path.join(d, "libavcodec-*"))[0], "rb").read()
    for header in (os.path.join(ROOT, "oracle", "h264_tables.h"),):
        src = open(header).read()
        tabs = {m.group(1): bytes(int(x) for x in re.findall(r"\d+", m.group(2)))
                for m in re.finditer(r"static const uint8_t (\w+)\[[^\]]*\](?:\[[^\]]*\])? = \{(.*?)\};", src, re.S)}
        for name in ("COEFF_TOKEN_LEN", "COEFF_TOKEN_BITS", "CHROMA_DC_COEFF_TOKEN_LEN", "CHROMA_DC_COEFF_TOKEN_BITS", "TOTAL_ZEROS_LEN",
                     "TOTAL_ZEROS_BITS", "CHROMA_DC_TOTAL_ZEROS_LEN", "CHROMA_DC_TOTAL_ZEROS_BITS", "RUN_BEFORE_LEN", "RUN_BEFORE_BITS",
                     "ZIGZAG4x4", "CHROMA_QP", "DEBLOCK_ALPHA", "DEBLOCK_BETA", "ZIGZAG8x8", "DEQUANT8_V", "CABAC_SIG8", "CABAC_LAST8"):
            assert blob.find(tabs[name]) >= 0, name


def test_cabac_tables_match_the_independent_decoder():
    """Tables 9-12..9-23 (m, n for ctxIdx 0..459, I slices and cabac_init_idc 0), 9-44 (rangeTabLPS) and 9-45 (state transitions)
    as committed in oracle/cabac_tables.h against the copies inside libavcodec"""
    if not avdec.available():
        pytest.skip("no libavcodec")
    import cv2
    d = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
    blob = open(glob.glob(os.path.join(d, "libavcodec-*"))[0], "rb").read()
    src = open(os.path.join(ROOT, "oracle", "cabac_tables.h")).read()
    tabs = {m.group(1): [int(x) for x in re.findall(r"-?\d+", m.group(2))]
            for m in re.finditer(r"static const u?int8_t (\w+)\[[^\]]*\] = \{(.*?)\};", src, re.S)}
    assert len(tabs["CABAC_INIT_I"]) == 920 and len(tabs["CABAC_INIT_P0"]) == 920
    for name in ("CABAC_INIT_I", "CABAC_INIT_P0"):         # the decoder's tables run to ctxIdx 1023; ours are their first 460 rows
        assert blob.find(np.array(tabs[name], np.int8).tobytes()) >= 0, name
    ns = blob.find(bytes([9, 8, 7, 7, 6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5]))      # ff_h264_cabac_tables: norm_shift, lps_range, mlps_state
    assert ns >= 0
    lps, ml = blob[ns + 512: ns + 1024], blob[ns + 1024: ns + 1280]
    for s_ in range(64):
        for q in range(4):
            assert lps[q * 128 + 2 * s_] == lps[q * 128 + 2 * s_ + 1] == tabs["CABAC_RANGE_LPS"][s_ * 4 + q]
        assert ml[128 + 2 * s_] // 2 == tabs["CABAC_NEXT_MPS"][s_] and ml[127 - 2 * s_] // 2 == tabs["CABAC_NEXT_LPS"][s_]
    # spot values of the standard itself
    assert tabs["CABAC_RANGE_LPS"][:4] == [128, 176, 208, 240] and tabs["CABAC_RANGE_LPS"][-4:] == [2, 2, 2, 2]
    assert tabs["CABAC_INIT_I"][:6] == [20, -15, 2, 54, 3, 74] and tabs["CABAC_INIT_P0"][22:26] == [23, 33, 23, 2]


def test_product_cabac_tables_equal_oracle_tables():
    def nums(text, name):
        m = re.search(name + r"\[[^\]]*\] = \{(.*?)\};", text, re.S)
        assert m, name
        return [int(x) for x in re.findall(r"-?\d+", m.group(1))]
    o = open(os.path.join(ROOT, "oracle", "cabac_tables.h")).read()
    p = open(os.path.join(ROOT, "media_b200", "csrc", "cabac_tables.cuh")).read()
    for a, b in (("CABAC_INIT_I", "c_cabac_init_i"), ("CABAC_INIT_P0", "c_cabac_init_p0"), ("CABAC_RANGE_LPS", "c_cabac_range_lps"),
                 ("CABAC_NEXT_LPS", "c_cabac_next_lps"), ("CABAC_NEXT_MPS", "c_cabac_next_mps")):
        assert nums(o, a) == nums(p, b) and len(nums(o, a)) > 0, (a, b)


def test_cabac_changes_only_the_entropy_coding(orc):
    """same frames, same QP: Main / High (CABAC) and Baseline (CAVLC) sessions reconstruct identically, CABAC is smaller, and the
    arithmetic coder run on the dumped bin list alone reproduces the slice payload"""
    w, h = 176, 144
    c = Content("A", w, h)
    e0, e1 = orc.Encoder(w, h), orc.Encoder(w, h, profile=1)
    for t in range(4):
        a0, a1 = e0.encode(c.frame(t), t == 0, 28), e1.encode(c.frame(t), t == 0, 28)
        assert np.array_equal(e0.recon(), e1.recon()) and len(a1) < len(a0)
        assert np.array_equal(e0.mb_info(), e1.mb_info())
        bins = e1.slice_bins(0)
        assert bins[-1] == (276 | 1 << 10) and int((bins & 1023 == 276).sum()) == e1.mbw * e1.mbh        # one end_of_slice_flag per MB
        out = np.zeros(bins.size * 4 + 64, np.uint8)
        n = orc.lib().orc_cabac_code_bins(bins.ctypes.data, bins.size, 28, int(t > 0), out.ctypes.data, out.size)
        payload = out[:n].tobytes()
        # Annex-B slice NAL: unescape, then the payload sits after the byte-aligned slice header
        nal = a1[a1.rfind(b"\x00\x00\x00\x01") + 5:]
        rbsp = nal.replace(b"\x00\x00\x03", b"\x00\x00")
        assert rbsp.endswith(payload)


def test_product_tables_equal_oracle_tables():
    """media_b200/csrc/h264_dev.cuh carries its own copies (the product never includes oracle/); they must agree."""
    def nums(text, name):
        m = re.search(name + r"\[[^\]]*\](?:\[[^\]]*\])? = \{(.*?)\};", text, re.S)
        assert m, name
        return [int(x) for x in re.findall(r"\d+", m.group(1))]
    o = open(os.path.join(ROOT, "oracle", "h264_tables.h")).read()
    p = open(os.path.join(ROOT, "media_b200", "csrc", "h264_dev.cuh")).read()
    pairs = [("COEFF_TOKEN_LEN", "c_coeff_token_len"), ("COEFF_TOKEN_BITS", "c_coeff_token_bits"), ("CHROMA_DC_COEFF_TOKEN_LEN", "c_cdc_token_len"),
             ("CHROMA_DC_COEFF_TOKEN_BITS", "c_cdc_token_bits"), ("TOTAL_ZEROS_LEN", "c_total_zeros_len"), ("TOTAL_ZEROS_BITS", "c_total_zeros_bits"),
             ("CHROMA_DC_TOTAL_ZEROS_LEN", "c_cdc_total_zeros_len"), ("CHROMA_DC_TOTAL_ZEROS_BITS", "c_cdc_total_zeros_bits"),
             ("RUN_BEFORE_LEN", "c_run_before_len"), ("RUN_BEFORE_BITS", "c_run_before_bits"), ("ZIGZAG4x4", "c_zigzag"), ("QUANT_MF", "c_quant_mf"),
             ("DEQUANT_V", "c_dequant_v"), ("CHROMA_QP", "c_chroma_qp"), ("DEBLOCK_ALPHA", "c_alpha"), ("DEBLOCK_BETA", "c_beta"), ("DEBLOCK_TC0", "c_tc0"),
             ("CBP_TO_CODENUM_INTER", "c_cbp_inter"), ("CBP_TO_CODENUM_INTRA", "c_cbp_intra"), ("LAMBDA_TAB", "c_lambda")]
    for a, b in pairs:
        assert nums(o, a) == nums(p, b), (a, b)
    # the 8x8 transform / CABAC ctxBlockCat 5 tables live beside their kernels
    t8 = open(os.path.join(ROOT, "media_b200", "csrc", "k_t8.cuh")).read(); cb = open(os.path.join(ROOT, "media_b200", "csrc", "k_cabac.cuh")).read()
    for a, b, text in (("ZIGZAG8x8", "c_zigzag8", t8), ("DEQUANT8_V", "c_dequant8_v", t8), ("QUANT8_MF", "c_quant8_mf", t8),
                       ("CABAC_SIG8", "c_cabac_sig8", cb), ("CABAC_LAST8", "c_cabac_last8", cb)):
        assert nums(o, a) == nums(text, b), (a, b)
    zz = nums(t8, "c_zigzag8"); izz = nums(t8, "c_izigzag8")
    assert sorted(zz) == list(range(64)) and all(izz[zz[i]] == i for i in range(64))


def test_transform_round_trip_and_quant(orc):
    L = orc.lib(); rng = np.random.default_rng(1)
    for qp in (0, 12, 26, 40, 51):
        for _ in range(50):
            res = rng.integers(-255, 256, 16).astype(np.int16)
            coef = np.zeros(16, np.int16); L.orc_dct4x4(res.ctypes.data, coef.ctypes.data)
            # forward core transform against the matrix definition Cf X Cf^T
            Cf = np.array([[1, 1, 1, 1], [2, 1, -1, -2], [1, -1, -1, 1], [1, -2, 2, -1]])
            assert np.array_equal(coef.reshape(4, 4), Cf @ res.reshape(4, 4).astype(np.int64) @ Cf.T)
            lz = np.zeros(16, np.int16); n = L.orc_quant4x4(coef.ctypes.data, lz.ctypes.data, qp, 1, 0)
            assert n == np.count_nonzero(lz)
            d = np.zeros(16, np.int32); L.orc_dequant4x4(lz.ctypes.data, d.ctypes.data, qp, 0)
            r = np.zeros(16, np.int32); L.orc_idct4x4(d.ctypes.data, r.ctypes.data)
            # reconstruction error bounded by the quantiser step (Qstep doubles every 6 QP, 0.625 at QP 0)
            step = 0.625 * 2 ** (qp / 6)
            assert np.abs(r - res).max() <= 1.5 * step + 1


def test_transform8x8_round_trip_and_definition(orc):
    """the 8x8 transform pair of the High profile: the inverse is 8.5.13's (checked through the decoder round trips below), the forward one
    must invert it up to the quantiser step, and a constant block must land in the DC coefficient alone"""
    L = orc.lib(); rng = np.random.default_rng(3)
    flat = np.full(64, 7, np.int16); coef = np.zeros(64, np.int32); L.orc_dct8x8(flat.ctypes.data, coef.ctypes.data)
    assert coef[0] == 64 * 7 and not coef[1:].any()
    for qp in (0, 12, 26, 35, 36, 40, 51):
        for _ in range(40):
            res = rng.integers(-255, 256, 64).astype(np.int16)
            L.orc_dct8x8(res.ctypes.data, coef.ctypes.data)
            lz = np.zeros(64, np.int16); n = L.orc_quant8x8(coef.ctypes.data, lz.ctypes.data, qp, 1)
            assert n == np.count_nonzero(lz)
            d = np.zeros(64, np.int32); L.orc_dequant8x8(lz.ctypes.data, d.ctypes.data, qp)
            r = np.zeros(64, np.int32); L.orc_idct8x8(d.ctypes.data, r.ctypes.data)
            step = 0.625 * 2 ** (qp / 6)
            assert np.abs(r - res).max() <= 2.0 * step + 1, (qp, np.abs(r - res).max())


@pytest.mark.parametrize("w,h,kind,qp,slices", [(176, 144, "A", 26, 1), (320, 240, "A", 32, 2), (320, 240, "B", 24, 1), (640, 368, "A", 40, 3), (208, 160, "D", 44, 1)])
def test_high_profile_8x8_transform_streams_decode(orc, w, h, kind, qp, slices):
    """High profile: PPS transform_8x8_mode_flag = 1, inter MBs choose between the 4x4 and the 8x8 transform (transform_size_8x8_flag,
    ctxBlockCat 5 residual blocks, deblocking of the 8x8 transform edges only); FFmpeg must reproduce the oracle's reconstruction"""
    if not avdec.available():
        pytest.skip("no libavcodec")
    o = orc.Encoder(w, h, num_slices=slices, profile=2); o4 = orc.Encoder(w, h, num_slices=slices, profile=2, no_t8x8=1)
    c = Content(kind, w, h)
    aus, recs, n8 = [], [], 0
    for t in range(5):
        f = c.frame(t)
        aus.append(o.encode(f, t == 0, qp)); recs.append(o.recon()); o4.encode(f, t == 0, qp)
        mi = o.mb_info(); t8 = (mi["i16_mode"] >> 2) & 1
        inter = (mi["mb_type"] == 0) | (mi["mb_type"] == 4)
        assert not t8[~inter & (mi["mb_type"] != 5)].any() and t8[mi["mb_type"] == 5].all()      # inter MBs by choice, Intra_8x8 MBs always
        assert ((mi["cbp"][(t8 == 1) & inter] & 15) != 0).all()
        for m in np.flatnonzero(t8):                               # the four nnz of an 8x8 block carry its level count
            assert all(len(set(mi["nnz"][m][4 * b: 4 * b + 4])) == 1 for b in range(4))
        n8 += int(t8.sum())
        assert not ((o4.mb_info()["i16_mode"] >> 2) & 1).any()
    assert n8 > 0 or kind == "D"
    dec = avdec.decode_stream(aus)
    assert len(dec) == 5 and all(np.array_equal(d, r) for d, r in zip(dec, recs))


def test_high_profile_random_geometries_decode(orc):
    """seeded sweep of High-profile streams (odd sizes, QP 0..51, 1-3 slices, both search ranges, contents A / B / D) through the independent decoder"""
    if not avdec.available():
        pytest.skip("no libavcodec")
    rng = np.random.default_rng(77); total8 = 0
    for trial in range(30):
        w = int(rng.integers(1, 24)) * 16 + int(rng.integers(0, 8)) * 2; h = int(rng.integers(1, 16)) * 16 + int(rng.integers(0, 8)) * 2
        qp = int(rng.choice([0, 8, 18, 26, 33, 36, 41, 47, 51])); slices = int(rng.integers(1, 4)); sr = int(rng.choice([16, 32]))
        kind = str(rng.choice(["A", "B", "D"]))
        o = orc.Encoder(w, h, num_slices=slices, search_range=sr, profile=2); c = Content(kind, w, h)
        aus, recs = [], []
        for t in range(4):
            aus.append(o.encode(c.frame(t), t == 0, qp)); recs.append(o.recon()); total8 += int(((o.mb_info()["i16_mode"] >> 2) & 1).sum())
        dec = avdec.decode_stream(aus)
        assert len(dec) == 4 and all(np.array_equal(a, b) for a, b in zip(dec, recs)), (trial, w, h, qp, slices, sr, kind)
    assert total8 > 100


@pytest.mark.parametrize("w,h,kind,qp,slices", [(176, 144, "A", 26, 1), (320, 240, "A", 40, 2), (320, 240, "B", 24, 1), (208, 160, "D", 33, 3), (48, 32, "A", 30, 1)])
def test_intra8x8_groundwork_streams_decode(orc, w, h, kind, qp, slices):
    """Intra_8x8 (8.3.2: reference sample filtering, nine predictors, mode prediction across Intra_4x4 / Intra_8x8 neighbours, I_NxN with
    transform_size_8x8_flag = 1) exists in the oracle only (OrcConfig.intra8x8, off by default; the CUDA path follows next round). The
    independent decoder must reproduce the reconstruction, key frames and intra MBs of P pictures alike."""
    if not avdec.available():
        pytest.skip("no libavcodec")
    o = orc.Encoder(w, h, num_slices=slices, profile=2, intra8x8=1); c = Content(kind, w, h)
    aus, recs, n8 = [], [], 0
    for t in range(4):
        aus.append(o.encode(c.frame(t), t in (0, 3), qp)); recs.append(o.recon())
        mi = o.mb_info(); i8 = mi["mb_type"] == 5; n8 += int(i8.sum())
        assert (((mi["i16_mode"][i8] >> 2) & 1) == 1).all()
        for m in np.flatnonzero(i8):
            assert all(len(set(mi["i4_mode"][m][4 * b: 4 * b + 4])) == 1 for b in range(4))
    assert n8 > 0
    dec = avdec.decode_stream(aus)
    assert len(dec) == 4 and all(np.array_equal(d, r) for d, r in zip(dec, recs))


def test_sad_satd_definitions(orc):
    L = orc.lib(); rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (16, 32), dtype=np.uint8); b = rng.integers(0, 256, (16, 32), dtype=np.uint8)
    assert L.orc_sad(a.ctypes.data, 32, b.ctypes.data, 32, 16, 16) == int(np.abs(a[:, :16].astype(int) - b[:, :16].astype(int)).sum())
    Hm = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]])
    d = a[:4, :4].astype(int) - b[:4, :4].astype(int)
    assert L.orc_satd4x4(a.ctypes.data, 32, b.ctypes.data, 32) == int(np.abs(Hm @ d @ Hm.T).sum()) // 2
    assert L.orc_satd16x16(a.ctypes.data, 32, a.ctypes.data, 32) == 0


def test_emulation_prevention_vectors(orc):
    L = orc.lib()
    def esc(b):
        i = np.frombuffer(bytes(b), np.uint8).copy(); o = np.zeros(len(b) * 2 + 4, np.uint8)
        n = L.orc_escape_rbsp(i.ctypes.data, len(b), o.ctypes.data); return bytes(o[:n])
    assert esc([0, 0, 0]) == bytes([0, 0, 3, 0])
    assert esc([0, 0, 1]) == bytes([0, 0, 3, 1])
    assert esc([0, 0, 4]) == bytes([0, 0, 4])
    assert esc([0, 0, 0, 0, 0]) == bytes([0, 0, 3, 0, 0, 3, 0])
    assert esc([1, 0, 0, 3, 0, 0, 2]) == bytes([1, 0, 0, 3, 3, 0, 0, 3, 2])


def test_level_table(orc):
    L = orc.lib()
    assert L.orc_level_for(1280, 720, 30) == 31 and L.orc_level_for(1920, 1080, 30) == 40
    assert L.orc_level_for(1920, 1080, 60) == 42 and L.orc_level_for(3840, 2160, 30) == 51 and L.orc_level_for(176, 144, 30) == 11


def test_qpel_planes_equal_the_normative_interpolation(orc):
    """the encoder's half-pel planes give the same samples as the direct 8.4.2.2.1 evaluation, including far outside the picture"""
    w, h = 48, 32
    e = orc.Encoder(w, h); c = Content("D", w, h)
    e.encode(c.frame(0), True, 30)
    ref = np.ascontiguousarray(e.plane(3, 0)); L = orc.lib()
    L.orc_dbg_build_halfpel(e.h)             # planes of the current reference (orc_encode builds them for every P frame)
    rng = np.random.default_rng(3)
    for _ in range(3000):
        xq, yq = int(rng.integers(-80, 4 * w + 80)), int(rng.integers(-80, 4 * h + 80))
        assert L.orc_dbg_qpel(e.h, xq, yq) == L.orc_interp_luma(ref.ctypes.data, w, w, h, xq, yq)


def test_colour_conversion_definition(orc):
    L = orc.lib(); rng = np.random.default_rng(4); w, h = 20, 12
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    out = np.zeros(w * h * 3 // 2, np.uint8); L.orc_rgba_to_i420(rgba.ctypes.data, w, h, out.ctypes.data)
    R, G, B = (rgba[..., i].astype(int) for i in range(3))
    assert np.array_equal(out[:w * h].reshape(h, w), ((66 * R + 129 * G + 25 * B + 128) >> 8) + 16)
    m = lambda P: (P.reshape(h // 2, 2, w // 2, 2).sum((1, 3)) + 2) >> 2
    r, g, b = m(R), m(G), m(B)
    assert np.array_equal(out[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), ((-38 * r - 74 * g + 112 * b + 128) >> 8) + 128)
    assert np.array_equal(out[w * h * 5 // 4:].reshape(h // 2, w // 2), ((112 * r - 94 * g - 18 * b + 128) >> 8) + 128)
    nv = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8); o2 = np.zeros_like(nv); L.orc_nv12_to_i420(nv.ctypes.data, w, h, o2.ctypes.data)
    assert np.array_equal(o2[w * h:w * h * 5 // 4], nv[w * h::2]) and np.array_equal(o2[w * h * 5 // 4:], nv[w * h + 1::2])


def test_static_content_converges_to_skip(orc):
    w, h = 160, 96
    e = orc.Encoder(w, h); c = Content("C", w, h)
    idr = e.encode(c.frame(0), True, 30)
    for t in range(1, 6):
        au = e.encode(c.frame(t), False, 30)
    assert (e.mb_info()["mb_type"] == 3).mean() > 0.8 and len(au) < len(idr) // 20


@needs_decoder
def test_intra4x4_every_mode_is_used_and_decodes(orc):
    """Intra_4x4 (8.3.1.2): across a few contents every one of the nine predictors gets chosen somewhere, I_NxN syntax
    (prev_intra4x4_pred_mode / rem_intra4x4_pred_mode, Intra cbp table) decodes to the oracle's reconstruction"""
    used = np.zeros(9, bool); w, h = 192, 128
    for kind, qp in (("A", 24), ("B", 30), ("D", 36)):
        e = orc.Encoder(w, h, num_slices=2); c = Content(kind, w, h)
        au = e.encode(c.frame(0), True, qp); mi = e.mb_info()
        i4 = mi["mb_type"] == 2
        assert i4.any()
        used[np.unique(mi["i4_mode"][i4])] = True
        assert (mi["i4_mode"][mi["mb_type"] == 1] == 0).all()
        dec = avdec.decode_stream([au])
        assert np.array_equal(dec[0], e.recon())
    assert used.all(), used


def test_intra4x4_predictors_against_the_formulas(orc):
    """orc_pred_i4 vs a direct numpy transcription of 8.3.1.2.1-9 for the modes with closed forms (V, H, DC, DDL, VL, HU)"""
    import ctypes as C
    L = orc.lib(); L.orc_pred_i4.restype = C.c_int; L.orc_pred_i4.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(5)
    for _ in range(100):
        img = rng.integers(0, 256, (8, 16), dtype=np.uint8); st = 16; bx, by = 4, 2
        T = img[by - 1, bx:bx + 8].astype(int); Lf = img[by:by + 4, bx - 1].astype(int)
        def run(mode, avail=15):
            out = np.zeros(16, np.uint8); assert L.orc_pred_i4(img.ctypes.data + by * st + bx, st, mode, avail, out.ctypes.data) == 1
            return out.reshape(4, 4)
        assert np.array_equal(run(0), np.tile(T[:4], (4, 1)))
        assert np.array_equal(run(1), np.tile(Lf[:, None], (1, 4)))
        assert (run(2) == (T[:4].sum() + Lf.sum() + 4) >> 3).all()
        ddl = np.array([[(T[x + y] + 2 * T[x + y + 1] + T[min(x + y + 2, 7)] + 2) >> 2 for x in range(4)] for y in range(4)])
        assert np.array_equal(run(3), ddl)
        vl = np.array([[((T[x + (y >> 1)] + T[x + (y >> 1) + 1] + 1) >> 1) if y % 2 == 0 else ((T[x + (y >> 1)] + 2 * T[x + (y >> 1) + 1] + T[x + (y >> 1) + 2] + 2) >> 2)
                        for x in range(4)] for y in range(4)])
        assert np.array_equal(run(7), vl)
        T2 = T.copy(); T2[4:] = T2[3]                  # top-right unavailable: p[4..7,-1] := p[3,-1]
        assert np.array_equal(run(3, 7), np.array([[(T2[x + y] + 2 * T2[x + y + 1] + T2[min(x + y + 2, 7)] + 2) >> 2 for x in range(4)] for y in range(4)]))
        out = np.zeros(16, np.uint8)
        assert L.orc_pred_i4(img.ctypes.data + by * st + bx, st, 4, 3, out.ctypes.data) == 0     # diagonal down-right needs the corner


@needs_decoder
def test_p8x8_partitions_decode_and_predictor_estimate_pays(orc):
    """P_8x8 macroblocks (sub_mb_type P_L0_8x8, partition MV prediction 8.4.1.3 / 6.4.11.7, per-partition bS) occur and decode;
    switching them off (oracle-only flag) still decodes, so both syntaxes are exercised"""
    w, h = 320, 192
    for no8 in (0, 1):
        e = orc.Encoder(w, h, num_slices=2, search_range=32, no_p8x8=no8); c = Content("A", w, h)
        aus, recs, n8 = [], [], 0
        for t in range(5):
            aus.append(e.encode(c.frame(t), t == 0, 22)); recs.append(e.recon())
            mi = e.mb_info(); n8 += int((mi["mb_type"] == 4).sum())
            inter = np.isin(mi["mb_type"], (0, 3))
            mv8 = mi["i4_mode"].view("<i2").reshape(-1, 4, 2)          # union with mv8[4][2]
            assert (mv8[inter] == mi["mv"][inter][:, None, :]).all()    # 16x16 / skip MBs carry their vector in all four partitions
        assert (n8 > 0) == (no8 == 0)
        dec = avdec.decode_stream(aus)
        assert len(dec) == 5 and all(np.array_equal(d, r) for d, r in zip(dec, recs))


def consumer_header_scan(au):
    """What the repo's own decoder-side consumer does with an access unit before feeding its ASIC
    (reference video_decoder/VideoDecoderNetint.cpp:737-792 DeviceDecSessionWrite, :794-841 FindNextNonVclNalu, :843-860
    FindNalStartCode): walk the non-VCL NALs from the front, collect SPS/PPS into a 4 KB header buffer, stop at the first VCL NAL.
    Returns (sps_found, pps_found, header_bytes, type of the NAL the scan stopped at)."""
    buf, sps, pps, hdr = bytes(au), False, False, 0
    def find_start(b):
        i = 0
        while not (b[i:i + 3] == b"\0\0\1" or b[i:i + 4] == b"\0\0\0\1"):
            i += 1
            if i + 3 > len(b):
                return -1
        return i
    while len(buf) > 3:
        i = find_start(buf)
        if i < 0:
            return sps, pps, hdr, -1
        if buf[i + 2] != 1:
            i += 1
        i += 3
        t = buf[i] & 0x1f
        if 1 <= t <= 5:                       # SLICE .. IDR_SLICE: VCL, stop
            return sps, pps, hdr, t
        while buf[i:i + 3] not in (b"\0\0\0", b"\0\0\1"):
            i += 1
            if i + 3 > len(buf):
                i = len(buf); break
        sps |= t == 7; pps |= t == 8
        if t in (7, 8):
            hdr += i
        buf = buf[i:]
        if sps and pps:
            nxt = find_start(buf)
            return sps, pps, hdr, (buf[nxt + (4 if buf[nxt + 2] != 1 else 3)] & 0x1f) if nxt >= 0 else -1
    return sps, pps, hdr, -1


def test_downstream_consumer_finds_the_parameter_sets(orc):
    """every IDR access unit carries SPS and PPS in front of the first VCL NAL (4-byte start codes, well inside
    the consumer's 4 KB header buffer); P access units start with a VCL NAL"""
    e = orc.Encoder(320, 180, num_slices=3); c = Content("A", 320, 180)
    idr = e.encode(c.frame(0), True, 26); p = e.encode(c.frame(1), False, 26)
    sps, pps, hdr, stop = consumer_header_scan(idr)
    assert sps and pps and 0 < hdr <= 4096 and stop == 5
    assert idr[:5] == b"\0\0\0\1\x67"
    assert consumer_header_scan(p) == (False, False, 0, 1)
    case = next(c for c in GOLDEN if "stream_hex" in c)
    assert consumer_header_scan(bytes.fromhex(case["stream_hex"][0]))[:2] == (True, True)


@needs_decoder
def test_scene_change_turns_a_p_frame_into_an_idr(orc):
    """a cut to unrelated content makes >= 2/5 of the macroblocks intra after the motion search: the picture is coded as an IDR
    (SPS + PPS in front, frame_num restarts) and the stream stays decodable; with the detector off it remains a P picture, and so does a
    cut within ORC_SC_MIN_DISTANCE = 10 pictures of the last IDR"""
    w, h = 256, 160
    a, b = Content("A", w, h, seed=1), Content("A", w, h, seed=99)
    for detect in (1, 0):
        e = orc.Encoder(w, h, scene_change=detect)
        frames = [a.frame(t) for t in range(3)] + [b.frame(t) for t in range(3, 11)] + [a.frame(11), a.frame(12)]      # cuts at 3 and 11
        aus, recs, kinds = [], [], []
        for t, f in enumerate(frames):
            aus.append(e.encode(f, t == 0, 28)); recs.append(e.recon()); kinds.append(e.last_was_idr())
        assert kinds == [True] + [False] * 10 + [bool(detect), False]
        assert (aus[11][4] == 0x67) == bool(detect) and aus[3][4] == 0x61
        dec = avdec.decode_stream(aus)
        assert len(dec) == 13 and all(np.array_equal(d, r) for d, r in zip(dec, recs))


@needs_decoder
@pytest.mark.parametrize("profile", [0, 2])
def test_background_detection_streams_decode_and_save_bits(orc, profile):
    """bEnableBackgroundDetection (VideoEncoderOpenH264.cpp:282), OUR definition: on a static scene with sensor noise and one moving
    object the static macroblocks are skipped -- the stream still decodes to the oracle's reconstruction, costs far fewer bits, and the moving
    object is still coded; complexity modes only remove tools (LOW: no Intra_4x4, no P_8x8)"""
    w, h, qp = 320, 192, 28
    c = Content("E", w, h)
    on, off = orc.Encoder(w, h, background_detection=1, profile=profile), orc.Encoder(w, h, profile=profile)
    aus, recs, b_on, b_off = [], [], 0, 0
    for t in range(8):
        f = c.frame(t)
        a = on.encode(f, t == 0, qp); aus.append(a); recs.append(on.recon())
        if t:
            b_on += len(a); b_off += len(off.encode(f, False, qp))
        else:
            off.encode(f, True, qp)
    dec = avdec.decode_stream(aus)
    assert len(dec) == 8 and all(np.array_equal(d, r) for d, r in zip(dec, recs))
    assert b_on < 0.7 * b_off, (b_on, b_off)
    mt = on.mb_info()["mb_type"]
    assert (mt == 3).mean() > 0.8 and (mt != 3).sum() >= 2              # mostly P_Skip, the moving object is not
    lo = orc.Encoder(w, h, complexity=0, profile=profile)
    a = Content("A", w, h)
    for t in range(3):
        lo.encode(a.frame(t), t == 0, qp)
        assert not np.isin(lo.mb_info()["mb_type"], [2, 4]).any()


def i8_filtered_edge(rec, st, o, avail):
    """numpy model of what the CUDA path does for one Intra_8x8 block: the raw edge L7..L0 X T0..T15 with the substitutions for missing
    neighbours, then the reference sample filter of 8.3.2.2.1 as ONE three-tap pass with doubled ends; returns (E', top, left, corner)"""
    top, left, corner, tr = bool(avail & 1), bool(avail & 2), bool(avail & 4), bool(avail & 8)
    t = [int(rec[o - st + i]) if top and (i < 8 or tr) else (int(rec[o - st + 7]) if top else 128) for i in range(16)]
    l = [int(rec[o + i * st - 1]) if left else 128 for i in range(8)]
    x = int(rec[o - st - 1]) if corner else 128
    raw = l[::-1] + [x] + t                                    # E index: L[i] = 7 - i, X = 8, T[i] = 9 + i
    E = list(raw)
    for k in range(25):
        a = raw[k - 1] if k > 0 else raw[0]
        d = raw[k + 1] if k < 24 else raw[24]
        if k == 7 and not corner: d = raw[7]                   # L'[0] without the corner: (3 l0 + l1 + 2) >> 2
        if k == 9 and not corner: a = raw[9]                   # T'[0] without the corner
        f = (a + 2 * raw[k] + d + 2) >> 2
        if k < 8: E[k] = f if left else 128
        elif k == 8:
            E[k] = f if (corner and top and left) else ((3 * x + t[0] + 2) >> 2 if corner and top else (3 * x + l[0] + 2) >> 2 if corner and left else x)
        else: E[k] = f if top else 128
    return E, top, left, corner


def test_intra8x8_lookup_tables_reproduce_the_oracle_predictors(orc):
    """media_b200/csrc/i8_tables.cuh (tools/make_i8_tables.py): the nine Intra_8x8 predictors as lookups into the one / two / three-tap tables of the
    filtered edge, against the oracle's direct transcription of 8.3.2.2.1-10 (orc_pred_i8), for every availability combination"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_i8_tables as mk
    gen = open(os.path.join(ROOT, "media_b200", "csrc", "i8_tables.cuh")).read()
    words = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", gen)]
    flat = [b for w in words for b in (w & 255, (w >> 8) & 255, (w >> 16) & 255, w >> 24)]
    assert flat == [v for m in range(9) for v in mk.TABLE[m]], "i8_tables.cuh is stale: run tools/make_i8_tables.py"
    rng = np.random.default_rng(88)
    st = 40
    L = orc.lib()
    for trial in range(60):
        rec = rng.integers(0, 256, st * 24, dtype=np.uint8) if trial % 3 else np.full(st * 24, rng.integers(0, 256), np.uint8)
        o = 9 * st + 9
        for avail in (0, 1, 2, 3, 7, 15, 5 | 2, 1 | 8, 3 | 8):
            if (avail & 4) and (avail & 3) != 3: continue
            if (avail & 8) and not (avail & 1): continue
            E, top, left, corner = i8_filtered_edge(rec, st, o, avail)
            Ed = [E[0]] + E + [E[24], E[24]]                                    # E[-1] .. E[26] with doubled ends
            F = {}
            for k in range(25):
                F[k] = E[k]; F[64 + k] = (Ed[k] + 2 * Ed[k + 1] + Ed[k + 2] + 2) >> 2
                if k < 24: F[32 + k] = (E[k] + E[k + 1] + 1) >> 1
            sT, sL = sum(E[9:17]), sum(E[0:8])
            F[95] = (sT + sL + 8) >> 4 if top and left else (sT + 4) >> 3 if top else (sL + 4) >> 3 if left else 128
            for m in range(9):
                want = np.zeros(64, np.uint8)
                ok = L.orc_pred_i8(rec.ctypes.data + o, st, m, avail, want.ctypes.data)
                allowed = m == 2 or ((m in (0, 3, 7)) and top) or ((m in (1, 8)) and left) or ((m in (4, 5, 6)) and top and left and corner)
                assert bool(ok) == allowed, (m, avail)
                if ok:
                    got = np.array([F[i] for i in mk.TABLE[m]], np.uint8)
                    assert np.array_equal(got, want), (trial, avail, m, got.reshape(8, 8), want.reshape(8, 8))


def test_key_pictures_with_their_own_slice_count_decode(orc):
    """OrcConfig.key_slices: pictures requested as key pictures take their own slice layout (here 5 against 2 for P pictures, CAVLC and CABAC);
    the independent decoder reproduces the oracle's reconstruction, and the NAL counts follow the picture kind"""
    if not avdec.available():
        pytest.skip("no libavcodec")
    w, h, qp = 176, 144, 28
    for profile in (0, 1):
        e = orc.Encoder(w, h, num_slices=2, key_slices=5, profile=profile); c = Content("A", w, h)
        aus, recs = [], []
        for t in range(5):
            au = e.encode(c.frame(t), t in (0, 3), qp); aus.append(au); recs.append(e.recon().copy())
            starts = [i for i in range(len(au) - 4) if au[i:i + 4] == b"\x00\x00\x00\x01"]
            assert sum((au[i + 4] & 31) in (1, 5) for i in starts) == (5 if t in (0, 3) else 2)
        dec = avdec.decode_stream(aus)
        assert len(dec) == 5
        for t in range(5):
            assert np.array_equal(dec[t], recs[t]), (profile, t)
