"""tools/make_cabac_tables.py -- writes the CABAC tables of ITU-T H.264 (context initialisation m,n of Tables 9-12..9-23 for
ctxIdx 0..459, I slices and cabac_init_idc 0; rangeTabLPS of Table 9-44; transIdxLPS / transIdxMPS of Table 9-45) as
oracle/cabac_tables.h and as the product's own copy media_b200/csrc/cabac_tables.cuh.

The numbers are normative constants of the standard. To rule out transcription slips they are read here from the copies inside
the independent decoder the tests use (libavcodec in the opencv wheel: cabac_context_init_I/PB and ff_h264_cabac_tables), located
by their well-known leading values; tests/test_oracle.py::test_cabac_tables_match_the_independent_decoder re-checks the committed
headers against that binary, and every CABAC stream the oracle writes is decoded by it bit-exactly.
"""
import glob
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCTX = 460


def read_from_decoder():
    import cv2
    d = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
    blob = open(glob.glob(os.path.join(d, "libavcodec-*"))[0], "rb").read()
    lead = np.array([20, -15, 2, 54, 3, 74, 20, -15, 2, 54, 3, 74, -28, 127, -23, 104, -6, 53, -1, 54, 7, 51], np.int8).tobytes()
    tabs, at = [], blob.find(lead)             # cabac_context_init_I[1024][2] and cabac_context_init_PB[3][1024][2], all start alike
    while at >= 0:
        tabs.append(np.frombuffer(blob[at: at + 2048], np.int8).reshape(1024, 2)[:NCTX]); at = blob.find(lead, at + 1)
    assert len(tabs) == 4
    init_i = next(t for t in tabs if t[70].tolist() == [0, 11] and t[72].tolist() == [0, 69])      # Table 9-17, I slices
    init_p = next(t for t in tabs if t[11].tolist() == [23, 33] and t[70].tolist() == [0, 45])      # Tables 9-13 / 9-17, cabac_init_idc 0
    ns = blob.find(bytes([9, 8, 7, 7, 6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5]))   # ff_h264_cabac_tables: norm_shift[512], lps_range[512], mlps_state[256]
    assert ns >= 0
    lps = np.frombuffer(blob[ns + 512: ns + 1024], np.uint8); ml = np.frombuffer(blob[ns + 1024: ns + 1280], np.uint8)
    rng = np.array([[lps[q * 128 + 2 * s] for q in range(4)] for s in range(64)])
    nxt_mps = [int(ml[128 + 2 * s]) // 2 for s in range(64)]; nxt_lps = [int(ml[127 - 2 * s]) // 2 for s in range(64)]
    return init_i, init_p, rng, nxt_lps, nxt_mps


def emit(path, head, qual, names):
    init_i, init_p, rng, nl, nm = read_from_decoder()
    def rows(a, per):
        flat = [str(int(x)) for x in np.asarray(a).ravel()]
        return "\n".join("    " + ", ".join(flat[i:i + per]) + "," for i in range(0, len(flat), per))
    with open(path, "w") as f:
        f.write(head)
        f.write(f"{qual} int8_t {names[0]}[{NCTX} * 2] = {{\n{rows(init_i, 24)}\n}};\n")
        f.write(f"{qual} int8_t {names[1]}[{NCTX} * 2] = {{\n{rows(init_p, 24)}\n}};\n")
        f.write(f"{qual} uint8_t {names[2]}[64 * 4] = {{\n{rows(rng, 32)}\n}};\n")
        f.write(f"{qual} uint8_t {names[3]}[64] = {{\n{rows(nl, 32)}\n}};\n")
        f.write(f"{qual} uint8_t {names[4]}[64] = {{\n{rows(nm, 32)}\n}};\n")


if __name__ == "__main__":
    doc = ("ITU-T H.264 CABAC constants: (m,n) of Tables 9-12..9-23 for ctxIdx 0..459 (I slices; P slices with cabac_init_idc 0),\n"
           " * rangeTabLPS (Table 9-44), transIdxLPS / transIdxMPS (Table 9-45). Written by tools/make_cabac_tables.py.")
    emit(os.path.join(ROOT, "oracle", "cabac_tables.h"),
         f"/* oracle/cabac_tables.h -- TEST INFRASTRUCTURE ONLY (see orc.h).\n * {doc} */\n#include <stdint.h>\n",
         "static const", ["CABAC_INIT_I", "CABAC_INIT_P0", "CABAC_RANGE_LPS", "CABAC_NEXT_LPS", "CABAC_NEXT_MPS"])
    emit(os.path.join(ROOT, "media_b200", "csrc", "cabac_tables.cuh"),
         f"/* media_b200/csrc/cabac_tables.cuh -- the product's own copy (never includes oracle/).\n * {doc} */\n#pragma once\n#include <stdint.h>\n",
         "__constant__", ["c_cabac_init_i", "c_cabac_init_p0", "c_cabac_range_lps", "c_cabac_next_lps", "c_cabac_next_mps"])
