"""tools/cabac_coder_bench.py -- the CABAC coder kernel alone on the bin list of a real 1080p P frame (made by the oracle here or
loaded from --bins): device ms per launch, cycles per bin, and with a -DCABAC_TIMING build (B200ENC_LIB) the producer / consumer phases."""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--bins", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "media_b200", "csrc", "variants", "bins_1080p_p.npy"))
ap.add_argument("--make", action="store_true", help="produce the bin list with the oracle (CPU) and save it")
ap.add_argument("--copies", type=int, default=32)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
if a.make:
    from oracle import orc_py
    from media_b200.synth import Content
    e = orc_py.Encoder(1920, 1080, profile=1); c = Content("A", 1920, 1080)
    e.encode(c.frame(0), True, 40); e.encode(c.frame(1), False, 36)
    np.save(a.bins, e.slice_bins(0)); print("saved", a.bins); sys.exit(0)
from media_b200 import enc
L = enc.lib()
L.b200k_cabac_code_bench.restype = C.c_int
L.b200k_cabac_code_bench.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
bins = np.load(a.bins)
ctx = bins & 1023; reg = (ctx < 0x3F8) & (ctx != 276)
nrec = int(np.where(reg, (bins >> 11) + 1, 1).sum())
ms = C.c_float(); st = (C.c_longlong * 8)()
for copies in (1, a.copies):
    enc.check(L.b200k_cabac_code_bench(0, bins.ctypes.data, bins.size, 36, 1, a.reps, copies, C.byref(ms), st))
    runs = a.reps * copies
    print(f"copies={copies}: {ms.value:.3f} ms per launch, {bins.size} entries, {nrec} bins -> {ms.value * 1e6 / nrec:.1f} ns per bin")
    if st[2]:
        print(f"  consumer: {st[0] / st[2]:.1f} cycles/bin total, {st[1] / st[2]:.1f} waiting;  producer: {st[3] / st[2]:.1f} cycles/bin total, "
              f"{st[4] / st[2]:.1f} waiting for room, {st[5] / st[2]:.1f} in turn loops; {st[7] / max(1, st[6]):.2f} turns per 32 entries, {st[3] / max(1, st[6]):.0f} cycles per step")
