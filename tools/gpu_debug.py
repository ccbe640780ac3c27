"""tools/gpu_debug.py -- stage-by-stage comparison of the CUDA path with the CPU oracle (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import orc_py
from media_b200 import enc
from media_b200.synth import Content


def first_diff(name, a, b, limit=5):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"   {name}: SHAPE {a.shape} vs {b.shape}"); return False
    if np.array_equal(a, b):
        return True
    idx = np.argwhere(a != b)
    print(f"   {name}: {len(idx)} diffs; first {[tuple(i) for i in idx[:limit]]} gpu={[a[tuple(i)] for i in idx[:limit]]} orc={[b[tuple(i)] for i in idx[:limit]]}")
    return False


def run(w, h, kind, nf, qp, ns, sr, verbose=True):
    g = enc.Session(w, h, const_qp=qp, num_slices=ns, search_range=sr, gop=1000, debug=1)
    o = orc_py.Encoder(w, h, num_slices=ns, search_range=sr)
    c = Content(kind, w, h)
    ok_all = True
    for t in range(nf):
        f = c.frame(t)
        t0 = time.time(); bs_g, info = g.encode(f); tg = time.time() - t0
        bs_o = o.encode(f, t == 0, qp)
        ok = bs_g == bs_o
        rec_ok = np.array_equal(g.recon(), o.recon())
        print(f"{w}x{h} {kind} qp{qp} sl{ns} sr{sr} frame {t}: gpu {len(bs_g)} B orc {len(bs_o)} B  bitstream {'OK' if ok else 'DIFF'}  recon {'OK' if rec_ok else 'DIFF'}  "
              f"(gpu call {tg*1e3:.2f} ms, kernels {g.kernel_ms():.3f} ms)")
        if not (ok and rec_ok):
            ok_all = False
            mbw = g.mbw
            if t > 0:
                for lv in (2, 1, 0):
                    first_diff(f"me{lv}", g.stage(f"me{lv}"), o.me_level(lv))
                first_diff("inter_cost", g.stage("inter_cost"), o.inter_cost())
            gi, oi = g.stage("mbinfo"), o.mb_info()
            for fld in ("mb_type", "i16_mode", "chroma_mode", "cbp", "mv", "nnz"):
                first_diff("mbinfo." + fld, gi[fld], oi[fld])
            gc, oc = g.stage("mbcoef"), o.mb_coef()
            for fld in ("luma", "luma_dc", "chroma_dc", "chroma_ac"):
                first_diff("coef." + fld, gc[fld], oc[fld])
            ny = g.mbw * g.mbh * 256
            for nm, which in (("src", 0), ("rec_pre", 1), ("rec", 2)):
                gs = g.stage(nm)
                for comp, (off, sz, ww) in enumerate(((0, ny, mbw * 16), (ny, ny // 4, mbw * 8), (ny * 5 // 4, ny // 4, mbw * 8))):
                    first_diff(f"{nm}[{comp}]", gs[off:off + sz].reshape(-1, ww), o.plane(which, comp))
            if not ok:
                n = min(len(bs_g), len(bs_o))
                d = next((i for i in range(n) if bs_g[i] != bs_o[i]), n)
                print(f"   bitstream first diff at byte {d}: gpu {bs_g[d:d+8].hex()} orc {bs_o[d:d+8].hex()}")
            break
    g.close(); o.close()
    return ok_all


if __name__ == "__main__":
    cases = [(64, 48, "A", 3, 26, 1, 16), (176, 144, "A", 4, 26, 1, 16), (176, 144, "D", 3, 30, 2, 16), (320, 180, "B", 4, 26, 3, 32),
             (322, 182, "A", 3, 10, 2, 64), (640, 360, "C", 3, 26, 4, 16), (1280, 720, "A", 3, 26, 1, 16), (1920, 1080, "B", 3, 30, 1, 16)]
    res = [run(*c) for c in cases]
    print("SUMMARY", res)
    sys.exit(0 if all(res) else 1)
