"""tools/profile_steps.py -- per-kernel CUDA-event times of IDR and P steps for a few session counts (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from media_b200 import enc
import bench

S_list = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "32", "96"])]
L = enc.lib()
pool = bench.make_pool()
fb = bench.W * bench.H * 3 // 2
dpool = []
for f in pool:
    p = L.b200enc_dev_alloc(0, fb); enc.check(L.b200enc_dev_upload(0, p, f.ctypes.data, fb)); dpool.append(p)
for S in S_list:
    ss = [enc.Session(bench.W, bench.H, fps=30, bitrate=4_000_000, gop=300, const_qp=-1, device=0) for _ in range(S)]
    b = enc.Batch(0, ss)
    b.set_profiling(True)
    for k in range(8):
        t0 = time.perf_counter()
        sizes = b.encode_ptrs([dpool[bench.pool_index(k, i)] for i in range(S)], 1)
        wall = (time.perf_counter() - t0) * 1e3
        kt = b.kernel_times()
        mi = ss[0].stage("mbinfo")
        if k in (0, 1, 4, 7):
            print(f"S={S} step {k}: wall {wall:.3f} ms dev {b.kernel_ms():.3f} ms qp {b._infos[0].qp} bytes0 {sizes[0]} types {np.bincount(mi['mb_type'], minlength=4).tolist()} | "
                  + " ".join(f"{n[2:]}={ms:.3f}" for n, ms in kt))
    for s in ss:
        s.close()
    b.close()
