"""tools/load_kernel_times.py [profile] [sessions] [groups] -- per-kernel CUDA-event times of a batch step UNDER LOAD: `groups` batches of
sessions / groups 1080p sessions each driven by an own host thread (the bench's regime), profiling on in every batch; prints the
mean time between the events around each kernel (queueing behind the other batches' kernels included), averaged over steps and batches."""
import sys, os, threading, collections, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from media_b200 import enc
from media_b200.synth import Content
profile = int(sys.argv[1]) if len(sys.argv) > 1 else 0
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8
steps, warm = 12, 4
w, h = 1920, 1080
L = enc.lib()
c = Content("A", w, h)
pool = [np.ascontiguousarray(c.frame(t)).ravel() for t in range(8)]
fb = pool[0].size
dpool = []
for f in pool:
    p = L.b200enc_dev_alloc(0, fb); enc.check(L.b200enc_dev_upload(0, p, f.ctypes.data, fb)); dpool.append(p)
groups = []
for g in range(G):
    ss = [enc.Session(w, h, bitrate=4_000_000, gop=300, device=0, profile=profile, num_slices=0 if profile else 1) for _ in range(S // G)]
    groups.append((ss, enc.Batch(0, ss)))
acc = [collections.defaultdict(float) for _ in range(G)]
def run(gi):
    ss, b = groups[gi]
    for k in range(warm + steps):
        if k == warm: b.set_profiling(True)
        b.encode_ptrs([dpool[(k + i) % len(dpool)] for i in range(len(ss))], 1)
        if k >= warm:
            for n, ms in b.kernel_times(): acc[gi][n] += ms
t0 = None
ths = [threading.Thread(target=run, args=(g,)) for g in range(G)]
t0 = time.perf_counter()
for t in ths: t.start()
for t in ths: t.join()
el = time.perf_counter() - t0
tot = collections.defaultdict(float)
for a in acc:
    for n, v in a.items(): tot[n] += v / (steps * G)
print(f"profile {profile}: {S} sessions in {G} batches, {S * (warm + steps) / el:.0f} frames/s incl. warm-up; mean ms per kernel per batch step under load (sum {sum(tot.values()):.2f}):")
print("  " + ", ".join(f"{n} {v:.2f}" for n, v in tot.items()))
