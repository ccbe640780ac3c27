#!/usr/bin/env python
"""bench.py -- 1080p H.264 encode throughput of the B200 path (BASELINE.json metric), one JSON line on stdout.

Workload (config.workload): S concurrent 1920x1080 sessions per GPU, Baseline IPPP, CBR 4 Mbps at 30 fps each
(BASELINE.json configs[1] scaled to the session count the metric's "real-time sessions" figure needs). A step is
one frame for every session of the GPU: S frames. `value` = frames/s with the input frames resident in HBM;
`e2e` = the same through the C ABI with HOST (pinned) input frames and host-visible bitstreams.
`--impl reference` times the CPU restatement of the path (oracle/, one single-threaded encoder per host core, as
the reference configures openh264 at video_codec/VideoEncoderOpenH264.cpp:294); libopenh264 itself is not in the
image, so its `kind` is "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

W, H, FPS, BITRATE, GOP = 1920, 1080, 30, 4_000_000, 300
METRIC = "1080p H.264 encode frames/s per GPU"
POOL_FRAMES = 16
PROFILE = 0   # 0 Constrained Baseline / CAVLC (BASELINE.json's configuration), 1 Main / CABAC, 2 High / CABAC
FMT, SLICES, SR, CQP, KIND, LABEL = 0, 1, 16, -1, "A", "Baseline CAVLC IPPP, CBR 4 Mbps @30fps each, gop 300, search +-16, 1 slice, content A (moving texture)"
# Secondary workloads (BASELINE.json configs 2-4); the default, and what the driver runs, is the 1080p session workload above.
WORKLOADS = {
    "1080p": {},
    "1080p-main": dict(PROFILE=1, SLICES=4, LABEL="Main profile CABAC IPPP (the wrapper's iEntropyCodingModeFlag = 1 with profile main), CBR 4 Mbps @30fps each, gop 300, search +-16, 4 slices (the CABAC default at 1080p), content A (moving texture)"),
    "1080p-main-1slice": dict(PROFILE=1, LABEL="Main profile CABAC IPPP, CBR 4 Mbps @30fps each, gop 300, search +-16, 1 slice, content A (moving texture)"),
    "1080p-high": dict(PROFILE=2, SLICES=4, LABEL="High profile CABAC IPPP with the 8x8 transform on inter MBs (profile high), CBR 4 Mbps @30fps each, gop 300, search +-16, 4 slices, content A (moving texture)"),
    "portrait720": dict(W=720, H=1280, CQP=26, sessions=1, groups=1, METRIC="720x1280 H.264 encode frames/s per GPU",
                        LABEL="ONE 720x1280 portrait stream (config 1): Baseline CAVLC, const QP 26, gop 300, search +-16, 1 slice, content A; latency-bound by construction"),
    "single": dict(sessions=1, groups=1, LABEL="ONE 1080p stream (config 2): Baseline CAVLC IPPP, CBR 4 Mbps @30fps, gop 300, search +-16, 1 slice, content A; latency-bound by construction"),
    "rgba720": dict(W=1280, H=720, FMT=2, KIND="B", sessions=64, groups=2, BITRATE=2_000_000, METRIC="720p RGBA->I420 + H.264 encode frames/s per GPU",
                    LABEL="RGBA8888 cloud-phone framebuffers (config 3), on-GPU RGBA->I420 + encode, Baseline CAVLC IPPP, CBR 2 Mbps @30fps each, gop 300, search +-16, content B (screen-like)"),
    "4k": dict(W=3840, H=2160, SLICES=8, SR=64, CQP=26, sessions=1, groups=1, METRIC="2160p H.264 encode frames/s per GPU",
               LABEL="ONE 3840x2160 stream (config 4): 8 slices, search +-64 + quarter-pel, const QP 26, content A; latency-bound by construction"),
}


def apply_workload(args):
    g = globals()
    for k, v in WORKLOADS[args.workload].items():
        if k in ("sessions", "groups"):
            if getattr(args, k) is None:
                setattr(args, k, v)
        else:
            g[k] = v
    if getattr(args, "slices", None) is not None:
        g["SLICES"] = args.slices
        g["LABEL"] += f" [slices overridden: {args.slices}]"
    if args.sessions is None:
        args.sessions = 128
    if args.groups is None:
        args.groups = 4


def frame_bytes():
    return W * H * 4 if FMT == 2 else W * H * 3 // 2


def new_session(enc, dev, **kw):
    return enc.Session(W, H, fps=FPS, bitrate=BITRATE, gop=GOP, const_qp=CQP, num_slices=SLICES, search_range=SR, input_format=FMT, device=dev, profile=PROFILE, **kw)


def make_pool(n=POOL_FRAMES):
    from media_b200.synth import Content, i420_to_rgba
    c = Content(KIND, W, H)
    fr = [c.frame(t) for t in range(n)]
    return [np.ascontiguousarray(i420_to_rgba(f, W, H)).ravel() for f in fr] if FMT == 2 else fr


def pool_index(step, sess, n=POOL_FRAMES):
    """ping-pong walk through the pool so consecutive frames of a session stay temporally adjacent"""
    t = (step + 3 * sess) % (2 * n - 2)
    return t if t < n else 2 * n - 2 - t


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop = gpu, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self.stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([x.strip() for x in line.split(",")])
        p.terminate()

    def summary(self):
        self.stop.set()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_b200(args):
    import torch
    from media_b200 import enc
    rank, local_rank, world = dist_env()
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if use_dist else 0
    try:        # keep this rank's caller threads and its pinned frame pool on the NUMA node of its GPU (8 ranks share a two-socket host)
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(dev))
    except Exception:
        pass
    S = args.sessions
    L = enc.lib()
    pool = make_pool()
    fb = frame_bytes()
    # device-resident pool and pinned host pool
    import ctypes as C
    dpool = []
    for f in pool:
        p = L.b200enc_dev_alloc(dev, fb)
        assert p, "device allocation failed"
        enc.check(L.b200enc_dev_upload(dev, p, f.ctypes.data, fb))
        dpool.append(p)
    hpool = []
    for f in pool:
        p = L.b200enc_host_alloc(fb)
        assert p, "pinned allocation failed"
        C.memmove(p, f.ctypes.data, fb)
        hpool.append(p)

    G = max(1, min(args.groups, S))          # independent batches (own stream each) driven by G host threads
    group_sizes = [S // G + (1 if i < S % G else 0) for i in range(G)]

    def new_groups():
        groups, sid = [], 0
        for n in group_sizes:
            ss = [new_session(enc, dev) for _ in range(n)]
            groups.append((ss, enc.Batch(dev, ss), list(range(sid, sid + n)))); sid += n
        return groups

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def run_group(grp, ptr_pool, device_input, first, count, acc):
        ss, batch, ids = grp
        dev_ms = launches = out_bytes = 0
        for k in range(first, first + count):
            sizes = batch.encode_ptrs([ptr_pool[pool_index(k, i)] for i in ids], device_input)
            dev_ms += batch.kernel_ms(); launches += batch.launches(); out_bytes += sum(sizes[j] for j in range(len(ids)))
        acc.append((dev_ms, launches, out_bytes))

    def run_all(groups, ptr_pool, device_input, first, count):
        acc = []
        if len(groups) == 1:
            run_group(groups[0], ptr_pool, device_input, first, count, acc)
        else:
            ths = [threading.Thread(target=run_group, args=(g, ptr_pool, device_input, first, count, acc)) for g in groups]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        return max(a[0] for a in acc), sum(a[1] for a in acc), sum(a[2] for a in acc)

    def timed(groups, ptr_pool, device_input, steps, warmup, step0=0):
        run_all(groups, ptr_pool, device_input, step0, warmup)
        barrier()
        t0 = time.perf_counter()
        dev_ms, launches, out_bytes = run_all(groups, ptr_pool, device_input, step0 + warmup, steps)
        barrier()
        el = time.perf_counter() - t0
        if use_dist:
            t = torch.tensor([el, dev_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el, dev_ms = t[0].item(), t[1].item()
        return el, dev_ms, launches, out_bytes

    sampler = ClockSampler(dev); sampler.start()
    groups = new_groups()
    sess = [x for g in groups for x in g[0]]
    batch = groups[0][1]
    el, dev_ms, launches, out_bytes = timed(groups, dpool, 1, args.steps, args.warmup)
    clocks = sampler.summary()
    value = world * S * args.steps / el
    # end to end: host pinned frames in, host-visible bitstreams out
    el_e, dev_ms_e, _, out_bytes_e = timed(groups, hpool, 0, args.steps, 1, step0=args.warmup + args.steps)
    e2e = world * S * args.steps / el_e
    # the reference's threading model: one caller thread per session, each blocked in b200enc_encode (what EncodeOneFrame
    # does); the per-GPU auto_batch scheduler coalesces them. Python threads add overhead, so this is a lower bound.
    thr = None
    if args.threads_e2e:
        for x in sess:
            x.close()
        sess = [new_session(enc, dev, auto_batch=1) for _ in range(S)]
        b0, f0, b1, f1 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        nst = max(4, args.steps)

        def caller(i, first, count):
            bs, n = C.c_void_p(), C.c_uint32()
            for k in range(first, first + count):
                L.b200enc_encode(sess[i].h, hpool[pool_index(k, i)], fb, C.byref(bs), C.byref(n), None)

        def run_callers(first, count):
            ths = [threading.Thread(target=caller, args=(i, first, count)) for i in range(S)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        run_callers(0, 3)
        barrier()
        L.b200enc_scheduler_stats(dev, C.byref(b0), C.byref(f0))
        t0 = time.perf_counter(); run_callers(3, nst); barrier(); el_t = time.perf_counter() - t0
        L.b200enc_scheduler_stats(dev, C.byref(b1), C.byref(f1))
        if use_dist:
            tt = torch.tensor([el_t], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); el_t = tt[0].item()
        thr = {"value": round(world * S * nst / el_t, 2), "unit": "frames/s", "caller_threads": S,
               "avg_batch": round((f1.value - f0.value) / max(1, b1.value - b0.value), 1),
               "note": "one Python thread per session blocked in b200enc_encode (auto_batch scheduler); host frames in, bitstreams out"}
        groups = new_groups(); batch = groups[0][1]; sess += [x for g in groups for x in g[0]]
    # per-kernel shares of one P step over ALL sessions of the GPU in a single batch (CUDA events around each launch on the
    # batch's stream); fresh sessions, so two untimed frames first (IDR + one P)
    for g_ in groups:
        g_[1].close()
    for x in sess:
        x.close()
    sess = [new_session(enc, dev) for _ in range(S)]
    batch = enc.Batch(dev, sess)
    groups = [(sess, batch, list(range(S)))]
    for k in range(3):
        if k == 2:
            batch.set_profiling(True)
        batch.encode_ptrs([dpool[pool_index(k, i)] for i in range(S)], 1)
    kt = batch.kernel_times()
    Sp = S                               # sessions in the profiled batch
    batch.set_profiling(False)
    tot = sum(ms for _, ms in kt) or 1.0
    top = max(kt, key=lambda x: x[1])
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        npx = ((W + 15) // 16 * 16) * ((H + 15) // 16 * 16)
        nmb = npx // 256
        # algorithmic bytes per P frame: src 1.5 + ref 1.5 + recon 1.5 B/px for ME+coding, +3.0 for the in-place deblock pass
        alg_bytes = {"k_me_fine": 4.5 * npx, "k_me_coarse": 2 * 0.3125 * npx, "k_deblock_wave": 3.0 * npx, "k_intra_wave": 3.0 * npx,
                     "k_cavlc_mb": nmb * 864.0, "k_ingest_planar": 3.0 * npx, "k_ingest_rgba": 5.5 * npx, "k_refplanes": 5.0 * npx}.get(top[0], 4.5 * npx) * Sp
        ach = alg_bytes / (top[1] * 1e-3) / 1e9
        gi, clk = C.c_double(), C.c_int()
        L.b200k_vabsdiff4_peak(dev, C.byref(gi), C.byref(clk))
        me_fine = dict(kt).get("k_me_fine", 0.0); me_coarse = dict(kt).get("k_me_coarse", 0.0)
        # implemented search, pixel absolute differences per MB: L2 81*64, L1 25*64, L0 26*256; SATD stage counted as 17*256
        absdiff_mb = (2 * SR // 4 + 1) ** 2 * 64 + 25 * 64 + 26 * 256 + 17 * 256 + 3 * 256      # + the intra estimate's three 16x16 SATDs
        me_ms = me_fine + me_coarse
        int_ach = (absdiff_mb / 4.0) * nmb * Sp / 32.0 / (me_ms * 1e-3) / 1e9 if me_ms > 0 else 0.0   # warp-instructions -> G lane... see DESIGN 5
        cpu = cpu_baseline_sample(threads=1, frames=args.cpu_frames) if world == 1 and not args.no_cpu else None
        # DRAM traffic and pipe utilisation of the dominant kernel from the committed `ncu --set full` capture (profiles/), scaled
        # from that capture's sessions per launch to this run's
        traffic, ncu_note = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_kernels.json")))
            k = prof["kernels"][top[0]]
            traffic = (k["dram_bytes_read"] + k["dram_bytes_write"]) * Sp / prof["sessions_per_launch"]
            ncu_note = {"source": "profiles/r01_ncu_kernels.json", "alu_pipe_pct_of_peak": k["alu_pipe_pct"], "sm_throughput_pct_of_peak": k["sm_throughput_pct"],
                        "warp_instructions_per_mb_1080p": round(k["warp_instructions"] / (8160 * prof["sessions_per_launch"]), 1)}
        except Exception:
            pass
        # instruction-issue roofline of the dominant kernel: executed warp-instructions per launch (committed ncu capture, scaled to this
        # run's sessions and macroblock count) over its live CUDA-event time, against SMs x 4 schedulers x the SM clock sampled under load
        issue = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_kernels.json")))
            wi = prof["kernels"][top[0]]["warp_instructions"] * Sp / prof["sessions_per_launch"] * nmb / 8160.0
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            peak_issue = sms * 4 * mhz * 1e6 / 1e9
            ach_issue = wi / (top[1] * 1e-3) / 1e9
            issue = {"kernel": top[0], "warp_instructions_per_launch": round(wi), "achieved_gwarp_instr_s": round(ach_issue, 1), "peak_gwarp_instr_s": round(peak_issue, 1),
                     "frac": round(ach_issue / peak_issue, 4), "sms": sms, "sm_mhz": mhz,
                     "note": "all executed warp-instructions (ncu smsp__inst_executed.sum of the committed capture) over the live kernel time; the kernel is issue-bound, this is the roofline that bounds it"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(el / args.steps * 1e3, 4), "device_ms_per_step": round(dev_ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "value_per_gpu": round(value / world, 2), "realtime_30fps_sessions": round(value / FPS, 1),
            "config": {"workload": f"{S} concurrent {W}x{H} sessions per GPU, {LABEL}; step = one frame of every session; {G} batch(es) of {group_sizes[0]} on own streams",
                       "sessions_per_gpu": S, "frames_per_step": S * world, "l2_policy": (f"inputs larger than L2: per-step working set ~{S * npx * 10 // 1000000} MB" if S * npx * 10 > 130e6 else
                                     f"working set ~{S * npx * 10 // 1000000} MB fits L2; every step encodes a different frame of the pool, reference planes are rewritten each step"),
                       "parallelism": f"sessions sharded over {world} GPU(s), no collective",
                       "timing": "value/ms_per_step: host clock between a device synchronize + barrier on both sides (max over ranks; an upper bound of the device time of "
                                 "the overlapping batch streams); device_ms_per_step: CUDA events on the batch streams (slowest batch group); kernel_ms: CUDA events per launch"},
            "e2e": {"value": round(e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": world * S * fb, "d2h_bytes_per_step": int(world * out_bytes_e / args.steps),
                    "ms_per_step": round(el_e / args.steps * 1e3, 4)},
            "gpu_launches": launches,
            "bitrate_mbps_per_session": round(out_bytes * 8 / (S * args.steps) * FPS / 1e6, 3),
            "kernel_ms": {k: round(v, 4) for k, v in kt},
            "roofline": {"bound": "hbm", "kernel": top[0], "share_of_step": round(top[1] / tot, 3), "achieved": round(ach, 1), "peak": hbm_peak,
                         "unit": "GB/s", "frac": round(ach / hbm_peak, 4), "traffic": traffic, "ncu": ncu_note,
                         "note": "the encode path is INT-ALU/latency bound, not HBM bound (SURVEY 8d); see roofline_int"},
            "roofline_int": {"kernels": "k_me_coarse+k_me_fine", "achieved_gwarp_instr_s": round(int_ach, 2),
                             "peak_gwarp_instr_s": round(gi.value / 32.0, 2), "peak_glane_instr_s": round(gi.value, 1), "sm_clock_mhz_in_peak_run": clk.value,
                             "frac": round(int_ach / (gi.value / 32.0), 4) if gi.value else None,
                             "unit": "G warp-instr/s of VABSDIFF4-equivalent work (px-absdiff/4/32)"},
            "roofline_issue": issue,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if thr:
            line["e2e_caller_threads"] = thr
    for s in sess:
        s.close()
    batch.close()
    if use_dist:
        dist.barrier(); dist.destroy_process_group()
    if line:
        print(json.dumps(line), flush=True)


def _cpu_worker(args):
    frames, qps = args
    from oracle import orc_py
    from media_b200.synth import Content
    c = Content(KIND, W, H)
    fs = [c.frame(t) for t in range(frames + 1)]
    e = orc_py.Encoder(W, H, num_slices=SLICES, search_range=SR, profile=PROFILE)
    e.encode(fs[0], True, qps[0])
    t0 = time.perf_counter()
    for t in range(1, frames + 1):
        e.encode(fs[t], False, qps[min(t, len(qps) - 1)])
    return time.perf_counter() - t0


def cpu_baseline_sample(threads, frames):
    """CPU restatement (oracle/, 'port'), `threads` single-threaded sessions in parallel, `frames` P frames each after one IDR."""
    import multiprocessing as mp
    qps = [34] * (frames + 1)
    if threads == 1:
        times = [_cpu_worker((frames, qps))]
    else:
        with mp.get_context("fork").Pool(threads) as pool:
            times = pool.map(_cpu_worker, [(frames, qps)] * threads)
    fps = sum(frames / t for t in times)
    return {"value": round(fps, 3), "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{threads} session(s) x {frames} P frames of {W}x{H} content {KIND} at QP 34 after one IDR (CPU restatement oracle/, gcc -O2, NOT openh264: libopenh264 is absent from the image)"}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = max(2, args.cpu_frames // 2)
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_baseline_sample(threads, frames))
    best = max(vals, key=lambda v: v["value"])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{W}x{H} sessions, {LABEL}; CPU restatement of the path, one single-threaded encoder per host core"},
        "cpu_baseline": best, "e2e": {"value": best["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--sessions", type=int, default=None)
    ap.add_argument("--groups", type=int, default=None)
    ap.add_argument("--slices", type=int, default=None, help="override the workload's slice count (0 = the engine's automatic count)")
    ap.add_argument("--cpu-frames", type=int, default=12)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--threads-e2e", action="store_true", help="also measure one caller thread per session through the auto_batch scheduler")
    args = ap.parse_args()
    apply_workload(args)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
