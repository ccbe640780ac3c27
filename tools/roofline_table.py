"""tools/roofline_table.py -- per-kernel roofline table of one P step from the committed ncu counters (profiles/r02_ncu_kernels.json):
algorithmic bytes (SURVEY 8d / DESIGN 5) over the kernel's time alone against the HBM peak, DRAM traffic, and executed warp-instructions
over the same time against the issue peak (SMs x 4 schedulers x SM clock). Writes markdown to stdout."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_kernels.json")))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6546.6)
S = prof["sessions_per_launch"]
W, H, LP = 1920, 1088, 32                     # coded size, luma border at search range 16
px, nmb = W * H, W * H // 256
ISSUE = 148 * 4 * 1.965                       # G warp-instr/s
padded = (W + 2 * LP) * (H + 2 * LP)
# algorithmic bytes per session-frame and the bound the design assigns
ALG = {
    "k_ingest_planar": (3.0 * px, "HBM", "1.5 B/px read + 1.5 B/px written"),
    "k_refplanes": (5.0 * padded, "HBM / ALU", "1 B read + 4 B written per padded sample"),
    "k_refchroma": (0.5 * px + 2.0 * (W // 2 + LP) * (H // 2 + LP), "HBM", "both chroma planes: 0.5 B/px read, the padded planes written"),
    "k_downsample": (2.5 * px, "HBM", "first of the two launches (level 1 of source and reference): 2 B/px read + 0.5 B/px written"),
    "k_me_coarse": (2 * 0.3125 * px, "INT (VABSDIFF4)", "pyramid levels of source and reference"),
    "k_me_fine": (4.5 * px + nmb * 816.0, "INT issue", "source 1.5 + reference 1.5 + reconstruction 1.5 B/px + 816 B levels per MB"),
    "k_intra_wave": (3.0 * px, "latency", "only intra MBs"),
    "k_deblock_bs": (nmb * (48.0 + 16.0), "LSU", "48 B MbInfo read (neighbours from L2), 16 B written per MB"),
    "k_deblock_wave": (3.0 * px + nmb * (16.0 + 192.0), "latency", "1.5 B/px read + written in place, 16 B boundary strengths and 192 B of hand-over messages per MB"),
    "k_cavlc_mb": (nmb * (816.0 + 48.0), "INT / LSU", "816 B levels + 48 B MbInfo per MB"),
    "k_slice_copy": (nmb * 64.0, "LSU", "the MB's bits, read and written (actual bitstream size)"),
}
print(f"| kernel | bound | time alone | algorithmic bytes / launch | achieved alg. GB/s | frac of HBM {HBM:.0f} GB/s | DRAM traffic / launch | G warp-instr/s | frac of issue peak {ISSUE:.0f} |")
print("|---|---|---|---|---|---|---|---|---|")
for name, k in prof["kernels"].items():
    if name not in ALG:
        continue
    b, bound, _ = ALG[name]
    t = k["time_us"] * 1e-6
    alg = b * S
    gbs = alg / t / 1e9
    traf = (k["dram_bytes_read"] + k["dram_bytes_write"])
    gi = k["warp_instructions"] / t / 1e9
    print(f"| `{name}` | {bound} | {k['time_us']:.0f} us | {alg / 1e6:.0f} MB | {gbs:.0f} | {gbs / HBM:.3f} | {traf / 1e6:.0f} MB | {gi:.0f} | {gi / ISSUE:.2f} |")
print()
for name, (b, bound, what) in ALG.items():
    print(f"* `{name}`: {b / 1e6:.2f} MB per 1080p session-frame = {what}.")
