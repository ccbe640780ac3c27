"""tools/ncu_summary.py -- markdown table of the per-kernel counters of an ncu report (`ncu --set full`), one row per kernel
(first profiled launch of each). usage: python tools/ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
rep = sys.argv[1]
json_out = sys.argv[2] if len(sys.argv) > 2 else None      # optional: also write the per-kernel counters as JSON (read by bench.py)
sessions = int(sys.argv[3]) if len(sys.argv) > 3 else 32     # sessions per launch in the profiled command
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp instr"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram written"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts")]
cols = [(m, n) for m, n in cols if m in ix]
print("| kernel | " + " | ".join(n for _, n in cols) + " |")
print("|---|" + "---|" * len(cols))
seen = set()
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    if name in seen:
        continue
    seen.add(name)
    cells = []
    for m, _ in cols:
        v, u = r[ix[m]], units[ix[m]]
        try:
            f = float(v); v = f"{f:,.0f}" if f >= 1000 else f"{f:.3g}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| {name} | " + " | ".join(cells) + " |")

if json_out:
    import json
    out = {"source": rep.split("/")[-1], "sessions_per_launch": sessions, "kernels": {}}
    seen = set()
    def num(r, m):
        try:
            return float(r[ix[m]])
        except (KeyError, ValueError):
            return None
    def to_bytes(r, m):
        v = num(r, m)
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[ix[m]], 1)
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        if name in seen:
            continue
        seen.add(name)
        out["kernels"][name] = {"dram_bytes_read": to_bytes(r, "dram__bytes_read.sum"), "dram_bytes_write": to_bytes(r, "dram__bytes_write.sum"),
                                "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                "sm_throughput_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                                "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                                "warp_instructions": num(r, "smsp__inst_executed.sum"), "registers": num(r, "launch__registers_per_thread"),
                                "time_us": (num(r, "gpu__time_duration.sum") or 0.0) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1.0),
                                "grid": num(r, "launch__grid_size")}
    json.dump(out, open(json_out, "w"), indent=1)
