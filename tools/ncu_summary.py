"""tools/ncu_summary.py -- markdown table of the per-kernel counters of an ncu report (`ncu --set full`), one row per kernel
(first profiled launch of each). usage: python tools/ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
rep = sys.argv[1]
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
