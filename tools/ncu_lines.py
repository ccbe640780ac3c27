"""tools/ncu_lines.py -- per-source-line instruction counts of one kernel from an ncu report.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [launch-skip]
Joins `ncu --page source --csv` (SASS order, per-instruction counters) with `nvdisasm -g` line info of the in-tree library."""
import csv, os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break                      # next launch of the same kernel
    if len(r) >= len(hdr) and r[ix["Instructions Executed"]].isdigit():
        sass.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), int(r[ix["L1 Wavefronts Shared Excessive"]] or 0) if "L1 Wavefronts Shared Excessive" in ix else 0))
tmp = tempfile.mkdtemp()
lib = os.environ.get("NCU_LINES_LIB", os.path.join(ROOT, "media_b200", "csrc", "libb200enc.so"))   # the build the report was taken with
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# locate the function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(os.environ.get("NCU_LINES_FUNC", kern), l))   # NCU_LINES_FUNC: mangled-name regex when templates share the name
lines = []; cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} profiled", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), (_, n, smp, exc) in zip(lines, sass):
    a = agg[(f, ln)]; a[0] += n; a[1] += smp; a[2] += exc
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
src_cache = {}
def src(f, ln):
    if f not in src_cache:
        p = os.path.join(ROOT, "media_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[f]
    return s[ln - 1].strip()[:110] if 0 < ln <= len(s) else ""
print(f"kernel {kern}: {tot} warp-instructions, {tots} samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("NCU_LINES_TOP", "40"))]:
    print(f"{a[0] / tot * 100:5.1f}% instr {a[1] / tots * 100:5.1f}% samples excess_smem_wavefronts {a[2]:>9d}  {f}:{ln}  {src(f, ln)}")
