"""tools/sass_summary.py -- per-kernel counts of the SASS mnemonics that matter for this path (TMA, packed-byte SIMD, dot products,
shared / global access widths, spills), from `cuobjdump -sass` of the in-tree sm_100a library. Writes profiles/<tag>_sass_summary.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "media_b200", "csrc", "libb200enc.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m: cur = m.group(1); continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur: regs[cur] = (int(m.group(1)), int(m.group(2)), int(re.search(r"LOCAL:(\d+)", line).group(1)) if "LOCAL:" in line else 0)
KEYS = ["UTMALDG", "SYNCS", "VABSDIFF4", "IDP.4A", "IDP.2A", "VIMNMX", "VIADD", "PRMT", "SHF", "LOP3", "REDUX", "SHFL", "MATCH", "LDG.E.128", "LDG.E.64", "STG.E.128", "LDS.128", "LDS.64", "LDS", "STS", "LDL", "STL", "BAR.SYNC", "NANOSLEEP", "ATOM", "RED."]
per = collections.OrderedDict(); name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); per[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1); per[name]["_total"] += 1
        for k in KEYS:
            if op.startswith(k) or (k in ("LDG.E.128", "LDG.E.64", "STG.E.128") and op.startswith(k.split(".")[0]) and k.split("E.")[1] in op.split(".")):
                per[name][k] += 1
def demangle(n):
    m = re.match(r"_ZN4b200(\d+)", n)
    if m:
        l = int(m.group(1)); i = n.index(m.group(1)) + len(m.group(1)); base = n[i:i + l]
        t = re.search(r"ILi(\d+)ELi(\d+)ELb(\d)", n)
        return base + (f"<{t.group(1)},{t.group(2)},{t.group(3)}>" if t else "")
    return n
out = [f"SASS summary of media_b200/csrc/libb200enc.so (sm_100a, `cuobjdump -sass` / `-res-usage`), tools/sass_summary.py {tag}", ""]
hdr = f"{'kernel':34s} {'instr':>6s} {'regs':>4s} {'smem':>6s} {'local':>5s} " + " ".join(f"{k:>9s}" for k in KEYS)
out.append(hdr)
for n, c in per.items():
    d = demangle(n)
    if d.startswith("b200k_") or "test" in d: continue
    r = regs.get(n, (0, 0, 0))
    out.append(f"{d[:34]:34s} {c['_total']:6d} {r[0]:4d} {r[1]:6d} {r[2]:5d} " + " ".join(f"{c[k]:9d}" for k in KEYS))
tot = collections.Counter()
for c in per.values(): tot.update(c)
out.append(""); out.append("whole library: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
p = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
open(p, "w").write("\n".join(out) + "\n"); print("\n".join(out))
