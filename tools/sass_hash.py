"""tools/sass_hash.py [library ...] -- sha256 over the SASS text of a build of libb200enc.so (`cuobjdump -sass`, instruction encodings and the source-path
`identifier` lines stripped): two builds with the same hash run the same device code. Used to show that the committed sources compile to exactly the
variant that went through the GPU parity suite (tools/variant_gate.sh), without a GPU."""
import hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or [os.path.join(ROOT, "media_b200", "csrc", "libb200enc.so")]
for lib in libs:
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    h = hashlib.sha256()
    n = 0
    for line in sass.splitlines():
        if not line.strip() or line.strip().startswith("identifier"):
            continue
        line = re.sub(r"/\*[0-9a-f]{16}\*/", "", line).rstrip()
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            n += 1
        h.update(line.encode() + b"\n")
    print(f"{h.hexdigest()}  {n} instructions  {os.path.relpath(lib, ROOT)}")
