"""Join an ncu SASS source page with nvdisasm line info: instructions executed / stall samples
per CUDA source line.   python scripts/ncu_lines.py report.ncu-rep <kernel-substring> [cubin]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep, kname = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "video-matting_b200", "csrc", "libvm_sm100a.so")
tmp = "/tmp/_cub"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iex, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
body = rows[2:]
base = int(body[0][0], 16)
counts = {int(r[0], 16) - base: (int(r[iex]), int(r[isamp]), r[isrc]) for r in body}
# find function in cubins
for cub in sorted(os.listdir(tmp)):
    if not cub.endswith(".cubin") or "-" in cub:
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    # split per function
    cur, line, fn = None, None, None
    per_line = collections.Counter(); per_samp = collections.Counter(); matched = 0
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1); continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and fn and kname in fn:
            off = int(m.group(1), 16)
            if off in counts:
                per_line[line] += counts[off][0]; per_samp[line] += counts[off][1]; matched += 1
    if matched:
        tot = sum(per_line.values()); tots = sum(per_samp.values()) or 1
        src_cache = {}
        print(f"{cub}: matched {matched} SASS instrs, {tot} warp-instr")
        for (f, ln), c in per_line.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 45):
            if f not in src_cache:
                p = os.path.join(ROOT, "video-matting_b200", "csrc", f)
                src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
            txt = src_cache[f][ln - 1].strip()[:95] if ln - 1 < len(src_cache[f]) else ""
            print(f"{100*c/tot:5.1f}% instr {100*per_samp[(f,ln)]/tots:5.1f}% stall | {f}:{ln:<4d} {txt}")
        break
