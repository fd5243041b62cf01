"""Instruction mix of the hot loop of the first kernel in an .ncu-rep: instructions whose executed
count is >= frac * max.   python scripts/ncu_loop.py rep [frac] [dump]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si, ie = hdr.index("Source"), hdr.index("Instructions Executed")
data = [(int(r[ie]), r[si].strip()) for r in rows[1:] if len(r) > ie and r[ie].isdigit()]
mx = max(d[0] for d in data)
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
tot = sum(d[0] for d in data)
hot = [d for d in data if d[0] >= frac * mx]
print(f"max exec {mx}, total warp-instr {tot}, hot-loop instrs {len(hot)} (sum {sum(d[0] for d in hot)} = {100*sum(d[0] for d in hot)/tot:.1f}% of all), per max-exec: {tot/mx:.1f}")
mix = collections.Counter()
for n, src in hot:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    mix[op.split(".")[0]] += n / mx
for op, c in mix.most_common(40):
    print(f"  {op:12s} {c:6.1f}")
if len(sys.argv) > 3:
    for n, src in data:
        print(f"{n:9d} {src[:100]}")
