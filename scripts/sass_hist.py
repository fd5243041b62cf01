"""Executed-instruction histogram of the first kernel of an .ncu-rep (needs --import-source on):
    python scripts/sass_hist.py report.ncu-rep [npixels]
Prints warp-instructions per opcode (and per pixel-warp if npixels is given)."""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si, ei, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
hist, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= ei or not r[ei].isdigit():
        if len(r) > 0 and r[0].startswith("Kernel Name"):
            break
        continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDG", "LDS", "STG", "STS", "F2F", "F2I", "I2F", "I2FP", "IMAD", "HADD2") else op.split(".")[0]
    n = int(r[ei]); hist[op] += n; tot += n
    samp[op] += int(r[ss]) if r[ss].isdigit() else 0
npx = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp-inst {tot}" + (f" = {tot * 32 / npx:.1f} thread-inst/px" if npx else ""))
stot = sum(samp.values()) or 1
for op, n in hist.most_common(45):
    print(f"  {op:14s} {n:12d} {100.0 * n / tot:5.1f}%  samples {100.0 * samp[op] / stot:5.1f}%" + (f"  {n * 32 / npx:6.1f}/px" if npx else ""))
