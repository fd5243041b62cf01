"""Top stall locations of the first kernel in an .ncu-rep: python scripts/ncu_hot.py rep [N] [ctx]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
si, ns, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [(int(r[ns]), k, r) for k, r in enumerate(rows[1:]) if len(r) > ie and r[ns].isdigit()]
tot = sum(d[0] for d in data)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 15
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for n, k, r in sorted(data, reverse=True)[:N]:
    st = sorted([(int(r[i]), hdr[i][6:]) for i in stall if r[i].isdigit() and int(r[i]) > 0], reverse=True)[:2]
    print(f"{100 * n / tot:5.1f}% idx{k:5d} exec={r[ie]:>9s} {r[si].strip()[:58]:58s} {st}")
    if ctx:
        for kk in range(max(0, k - ctx), k):
            rr = rows[1 + kk]
            print("        ", rr[ns].rjust(6), rr[ie].rjust(9), rr[si].strip()[:80])
