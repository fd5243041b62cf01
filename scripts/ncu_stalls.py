"""Stall breakdown and memory-pipe metrics of the first kernel of an .ncu-rep (appended to the summaries under profiles/):
    python scripts/ncu_stalls.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, r = rows[0], rows[2]
st = []
for i, k in enumerate(h):
    if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        try:
            st.append((float(r[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1.0
print("  warp-cycles per issued instruction by stall reason (share of the warp time):")
for v, k in sorted(st, reverse=True)[:8]:
    print(f"    {k:28s} {v:7.2f}  {100 * v / tot:5.1f} %")
for k in ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
          "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum"]:
    if k in h:
        print(f"  {k:85s} {r[h.index(k)]}")
