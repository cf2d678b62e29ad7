"""Summarise an .ncu-rep (raw page + source page) into text: python tools/ncu_summary.py file.ncu-rep [out.txt]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__cluster_max_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum"]
print(f"# {rep}", file=out)
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]: {[r[i] for r in data]}", file=out)
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        print(f"{h.replace('smsp__average_warps_issue_stalled_', 'stall/').replace('_per_issue_active.ratio', '')}: {[r[i] for r in data]}", file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = next(k for k, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr, data = rows[hidx], rows[hidx + 1:]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, samp = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    tot[op] += int(r[iE]); samp[op] += int(r[iN])
total = sum(tot.values())
print(f"dynamic warp instructions: {total}", file=out)
for op, n in tot.most_common(18):
    print(f"  {op:8s} {n:13d} {100 * n / total:5.1f}%  stall samples {samp[op]}", file=out)
