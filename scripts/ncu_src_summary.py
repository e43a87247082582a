"""Summarise one `ncu --set full --import-source on` capture: headline metrics, opcode histogram of executed instructions,
and executed instructions per 500-instruction stretch of SASS (where the issue slots go).
    python scripts/ncu_src_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
print("# %s" % rep)
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print("%-90s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data)
print("\n# SASS: %d instructions, %d executed (warp level)" % (len(data), tot))
ops = collections.Counter()
for r in data:
    t = r[isrc].split()
    ops[t[1] if t[0].startswith("@") else t[0]] += int(r[ie])
print("# opcode histogram:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(22)))
print("# executed instructions per stretch of 500 SASS instructions (share, stall samples, first instruction)")
for i in range(0, len(data), 500):
    c = sum(int(r[ie]) for r in data[i:i + 500])
    s = sum(int(r[isamp]) for r in data[i:i + 500])
    print("%6d %10d %5.1f%% %6d  %s" % (i, c, 100.0 * c / tot, s, data[i][isrc].strip()[:60]))
