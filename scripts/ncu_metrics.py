"""Headline metrics of `ncu --set full` captures as one JSON object (profiles/r02_ncu_metrics.json): per report the kernel
name, duration, DRAM bytes read / written (the `traffic` of bench.py's roofline), tensor-pipe and issue utilisation.
    python scripts/ncu_metrics.py name=path.ncu-rep [name=path ...] > metrics.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_elapsed_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read", "launch__registers_per_thread": "registers",
        "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_instructions",
        # the L1 / shared-memory data pipe: LSU wavefronts (global + shared accesses of the warps) and the tensor core's
        # operand reads share it -- the bound of the N <= 128 conv layers found in round 2
        "l1tex__data_pipe_lsu_wavefronts.sum": "l1_lsu_wavefronts",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum": "l1_tensor_operand_wavefronts",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_lsu_wavefronts_pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "l1_tensor_operand_wavefronts_pct",
        "sm__cycles_elapsed.max": "sm_cycles"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
out = {}
for arg in sys.argv[1:]:
    name, path = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        out[name] = {"error": "no kernel in %s" % path}
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None, "report": path}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            d[WANT[h]] = x * SCALE.get(u, 1.0) if WANT[h] in ("time_us", "dram_read", "dram_write", "l2_to_sm_read") else x
    if "dram_read" in d and "dram_write" in d:
        d["traffic_bytes"] = d["dram_read"] + d["dram_write"]
    if "l1_lsu_wavefronts_pct" in d and "l1_tensor_operand_wavefronts_pct" in d:
        d["l1_data_pipe_busy_pct"] = d["l1_lsu_wavefronts_pct"] + d["l1_tensor_operand_wavefronts_pct"]
    out[name] = d
print(json.dumps(out, indent=1, sort_keys=True))
