"""Summarise one `ncu --set full --import-source on` capture of ocp_ipm_kernel for profiles/:
headline metrics, warp-stall sampling totals, and samples attributed to the kernel-body source line that
(transitively) inlined each SASS instruction (nvdisasm -gi on the cubin of the same build).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep build/obj_libb200mpc.so/ocp_inst0.o B [kernel-name-fragment source.cuh] > profiles/x_summary.json
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile

rep, so, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
KERNEL = sys.argv[4] if len(sys.argv) > 4 else "ocp_ipm_kernelILi3ELi8ELi20"     # mangled-name fragment of the captured kernel
SRC = sys.argv[5] if len(sys.argv) > 5 else "ocp_ipm.cuh"                          # the source file its body lives in
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
doc = {"report": os.path.basename(rep), "batch": B, "metrics": {k: {"value": m[k][0], "unit": m[k][1]} for k in keep if k in m}}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
data = [r for r in rows[2:] if len(r) >= len(h2)]
ix = {h: i for i, h in enumerate(h2)}
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
alls = sum(int(r[ix["# Samples"]] or 0) for r in data)
doc["static_instructions"] = len(data)
doc["warp_instructions_per_instance"] = round(sum(int(r[ix["Instructions Executed"]] or 0) for r in data) / B)
doc["stall_samples_pct"] = {s: round(100.0 * v / alls, 2) for s, v in sorted(tot.items(), key=lambda x: -x[1]) if v}
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    dis, start = [], []
    for cub in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):     # one cubin per translation unit
        dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout.split("\n")
        start = [i for i, l in enumerate(dis) if ".section" in l and ".text." in l and KERNEL in l]
        if start:
            break
start = start[0]
end = next(i for i in range(start + 1, len(dis)) if dis[i].startswith("\t.section"))
ins, pending, outer = [], [], None
for l in dis[start:end]:
    mm = re.search(r'//## File ".*?", line (\d+)( inlined at ".*?", line (\d+))?', l)
    if mm:
        pending.append((int(mm.group(1)), int(mm.group(3)) if mm.group(3) else None))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+.*?;", l):
        if pending:
            outer = pending[-1][1] if pending[-1][1] is not None else pending[-1][0]
            pending = []
        ins.append(outer)
if len(ins) == len(data):
    srcl = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "car_racing_b200", "csrc", SRC)).read().split("\n")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for o, r in zip(ins, data):
        a = agg[o]
        a[0] += 1
        a[1] += int(r[ix["Instructions Executed"]] or 0)
        a[2] += int(r[ix["# Samples"]] or 0)
        a[3] += int(r[ix["stall_no_inst"]] or 0)
    doc["by_kernel_body_line"] = [
        {"line": k, "source": srcl[k - 1].strip()[:90] if k and k <= len(srcl) else "", "static_instructions": a[0],
         "executed_per_instance": round(a[1] / B), "samples_pct": round(100.0 * a[2] / alls, 2),
         "no_inst_pct_of_its_samples": round(100.0 * a[3] / max(a[2], 1), 1)}
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:16]]
else:
    doc["by_kernel_body_line"] = "SASS of this build does not match the capture (%d vs %d instructions)" % (len(ins), len(data))
print(json.dumps(doc, indent=1))
