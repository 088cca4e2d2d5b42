"""Counted FP64 work per solve from ncu's SASS instruction counters (one launch per kernel, on a B200):

    ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none -k regex:ocp_ipm -s 1 -c 1 --csv --log-file gpurun_out/cnt_cbf.csv python tools/one_launch.py --B 1024
    python tools/fp64_counts.py gpurun_out/cnt_cbf.csv:1024 gpurun_out/cnt_lmpc.csv:512 ... > profiles/fp64_counts.json

flops = 2 * DFMA + DMUL + DADD thread-level instructions (predicated-on), divided by the batch of the launch.  bench.py puts
`flops_per_solve x solves/s` against 148 SM x 64 FMA/clk x 2 x clock in `roofline.fp64`; the file records the hash of the
library the counts were taken from."""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"ocp_ipm_kernel<3, 8, 20>": "ocp_ipm_kernel<3,QDIAG,20>", "ocp_ipm_kernel<3, 0, 20>": "ocp_ipm_kernel<3,0,20>", "ocp_ipm_kernel<0, 3, 0>": "ocp_ipm_kernel<0,3,0>",
         "lmpc_kernel": "lmpc_kernel", "ilqr_kernel": "ilqr_kernel"}
def source_sha16():
    """Identity of the kernels the counts belong to: hash of the CUDA sources and the C-ABI header (the built library's own
    hash is not reproducible: two nvcc runs on the same sources give different files)."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "car_racing_b200", "csrc")
    for f in sorted(os.listdir(csrc)) + ["../../include/b200mpc.h"]:
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


doc = {"src_sha16": source_sha16(),
       "definition": "flops = 2*dfma + dmul + dadd (smsp__sass_thread_inst_executed_op_*_pred_on.sum) / batch of the launch",
       "kernels": {}}
traffic = {"kernels": {}}
for spec in sys.argv[1:]:
    path, B = spec.rsplit(":", 1)
    B = int(B)
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    if not rows:
        continue
    hdr = rows[0]
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        per.setdefault(r[kn], {})[r[mn]] = float(r[mv].replace(",", ""))
    for kname, m in per.items():
        key = next((v for k, v in NAMES.items() if k in kname), None)
        if key is None:
            key = re.sub(r"\(.*", "", kname)
        g = lambda n: m.get(n, 0.0)
        dfma, dmul, dadd = (g("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % x) for x in ("dfma", "dmul", "dadd"))
        doc["kernels"][key] = {"batch": B, "dfma": dfma, "dmul": dmul, "dadd": dadd, "flops_per_solve": (2 * dfma + dmul + dadd) / B,
                               "warp_instructions_per_solve": g("smsp__inst_executed.sum") / B, "report": os.path.basename(path)}
        if "dram__bytes_read.sum" in m:
            traffic["kernels"][key] = {"batch": B, "dram_bytes_per_launch": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
                                       "report": os.path.basename(path)}
doc["dram_traffic"] = traffic["kernels"]
print(json.dumps(doc, indent=1))
