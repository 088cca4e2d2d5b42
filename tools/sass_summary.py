"""Per-kernel SASS facts of the built library (run here, no GPU needed): instruction count, FP64 / TMA / mbarrier / warp-reduce /
system-scope mnemonics, registers, stack.  Written to profiles/sass_summary.txt so that the evidence for "TMA staging, mbarrier,
no tensor cores, FP64" can be read without disassembling the .so.

    python tools/sass_summary.py [car_racing_b200/libb200mpc.so] > profiles/sass_summary.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "car_racing_b200", "libb200mpc.so")
KEYS = ["DFMA", "DMUL", "DADD", "MUFU", "UBLKCP", "SYNCS", "REDUX", "SHFL", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "NANOSLEEP", "BAR", "HMMA", "UTCHMMA", "BRA"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except OSError:
        return n


res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", line))
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            counts[cur][base] += 1
        if ".SYS" in op or ".STRONG.SYS" in line:
            counts[cur]["sys-scope"] += 1
sha = hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16]
print(f"# {os.path.relpath(lib, ROOT)}  sha256[:16] = {sha}   (tools/sass_summary.py; cuobjdump -sass / -res-usage, sm_100a)")
print("# kernel | SASS instructions | registers | stack B | " + " ".join(KEYS) + " | sys-scope")
seen = set()
for fn, c in counts.items():
    name = demangle(fn)
    short = re.sub(r"\(.*", "", name).replace("b200mpc::", "").replace("void ", "")
    if short in seen:
        continue
    seen.add(short)
    u = usage.get(fn, {})
    print(f"{short} | {c['_total']} | {u.get('REG', '?')} | {u.get('STACK', '?')} | " + " ".join(str(c[k]) for k in KEYS) + f" | {c['sys-scope']}")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, REDUX = warp integer reduce, HMMA/UTCHMMA = tensor-core MMA (none: FP64, 6-14 wide banded blocks)")
