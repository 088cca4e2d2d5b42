"""TEST INFRASTRUCTURE: builds tests/host_emulation/_build/libb200mpc_emu.so -- the WHOLE C-ABI library
(car_racing_b200/csrc/capi.cu + every kernel header) compiled by g++ against the stand-in cuda_runtime.h beside this file,
so that the product's own Python API can be exercised end to end on a machine without a GPU: kernels run with one host
thread per CUDA thread, "device" memory is host memory.

The sources are used as they are, except for two mechanical substitutions made on copies under _build/src/:
  * `kernel<<<grid, block, smem, stream>>>(args);`            ->  `emu_launch(grid, block, smem, [&]() { kernel(args); });`
  * `extern __shared__ __align__(16) double sm[];`            ->  `double *sm = emu_dynamic_smem;`
Nothing on the product path knows about this library: only tests load it (by pointing car_racing_b200._capi at it inside a
fixture).  It is a checker of the kernels' logic, not a fallback."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "car_racing_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libb200mpc_emu.so")

_LAUNCH = re.compile(r"(\b[A-Za-z_]\w*(?:<[^<>;]*>)?)<<<(.*?)>>>\((.*?)\);", re.S)


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def _rewrite_launch(m):
    cfg = _split_top(m.group(2))
    grid, block, smem = cfg[0], cfg[1], (cfg[2] if len(cfg) > 2 else "0")
    return "emu_launch(%s, %s, %s, [&]() { %s(%s); });" % (grid, block, smem, m.group(1), m.group(3))


def transform(text):
    text, n_launch = _LAUNCH.subn(_rewrite_launch, text)
    text, n_smem = re.subn(r"extern\s+__shared__\s+__align__\(16\)\s+double\s+sm\[\];", "double *sm = emu_dynamic_smem;", text)
    return text, n_launch, n_smem


def build(force=False):
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    deps = [os.path.join(CSRC, f) for f in srcs] + [os.path.join(HERE, "cuda_runtime.h"), os.path.join(ROOT, "include", "b200mpc.h"),
                                                    os.path.abspath(__file__)]
    lib = LIB
    if os.environ.get("B200MPC_EMU_CXXFLAGS"):                      # a variant: its own file, never served from the cache
        lib, force = LIB[:-3] + "_variant.so", True
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    sdir = os.path.join(OUT, "src", "car_racing_b200", "csrc")
    os.makedirs(sdir, exist_ok=True)
    os.makedirs(os.path.join(OUT, "src", "include"), exist_ok=True)
    with open(os.path.join(ROOT, "include", "b200mpc.h")) as f:
        open(os.path.join(OUT, "src", "include", "b200mpc.h"), "w").write(f.read())
    launches = smem = 0
    for f in srcs:
        text, a, b = transform(open(os.path.join(CSRC, f)).read())
        launches, smem = launches + a, smem + b
        open(os.path.join(sdir, f if f.endswith(".cuh") else f[:-3] + ".cpp"), "w").write(text)
    assert launches >= 10 and smem >= 4, (launches, smem)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas",
           "-DB200MPC_HOST_EMULATION", "-DEMU_WITH_LAUNCH", "-DEMU_WARPS", "-DEMU_RUNTIME_API", "-I", HERE,
           os.path.join(sdir, "capi.cpp"), "-o", lib]
    extra = os.environ.get("B200MPC_EMU_CXXFLAGS", "").split()     # e.g. the -D flags of a kernel variant (tools/variants.sh)
    if extra:
        cmd[-2:-2] = extra
    if os.environ.get("B200MPC_EMU_TSAN"):
        cmd[1:1] = ["-g", "-fsanitize=thread"]
    if os.environ.get("B200MPC_EMU_ASAN"):      # "device" buffers are malloc'ed: out-of-bounds global accesses of a kernel are caught
        cmd[1:1] = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"]
    tmp = lib + ".tmp%d" % os.getpid()                              # atomic: a concurrent reader never sees a half-written library
    cmd[cmd.index("-o") + 1] = tmp
    subprocess.run(cmd, check=True)
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    print(build(force=True))
