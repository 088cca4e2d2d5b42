import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    """A CUDA device the library can use: the built libb200mpc.so loads and b200mpc_create succeeds."""
    try:
        from car_racing_b200 import _capi
        h = _capi.Handle()
        h.ptr
        h.close()
        return True, ""
    except Exception as e:           # library not built, or B200MPC_ERR_NODEVICE
        return False, str(e)


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing them (with `-m gpu`
    explicitly requested they run and fail loudly: a GPU box whose extension does not load must not look green)."""
    if "gpu" in (config.getoption("markexpr", default="") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    ok, why = _gpu_available()
    if not ok:
        skip = pytest.mark.skip(reason="needs a CUDA device: " + why[:120])
        for it in gpu_items:
            it.add_marker(skip)


# The CPU tier compiles the kernels' sources with g++ (tests/host_emulation/); the two large translation units take 1-2 minutes each on a cold
# checkout.  They are started in the background when the session starts, so that they overlap each other and the tests that do not need them;
# the tests that do call wait_prebuilt(name) before their own staleness check.  Nothing is built for a `-m gpu` session.
_PREBUILD = {}


def _hot_kernel_lib():
    emu = os.path.join(ROOT, "tests", "host_emulation")
    lib = os.path.join(emu, "_build", "libocp_ipm_emu.so")
    src = [os.path.join(emu, "ocp_ipm_host.cpp"), os.path.join(emu, "cuda_runtime.h"),
           os.path.join(ROOT, "car_racing_b200", "csrc", "ocp_ipm.cuh"), os.path.join(ROOT, "include", "b200mpc.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in src):
        import subprocess
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        tmp = lib + ".tmp%d" % os.getpid()
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                        "-I", emu, src[0], "-o", tmp], check=True)
        os.replace(tmp, lib)
    return lib


def _emu_library():
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_emu_library", os.path.join(ROOT, "tests", "host_emulation", "build_emu_library.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def pytest_sessionstart(session):
    import threading
    if session.config.getoption("markexpr", default="").strip() == "gpu" or session.config.getoption("collectonly", default=False):
        return
    for name, fn in (("b200mpc_emu", _emu_library), ("ocp_ipm_emu", _hot_kernel_lib)):
        box = {}

        def work(fn=fn, box=box):
            try:
                box["path"] = fn()
            except BaseException as e:          # re-raised in the test that asks for the library
                box["error"] = e
        th = threading.Thread(target=work, daemon=True)
        th.start()
        _PREBUILD[name] = (th, box)


def wait_prebuilt(name):
    """Path of a host-emulation library: joins the background build started at session start (or builds it now)."""
    if name in _PREBUILD:
        th, box = _PREBUILD[name]
        th.join()
        if "error" in box:
            raise box["error"]
    return {"b200mpc_emu": _emu_library, "ocp_ipm_emu": _hot_kernel_lib}[name]()


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def crb():
    import car_racing_b200
    return car_racing_b200
