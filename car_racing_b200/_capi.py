"""ctypes binding of libb200mpc.so (include/b200mpc.h).

There is no CPU fallback: importing this module without the built shared library, or creating
a handle without a CUDA device, raises.  Build with `python __graft_entry__.py` (or
`__graft_entry__.build()`), which runs nvcc for sm_100a.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200mpc.so")
NMAX, MMAX = 64, 8


class CbfParams(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("M", C.c_int32), ("xt_per_stage", C.c_int32), ("flags", C.c_int32),
        ("A", C.c_double * 36), ("B", C.c_double * 12), ("Q", C.c_double * 36), ("R", C.c_double * 4),
        ("umax", C.c_double * 2), ("vmin", C.c_double), ("vmax", C.c_double), ("width", C.c_double),
        ("alpha", C.c_double), ("margin", C.c_double), ("L", C.c_double), ("W", C.c_double),
        ("slack_w", C.c_double),
    ]


class IpmOptions(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("max_iter", C.c_int32), ("acceptable_iter", C.c_int32),
        ("acceptable_tol", C.c_double), ("mu_init", C.c_double), ("rho", C.c_double),
        ("bound_push", C.c_double), ("bound_frac", C.c_double), ("max_grad", C.c_double),
        ("start", C.c_int32), ("max_reset", C.c_int32),
    ]


START_ROLLOUT, START_ZERO = 0, 1
STATUS_NAMES = {0: "solved", 1: "max_iter", 2: "line search failed", 3: "inertia correction failed",
                4: "x0 violates its stage-0 bound rows (the reference's NLP is infeasible)"}


class IlqrParams(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("max_iter", C.c_int32),
        ("A", C.c_double * 36), ("B", C.c_double * 12), ("Q", C.c_double * 36), ("R", C.c_double * 4),
        ("L", C.c_double), ("W", C.c_double),
    ]


class LmpcParams(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("K", C.c_int32),
        ("Q", C.c_double * 36), ("R", C.c_double * 4), ("dR", C.c_double * 4), ("xtrk", C.c_double * 6),
        ("umax", C.c_double * 2), ("vmax", C.c_double), ("width", C.c_double),
    ]


LMPC_NMAX, LMPC_KMAX = 16, 64
SYSID_LMAX, SYSID_PMAX = 4, 64


class SysidParams(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("num_laps", C.c_int32), ("max_num_point", C.c_int32), ("num_segments", C.c_int32),
        ("lap_rows", C.c_int32 * SYSID_LMAX), ("lap_stride", C.c_int32), ("reserved", C.c_int32),
        ("dt", C.c_double), ("h", C.c_double), ("lap_length", C.c_double),
    ]


class PlantParams(C.Structure):
    _fields_ = [
        ("n_sub", C.c_int32), ("num_segments", C.c_int32), ("wrap_lap", C.c_int32), ("reserved", C.c_int32),
        ("delta_t", C.c_double), ("lap_length", C.c_double),
        ("m", C.c_double), ("lf", C.c_double), ("lr", C.c_double), ("Iz", C.c_double), ("Df", C.c_double), ("Cf", C.c_double),
        ("Bf", C.c_double), ("Dr", C.c_double), ("Cr", C.c_double), ("Br", C.c_double),
    ]


class PlannerSelectParams(C.Structure):
    _fields_ = [
        ("C", C.c_int32), ("N", C.c_int32), ("num_veh", C.c_int32), ("old_direction_flag", C.c_int32),
        ("N_ctrl", C.c_int32), ("M_ctrl", C.c_int32),
        ("veh_length", C.c_double), ("veh_width", C.c_double), ("lap_length", C.c_double),
    ]


class RolloutParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("num_segments", C.c_int32), ("timestep", C.c_double), ("lap_length", C.c_double)]


class PlannerPrepareParams(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("num_veh", C.c_int32), ("num_opt", C.c_int32), ("reserved", C.c_int32),
        ("prediction_factor", C.c_double), ("track_width", C.c_double), ("lap_length", C.c_double),
        ("veh_length", C.c_double), ("veh_width", C.c_double), ("safety_margin", C.c_double), ("vx_max", C.c_double),
        ("w_ey_rate", C.c_double), ("w_progress", C.c_double), ("w_track", C.c_double),
    ]


RECORD_DTYPE = np.dtype([("cost", "<f8"), ("u0", "<f8", (2,)), ("status", "<i4"), ("iters", "<i4")])
assert RECORD_DTYPE.itemsize == 32

EXPORTS = [
    "b200mpc_version", "b200mpc_default_ipm_options", "b200mpc_create", "b200mpc_create_ex", "b200mpc_destroy", "b200mpc_last_error",
    "b200mpc_stream", "b200mpc_launch_count", "b200mpc_cbf_record_doubles", "b200mpc_cbf_record_doubles_ex", "b200mpc_cbf_solve",
    "b200mpc_cbf_solve_device", "b200mpc_ilqr_record_doubles", "b200mpc_ilqr_solve", "b200mpc_ilqr_solve_device",
    "b200mpc_argmin_cost_device", "b200mpc_lmpc_record_doubles", "b200mpc_lmpc_solve", "b200mpc_lmpc_solve_device",
    "b200mpc_lmpc_sysid", "b200mpc_lmpc_sysid_device", "b200mpc_plant_step", "b200mpc_plant_step_device",
    "b200mpc_cbf_solve_async", "b200mpc_synchronize", "b200mpc_host_alloc", "b200mpc_host_free", "b200mpc_planner_select_device", "b200mpc_plan_and_track", "b200mpc_ilqr_solve_async", "b200mpc_lmpc_solve_async",
    "b200mpc_planner_prepare_device", "b200mpc_planner_prepare", "b200mpc_plan_and_track_prepared",
    "b200mpc_rival_rollout", "b200mpc_rival_rollout_device", "b200mpc_curv_to_glob", "b200mpc_curv_to_glob_device",
    "b200mpc_comm_create", "b200mpc_comm_export", "b200mpc_comm_connect", "b200mpc_comm_destroy", "b200mpc_comm_publish_next",
    "b200mpc_comm_argmin",
]
COMM_HANDLE_BYTES, COMM_MAX_WORLD = 64, 16

_lib = None


def lib():
    """Load libb200mpc.so (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python __graft_entry__.py`). "
            "car_racing_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_int
    L.b200mpc_version.restype = C.c_int
    L.b200mpc_default_ipm_options.argtypes = [C.POINTER(IpmOptions)]
    L.b200mpc_default_ipm_options.restype = None
    L.b200mpc_create.argtypes = [ip, ip, C.POINTER(vp)]
    L.b200mpc_create_ex.argtypes = [ip, ip, ip, C.POINTER(vp)]
    L.b200mpc_destroy.argtypes = [vp]
    L.b200mpc_destroy.restype = None
    L.b200mpc_last_error.argtypes = [vp]
    L.b200mpc_last_error.restype = C.c_char_p
    L.b200mpc_stream.argtypes = [vp]
    L.b200mpc_stream.restype = C.c_uint64
    L.b200mpc_launch_count.argtypes = [vp]
    L.b200mpc_launch_count.restype = C.c_uint64
    L.b200mpc_cbf_record_doubles.argtypes = [ip, ip, ip]
    L.b200mpc_cbf_record_doubles_ex.argtypes = [ip, ip, ip, ip]
    L.b200mpc_ilqr_record_doubles.argtypes = [ip]
    cbf_args = [vp, C.POINTER(CbfParams), C.POINTER(IpmOptions), ip, dp, dp, dp, dp, dp, dp]
    L.b200mpc_cbf_solve.argtypes = cbf_args
    L.b200mpc_cbf_solve_device.argtypes = cbf_args
    L.b200mpc_cbf_solve_async.argtypes = cbf_args
    L.b200mpc_synchronize.argtypes = [vp]
    L.b200mpc_host_alloc.argtypes = [C.c_size_t]
    L.b200mpc_host_alloc.restype = C.c_void_p
    L.b200mpc_host_free.argtypes = [vp]
    L.b200mpc_host_free.restype = None
    L.b200mpc_planner_select_device.argtypes = [vp, C.POINTER(PlannerSelectParams)] + [dp] * 10
    L.b200mpc_plan_and_track.argtypes = [vp, C.POINTER(CbfParams), C.POINTER(CbfParams), C.POINTER(IpmOptions),
                                         C.POINTER(PlannerSelectParams)] + [dp] * 14
    c2g_args = [vp, ip, ip, C.c_double, dp, dp, ip, dp, ip, dp]
    L.b200mpc_curv_to_glob.argtypes = c2g_args
    L.b200mpc_curv_to_glob_device.argtypes = c2g_args
    roll_args = [vp, C.POINTER(RolloutParams), ip] + [dp] * 5
    L.b200mpc_rival_rollout.argtypes = roll_args
    L.b200mpc_rival_rollout_device.argtypes = roll_args
    prep_args = [vp, C.POINTER(PlannerPrepareParams)] + [dp] * 13
    L.b200mpc_planner_prepare_device.argtypes = prep_args
    L.b200mpc_planner_prepare.argtypes = prep_args
    L.b200mpc_plan_and_track_prepared.argtypes = [vp, C.POINTER(CbfParams), C.POINTER(CbfParams), C.POINTER(IpmOptions),
                                                  C.POINTER(PlannerSelectParams), C.POINTER(PlannerPrepareParams)] + \
        [dp] * 5 + [ip] + [dp] * 18
    ilqr_args = [vp, C.POINTER(IlqrParams), ip, dp, dp, dp, dp]
    L.b200mpc_ilqr_solve.argtypes = ilqr_args
    L.b200mpc_ilqr_solve_device.argtypes = ilqr_args
    L.b200mpc_ilqr_solve_async.argtypes = ilqr_args
    L.b200mpc_argmin_cost_device.argtypes = [vp, dp, ip, ip, dp]
    L.b200mpc_lmpc_record_doubles.argtypes = [ip, ip]
    lmpc_args = [vp, C.POINTER(LmpcParams), C.POINTER(IpmOptions), ip, dp, dp, dp, dp, dp, dp]
    L.b200mpc_lmpc_solve.argtypes = lmpc_args
    L.b200mpc_lmpc_solve_device.argtypes = lmpc_args
    L.b200mpc_lmpc_solve_async.argtypes = lmpc_args
    sysid_args = [vp, C.POINTER(SysidParams), ip, dp, dp, dp, dp, ip, ip, dp, dp]
    L.b200mpc_lmpc_sysid.argtypes = sysid_args
    L.b200mpc_lmpc_sysid_device.argtypes = sysid_args
    L.b200mpc_comm_create.argtypes = [vp, ip, ip, ip, ip, C.POINTER(vp)]
    L.b200mpc_comm_export.argtypes = [vp, dp]
    L.b200mpc_comm_connect.argtypes = [vp, dp]
    L.b200mpc_comm_destroy.argtypes = [vp]
    L.b200mpc_comm_destroy.restype = None
    L.b200mpc_comm_publish_next.argtypes = [vp, vp, ip]
    L.b200mpc_comm_argmin.argtypes = [vp, vp, ip, ip, dp, dp]
    plant_args = [vp, C.POINTER(PlantParams), ip, dp, ip, ip, dp, dp, ip, dp, dp, dp]
    L.b200mpc_plant_step.argtypes = plant_args
    L.b200mpc_plant_step_device.argtypes = plant_args
    _lib = L
    return L


class B200MPCError(RuntimeError):
    pass


class Handle:
    """Opaque solver handle (device buffers + stream).  Not fork-safe; pickles without the
    native handle and re-creates it lazily (controller objects are pickled with the simulator,
    car_racing/tests/mpccbf_test.py:45-46)."""

    def __init__(self, device=-1, max_batch=1024, high_priority=False):
        self.device, self.max_batch, self.high_priority = device, max_batch, bool(high_priority)
        self._h = None

    def _ensure(self):
        if self._h is None:
            h = C.c_void_p()
            rc = lib().b200mpc_create_ex(self.device, self.max_batch, int(getattr(self, "high_priority", False)), C.byref(h))
            if rc != 0:
                raise B200MPCError(f"b200mpc_create failed ({rc}): {lib().b200mpc_last_error(None).decode()}")
            self._h = h
        return self._h

    @property
    def ptr(self):
        return self._ensure()

    def check(self, rc, what):
        if rc != 0:
            raise B200MPCError(f"{what} failed ({rc}): {lib().b200mpc_last_error(self._h).decode()}")

    @property
    def stream(self):
        return int(lib().b200mpc_stream(self.ptr))

    def synchronize(self):
        self.check(lib().b200mpc_synchronize(self.ptr), "b200mpc_synchronize")

    @property
    def launch_count(self):
        return int(lib().b200mpc_launch_count(self.ptr))

    def close(self):
        if self._h is not None:
            lib().b200mpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        return {"device": self.device, "max_batch": self.max_batch, "high_priority": getattr(self, "high_priority", False)}

    def __setstate__(self, st):
        self.device, self.max_batch, self.high_priority = st["device"], st["max_batch"], st.get("high_priority", False)
        self._h = None


class PinnedArray:
    """A numpy view of page-locked host memory (b200mpc_host_alloc) -- what the asynchronous calls need for their
    copies to overlap with compute.  Keep the object alive while `.a` is in use."""

    def __init__(self, shape, dtype=np.float64):
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        self._p = lib().b200mpc_host_alloc(max(n, 1))
        if not self._p:
            raise B200MPCError("b200mpc_host_alloc failed")
        buf = (C.c_char * max(n, 1)).from_address(self._p)
        self.a = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        self.a[...] = 0

    def __del__(self):
        try:
            p, self._p = self._p, None
            if p:
                self.a = None
                lib().b200mpc_host_free(p)
        except Exception:
            pass


def default_options(**kw):
    o = IpmOptions()
    lib().b200mpc_default_ipm_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown IPM option {k!r}")
        setattr(o, k, v)
    return o


def _fill(dst, src, n):
    a = np.ascontiguousarray(src, dtype=np.float64).ravel()
    if a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    C.memmove(dst, a.ctypes.data, a.nbytes)


FLAG_STAGE_BOUNDS, FLAG_EY_RATE, FLAG_RIVAL_SIZE = 1, 2, 4


def make_cbf_params(prm, M, xt_per_stage, flags=0):
    p = CbfParams()
    p.N, p.M, p.xt_per_stage, p.flags = int(prm["N"]), int(M), int(bool(xt_per_stage)), int(flags)
    _fill(p.A, prm["A"], 36); _fill(p.B, prm["B"], 12); _fill(p.Q, prm["Q"], 36); _fill(p.R, prm["R"], 4)
    _fill(p.umax, prm["umax"], 2)
    for k in ("vmin", "vmax", "width", "alpha", "margin", "L", "W", "slack_w"):
        setattr(p, k, float(prm[k]))
    return p


def make_ilqr_params(prm):
    p = IlqrParams()
    p.N, p.max_iter = int(prm["N"]), int(prm["max_iter"])
    _fill(p.A, prm["A"], 36); _fill(p.B, prm["B"], 12); _fill(p.Q, prm["Q"], 36); _fill(p.R, prm["R"], 4)
    p.L, p.W = float(prm["L"]), float(prm["W"])
    return p


def make_lmpc_params(prm, K):
    p = LmpcParams()
    p.N, p.K = int(prm["N"]), int(K)
    _fill(p.Q, prm["Q"], 36); _fill(p.R, prm["R"], 4); _fill(p.dR, prm["dR"], 4); _fill(p.xtrk, prm["xtrk"], 6)
    _fill(p.umax, prm["umax"], 2)
    p.vmax, p.width = float(prm["vmax"]), float(prm["width"])
    return p


DEFAULT_DYN = (1.98, 0.125, 0.125, 0.024, 0.8 * 1.98 * 9.81 / 2.0, 1.25, 1.0, 0.8 * 1.98 * 9.81 / 2.0, 1.25, 1.0)   # base.py:659-672


def make_plant_params(num_segments, lap_length, timestep=0.1, delta_t=0.001, dyn=DEFAULT_DYN, wrap_lap=False):
    p = PlantParams()
    i = 0
    while (i + 1) * delta_t <= timestep:      # the reference's loop condition (base.py:909)
        i += 1
    p.n_sub, p.num_segments, p.wrap_lap = i, int(num_segments), int(bool(wrap_lap))
    p.delta_t, p.lap_length = float(delta_t), float(lap_length)
    for k, v in zip(("m", "lf", "lr", "Iz", "Df", "Cf", "Bf", "Dr", "Cr", "Br"), dyn):
        setattr(p, k, float(v))
    return p
