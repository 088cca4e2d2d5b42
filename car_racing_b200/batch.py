"""Explicit batch API: stacked numpy arrays in, stacked numpy arrays out, one kernel launch.

The packers lay each instance out as ONE contiguous record (the warp that solves the instance
pulls it with a single TMA bulk copy); see include/b200mpc.h for the layout.
"""
import ctypes as C

import numpy as np

from . import _capi

_default_handle = None


def default_handle():
    global _default_handle
    if _default_handle is None:
        _default_handle = _capi.Handle()
    return _default_handle


def cbf_record_doubles(N, M, xt_per_stage, flags=0):
    hdr = (6 + M + 1) & ~1
    n = hdr + (6 * (N + 1) if xt_per_stage else 6) + 2 * M * (N + 1)
    n = (n + 1) & ~1
    if flags & _capi.FLAG_STAGE_BOUNDS:
        n += 4 * (N + 1)
    if flags & _capi.FLAG_EY_RATE:
        n += (N + 1) & ~1
    if flags & _capi.FLAG_RIVAL_SIZE:
        n += 2 * M
    return n


def ilqr_record_doubles(N):
    return (14 + 2 * (N + 1) + 1) & ~1


def lmpc_record_doubles(N, K):
    return (8 + 54 * N + 7 * K + 1) & ~1


NO_BOUND = 1e300


def cbf_flags(M, xlb=None, xub=None, wd=None, sizes=None):
    """Flags of the packed record from which optional blocks are given.  The kernel has the planner blocks (per-stage
    bounds + ey-rate weights) as ONE code path, so either of them switches both on (the missing one is packed as
    "no bound" / zero weights); per-rival sizes only exist with rivals."""
    if (xlb is None) != (xub is None):
        raise ValueError("xlb and xub must be given together")
    planner = xlb is not None or wd is not None
    return ((_capi.FLAG_STAGE_BOUNDS | _capi.FLAG_EY_RATE) if planner else 0) | \
        (_capi.FLAG_RIVAL_SIZE if (sizes is not None and M > 0) else 0)


def pack_cbf(x0, xt, obs, lap_off, N, out=None, xlb=None, xub=None, wd=None, sizes=None):
    """x0 (B,6); xt (6,), (B,6) or (B,N+1,6); obs (B,M,2,N+1) = rows 4,5 of each rival's predicted
    trajectory (control.py:509-511); lap_off (B,M) or None.  Optional planner blocks: xlb/xub (B,N+1,2)
    per-stage bounds on (vx, ey) (+-inf = none), wd (B,N) ey-rate weights.  Optional sizes (B,M,2) or (M,2):
    (l_agent+l_obs, w_agent+w_obs) of every rival (control.py:530-535) -- without it prm["L"], prm["W"] apply to all.
    Returns (records (B,stride), M, xt_per_stage) -- flags follow from which optional blocks are given (cbf_flags)."""
    x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, dtype=np.float64)))
    B = x0.shape[0]
    obs = np.asarray(obs, dtype=np.float64)
    obs = obs.reshape(B, -1, 2, N + 1) if obs.size else np.zeros((B, 0, 2, N + 1))
    M = obs.shape[1]
    if M > _capi.MMAX:
        raise ValueError(f"at most {_capi.MMAX} rivals per instance are supported, got {M}")
    xt = np.asarray(xt, dtype=np.float64)
    per_stage = xt.ndim == 3
    flags = cbf_flags(M, xlb, xub, wd, sizes)
    stride = cbf_record_doubles(N, M, per_stage, flags)
    rec = np.zeros((B, stride)) if out is None else out
    hdr = (6 + M + 1) & ~1
    rec[:, 0:6] = x0
    if M and lap_off is not None:
        rec[:, 6:6 + M] = np.asarray(lap_off, dtype=np.float64).reshape(B, M)
    if per_stage:
        rec[:, hdr:hdr + 6 * (N + 1)] = xt.reshape(B, 6 * (N + 1))
        o = hdr + 6 * (N + 1)
    else:
        rec[:, hdr:hdr + 6] = xt.reshape(-1, 6)
        o = hdr + 6
    if M:
        rec[:, o:o + 2 * M * (N + 1)] = obs.reshape(B, 2 * M * (N + 1))
    o = cbf_record_doubles(N, M, per_stage, 0)
    if flags & _capi.FLAG_STAGE_BOUNDS:
        if xlb is None:                       # only wd given: no per-stage bounds at all
            lo, hi = np.full((B, N + 1, 2), -NO_BOUND), np.full((B, N + 1, 2), NO_BOUND)
        else:
            lo = np.clip(np.asarray(xlb, dtype=np.float64).reshape(B, N + 1, 2), -NO_BOUND, NO_BOUND)
            hi = np.clip(np.asarray(xub, dtype=np.float64).reshape(B, N + 1, 2), -NO_BOUND, NO_BOUND)
        rec[:, o:o + 4 * (N + 1)] = np.concatenate([lo, hi], axis=2).reshape(B, 4 * (N + 1))
        o += 4 * (N + 1)
    if flags & _capi.FLAG_EY_RATE:
        if wd is not None:
            rec[:, o:o + N] = np.asarray(wd, dtype=np.float64).reshape(B, N)
        o += (N + 1) & ~1
    if flags & _capi.FLAG_RIVAL_SIZE:
        sz = np.broadcast_to(np.asarray(sizes, dtype=np.float64), (B, M, 2))
        if not (sz > 0).all():
            raise ValueError("rival sizes must be positive")
        rec[:, o:o + 2 * M] = sz.reshape(B, 2 * M)
    return rec, M, per_stage


def pack_ilqr(x0, xt, obs, lap_off, N):
    """x0 (B,6); xt (6,)|(B,6); obs (B,2,N+1) = rows 4,5 of the rival's prediction; lap_off (B,)|None."""
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    B = x0.shape[0]
    stride = ilqr_record_doubles(N)
    rec = np.zeros((B, stride))
    rec[:, 0:6] = x0
    rec[:, 6:12] = np.asarray(xt, dtype=np.float64).reshape(-1, 6)
    if lap_off is not None:
        rec[:, 12] = np.asarray(lap_off, dtype=np.float64).reshape(B)
    rec[:, 14:14 + 2 * (N + 1)] = np.asarray(obs, dtype=np.float64).reshape(B, 2 * (N + 1))
    return rec


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def solve_cbf_packed(records, prm, M, xt_per_stage, want=("aux", "x", "u", "sigma"), handle=None, flags=0, **opt):
    """Host-pointer call: H2D + kernel + D2H inside b200mpc_cbf_solve."""
    h = handle or default_handle()
    records = np.ascontiguousarray(records, dtype=np.float64)
    B, N = records.shape[0], int(prm["N"])
    if records.shape[1] != cbf_record_doubles(N, M, xt_per_stage, flags):
        raise ValueError("record stride does not match (N, M, xt_per_stage, flags)")
    p = _capi.make_cbf_params(prm, M, xt_per_stage, flags)
    o = _capi.default_options(**opt)
    rec = np.zeros(B, dtype=_capi.RECORD_DTYPE)
    aux = np.zeros((B, 4)) if "aux" in want else None
    x = np.zeros((B, N + 1, 6)) if "x" in want else None
    u = np.zeros((B, N, 2)) if "u" in want else None
    sg = np.zeros((B, M, N + 1)) if ("sigma" in want and M > 0) else None
    rc = _capi.lib().b200mpc_cbf_solve(h.ptr, C.byref(p), C.byref(o), B, _ptr(records), _ptr(rec), _ptr(aux), _ptr(x),
                                       _ptr(u), _ptr(sg))
    h.check(rc, "b200mpc_cbf_solve")
    out = dict(u0=rec["u0"].copy(), cost=rec["cost"].copy(), status=rec["status"].copy(), iters=rec["iters"].copy(),
               record=rec)
    if aux is not None:
        out.update(kkt_err=aux[:, 0], elastic_max=aux[:, 1], n_refactor=aux[:, 2].astype(int),
                   n_backtrack=aux[:, 3].astype(int))
    if x is not None:
        out["x"] = x
    if u is not None:
        out["u"] = u
    if "sigma" in want:
        out["sigma"] = sg if sg is not None else np.zeros((B, 0, N + 1))
    return out


class _Pipeline:
    """`depth` handles (one CUDA stream + staging buffers each) driven round-robin through an *_async entry point, pinned
    host buffers.  Iteration counts are heavy-tailed (MPC-CBF: mean 37, max 200; iLQR: mean 5, max 41), so a lone batch that
    fills the GPU exactly once leaves most SMs idle while its last instances finish; with the next batches already enqueued
    their instances take the vacated warp slots.  Per-batch latency is unchanged.

        t = pipe.submit(records)          # (B, stride) float64; copied into the slot's pinned buffer
        ...
        rec = pipe.result(t)              # structured (B,) array: cost, u0, status, iters (waits for that batch)
    """

    def __init__(self, B, stride, depth, device):
        self.B, self.stride, self.depth = int(B), int(stride), int(depth)
        self.handles = [_capi.Handle(device=device, max_batch=B) for _ in range(self.depth)]
        self.pin_in = [_capi.PinnedArray((B, self.stride)) for _ in range(self.depth)]
        self.pin_out = [_capi.PinnedArray((B,), _capi.RECORD_DTYPE) for _ in range(self.depth)]
        self.busy = [False] * self.depth
        self.n_submitted = 0

    def _enqueue(self, k):
        raise NotImplementedError

    def submit(self, records, copy=True):
        """Enqueue one batch; returns its ticket.  The slot's previous batch must have been collected."""
        k = self.n_submitted % self.depth
        if self.busy[k]:
            raise RuntimeError("pipeline: collect result(%d) before submitting again" % (self.n_submitted - self.depth))
        if copy:
            self.pin_in[k].a[...] = records
        self._enqueue(k)
        self.busy[k] = True
        self.n_submitted += 1
        return self.n_submitted - 1

    def result(self, ticket, copy=True):
        k = ticket % self.depth
        self.handles[k].synchronize()
        self.busy[k] = False
        return self.pin_out[k].a.copy() if copy else self.pin_out[k].a

    @property
    def launch_count(self):
        return sum(h.launch_count for h in self.handles)

    def close(self):
        for h in self.handles:
            h.close()


class CbfPipeline(_Pipeline):
    """Several MPC-CBF / MPC-LTI / planner-QP batches in flight (b200mpc_cbf_solve_async)."""

    def __init__(self, prm, M, B, depth=4, xt_per_stage=False, flags=0, device=-1, **opt):
        self.N, self.M = int(prm["N"]), int(M)
        self.p = _capi.make_cbf_params(prm, M, xt_per_stage, flags)
        self.o = _capi.default_options(**opt)
        super().__init__(B, cbf_record_doubles(self.N, self.M, xt_per_stage, flags), depth, device)

    def _enqueue(self, k):
        h = self.handles[k]
        rc = _capi.lib().b200mpc_cbf_solve_async(h.ptr, C.byref(self.p), C.byref(self.o), self.B, _ptr(self.pin_in[k].a),
                                                 _ptr(self.pin_out[k].a), None, None, None, None)
        h.check(rc, "b200mpc_cbf_solve_async")


class IlqrPipeline(_Pipeline):
    """Several iLQR batches in flight (b200mpc_ilqr_solve_async; records from pack_ilqr)."""

    def __init__(self, prm, B, depth=4, device=-1):
        self.N = int(prm["N"])
        self.p = _capi.make_ilqr_params(prm)
        super().__init__(B, ilqr_record_doubles(self.N), depth, device)

    def _enqueue(self, k):
        h = self.handles[k]
        rc = _capi.lib().b200mpc_ilqr_solve_async(h.ptr, C.byref(self.p), self.B, _ptr(self.pin_in[k].a), _ptr(self.pin_out[k].a),
                                                  None, None)
        h.check(rc, "b200mpc_ilqr_solve_async")


class LmpcPipeline(_Pipeline):
    """Several LMPC batches in flight (b200mpc_lmpc_solve_async; records from pack_lmpc)."""

    def __init__(self, prm, K, B, depth=4, device=-1, **opt):
        self.N, self.K = int(prm["N"]), int(K)
        self.p = _capi.make_lmpc_params(prm, K)
        self.o = _capi.default_options(**opt)
        super().__init__(B, lmpc_record_doubles(self.N, self.K), depth, device)

    def _enqueue(self, k):
        h = self.handles[k]
        rc = _capi.lib().b200mpc_lmpc_solve_async(h.ptr, C.byref(self.p), C.byref(self.o), self.B, _ptr(self.pin_in[k].a),
                                                  _ptr(self.pin_out[k].a), None, None, None, None)
        h.check(rc, "b200mpc_lmpc_solve_async")


def solve_cbf_batch(x0, xt, obs, lap_off, prm, want=("aux", "x", "u", "sigma"), handle=None, xlb=None, xub=None, wd=None,
                    sizes=None, **opt):
    """Batched control.mpccbf / mpc_lti / mpc_multi_agents solve (control.py:476-607,198-248,251-473); with
    xlb/xub/wd also the planner's candidate QP (planning/overtake_traj_planner.py:248-379)."""
    records, M, per_stage = pack_cbf(x0, xt, obs, lap_off, int(prm["N"]), xlb=xlb, xub=xub, wd=wd, sizes=sizes)
    flags = cbf_flags(M, xlb, xub, wd, sizes)
    return solve_cbf_packed(records, prm, M, per_stage, want=want, handle=handle, flags=flags, **opt)


def solve_ilqr_batch(x0, xt, obs, lap_off, prm, want=("x", "u"), handle=None):
    """Batched control.ilqr (control.py:64-195).  prm: A,B,Q,R,N,max_iter,L,W."""
    h = handle or default_handle()
    N = int(prm["N"])
    records = pack_ilqr(x0, xt, obs, lap_off, N)
    B = records.shape[0]
    p = _capi.make_ilqr_params(prm)
    rec = np.zeros(B, dtype=_capi.RECORD_DTYPE)
    x = np.zeros((B, N + 1, 6)) if "x" in want else None
    u = np.zeros((B, N, 2)) if "u" in want else None
    rc = _capi.lib().b200mpc_ilqr_solve(h.ptr, C.byref(p), B, _ptr(records), _ptr(rec), _ptr(x), _ptr(u))
    h.check(rc, "b200mpc_ilqr_solve")
    out = dict(u0=rec["u0"].copy(), cost=rec["cost"].copy(), converged=(rec["status"] == 0), iters=rec["iters"].copy(),
               record=rec)
    if x is not None:
        out["x"] = x
    if u is not None:
        out["u"] = u
    return out


def pack_lmpc(x0, u_old, A, B, Cm, SS, Qfun, N):
    """x0 (Bn,6); u_old (Bn,2); A (Bn,N,6,6); B (Bn,N,6,2); Cm (Bn,N,6); SS (Bn,6,K); Qfun (Bn,K)."""
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    Bn = x0.shape[0]
    SS = np.asarray(SS, dtype=np.float64).reshape(Bn, 6, -1)
    K = SS.shape[2]
    if not (2 <= N <= _capi.LMPC_NMAX and 1 <= K <= _capi.LMPC_KMAX):
        raise ValueError(f"LMPC sizes out of range: N={N} (2..{_capi.LMPC_NMAX}), K={K} (1..{_capi.LMPC_KMAX})")
    rec = np.zeros((Bn, lmpc_record_doubles(N, K)))
    rec[:, 0:6] = x0
    rec[:, 6:8] = np.asarray(u_old, dtype=np.float64).reshape(Bn, 2)
    o = 8
    rec[:, o:o + 36 * N] = np.asarray(A, dtype=np.float64).reshape(Bn, 36 * N); o += 36 * N
    rec[:, o:o + 12 * N] = np.asarray(B, dtype=np.float64).reshape(Bn, 12 * N); o += 12 * N
    rec[:, o:o + 6 * N] = np.asarray(Cm, dtype=np.float64).reshape(Bn, 6 * N); o += 6 * N
    rec[:, o:o + 6 * K] = SS.reshape(Bn, 6 * K); o += 6 * K
    rec[:, o:o + K] = np.asarray(Qfun, dtype=np.float64).reshape(Bn, K)
    return rec, K


def solve_lmpc_batch(x0, u_old, A, B, Cm, SS, Qfun, prm, want=("aux", "x", "u", "lambda"), handle=None, **opt):
    """Batched control.lmpc QP (control.py:610-730): LTV model, safe-set convex hull terminal constraint."""
    h = handle or default_handle()
    N = int(prm["N"])
    records, K = pack_lmpc(x0, u_old, A, B, Cm, SS, Qfun, N)
    Bn = records.shape[0]
    p = _capi.make_lmpc_params(prm, K)
    o = _capi.default_options(**opt)
    rec = np.zeros(Bn, dtype=_capi.RECORD_DTYPE)
    aux = np.zeros((Bn, 4)) if "aux" in want else None
    x = np.zeros((Bn, N + 1, 6)) if "x" in want else None
    u = np.zeros((Bn, N, 2)) if "u" in want else None
    lam = np.zeros((Bn, K)) if "lambda" in want else None
    rc = _capi.lib().b200mpc_lmpc_solve(h.ptr, C.byref(p), C.byref(o), Bn, _ptr(records), _ptr(rec), _ptr(aux), _ptr(x), _ptr(u),
                                        _ptr(lam))
    h.check(rc, "b200mpc_lmpc_solve")
    out = dict(u0=rec["u0"].copy(), cost=rec["cost"].copy(), status=rec["status"].copy(), iters=rec["iters"].copy(), record=rec)
    if aux is not None:
        out.update(kkt_err=aux[:, 0], n_backtrack=aux[:, 3].astype(int))
    if x is not None:
        out["x"] = x
    if u is not None:
        out["u"] = u
    if lam is not None:
        out["lambda"] = lam
    return out


def pack_laps(ss_xcurv, u_ss, time_ss, used_laps):
    """Stored laps of the reference (ss_xcurv (T,6,L), u_ss (T,2,L), time_ss (L); base.py:430-435) -> the kernel's
    structure-of-arrays block [lap][vx, vy, wz, delta, a][lap_stride] and the per-lap row counts."""
    used_laps = list(used_laps)
    rows = [int(time_ss[l]) for l in used_laps]
    stride = (max(rows) + 2 + 1) & ~1
    if stride > np.asarray(ss_xcurv).shape[0]:
        stride = np.asarray(ss_xcurv).shape[0]
    laps = np.zeros((len(used_laps), 5, stride))
    for k, l in enumerate(used_laps):
        laps[k, 0:3, :] = np.asarray(ss_xcurv, dtype=np.float64)[:stride, 0:3, l].T
        laps[k, 3:5, :] = np.asarray(u_ss, dtype=np.float64)[:stride, :, l].T
    return laps, rows, stride


def estimate_abc_batch(lin_points, lin_input, ss_xcurv, u_ss, time_ss, used_laps, point_and_tangent, dt, max_num_point=40,
                       h=5.0, want=("idx", "status"), handle=None, out=None, out_stride=None, out_offset=0):
    """Batched LMPCRacingGame.estimate_ABC (base.py:585-622).  lin_points (Bn,N+1,6)|(N+1,6), lin_input (Bn,N,2)|(N,2).
    Returns A (Bn,N,6,6), B (Bn,N,6,2), C (Bn,N,6) [+ idx, status].  With `out` (Bn,out_stride) the model block is written
    into caller records (e.g. LMPC records, out_offset 8) instead."""
    hd = handle or default_handle()
    lp = np.asarray(lin_points, dtype=np.float64)
    li = np.asarray(lin_input, dtype=np.float64)
    if lp.ndim == 2:
        lp, li = lp[None], li[None]
    Bn, N = li.shape[0], li.shape[1]
    lin = np.ascontiguousarray(np.concatenate([lp[:, :N, :], li], axis=2))       # (Bn, N, 8)
    laps, rows, stride = pack_laps(ss_xcurv, u_ss, time_ss, used_laps)
    pat = np.asarray(point_and_tangent, dtype=np.float64)
    seg = np.ascontiguousarray(pat[:, 3:6])
    p = _capi.SysidParams()
    p.N, p.num_laps, p.max_num_point, p.num_segments = N, len(rows), int(max_num_point), seg.shape[0]
    for k, r in enumerate(rows):
        p.lap_rows[k] = r
    p.lap_stride = stride
    p.dt, p.h, p.lap_length = float(dt), float(h), float(pat[-1, 3] + pat[-1, 4])
    own = out is None
    if own:
        out_stride, out_offset = 54 * N, 0
        out = np.zeros((Bn, out_stride))
    idx = np.zeros((Bn, N, len(rows), int(max_num_point)), dtype=np.int32) if "idx" in want else None
    st = np.zeros((Bn, N), dtype=np.int32) if "status" in want else None
    rc = _capi.lib().b200mpc_lmpc_sysid(hd.ptr, C.byref(p), Bn, _ptr(lin), _ptr(laps), _ptr(seg), _ptr(out), int(out_stride),
                                        int(out_offset), _ptr(idx), _ptr(st))
    hd.check(rc, "b200mpc_lmpc_sysid")
    blk = out[:, out_offset:out_offset + 54 * N]
    res = dict(A=blk[:, :36 * N].reshape(Bn, N, 6, 6), B=blk[:, 36 * N:48 * N].reshape(Bn, N, 6, 2),
               C=blk[:, 48 * N:54 * N].reshape(Bn, N, 6))
    if idx is not None:
        res["idx"] = idx
    if st is not None:
        res["status"] = st
    return res


def plant_step_batch(xcurv, xglob, u, draws, point_and_tangent, timestep=0.1, dyn=_capi.DEFAULT_DYN, wrap_lap=False, handle=None):
    """Batched DynamicBicycleModel.forward_dynamics (base.py:897-942).  xcurv, xglob (Bn,6); u (Bn,2); draws (Bn,3) standard
    normal draws or None (zero_noise_flag).  Returns (xcurv_next, xglob_next[, laps])."""
    hd = handle or default_handle()
    xc = np.ascontiguousarray(np.atleast_2d(np.asarray(xcurv, dtype=np.float64))).copy()
    xg = np.ascontiguousarray(np.atleast_2d(np.asarray(xglob, dtype=np.float64))).copy()
    uu = np.ascontiguousarray(np.atleast_2d(np.asarray(u, dtype=np.float64)))
    Bn = xc.shape[0]
    dr = None if draws is None else np.ascontiguousarray(np.asarray(draws, dtype=np.float64).reshape(Bn, 3))
    pat = np.asarray(point_and_tangent, dtype=np.float64)
    seg = np.ascontiguousarray(pat[:, 3:6])
    p = _capi.make_plant_params(seg.shape[0], pat[-1, 3] + pat[-1, 4], timestep=timestep, dyn=dyn, wrap_lap=wrap_lap)
    laps = np.zeros(Bn, dtype=np.int32) if wrap_lap else None
    rc = _capi.lib().b200mpc_plant_step(hd.ptr, C.byref(p), Bn, _ptr(xc), 6, 0, _ptr(xg), _ptr(uu), 2, _ptr(dr), _ptr(seg), _ptr(laps))
    hd.check(rc, "b200mpc_plant_step")
    return (xc, xg, laps) if wrap_lap else (xc, xg)


def rival_rollout_batch(xcurv, xglob, point_and_tangent, lap_length, timestep, n, with_glob=False, handle=None):
    """Batched offboard.DynamicBicycleModel.get_trajectory_nsteps (racing/offboard.py:80-94): zero-input Frenet rollout of Bn
    rivals over n steps in one launch.  xcurv, xglob (Bn,6).  Returns xcurv_nsteps (Bn,6,n) and, on request, xglob_nsteps."""
    hd = handle or default_handle()
    xc = np.ascontiguousarray(np.atleast_2d(np.asarray(xcurv, dtype=np.float64)))
    xg = np.ascontiguousarray(np.atleast_2d(np.asarray(xglob, dtype=np.float64)))
    if xc.shape != xg.shape or xc.shape[1] != 6:
        raise ValueError("xcurv and xglob must both be (B, 6)")
    Bn = xc.shape[0]
    seg = np.ascontiguousarray(np.asarray(point_and_tangent, dtype=np.float64)[:, 3:6])
    p = _capi.RolloutParams()
    p.n, p.num_segments, p.timestep, p.lap_length = int(n), seg.shape[0], float(timestep), float(lap_length)
    out_c = np.zeros((Bn, 6, int(n)))
    out_g = np.zeros((Bn, 6, int(n))) if with_glob else None
    rc = _capi.lib().b200mpc_rival_rollout(hd.ptr, C.byref(p), Bn, _ptr(xc), _ptr(xg), _ptr(seg), _ptr(out_c), _ptr(out_g))
    hd.check(rc, "b200mpc_rival_rollout")
    return (out_c, out_g) if with_glob else out_c


def curv_to_glob_batch(s, ey, point_and_tangent, lap_length, handle=None):
    """Batched racing_env.get_global_position / get_orientation (utils/racing_env.py:6-127): s, ey of any (equal) shape ->
    (x, y, psi) arrays of that shape, one launch.  The controllers' global-frame logs (utils/base.py:500-509, 573-580) and
    the planner's plot copies (planner_helper.py:208-220) call the scalar reference functions once per point."""
    hd = handle or default_handle()
    sa = np.ascontiguousarray(np.asarray(s, dtype=np.float64))
    ea = np.ascontiguousarray(np.asarray(ey, dtype=np.float64))
    if sa.shape != ea.shape or sa.size < 1:
        raise ValueError("s and ey must have the same non-empty shape")
    pat = np.ascontiguousarray(np.asarray(point_and_tangent, dtype=np.float64)[:, :6])
    out = np.zeros((sa.size, 3))
    rc = _capi.lib().b200mpc_curv_to_glob(hd.ptr, sa.size, pat.shape[0], float(lap_length), _ptr(pat), _ptr(sa), 1, _ptr(ea), 1,
                                          _ptr(out))
    hd.check(rc, "b200mpc_curv_to_glob")
    return out[:, 0].reshape(sa.shape), out[:, 1].reshape(sa.shape), out[:, 2].reshape(sa.shape)
