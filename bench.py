#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: MPC-CBF solves/sec (N=20, 6-state, 3 rivals); the other BASELINE configs on request.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--batch B]

A "step" is one pass of the hot path over one batch of synthetic scenarios of the chosen BASELINE.json config
(car_racing_b200/scenarios.py, seed 1 + rank):
    --config 2 (default)  MPC-CBF N=20, 3 static rivals, l_shape, 1024 scenarios per GPU          (configs[1], the metric's config)
    --config 3            overtake planner: 64 candidate QPs x 2 moving rivals, goggle, per GPU  (configs[2])
    --config 4            LMPC N=12, 44 safe-set points, ellipse, 512 scenarios per GPU           (configs[3]: 4096 over 8 GPUs)
    --config 5            iLQR N=50, 1024 scenarios per GPU                                      (configs[4]: 8192 over 8 GPUs)
With N>1 (torchrun, one rank per GPU) every rank solves its own shard (weak scaling) and the step ends with the exchange
of the 32-byte records: the solver kernels' epilogue stores them into every rank's gathered buffer over NVLink
(b200mpc_comm_*, csrc/exchange.cuh), then one argmin kernel per rank.  If the peers' windows cannot be mapped (no CUDA IPC),
the exchange falls back to one NCCL all-gather per step and the line says so (`config.exchange`).

`value`    : whole-job solves/s with the packed inputs already resident in HBM (kernels only), exchange included.
`e2e`      : the same metric through the host-pointer C-ABI call (pinned host buffers; H2D of the records, kernel, D2H of
             the 32-byte result records inside the timed region; N>1: the exchange and the D2H of its argmin as well).
`roofline` : algorithmic HBM bytes / kernel time against the measured copy bandwidth -- stated honestly: these kernels
             are bound by FP64 issue and instruction delivery, not by HBM; `roofline.fp64` puts the counted FP64 work beside it.
`--impl reference` times the CPU implementation of the path on the host cores: CasADi/IPOPT is probed for, but it is not
installable here, so the arm times the oracle port (oracle/*.c) and says so.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# every stream its own hardware queue (D solver streams + the exchange stream + torch's + NCCL's): with the default of 8
# connections streams alias, and the exchange window's waiting kernels would block unrelated streams behind them
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

OUT = sys.stdout


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
class Work:
    """One BASELINE config: host records of this rank's shard, the three ways to run it (device pointers, host pointers,
    CPU oracle) and its algorithmic bytes per solve."""
    key = metric = kernel = None
    default_batch = 1024
    # A batch ends with its own stragglers (heavy-tailed iteration counts), so one stream turns a batch around in the time of
    # its slowest instance; D streams overlap D batches until the SMs are full.  MPC-CBF / LMPC fill the SMs at D = 6; an iLQR
    # solve is so short (mean 5 iterations, the slowest instance of a shard 25-45) that the SMs are full only at D = 16.
    default_inflight = 6

    def config(self, B, world, D, exchange, l2="rotate"):
        """The `config` object of the JSON line -- the same for both arms (--impl ours / reference)."""
        return {"workload": self.workload % B, "baseline_config": self.key, "global_batch": B * world,
                "parallelism": "dp%d: independent scenario shards, one exchange of the 32-byte records + argmin per step" % world,
                "exchange": exchange, "batches_in_flight": D,
                "pipelining": ("steps are issued round-robin on %d streams (one handle each): the stragglers of one batch overlap "
                               "the next batches; timed first start event -> last end event, flushes included" % D)
                if D > 1 else "none: one batch at a time",
                "l2": L2_TEXT[l2],
                "solver": self.solver}


class Work2(Work):
    key, default_batch, kernel = 2, 1024, "ocp_ipm_kernel<3,QDIAG,20>"
    metric = "MPC-CBF solves/sec (N=20, 6-state, 3 obs)"
    workload = "MPC-CBF N=20, 3 static rivals, l_shape, batch=%d random x0 per GPU (BASELINE config 2)"
    solver = "FP64 barrier-SQP (IPOPT conventions), Riccati KKT, tol 1e-8"
    alg_out = 536

    def __init__(self, B, seed):
        from car_racing_b200 import _capi, batch, scenarios
        self.raw = scenarios.mpccbf_scenarios(B, N=20, M=3, seed=seed)
        self.prm = scenarios.default_cbf_params(N=20)
        self.rec, M, ps = batch.pack_cbf(*self.raw, 20)
        self.p, self.o = _capi.make_cbf_params(self.prm, M, ps), _capi.default_options()
        self.alg_bytes = self.rec.shape[1] * 8 + self.alg_out

    def device(self, L, h, B, d_in, d_rec):
        import ctypes as C
        return L.b200mpc_cbf_solve_device(h.ptr, C.byref(self.p), C.byref(self.o), B, d_in, d_rec, None, None, None, None)

    def host(self, L, h, B, p_in, p_out, blocking=False):
        import ctypes as C
        fn = L.b200mpc_cbf_solve if blocking else L.b200mpc_cbf_solve_async
        return fn(h.ptr, C.byref(self.p), C.byref(self.o), B, p_in, p_out, None, None, None, None)

    def cpu(self, orc, n, nthreads):
        x0, xt, obs, lo = self.raw
        return orc.solve_cbf_batch(x0[:n], xt, obs[:n], lo[:n], self.prm, nthreads=nthreads)


class Work3(Work2):
    key, default_batch, kernel = 3, 64, "ocp_ipm_kernel<0,3,0>"
    metric = "overtake-planner candidate solves/sec (N=10, 64 candidates, 2 rivals)"
    workload = "overtake planner: %d lane/side/gap candidate QPs x 2 moving rivals, goggle, N=10, per GPU (BASELINE config 3)"
    solver = "FP64 barrier method (IPOPT conventions) on the candidate QP, Riccati KKT, tol 1e-8"

    def __init__(self, B, seed):
        from car_racing_b200 import _capi, batch, planning, scenarios
        sc = scenarios.planner_scenarios(C=B, N=10, seed=seed)
        self.kw, _ = planning.pack_candidates(sc["x0"], sc["s_ref"], sc["ey_ref"], sc["xlb"], sc["xub"], 10)
        self.prm = planning.planner_params(scenarios.LTI_A, scenarios.LTI_B, 10)
        kw = self.kw
        self.rec, M, ps = batch.pack_cbf(kw["x0"], kw["xt"], kw["obs"], None, 10, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
        self.p = _capi.make_cbf_params(self.prm, 0, True, _capi.FLAG_STAGE_BOUNDS | _capi.FLAG_EY_RATE)
        self.o = _capi.default_options()
        self.alg_bytes = self.rec.shape[1] * 8 + 6 * 11 * 8 + 16        # in: the packed record; out: trajectory 6 x 11 + costs

    def cpu(self, orc, n, nthreads):
        kw = self.kw
        return orc.solve_cbf_batch(kw["x0"][:n], kw["xt"][:n], kw["obs"][:n], None, self.prm, xlb=kw["xlb"][:n], xub=kw["xub"][:n],
                                   wd=kw["wd"][:n], nthreads=nthreads)


class Work4(Work):
    key, default_batch, kernel = 4, 512, "lmpc_kernel"
    metric = "LMPC solves/sec (N=12, 44 safe-set points)"
    workload = "LMPC N=12, learned safe-set terminal constraint (44 points), ellipse, batch=%d scenarios per GPU (BASELINE config 4: 4096 over 8 GPUs)"
    solver = "FP64 barrier method (IPOPT conventions), condensed KKT, tol 1e-8"

    def __init__(self, B, seed):
        from car_racing_b200 import _capi, batch, scenarios
        self.raw = scenarios.lmpc_scenarios(B, seed=seed)
        self.prm = scenarios.default_lmpc_params()
        self.rec, K = batch.pack_lmpc(*self.raw, int(self.prm["N"]))
        self.p, self.o = _capi.make_lmpc_params(self.prm, K), _capi.default_options()
        self.alg_bytes = self.rec.shape[1] * 8 + (13 * 6 + 12 * 2) * 8

    def device(self, L, h, B, d_in, d_rec):
        import ctypes as C
        return L.b200mpc_lmpc_solve_device(h.ptr, C.byref(self.p), C.byref(self.o), B, d_in, d_rec, None, None, None, None)

    def host(self, L, h, B, p_in, p_out, blocking=False):
        import ctypes as C
        fn = L.b200mpc_lmpc_solve if blocking else L.b200mpc_lmpc_solve_async
        return fn(h.ptr, C.byref(self.p), C.byref(self.o), B, p_in, p_out, None, None, None, None)

    def cpu(self, orc, n, nthreads):
        return orc.solve_lmpc_batch(*[a[:n] for a in self.raw], self.prm, nthreads=nthreads)


class Work5(Work):
    key, default_batch, kernel, default_inflight = 5, 1024, "ilqr_kernel", 16
    metric = "iLQR solves/sec (N=50, 1 rival)"
    workload = "iLQR N=50, LTI bicycle model + 1 rival, batch=%d random (x0, obstacle) scenarios per GPU (BASELINE config 5: 8192 over 8 GPUs)"
    solver = "FP64 iLQR as control.ilqr (max_iter 150, regularised backward pass)"

    def __init__(self, B, seed):
        from car_racing_b200 import _capi, batch, scenarios
        self.raw = scenarios.ilqr_scenarios(B, N=50, seed=seed)
        p = scenarios.default_cbf_params()
        self.prm = dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=50, max_iter=150, L=0.4, W=0.2)
        self.rec = batch.pack_ilqr(*self.raw, 50)
        self.p = _capi.make_ilqr_params(self.prm)
        self.alg_bytes = self.rec.shape[1] * 8 + 32

    def device(self, L, h, B, d_in, d_rec):
        import ctypes as C
        return L.b200mpc_ilqr_solve_device(h.ptr, C.byref(self.p), B, d_in, d_rec, None, None)

    def host(self, L, h, B, p_in, p_out, blocking=False):
        import ctypes as C
        fn = L.b200mpc_ilqr_solve if blocking else L.b200mpc_ilqr_solve_async
        return fn(h.ptr, C.byref(self.p), B, p_in, p_out, None, None)

    def cpu(self, orc, n, nthreads):
        x0, xt, obs, lo = self.raw
        return orc.solve_ilqr_batch(x0[:n], xt, obs[:n], lo[:n], self.prm, nthreads=nthreads)


WORKS = {2: Work2, 3: Work3, 4: Work4, 5: Work5}
L2_BYTES = 126 * 1024 * 1024
L2_TEXT = {
    "rotate": ("inputs larger than L2: every step reads its records from the next of R device copies (R x record bytes > 1.15 x the 126 MB "
               "L2, round-robin over the whole run), so a copy is touched again only after more than an L2's worth of other inputs; the "
               "end-to-end leg (fresh H2D every step) writes a 256 MiB buffer before every step instead"),
    "flush": "256 MiB buffer written before every timed step, on the step's stream (inputs are << L2)"}


def reference_solver_probe():
    """is the reference's own solver (CasADi + IPOPT) or an install of the reference present on this box?"""
    probe = {"casadi": False, "reference_install": os.path.isdir(os.path.join(ROOT, "baseline", "_ref")),
             "reference_checkout": os.path.isdir("/root/reference")}
    try:
        import casadi  # noqa: F401
        probe["casadi"] = True
    except Exception:
        pass
    return probe


def cpu_port_rate(work, n, nthreads, budget_s=10.0, max_passes=16):
    """The CPU restatement (oracle/) on `nthreads` host threads over the first n instances of the same workload; whole
    passes until ~budget_s of wall time.  Returns (solves/s, seconds, passes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.lib()
    t = time.perf_counter()
    passes = 0
    while True:
        work.cpu(orc, n, nthreads)
        passes += 1
        dt = time.perf_counter() - t
        if dt >= budget_s or passes >= max_passes:
            break
    return n * passes / dt, dt, passes


def cpu_baseline_obj(work, v, cores, sample, probe):
    return {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample,
            "build": "oracle/Makefile: gcc -O3 -march=x86-64-v3 -ffp-contract=off, pthreads",
            "reference_solver_probe": probe,
            "note": ("CasADi/IPOPT is not importable here and the reference has no compiled sources: the CPU arm is the oracle port "
                     "of the same problem and interior-point method" if work.key != 5 else
                     "C port of control.ilqr (pinned to the reference's own outputs, tests/golden/ilqr_golden.npz); the reference's "
                     "numpy control.ilqr itself needs the /root/reference checkout, which does not travel to the GPU box "
                     "(measured in the build container: p50 20 ms per solve on one core, SURVEY.md 8(d))")}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this metric on all host threads (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.batch
    work = WORKS[args.config](B, 1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.lib()
    probe = reference_solver_probe()
    for _ in range(args.warmup):
        work.cpu(orc, min(64, B), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        work.cpu(orc, B, cores)
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    line = {"impl": "reference", "metric": work.metric, "value": val, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": work.config(B, max(world, args.gpus), max(1, args.inflight), exchange_name(max(world, args.gpus), None), args.l2),
            "cpu_baseline": cpu_baseline_obj(work, val, cores, "%d instances per step, %d pthreads" % (B, cores), probe),
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=OUT, flush=True)


def exchange_name(world, px):
    if world <= 1:
        return "none (one GPU)"
    if px is None or px == "p2p":
        return "p2p: records stored into every rank's window from the solver epilogue (b200mpc_comm_*), 1 argmin kernel per rank"
    return "nccl: all_gather_into_tensor of the records + argmin kernel (peer windows could not be mapped: %s)" % px


def _claim_stdout():
    """Rank 0 must print ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter) is
    sent to stderr; the returned file object is the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def fp64_roofline(work, solves_per_s, sm_mhz):
    """Counted FP64 work of the kernel (ncu: smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on, tools/fp64_counts.py)
    against 148 SMs x 64 FMA lanes x 2 flop x clock.  The counts are per kernel build: the file records the library's hash."""
    p = os.path.join(ROOT, "profiles", "fp64_counts.json")
    if not os.path.exists(p):
        return None
    doc = json.load(open(p))
    ent = doc.get("kernels", {}).get(work.kernel)
    if not ent:
        return None
    h = hashlib.sha256()            # identity of the kernels: the CUDA sources + the header (tools/fp64_counts.py: source_sha16)
    csrc = os.path.join(ROOT, "car_racing_b200", "csrc")
    for f in sorted(os.listdir(csrc)) + ["../../include/b200mpc.h"]:
        h.update(open(os.path.join(csrc, f), "rb").read())
    sha = h.hexdigest()[:16]
    peak = 148 * 64 * 2 * (sm_mhz or 1965.0) * 1e6 / 1e12
    ach = ent["flops_per_solve"] * solves_per_s / 1e12
    return {"flops_per_solve": ent["flops_per_solve"], "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_def": "148 SM x 64 FP64 FMA/clk x 2 x %.0f MHz" % (sm_mhz or 1965.0),
            "source": "profiles/fp64_counts.json (%s)" % ent.get("report", "?"), "counts_from_these_sources": doc.get("src_sha16") == sha}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKS), help="BASELINE.json config (2 = the metric's)")
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU per step (default: the config's)")
    ap.add_argument("--inflight", type=int, default=None,
                    help="batches in flight (streams driven round-robin); 1 = strictly one batch at a time; default: the config's "
                         "(6; 16 for the iLQR config, whose short solves are bound by each batch's own stragglers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: force the NCCL all-gather form of the exchange")
    ap.add_argument("--same-shards", action="store_true",
                    help="diagnostic: every rank solves the SAME scenarios (seed 1) instead of its own shard (seed 1 + rank) -- "
                         "separates the cost of the exchange from the unequal work of random shards")
    ap.add_argument("--l2", default="rotate", choices=["rotate", "flush"],
                    help="how the timed steps are kept from re-reading their inputs out of L2: rotate through device copies of the "
                         "records that together exceed L2 (default), or write a 256 MiB buffer before every step")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = WORKS[args.config].default_batch
    if args.inflight is None:
        args.inflight = WORKS[args.config].default_inflight
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global OUT
    OUT = _claim_stdout()
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    import ctypes as C
    import torch
    import torch.distributed as dist
    from car_racing_b200 import _capi, sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    D = max(1, args.inflight)
    work = WORKS[args.config](B, seed=1 if args.same_shards else 1 + rank)      # each rank: its own shard of scenarios
    rec_host = np.ascontiguousarray(work.rec)
    hs = [_capi.Handle(device=local_rank, max_batch=B) for _ in range(D)]   # one stream + staging buffers per slot
    L = _capi.lib()
    exts = [torch.cuda.ExternalStream(h.stream, device=dev) for h in hs]
    d_in = torch.from_numpy(rec_host).to(dev)
    # --l2 rotate: R copies of this rank's records (different addresses, same contents), together > 1.15 x L2; step i reads copy i mod R
    R_COPIES = max(2, -(-int(1.15 * L2_BYTES) // rec_host.nbytes)) if args.l2 == "rotate" else 1
    d_ins = [d_in] + [d_in.clone() for _ in range(R_COPIES - 1)]
    n_steps_issued = [0]

    def next_input():
        p = d_ins[n_steps_issued[0] % R_COPIES].data_ptr()
        n_steps_issued[0] += 1
        return p
    d_rec = [torch.zeros((B, 4), dtype=torch.float64, device=dev) for _ in range(D)]
    d_all = [torch.zeros((world * B, 4), dtype=torch.float64, device=dev) for _ in range(D)]
    d_arg = [torch.full((1,), -7, dtype=torch.int32, device=dev) for _ in range(D)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    torch.cuda.synchronize()

    # ---- N>1: the exchange.  p2p: the library's window (the solver epilogue stores the records into every rank's gathered
    #      buffer); its argmin kernel runs on the exchange handle's stream and waits on the device for the arrivals.
    #      One exchange stream PER SLOT: an argmin kernel waits for its step to finish on every rank, and steps finish out of
    #      order (each batch ends with its own stragglers) -- on one stream the consumers would serve the slots in issue order
    #      and a slot whose step is long done would wait behind one that is not (measured with one stream: iLQR, 2 GPUs,
    #      0.91 ms per step against 0.59 on one GPU).
    hcs = [_capi.Handle(device=local_rank, max_batch=B) for _ in range(D)] if world > 1 else []
    ext_cs = [torch.cuda.ExternalStream(h.stream, device=dev) for h in hcs]
    hc = hcs[0] if hcs else None
    px, px_why = None, "p2p"
    if world > 1 and args.exchange == "p2p":
        try:
            px = sharding.PeerExchange(hc, rank, world, max_batch=B, slots=D)
        except Exception as e:                       # CUDA IPC / peer access unavailable on this box
            px, px_why = None, str(e).replace("\n", " ")[:160]
        ok = torch.tensor([1 if px is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if px is not None:
                px.close()
            px = None
    elif world > 1:
        px_why = "--exchange nccl"
    ag_done = [None] * D

    def step_device(k):
        """One step = one batch of B instances: L2 flush, the solve kernel on slot k's stream; (N>1) the exchange."""
        h = hs[k]
        with torch.cuda.stream(exts[k]):
            if args.l2 == "flush":
                flush.zero_()
            if px is not None:
                px.publish_next(h, k)
            elif world > 1 and ag_done[k] is not None:
                exts[k].wait_event(ag_done[k])            # the all-gather that last read this result buffer
            h.check(work.device(L, h, B, next_input(), d_rec[k].data_ptr()), "solve_device")
            if world > 1 and px is None:
                ev = torch.cuda.Event()
                ev.record(exts[k])
        if px is not None:
            px.argmin(hcs[k], k, d_arg[k].data_ptr(), d_all[k].data_ptr())
        elif world > 1:
            with torch.cuda.stream(ext_cs[0]):        # NCCL completes collectives in issue order: one stream
                ext_cs[0].wait_event(ev)
                dist.all_gather_into_tensor(d_all[k], d_rec[k])
                hc.check(L.b200mpc_argmin_cost_device(hc.ptr, d_all[k].data_ptr(), world * B, 0, d_arg[k].data_ptr()), "argmin")
                ag_done[k] = torch.cuda.Event()
                ag_done[k].record(ext_cs[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def launches():
        return sum(h.launch_count for h in hs) + sum(h.launch_count for h in hcs)

    for i in range(max(args.warmup, D)):
        step_device(i % D)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- (a) one batch at a time: per-step CUDA events on the launching stream (the kernel's launch duration for the
    #      roofline, and the latency of one batch); no exchange here
    evs = []
    for _ in range(args.steps):
        with torch.cuda.stream(exts[0]):
            if args.l2 == "flush":
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(exts[0])
            hs[0].check(work.device(L, hs[0], B, next_input(), d_rec[0].data_ptr()), "solve_device")
            e1.record(exts[0])
            evs.append((e0, e1))
    barrier()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # ---- (b) the timed region of the contract: exactly K steps, D batches in flight (slot = step mod D, each slot its
    #      own stream), L2 flush before every step, bracketed by barrier + synchronize; device time = first start event
    #      to the last end event over all streams (exchange stream included)
    launches0 = launches()
    barrier()
    wall0 = time.perf_counter()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(D)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(D)]
    for k in range(D):
        starts[k].record(exts[k])
    for i in range(args.steps):
        step_device(i % D)
    for k in range(D):
        ends[k].record(exts[k])
    for ec in ext_cs:
        end_c = torch.cuda.Event(enable_timing=True)
        end_c.record(ec)
        ends.append(end_c)
    barrier()
    wall = time.perf_counter() - wall0
    launches_dev = launches() - launches0
    t_dev = max(sa.elapsed_time(eb) for sa in starts for eb in ends) * 1e-3
    status = sharding.tensor_to_records(d_rec[0])["status"]
    conv = float((status == 0).mean())

    # ---- the exchange delivered what a host-side gather + argmin gives: every slot's gathered buffer holds this rank's
    #      records at its shard, all ranks hold the same buffer, and the device argmin equals numpy's first-min
    exchange_ok = None
    if world > 1:
        exchange_ok = True
        for k in range(min(D, args.steps)):
            allr = sharding.tensor_to_records(d_all[k])
            mine = sharding.tensor_to_records(d_rec[k])
            same_shard = bool((allr[rank * B:(rank + 1) * B].tobytes() == mine.tobytes()))
            arg_ok = int(d_arg[k].item()) == sharding.argmin_first(allr)
            digest = torch.tensor([float(np.frombuffer(hashlib.sha256(allr.tobytes()).digest()[:6], dtype=np.uint16).sum())],
                                  dtype=torch.float64, device=dev)
            lo, hi = digest.clone(), digest.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            exchange_ok = exchange_ok and same_shard and arg_ok and float(lo.item()) == float(hi.item())

    # ---- (c) end to end through the host-pointer C-ABI, the same D slots: every step copies its records from pinned
    #      host memory (H2D), solves, copies the 32-byte result records back (D2H) and the host reads them; N>1: the step
    #      also publishes into the exchange window, and the argmin over the gathered records is copied back and read
    pin_in = [torch.from_numpy(rec_host).pin_memory() for _ in range(D)]
    pin_out = [torch.zeros((B, 4), dtype=torch.float64).pin_memory() for _ in range(D)]
    pin_arg = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(D)]
    arg_ev = [None] * D

    def e2e_run(nsteps):
        acc = 0.0
        for i in range(nsteps):
            k = i % D
            if i >= D:
                hs[k].synchronize()
                acc += float(pin_out[k][0, 0])                # the step's result is read on the host
                if arg_ev[k] is not None:
                    arg_ev[k].synchronize()
                    acc += float(pin_arg[k][0])
            with torch.cuda.stream(exts[k]):
                flush.zero_()
            if px is not None:
                px.publish_next(hs[k], k)
            hs[k].check(work.host(L, hs[k], B, pin_in[k].data_ptr(), pin_out[k].data_ptr(), blocking=(D == 1)), "solve(host)")
            if px is not None:
                px.argmin(hcs[k], k, d_arg[k].data_ptr(), d_all[k].data_ptr())     # the same exchange work as the `value` leg
                with torch.cuda.stream(ext_cs[k]):
                    pin_arg[k].copy_(d_arg[k], non_blocking=True)
                    arg_ev[k] = torch.cuda.Event()
                    arg_ev[k].record(ext_cs[k])
        for k in range(D):
            hs[k].synchronize()
            acc += float(pin_out[k][0, 0])
            if arg_ev[k] is not None:
                arg_ev[k].synchronize()
                acc += float(pin_arg[k][0])
        return acc

    e2e_run(max(args.warmup, D))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    # p50 latency of a single solve (B=1) and of one whole batch through the blocking call
    lat, lat_b = [], []
    for k in range(220):      # 20 warm-ups, 200 timed single-instance calls
        t0 = time.perf_counter()
        work.host(L, hs[0], 1, pin_in[0].data_ptr(), pin_out[0].data_ptr(), blocking=True)
        if k >= 20:
            lat.append(1e3 * (time.perf_counter() - t0))
    for k in range(8):
        t0 = time.perf_counter()
        work.host(L, hs[0], B, pin_in[0].data_ptr(), pin_out[0].data_ptr(), blocking=True)
        if k >= 2:
            lat_b.append(1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop()
    launches_total = launches() - launches0

    # ---- max over ranks
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, wall = [float(v) for v in tt.tolist()]
        cv = torch.tensor([conv, 1.0 if exchange_ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(cv, op=dist.ReduceOp.MIN)
        conv, exchange_ok = float(cv[0].item()), bool(cv[1].item() > 0.5)
    total = B * world * args.steps
    value = total / t_dev
    hbm_peak, sm_max, peak_src = peaks()
    achieved = work.alg_bytes * B / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        ent = t.get("kernels", {}).get(work.kernel) or (t if work.key == 2 and "dram_bytes_per_launch" in t else None)
        if ent:
            traffic, traffic_src = ent.get("dram_bytes_per_launch"), "profiles/dram_traffic.json (ncu counters dram__bytes_read.sum + dram__bytes_write.sum of one B=%s launch, tools/fp64_counts.py)" % ent.get("batch", "1024")
    cfg = work.config(B, world, D, exchange_name(world, "p2p" if px is not None else px_why), args.l2)
    line = {
        "metric": work.metric, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic" + (" (diagnostic: the same shard on every rank)" if args.same_shards else ""),
        "config": cfg, "converged_frac": conv, "input_copies_rotated": R_COPIES,
        "e2e": {"value": total / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(rec_host.nbytes),
                "d2h_bytes_per_step": int(B * 32 + (4 if px is not None else 0)), "ms_per_step": 1e3 * t_e2e / args.steps,
                "batches_in_flight": D, "p50_latency_ms_batch1": float(np.median(lat)),
                "p10_p90_latency_ms_batch1": [float(np.percentile(lat, 10)), float(np.percentile(lat, 90))],
                "p50_latency_ms_one_batch": float(np.median(lat_b)),
                "note": ("per rank: host records -> H2D -> solve (its epilogue stores the records into every rank's window) -> D2H -> host "
                         "read; + the argmin over the gathered records -> D2H -> host read" if px is not None else
                         "host records -> H2D -> solve -> D2H -> host read" +
                         ("; the NCCL form of the exchange is part of `value` only" if world > 1 else ""))},
        "one_batch_at_a_time": {"value": B * world / (kernel_ms * 1e-3), "unit": "solves/s", "ms_per_step": kernel_ms,
                                "note": "the same kernel, steps serialised on one stream (per-step CUDA events), no exchange; its launch duration is the roofline's"},
        "gpu_launches": int(launches_dev),
        "gpu_launches_incl_e2e": int(launches_total),
        "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": work.kernel,
                     "note": "algorithmic bytes = %d B/solve x %d solves per launch / mean launch duration (one batch at a time); the kernel "
                             "is bound by FP64 issue and instruction delivery, not by HBM: see roofline.fp64, profiles/ and DESIGN.md section 5" % (work.alg_bytes, B)},
    }
    if exchange_ok is not None:
        line["exchange_verified"] = exchange_ok
    f64 = fp64_roofline(work, value / world, clocks.get("sm_mhz") or sm_max)
    if f64:
        line["roofline"]["fp64"] = f64
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = min(B, 1024)
        v, dt, passes = cpu_port_rate(work, n, cores)
        line["cpu_baseline"] = cpu_baseline_obj(work, v, cores, "the same %d scenarios, %d passes, %d pthreads (%.1f s)" % (n, passes, cores, dt),
                                                reference_solver_probe())
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), file=OUT, flush=True)
    if px is not None:
        barrier()
        px.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
