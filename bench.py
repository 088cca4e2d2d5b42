#!/usr/bin/env python
"""bench.py -- MPC-CBF solves/sec (N=20, 6-state, 3 rivals), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one pass of the hot path over one batch of B=1024 synthetic config-2 scenarios
(car_racing_b200/scenarios.py, seed 1).  With N>1 (torchrun, one rank per
GPU) every rank solves its own 1024-scenario shard (weak scaling) and the step ends with one NCCL
all-gather of the 32-byte records + the argmin kernel.

`value`    : whole-job solves/s with the packed inputs already resident in HBM (kernel only).
`e2e`      : the same metric through the host-pointer C-ABI call (pinned host buffers; H2D of the
             records, kernel, D2H of the 32-byte result records inside the timed region).
`roofline` : algorithmic HBM bytes / kernel time against the measured copy bandwidth -- stated
             honestly: this kernel is FP64-latency / shared-memory bound, not HBM bound.
`--impl reference` times the CPU implementation of the path (the oracle port: CasADi/IPOPT is not
installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT = sys.stdout
METRIC = "MPC-CBF solves/sec (N=20, 6-state, 3 obs)"
ALG_BYTES_PER_SOLVE = 1136 + 536      # 1136 B in (one packed record) + 536 B out
N_H, M_OBS = 20, 3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


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


def workload(B, seed):
    from car_racing_b200 import batch, scenarios
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N_H, M=M_OBS, seed=seed)
    prm = scenarios.default_cbf_params(N=N_H)
    rec, M, ps = batch.pack_cbf(x0, xt, obs, lap_off, N_H)
    return (x0, xt, obs, lap_off), prm, rec


def cpu_port_rate(B_sample, nthreads, seed=1):
    """The CPU restatement (oracle/ocp_oracle.c) on `nthreads` host threads; returns (solves/s, seconds, passes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from car_racing_b200 import scenarios
    (x0, xt, obs, lap_off), prm, _ = workload(B_sample, seed)
    orc.lib()
    t = time.perf_counter()
    passes = 0
    while True:       # bounded sample: whole passes over the batch until ~10 s of wall time (at most 16 passes)
        orc.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=nthreads)
        passes += 1
        dt = time.perf_counter() - t
        if dt >= 10.0 or passes >= 16:
            break
    return B_sample * passes / dt, dt, passes


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this metric.  CasADi/IPOPT cannot be installed
    here, so the arm times the oracle port on all host threads.  Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.batch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from car_racing_b200 import scenarios
    (x0, xt, obs, lap_off), prm, _ = workload(B, 1)
    orc.lib()
    for _ in range(args.warmup):
        orc.solve_cbf_batch(x0[:64], xt, obs[:64], lap_off[:64], prm, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=cores)
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "MPC-CBF N=20, 3 static rivals, l_shape, batch=%d random x0 per GPU (BASELINE config 2)" % B,
                       "note": "CasADi/IPOPT not installable offline; CPU port of the same NLP + interior point (oracle/ocp_oracle.c)"},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": "%d instances per step, %d pthreads" % (B, cores)},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=OUT, flush=True)


def _claim_stdout():
    """Rank 0 must print ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter) is
    sent to stderr; the returned file object is the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU per step")
    ap.add_argument("--inflight", type=int, default=6,
                    help="batches in flight (streams driven round-robin); 1 = strictly one batch at a time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange-priority", default="normal", choices=["high", "normal"],
                    help="N>1: stream priority of the exchange step (NCCL all-gather + argmin); measured on 2 B200: high is "
                         "slower (profiles/r03e_exchange_priority.json)")
    args = ap.parse_args()
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
    import car_racing_b200 as crb
    from car_racing_b200 import _capi, sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        # The exchange is 32 KB per rank and step, but its kernels share the SMs with up to D in-flight solver grids (7 resident
        # CTAs per SM).  Tried: exchange on high-priority streams (NCCL's and the argmin handle's) so that the block scheduler
        # places its CTAs first -- measured SLOWER on 2 B200 (696-703 k against 733-750 k solves/s, r03e), so normal is the default.
        pg_opts = None
        if args.exchange_priority == "high":
            pg_opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)
    B = args.batch
    D = max(1, args.inflight)
    _, prm, rec_host = workload(B, seed=1 + rank)            # each rank: its own shard of scenarios
    hs = [_capi.Handle(device=local_rank, max_batch=B) for _ in range(D)]   # one stream + staging buffers per slot
    L = _capi.lib()
    p = _capi.make_cbf_params(prm, M_OBS, False)
    o = _capi.default_options()
    exts = [torch.cuda.ExternalStream(h.stream, device=dev) for h in hs]
    d_in = torch.from_numpy(rec_host).to(dev)
    d_rec = [torch.zeros((B, 4), dtype=torch.float64, device=dev) for _ in range(D)]
    d_all = [torch.zeros((world * B, 4), dtype=torch.float64, device=dev) for _ in range(D)]
    d_arg = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(D)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    torch.cuda.synchronize()

    # N>1: the exchange step (all-gather of the 32-byte records + argmin) runs on its own stream.  NCCL completes collectives
    # in issue order, so an all-gather enqueued on a solver stream would hold that stream until every EARLIER batch's
    # stragglers have finished on every rank; here the solver stream only records an event and moves on to its next batch
    # (result buffers are double-buffered per slot).
    hc = _capi.Handle(device=local_rank, max_batch=B, high_priority=args.exchange_priority == "high") if world > 1 else None
    ext_c = torch.cuda.ExternalStream(hc.stream, device=dev) if world > 1 else None
    d_rec2 = [[d_rec[k], torch.zeros_like(d_rec[k])] for k in range(D)]
    ag_done = [[None, None] for _ in range(D)]
    n_issued = [0] * D

    def step_device(k):
        """One step = one batch of B instances: L2 flush, the solve kernel on slot k's stream; (N>1) all-gather of the
        32-byte records + argmin on the exchange stream."""
        h = hs[k]
        b = n_issued[k] & 1
        n_issued[k] += 1
        with torch.cuda.stream(exts[k]):
            flush.zero_()
            if ag_done[k][b] is not None:
                exts[k].wait_event(ag_done[k][b])          # the exchange that last read this result buffer
            rc = L.b200mpc_cbf_solve_device(h.ptr, C.byref(p), C.byref(o), B, d_in.data_ptr(), d_rec2[k][b].data_ptr(), None, None, None, None)
            h.check(rc, "b200mpc_cbf_solve_device")
            if world > 1:
                ev = torch.cuda.Event()
                ev.record(exts[k])
        if world > 1:
            with torch.cuda.stream(ext_c):
                ext_c.wait_event(ev)
                dist.all_gather_into_tensor(d_all[k], d_rec2[k][b])
                hc.check(L.b200mpc_argmin_cost_device(hc.ptr, d_all[k].data_ptr(), world * B, 0, d_arg[k].data_ptr()), "argmin")
                ag_done[k][b] = torch.cuda.Event()
                ag_done[k][b].record(ext_c)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def launches():
        return sum(h.launch_count for h in hs) + (hc.launch_count if hc is not None else 0)

    for i in range(max(args.warmup, D)):
        step_device(i % D)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- (a) one batch at a time: per-step CUDA events on the launching stream (the kernel's launch duration for the
    #      roofline, and the latency of one 1024-instance batch)
    evs = []
    for _ in range(args.steps):
        with torch.cuda.stream(exts[0]):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(exts[0])
            rc = L.b200mpc_cbf_solve_device(hs[0].ptr, C.byref(p), C.byref(o), B, d_in.data_ptr(), d_rec[0].data_ptr(), None, None, None, None)
            hs[0].check(rc, "b200mpc_cbf_solve_device")
            e1.record(exts[0])
            evs.append((e0, e1))
    barrier()
    ms_serial = [a.elapsed_time(b) for a, b in evs]
    kernel_ms = float(np.mean(ms_serial))

    # ---- (b) the timed region of the contract: exactly K steps, D batches in flight (slot = step mod D, each slot its
    #      own stream), L2 flush before every step, bracketed by barrier + synchronize; device time = first start event
    #      to the last end event over all streams
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
    if world > 1:
        end_c = torch.cuda.Event(enable_timing=True)
        end_c.record(ext_c)
        ends.append(end_c)
    barrier()
    wall = time.perf_counter() - wall0
    launches_dev = launches() - launches0
    t_dev = max(sa.elapsed_time(eb) for sa in starts for eb in ends) * 1e-3
    status = sharding.tensor_to_records(d_rec[0])["status"]
    conv = float((status == 0).mean())

    # ---- (c) end to end through the host-pointer C-ABI, the same D slots: every step copies its records from pinned
    #      host memory (H2D), solves, copies the 32-byte result records back (D2H) and the host reads them
    pin_in = [torch.from_numpy(rec_host).pin_memory() for _ in range(D)]
    pin_out = [torch.zeros((B, 4), dtype=torch.float64).pin_memory() for _ in range(D)]

    def e2e_run(nsteps):
        acc = 0.0
        for i in range(nsteps):
            k = i % D
            if i >= D:
                hs[k].synchronize()
                acc += float(pin_out[k][0, 0])                # the step's result is read on the host
            with torch.cuda.stream(exts[k]):
                flush.zero_()
            fn = L.b200mpc_cbf_solve_async if D > 1 else L.b200mpc_cbf_solve
            rc = fn(hs[k].ptr, C.byref(p), C.byref(o), B, pin_in[k].data_ptr(), pin_out[k].data_ptr(), None, None, None, None)
            hs[k].check(rc, "b200mpc_cbf_solve(_async)")
        for k in range(D):
            hs[k].synchronize()
            acc += float(pin_out[k][0, 0])
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
        L.b200mpc_cbf_solve(hs[0].ptr, C.byref(p), C.byref(o), 1, pin_in[0].data_ptr(), pin_out[0].data_ptr(), None, None, None, None)
        if k >= 20:
            lat.append(1e3 * (time.perf_counter() - t0))
    for k in range(8):
        t0 = time.perf_counter()
        L.b200mpc_cbf_solve(hs[0].ptr, C.byref(p), C.byref(o), B, pin_in[0].data_ptr(), pin_out[0].data_ptr(), None, None, None, None)
        if k >= 2:
            lat_b.append(1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop()
    launches_total = launches() - launches0

    # ---- max over ranks
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, wall = [float(v) for v in tt.tolist()]
        cv = torch.tensor([conv], dtype=torch.float64, device=dev)
        dist.all_reduce(cv, op=dist.ReduceOp.MIN)
        conv = float(cv.item())
    total = B * world * args.steps
    value = total / t_dev
    hbm_peak, peak_src = peaks()
    achieved = ALG_BYTES_PER_SOLVE * B / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "MPC-CBF N=20, 3 static rivals, l_shape, batch=%d random x0 per GPU (BASELINE config 2)" % B,
                   "global_batch": B * world, "parallelism": "dp%d (independent scenario shards + 1 all-gather of 32-B records and argmin per step, on an exchange stream%s)" % (world, ", high priority" if world > 1 and args.exchange_priority == "high" else ""),
                   "l2": "256 MiB buffer written before every timed step, on the step's stream (inputs are 1.1 MB << L2)",
                   "batches_in_flight": D,
                   "pipelining": ("steps are issued round-robin on %d streams (one handle each): the stragglers of one 1024-instance "
                                  "batch overlap the next batches; timed first start event -> last end event, flushes included" % D)
                   if D > 1 else "none: one batch at a time",
                   "solver": "FP64 barrier-SQP (IPOPT conventions), Riccati KKT, tol 1e-8", "converged_frac": conv},
        "e2e": {"value": total / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(rec_host.nbytes),
                "d2h_bytes_per_step": int(B * 32), "ms_per_step": 1e3 * t_e2e / args.steps,
                "batches_in_flight": D, "p50_latency_ms_batch1": float(np.median(lat)), "p10_p90_latency_ms_batch1": [float(np.percentile(lat, 10)), float(np.percentile(lat, 90))],
                "p50_latency_ms_one_batch": float(np.median(lat_b)),
                "note": ("per rank: host records -> H2D -> solve -> D2H -> host read; the cross-rank exchange step (all-gather + argmin) is part of "
                         "`value` only") if world > 1 else "host records -> H2D -> solve -> D2H -> host read"},
        "one_batch_at_a_time": {"value": B * world / (kernel_ms * 1e-3), "unit": "solves/s", "ms_per_step": kernel_ms,
                                "note": "the same kernel, steps serialised on one stream (per-step CUDA events); its launch duration is the roofline's"},
        "gpu_launches": int(launches_dev),
        "gpu_launches_incl_e2e": int(launches_total),
        "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "note": "algorithmic bytes = 1672 B/solve x %d; the kernel is FP64-latency/shared-memory bound, see profiles/ and DESIGN.md" % B},
    }
    # what actually bounds the kernel (ncu, crowded launch; measured once per kernel version, not live): see DESIGN.md section 5
    sp = os.path.join(ROOT, "profiles", "r01x_ncu_crowded_b8192_summary.json")
    if os.path.exists(sp):
        m = json.load(open(sp))["metrics"]
        g = lambda k: float(m[k]["value"]) if k in m else None
        line["roofline"]["ncu_crowded_launch"] = {
            "source": "profiles/r01x_ncu_crowded_b8192_summary.json",
            "issue_slots_busy_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_busy_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "icache_hit_pct": g("sm__icc_request_hit_rate.pct"),
            "bound": "dependent-issue latency (stall_wait 45 %), after the instruction-fetch bound was removed"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, dt, passes = cpu_port_rate(min(B, 1024), cores)
        line["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                                "sample": "the same %d scenarios, %d passes, %d pthreads (%.1f s)" % (min(B, 1024), passes, cores, dt)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
