"""Worker of tests/test_exchange_gpu.py: one process per rank (several ranks may share one GPU -- CUDA IPC maps a window of
another process on the same device as well), gloo only for the hand-over of the window handles and for the checks.

    RANK=r WORLD_SIZE=w MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tests/helpers/exchange_worker.py [--free-running]

Every use: publish_next + the real MPC-CBF / iLQR solve of this rank's shard, then the window's argmin kernel; the gathered
records must equal what an all_gather of the host copies gives, on every rank, and the argmin must be numpy's first-min.
Without --free-running a barrier separates solve and argmin (ranks sharing one GPU time-slice it: a kernel spinning on a
counter would otherwise wait out its time slice); with it the ranks run unsynchronised for many uses of few slots, which is
what exercises the arrival counters, the acknowledgements and the flow control (needs one GPU per rank)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from car_racing_b200 import _capi, batch, scenarios, sharding   # noqa: E402


def main():
    free = "--free-running" in sys.argv
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    L = _capi.lib()
    B, slots, uses = 96, 2, (40 if free else 6)
    h, hc = _capi.Handle(device=dev, max_batch=B), _capi.Handle(device=dev, max_batch=B)
    px = sharding.PeerExchange(hc, rank, world, max_batch=B, slots=slots)
    prm = scenarios.default_cbf_params(N=20)
    x0, xt, obs, lo = scenarios.mpccbf_scenarios(B, N=20, M=3, seed=100 + rank)
    p, o = _capi.make_cbf_params(prm, 3, False), _capi.default_options()
    ip = scenarios.default_cbf_params()
    iprm = dict(A=ip["A"], B=ip["B"], Q=ip["Q"], R=ip["R"], N=50, max_iter=150, L=0.4, W=0.2)
    ix0, ixt, iobs, ilo = scenarios.ilqr_scenarios(B, N=50, seed=200 + rank)
    irec_in = batch.pack_ilqr(ix0, ixt, iobs, ilo, 50)
    iparams = _capi.make_ilqr_params(iprm)
    d_arg = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_all = torch.zeros((world * B, 4), dtype=torch.float64, device="cuda")
    P = batch._ptr
    bad = 0
    pending = []
    for use in range(uses):
        slot = use % slots
        rec = np.zeros(B, dtype=_capi.RECORD_DTYPE)
        px.publish_next(h, slot)
        if use % 3 == 2:          # the iLQR kernel carries the same epilogue
            h.check(L.b200mpc_ilqr_solve(h.ptr, C.byref(iparams), B, P(irec_in), P(rec), None, None), "ilqr")
        else:
            rin, M, ps = batch.pack_cbf(x0 + 0.003 * use, xt, obs, lo, 20)
            h.check(L.b200mpc_cbf_solve(h.ptr, C.byref(p), C.byref(o), B, P(rin), P(rec), None, None, None, None), "cbf")
        if not free:
            dist.barrier()
        px.argmin(hc, slot, d_arg.data_ptr(), d_all.data_ptr())
        hc.synchronize()
        got = sharding.tensor_to_records(d_all).copy()
        arg = int(d_arg.item())
        pending.append((rec, got, arg))
    for use, (rec, got, arg) in enumerate(pending):       # the checks use collectives: after the free-running part
        parts = [None] * world
        dist.all_gather_object(parts, rec.tobytes())
        want = np.frombuffer(b"".join(parts), dtype=_capi.RECORD_DTYPE)
        if got.tobytes() != want.tobytes() or arg != sharding.argmin_first(want):
            bad += 1
            print(f"[rank {rank}] use {use}: gathered records or argmin differ (arg {arg} vs {sharding.argmin_first(want)})", flush=True)
    t = torch.tensor([bad])
    dist.all_reduce(t)
    dist.barrier()
    px.close()
    if rank == 0:
        print("EXCHANGE_OK" if int(t.item()) == 0 else "EXCHANGE_FAILED", "ranks", world, "gpus", ngpu, "uses", uses, "free" if free else "stepped", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
