"""Multi-GPU plumbing: one process per GPU, contiguous block partition of the instance batch, one exchange of the
32-byte per-instance records, then the same first-min argmin on every rank
(overtake_traj_planner.py:244: `direction_flag = cost_selection.index(min(cost_selection))`).
The solve itself needs no collective: instances are independent.

Two forms of the exchange:
  * PeerExchange -- the product path on GPUs: the library's own window (b200mpc_comm_*, csrc/exchange.cuh); the solver
    kernels' epilogue stores each record into every rank's gathered buffer over NVLink, no collective kernel.
    torch.distributed is only used once, to hand the 64-byte window handles around.
  * all_gather_records -- torch.distributed.all_gather_into_tensor (NCCL on GPUs, gloo in the CPU tests) for callers
    that hold their records in torch tensors."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._capi import RECORD_DTYPE


class PeerExchange:
    """The exchange window of this rank (include/b200mpc.h: b200mpc_comm_*).

        px = PeerExchange(handle, rank, world, max_batch=B, slots=D)      # collective: every rank, same arguments
        px.publish_next(handle_k, slot)                                   # the next solve on handle_k publishes into `slot`
        L.b200mpc_cbf_solve_device(handle_k.ptr, ...)                     #   (every rank solves the same B in that step)
        px.argmin(handle_c, slot, d_arg_ptr, d_all_ptr)                   # on handle_c's stream: wait, argmin, copy, release

    `allgather` hands the 64-byte handles around: a callable bytes -> list of `world` bytes objects; default
    torch.distributed.all_gather_object on the default group."""

    def __init__(self, handle, rank, world, max_batch, slots, allgather=None):
        self.rank, self.world, self.max_batch, self.slots = int(rank), int(world), int(max_batch), int(slots)
        L = _capi.lib()
        c = C.c_void_p()
        handle.check(L.b200mpc_comm_create(handle.ptr, self.rank, self.world, self.max_batch, self.slots, C.byref(c)),
                     "b200mpc_comm_create")
        self._c = c
        mine = (C.c_char * _capi.COMM_HANDLE_BYTES)()
        if L.b200mpc_comm_export(self._c, mine) != 0:
            raise _capi.B200MPCError("b200mpc_comm_export failed")
        if allgather is None:
            def allgather(b):
                out = [None] * self.world
                dist.all_gather_object(out, b)
                return out
        blobs = allgather(bytes(mine.raw)) if self.world > 1 else [bytes(mine.raw)]
        if len(blobs) != self.world or any(len(b) != _capi.COMM_HANDLE_BYTES for b in blobs):
            raise ValueError("allgather must return `world` handles of %d bytes" % _capi.COMM_HANDLE_BYTES)
        rc = L.b200mpc_comm_connect(self._c, b"".join(blobs))
        if rc != 0:
            msg = L.b200mpc_last_error(None).decode()
            self.close()
            raise _capi.B200MPCError("b200mpc_comm_connect failed (%d): %s" % (rc, msg))

    def publish_next(self, handle, slot):
        handle.check(_capi.lib().b200mpc_comm_publish_next(handle.ptr, self._c, int(slot)), "b200mpc_comm_publish_next")

    def argmin(self, handle, slot, d_out, d_all=None, max_status=0):
        """d_out: device pointer of one int32; d_all: optional device pointer of world * B records."""
        handle.check(_capi.lib().b200mpc_comm_argmin(handle.ptr, self._c, int(slot), int(max_status), d_out, d_all),
                     "b200mpc_comm_argmin")

    def close(self):
        c, self._c = getattr(self, "_c", None), None
        if c:
            _capi.lib().b200mpc_comm_destroy(c)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_range(B, rank, world):
    """Rank r owns instances [lo, hi): contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def records_to_tensor(rec, device=None):
    """numpy record array -> (n, 4) float64 tensor view (cost, u0[0], u0[1], packed status/iters)."""
    t = torch.from_numpy(np.ascontiguousarray(rec).view(np.float64).reshape(-1, 4))
    return t.to(device) if device is not None else t


def tensor_to_records(t):
    return t.detach().cpu().contiguous().numpy().reshape(-1).view(RECORD_DTYPE)


def all_gather_records(local, B, group=None):
    """local: (n_local, 4) float64 tensor of this rank's records.  Returns the (B, 4) tensor of all
    records in instance order on every rank.  Uneven shards are padded to the largest one."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    nmax = max(sizes)
    buf = local
    if local.shape[0] < nmax:
        buf = torch.zeros((nmax, 4), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    out = torch.empty((world * nmax, 4), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if all(s == nmax for s in sizes):
        return out
    return torch.cat([out[r * nmax: r * nmax + sizes[r]] for r in range(world)], dim=0)


def argmin_first(records, max_status=0):
    """Index of the lowest cost among instances with status <= max_status; lowest index wins ties; -1 if none."""
    rec = tensor_to_records(records) if torch.is_tensor(records) else records
    ok = rec["status"] <= max_status
    if not ok.any():
        return -1
    cost = np.where(ok, rec["cost"], np.inf)
    return int(np.argmin(cost))      # numpy returns the first minimum
