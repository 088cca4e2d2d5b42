"""Multi-GPU plumbing: one process per GPU, contiguous block partition of the instance batch, one
all-gather of the 32-byte per-instance records (torch.distributed: NCCL over NVLink on GPUs, gloo in
the CPU tests), then the same first-min argmin on every rank
(overtake_traj_planner.py:244: `direction_flag = cost_selection.index(min(cost_selection))`).
The solve itself needs no collective: instances are independent."""
import numpy as np
import torch
import torch.distributed as dist

from ._capi import RECORD_DTYPE


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
