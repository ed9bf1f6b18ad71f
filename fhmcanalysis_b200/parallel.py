"""State-point sharding across the GPUs of one box (one process per GPU, torch.distributed).

The path has no exchange step: state points are independent, so each rank takes a contiguous slice of the
state-point range, runs the same kernels on it, and the only collective is the final gather of the packed result
records (NCCL over NVLink on GPUs; gloo in the CPU tests of this host-side logic)."""
import numpy as np


def shard_bounds(n_states, world, rank):
    """Contiguous, balanced slice [lo, hi) of range(n_states) owned by ``rank`` (sizes differ by at most 1)."""
    n_states, world, rank = int(n_states), int(world), int(rank)
    base, rem = divmod(n_states, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_states, world):
    return [shard_bounds(n_states, world, r)[1] - shard_bounds(n_states, world, r)[0] for r in range(world)]


FLOAT_FIELDS = ("lnnorm", "fe", "avg")
INT_FIELDS = ("status", "nphase", "nmin", "bounds", "max_idx", "min_idx")


def pack_records(rec):
    """dict of per-state-point tensors [S, ...] -> (float64 [S, F], int32 [S, I]) rows, one per state point."""
    import torch
    S = rec["lnnorm"].shape[0]
    f = torch.cat([rec[k].reshape(S, -1) for k in FLOAT_FIELDS if rec.get(k) is not None], dim=1).contiguous()
    i = torch.cat([rec[k].reshape(S, -1).to(torch.int32) for k in INT_FIELDS], dim=1).contiguous()
    return f, i


def unpack_records(f, i, pmax, n_sel):
    S = f.shape[0]
    out, c = {}, 0
    out["lnnorm"] = f[:, 0]
    out["fe"] = f[:, 1:1 + pmax]
    c = 1 + pmax
    out["avg"] = f[:, c:c + pmax * n_sel].reshape(S, pmax, n_sel) if n_sel else None
    widths = (("status", 1), ("nphase", 1), ("nmin", 1), ("bounds", 2 * pmax), ("max_idx", pmax), ("min_idx", pmax + 1))
    c = 0
    for k, w in widths:
        out[k] = i[:, c:c + w]
        c += w
    out["status"], out["nphase"], out["nmin"] = out["status"][:, 0], out["nphase"][:, 0], out["nmin"][:, 0]
    out["bounds"] = out["bounds"].reshape(S, pmax, 2)
    return out


def all_gather_records(f, i, n_states, group=None):
    """Gather the per-rank packed rows into the full [n_states, ...] arrays on every rank (one all_gather per
    dtype; shards are padded to the largest shard so all_gather_into_tensor can be used)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_states, world)
    smax = max(sizes)
    outs = []
    for t in (f, i):
        pad = torch.zeros((smax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        full = torch.empty((world * smax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, pad, group=group)
        outs.append(torch.cat([full[r * smax:r * smax + sizes[r]] for r in range(world)], dim=0))
    return outs[0], outs[1]


def sweep_sharded(make_device_hist, mu1, beta=None, dmu=None, pmax=4, lanes=0, gather=True, group=None):
    """Run a flat state-point list sharded over the ranks of ``group`` (call from every rank).

    make_device_hist: callable returning this rank's engine.DeviceHistogram (the blob is replicated by plain H2D).
    Returns the gathered dict of NumPy arrays on every rank (or this rank's shard when gather=False)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mu1 = np.asarray(mu1, dtype=np.float64)
    S = len(mu1)
    lo, hi = shard_bounds(S, world, rank)

    def cut(x):
        if x is None:
            return None
        x = np.asarray(x, dtype=np.float64)
        return x if x.size == 1 else x[lo:hi]

    dh = make_device_hist()
    res = dh.sweep(mu1[lo:hi], cut(beta), cut(dmu), pmax=pmax, lanes=lanes)
    rec = {k: getattr(res, k) for k in FLOAT_FIELDS + INT_FIELDS}
    f, i = pack_records(rec)
    if gather and world > 1:
        f, i = all_gather_records(f, i, S, group)
    out = unpack_records(f, i, pmax, dh.n_sel)
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
