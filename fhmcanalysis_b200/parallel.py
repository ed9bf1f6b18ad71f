"""State-point sharding across the GPUs of one box (one process per GPU, torch.distributed).

The path has no exchange step: state points are independent, so each rank takes a contiguous slice of the state-point
range, runs the same kernels on it, and the only communication is the final gather of the result records.  Two gathers:

* ``sweep_sharded_compact`` (pure mu sweeps, the headline metric): the gather is FUSED INTO THE SWEEP KERNEL.  Every rank
  allocates the gathered buffer in symmetric memory (torch.distributed._symmetric_memory: the buffers of all ranks are
  mapped into each other's address space over NVLink / NVSwitch) and ``k_sweep_prod2<compact>`` writes each finished
  record -- 60 bytes for a two-phase state point with two averaged quantities -- straight into block ``rank`` of EVERY
  rank's buffer with plain stores while it is still walking the next state points.  No collective follows, only a barrier.
  Where symmetric memory is not available the records are written locally and gathered with ONE NCCL
  ``all_gather_into_tensor`` of the compact blocks (gloo in the CPU tests of this host-side logic).
* ``sweep_sharded`` / ``find_phase_eq_sharded`` / ``reweight_2d_sharded`` / ``sweep_grid_sharded``: full records (or the
  solver's / the 2-D kernel's outputs) gathered with padded ``all_gather_into_tensor`` calls, one per dtype.  The two sweeps
  first agree (one MAX all-reduce of two integers) on the largest phase / minima count any record has and gather only that
  many per-phase columns: the capacity pmax = 8 of a Taylor grid holds two or three phases, 96-120 instead of 216 bytes per
  state point on the wire.

Layout of a gathered compact result: ``world`` blocks of ``fhmc_pack_soa16_bytes(smax, pmax, n_sel)`` bytes, block r = the
narrow phase-major records of rank r's shard (``smax`` = largest shard).  ``ShardedRecords.host()`` concatenates them.
"""
import numpy as np


def shard_bounds(n_states, world, rank):
    """Contiguous, balanced slice [lo, hi) of range(n_states) owned by ``rank`` (sizes differ by at most 1)."""
    n_states, world, rank = int(n_states), int(world), int(rank)
    base, rem = divmod(n_states, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_states, world):
    return [shard_bounds(n_states, world, r)[1] - shard_bounds(n_states, world, r)[0] for r in range(world)]


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


FLOAT_FIELDS = ("lnnorm", "fe", "avg")
INT_FIELDS = ("status", "nphase", "nmin", "bounds", "max_idx", "min_idx")


def live_widths(nphase, nmin, pmax, group=None):
    """(P, M): the largest phase count and the largest minima count over ALL ranks' records, clamped to the capacities
    pmax / pmax + 1 -- the widths worth gathering (one MAX all-reduce of two integers; synchronises).  ``nphase`` / ``nmin``
    are this rank's tensors (possibly empty)."""
    import torch
    import torch.distributed as dist
    if nphase.numel():
        w = torch.stack([nphase.max(), nmin.max()]).to(torch.int64)
    else:
        w = torch.zeros(2, dtype=torch.int64, device=nphase.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
    P, M = (int(x) for x in w.tolist())
    return max(1, min(int(pmax), P)), max(1, min(int(pmax) + 1, M))


def pack_records(rec, widths=None):
    """dict of per-state-point tensors [S, ...] -> (float64 [S, F], int32 [S, I]) rows, one per state point.
    widths = (P, M) from live_widths(): only the first P phase slots (fe, avg, bounds, max_idx) and the first M entries of
    min_idx are packed -- with pmax = 8 and two-phase records that is 96 instead of 216 bytes per state point on the wire."""
    import torch
    S = rec["lnnorm"].shape[0]
    if widths is not None:
        P, M = widths
        rec = dict(rec)
        for k in ("fe", "avg", "bounds", "max_idx"):
            if rec.get(k) is not None:
                rec[k] = rec[k][:, :P]
        rec["min_idx"] = rec["min_idx"][:, :M]

    def rows(x):   # (S may be 0: an explicit width instead of -1)
        return x.reshape(S, int(np.prod(x.shape[1:])))
    f = torch.cat([rows(rec[k]) for k in FLOAT_FIELDS if rec.get(k) is not None], dim=1).contiguous()
    i = torch.cat([rows(rec[k]).to(torch.int32) for k in INT_FIELDS], dim=1).contiguous()
    return f, i


def unpack_records(f, i, pmax, n_sel, widths=None):
    """Inverse of pack_records.  With ``widths`` the per-phase arrays come back with P (min_idx: M) columns."""
    S = f.shape[0]
    mcols = pmax + 1
    if widths is not None:
        pmax, mcols = widths
    out, c = {}, 0
    out["lnnorm"] = f[:, 0]
    out["fe"] = f[:, 1:1 + pmax]
    c = 1 + pmax
    out["avg"] = f[:, c:c + pmax * n_sel].reshape(S, pmax, n_sel) if n_sel else None
    cols = (("status", 1), ("nphase", 1), ("nmin", 1), ("bounds", 2 * pmax), ("max_idx", pmax), ("min_idx", mcols))
    c = 0
    for k, w in cols:
        out[k] = i[:, c:c + w]
        c += w
    out["status"], out["nphase"], out["nmin"] = out["status"][:, 0], out["nphase"][:, 0], out["nmin"][:, 0]
    out["bounds"] = out["bounds"].reshape(S, pmax, 2)
    return out


def all_gather_rows(t, n_states, group=None):
    """Gather per-rank row blocks [S_r, ...] into the full [n_states, ...] tensor on every rank (shards are padded to the
    largest shard so that one all_gather_into_tensor does it; an EMPTY shard still joins the collective)."""
    import torch
    import torch.distributed as dist
    world, _ = _world(group)
    if world == 1:
        return t
    sizes = shard_sizes(n_states, world)
    smax = max(max(sizes), 1)
    if min(sizes) == smax:   # equal shards: the collective writes the final tensor, no padding and no concatenation
        full = torch.empty((world * smax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t.contiguous(), group=group)
        return full
    pad = torch.zeros((smax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    full = torch.empty((world * smax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    return torch.cat([full[r * smax:r * smax + sizes[r]] for r in range(world)], dim=0)


def all_gather_records(f, i, n_states, group=None):
    return all_gather_rows(f, n_states, group), all_gather_rows(i, n_states, group)


def _empty_records(pmax, n_sel, device):
    """Zero-row record tensors of a rank whose shard is empty (it still joins the collectives)."""
    import torch
    z = lambda *shape, dt=torch.float64: torch.zeros(shape, dtype=dt, device=device)
    i32 = torch.int32
    return {"lnnorm": z(0), "fe": z(0, pmax), "avg": z(0, pmax, n_sel) if n_sel else None, "status": z(0, dt=i32), "nphase": z(0, dt=i32),
            "nmin": z(0, dt=i32), "bounds": z(0, pmax, 2, dt=i32), "max_idx": z(0, pmax, dt=i32), "min_idx": z(0, pmax + 1, dt=i32)}


def _cut(x, lo, hi):
    if x is None:
        return None
    x = np.asarray(x, dtype=np.float64)
    return x if x.size == 1 else x[lo:hi]


def _deliver(out, to_host, pmax=None):
    """Gathered records as NumPy arrays (to_host; per-phase arrays padded back to the capacity ``pmax`` with NaN / -1 where
    only the live widths were gathered) or as the device tensors the collective left behind (``out['widths']`` = (P, M))."""
    if not to_host:
        return out
    res = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items() if k != "widths"}
    if pmax is not None and out.get("widths") is not None:
        def pad(a, n, fill):
            if a is None or a.shape[1] >= n:
                return a
            full = np.full((a.shape[0], n) + a.shape[2:], fill, dtype=a.dtype)
            full[:, :a.shape[1]] = a
            return full
        for k, n, fill in (("fe", pmax, np.nan), ("avg", pmax, np.nan), ("bounds", pmax, -1), ("max_idx", pmax, -1), ("min_idx", pmax + 1, -1)):
            res[k] = pad(res.get(k), n, fill)
    return res


def sweep_sharded(make_device_hist, mu1, beta=None, dmu=None, pmax=4, lanes=0, gather=True, group=None, to_host=True):
    """Run a flat state-point list sharded over the ranks of ``group`` (call from every rank); FULL records.

    make_device_hist: callable returning this rank's engine.DeviceHistogram (the blob is replicated by plain H2D).
    Returns the gathered dict of NumPy arrays on every rank (or this rank's shard when gather=False)."""
    import torch
    world, rank = _world(group)
    mu1 = np.atleast_1d(np.asarray(mu1, dtype=np.float64))
    S = len(mu1)
    lo, hi = shard_bounds(S, world, rank)
    dh = make_device_hist()
    widths = None
    if hi > lo:
        res = dh.sweep(mu1[lo:hi], _cut(beta, lo, hi), _cut(dmu, lo, hi), pmax=pmax, lanes=lanes)
        rec = {k: getattr(res, k) for k in FLOAT_FIELDS + INT_FIELDS}
    else:   # more ranks than state points: nothing to compute here, but the collectives below need every rank
        rec = _empty_records(pmax, dh.n_sel, dh.device)
    if gather and world > 1:
        widths = live_widths(rec["nphase"], rec["nmin"], pmax, group)     # only the phase slots that exist go on the wire
    f, i = pack_records(rec, widths)
    if gather and world > 1:
        f, i = all_gather_records(f, i, S, group)
    out = unpack_records(f, i, pmax, dh.n_sel, widths)
    out["widths"] = widths
    return _deliver(out, to_host, pmax)


def sweep_grid_sharded(make_device_hist, mu1, betas, dmus, pmax=4, group=None, to_host=True):
    """(beta x dmu) Taylor grid (temp_dmu_extrap_multi order: dmu fastest) at fixed mu1, the BETA rows sharded over the ranks
    (BASELINE config 3).  Returns the gathered records as NumPy arrays of nb*nd state points on every rank."""
    import torch
    world, rank = _world(group)
    betas = np.atleast_1d(np.asarray(betas, dtype=np.float64))
    dmus = np.atleast_1d(np.asarray(dmus, dtype=np.float64))
    nb, nd = len(betas), len(dmus)
    lo, hi = shard_bounds(nb, world, rank)
    dh = make_device_hist()
    widths = None
    if hi > lo:
        res = dh.sweep(np.atleast_1d(mu1), betas[lo:hi], dmus, grid=True, pmax=pmax)
        rec = {k: getattr(res, k) for k in FLOAT_FIELDS + INT_FIELDS}
    else:
        rec = _empty_records(pmax, dh.n_sel, dh.device)
    if world > 1:
        # only the phase slots that exist anywhere on the grid go on the wire (pmax = 8, two or three phases: 96-120 of 216 B)
        widths = live_widths(rec["nphase"], rec["nmin"], pmax, group)
    f, i = pack_records(rec, widths)
    if world > 1:
        # shards are whole beta rows: gather row blocks of nd state points each
        n_f, n_i = f.shape[1], i.shape[1]
        f = all_gather_rows(f.reshape(-1, nd, n_f), nb, group).reshape(-1, n_f)
        i = all_gather_rows(i.reshape(-1, nd, n_i), nb, group).reshape(-1, n_i)
    out = unpack_records(f, i, pmax, dh.n_sel, widths)
    out["widths"] = widths
    return _deliver(out, to_host, pmax)


def find_phase_eq_sharded(make_device_hist, mu_guess, betas, dmu=None, lnz_tol=1e-10, pmax=4, group=None, to_host=True, **kw):
    """Batched coexistence solves (K4, BASELINE config 4) with the temperatures sharded over the ranks; every rank returns
    the gathered dict: mu_coex, dfe, iters, status/code/converged, nphase, fe, avg, bounds."""
    import torch
    from . import _lib
    world, rank = _world(group)
    betas = np.atleast_1d(np.asarray(betas, dtype=np.float64))
    T = len(betas)
    lo, hi = shard_bounds(T, world, rank)
    dh = make_device_hist()
    g = np.broadcast_to(np.atleast_1d(np.asarray(mu_guess, dtype=np.float64)), betas.shape)
    n_f, n_i = 2 + 1 + pmax + pmax * dh.n_sel, 1 + 3 + 2 * pmax + pmax + pmax + 1
    if hi > lo:
        res = dh.find_phase_eq(g[lo:hi].copy(), beta=betas[lo:hi], dmu=_cut(dmu, lo, hi), lnz_tol=lnz_tol, pmax=pmax, **kw)
        f, i = pack_records({k: getattr(res, k) for k in FLOAT_FIELDS + INT_FIELDS})
        f = torch.cat([res.extra["mu_coex"][:, None], res.extra["dfe"][:, None], f], dim=1).contiguous()
        i = torch.cat([res.extra["iters"][:, None].to(torch.int32), i], dim=1).contiguous()
    else:
        f = torch.zeros((0, n_f), dtype=torch.float64, device=dh.device)
        i = torch.zeros((0, n_i), dtype=torch.int32, device=dh.device)
    f, i = all_gather_records(f, i, T, group)
    out = unpack_records(f[:, 2:], i[:, 1:], max(pmax, 2), dh.n_sel)
    out["mu_coex"], out["dfe"], out["iters"] = f[:, 0], f[:, 1], i[:, 0]
    if not to_host:
        return out
    out = _deliver(out, True)
    out["status"] = out["status"].view(np.uint32)
    out["code"] = (out["status"] & _lib.ST_CODE_MASK).astype(np.int32)
    out["converged"] = (out["code"] == 0) & ((out["status"] & _lib.ST_JUMP) == 0)
    return out


def reweight_2d_sharded(lnpi, bounds, op1, op2, a1, a2, props=None, device=None, group=None, product=None, to_host=True):
    """K5 (BASELINE config 5): the (a1, a2) state points of a 2-D joint-histogram reweight sharded over the ranks, the
    histogram replicated; returns the gathered [S, 3 + n_prop] array (lnZ, <op1>, <op2>, <prop>...) on every rank."""
    import torch
    from . import engine
    world, rank = _world(group)
    if not isinstance(a1, torch.Tensor):   # (device tensors are sliced as they are: inputs already resident in HBM)
        a1 = np.atleast_1d(np.asarray(a1, dtype=np.float64))
        a2 = np.atleast_1d(np.asarray(a2, dtype=np.float64))
    S = int(a1.shape[0])
    lo, hi = shard_bounds(S, world, rank)
    n_prop = 0 if props is None else len(props)
    dev = engine.require_cuda(device)
    if hi > lo:
        out = engine.reweight_2d(lnpi, bounds, op1, op2, a1[lo:hi], a2[lo:hi], props, device=dev, return_device=True, product=product)
    else:
        out = torch.zeros((0, 3 + n_prop), dtype=torch.float64, device=dev)
    full = all_gather_rows(out, S, group)
    return full.cpu().numpy() if to_host else full


# ----------------------------------------------------------------------------------------------------------------------
# compact records, gather fused into the sweep kernel
# ----------------------------------------------------------------------------------------------------------------------
class ShardedRecords(object):
    """Gathered compact records of a sharded mu sweep: ``buf`` = world blocks of narrow phase-major records (see module
    docstring), on the device of this rank."""

    def __init__(self, buf, n_states, world, pmax, n_sel, block_bytes, fused):
        self.buf, self.n_states, self.world, self.pmax, self.n_sel = buf, int(n_states), int(world), int(pmax), int(n_sel)
        self.block_bytes, self.fused = int(block_bytes), bool(fused)
        self.sizes = shard_sizes(self.n_states, self.world)
        self.smax = max(max(self.sizes), 1)
        self.gathered = True     # False: only this rank's block is filled (sweep_sharded_compact(gather=False))

    def views(self, rank):
        """[smax, ...] views (status, nphase, fe, avg, bounds) of rank ``rank``'s block; rows >= sizes[rank] are padding."""
        from .engine import soa16_views
        return soa16_views(self.buf[rank * self.block_bytes:(rank + 1) * self.block_bytes], self.smax, self.pmax, self.n_sel)

    def host(self):
        """dict of NumPy arrays over all n_states state points: status (uint16), code, safe, nphase, fe, avg, bounds.
        Phase slots p >= nphase[s] are not transported (the kernels only store what exists); they read NaN / -1 here."""
        import torch
        hb = self.buf.cpu() if self.buf.is_cuda else self.buf
        parts = []
        from .engine import soa16_views
        for r in range(self.world):
            v = soa16_views(hb[r * self.block_bytes:(r + 1) * self.block_bytes], self.smax, self.pmax, self.n_sel)
            parts.append({k: (x[:self.sizes[r]] if x is not None else None) for k, x in v.items()})
        out = {}
        for k in parts[0]:
            out[k] = None if parts[0][k] is None else torch.cat([p[k] for p in parts], dim=0).numpy()
        out["status"] = out["status"].view(np.uint16)
        out["code"] = (out["status"] & 0xFF).astype(np.int32)
        out["safe"] = (out["status"] & 0x100) != 0
        live = np.arange(self.pmax)[None, :] < np.where(out["code"] == 0, out["nphase"].astype(np.int64), 0)[:, None]
        out["fe"] = np.where(live, out["fe"], np.nan)
        if out["avg"] is not None:
            out["avg"] = np.where(live[:, :, None], out["avg"], np.nan)
        out["bounds"] = np.where(live[:, :, None], out["bounds"], -1).astype(np.int16)
        return out


class CompactGather(object):
    """Reusable state of sweep_sharded_compact for one (n_states, pmax, n_sel): the gathered buffer (symmetric memory when
    the NVLink-fused path is available) and, for that path, the peers' block pointers."""

    def __init__(self, dh, n_states, pmax=4, group=None, fused=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.world, self.rank = _world(group)
        self.group, self.pmax, self.n_sel, self.n_states = group, int(pmax), dh.n_sel, int(n_states)
        self.sizes = shard_sizes(self.n_states, self.world)
        self.smax = max(max(self.sizes), 1)
        L = _lib.load()
        self.block_bytes = (int(L.fhmc_pack_soa16_bytes(self.smax, self.pmax, self.n_sel)) + 255) & ~255
        self.fused, self.handle, self.why = False, None, ""
        nbytes = self.world * self.block_bytes
        dev = dh.device
        want = (self.world > 1 and dev.type == "cuda") if fused is None else bool(fused)
        if want and self.world > 1 and self.world <= 8:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
                self.handle = symm_mem.rendezvous(buf, grp)
                self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
                self.buf = buf
                self.fused = True
            except Exception as e:   # no symmetric memory on this system / build: NCCL gather of the compact blocks instead
                self.why = repr(e)
                if fused:
                    raise
        if not self.fused:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.local = torch.empty(self.block_bytes, dtype=torch.uint8, device=dev)

    def barrier(self):
        """Device-side barrier over the ranks on the current stream (fused path) or a collective barrier."""
        import torch.distributed as dist
        if self.fused:
            self.handle.barrier(channel=0)
        elif self.world > 1:
            dist.barrier(group=self.group)


def sweep_sharded_compact(dh, mu1, pmax=4, group=None, state=None, fused=None, pre_barrier=True, gather=True):
    """Pure mu sweep sharded over the ranks with the gather FUSED into the sweep kernel (see module docstring).  Call from
    every rank with the same ``mu1``; asynchronous on the current stream.  Returns (ShardedRecords, state) -- pass
    ``state`` back in to reuse the buffers (and the symmetric-memory rendezvous) for the next sweep of the same size.

    pre_barrier: ranks may still be reading the buffer of the previous sweep; a barrier first makes overwriting safe.
    gather=False: the records STAY SHARDED -- every rank fills only its own block of its buffer (``ShardedRecords.views(rank)``),
    no peer stores, no barrier, no collective: the state points are independent, so nothing on the data path needs an
    exchange; the gather is a convenience for callers that want every record everywhere."""
    import torch
    import torch.distributed as dist
    mu1 = np.atleast_1d(np.asarray(mu1, dtype=np.float64)) if not isinstance(mu1, torch.Tensor) else mu1
    S = int(mu1.shape[0])
    if state is None or state.n_states != S or state.pmax != pmax or state.n_sel != dh.n_sel:
        state = CompactGather(dh, S, pmax, group, fused)
    world, rank = state.world, state.rank
    lo, hi = shard_bounds(S, world, rank)
    shard = mu1[lo:hi]
    if not gather:
        if hi > lo:
            dh.sweep_compact(shard, pmax=pmax, dst=state.buf[rank * state.block_bytes:(rank + 1) * state.block_bytes], n_total=state.smax,
                             first=0, fill_dead=False)
        rec = ShardedRecords(state.buf, S, world, pmax, dh.n_sel, state.block_bytes, state.fused)
        rec.gathered = False
        return rec, state
    if state.fused:
        if pre_barrier:
            state.barrier()
        if hi > lo:
            off = rank * state.block_bytes
            # (dead phase slots are NOT written: they would double the NVLink traffic; host() masks them)
            dh.sweep_compact(shard, pmax=pmax, dst=[p + off for p in state.peer_ptrs], n_total=state.smax, first=0, fill_dead=False)
        state.barrier()   # every rank's stores have landed in every buffer once all kernels are done
    elif world == 1:
        dh.sweep_compact(shard, pmax=pmax, dst=state.buf, n_total=state.smax, first=0, fill_dead=False)
    else:
        if hi > lo:
            dh.sweep_compact(shard, pmax=pmax, dst=state.local, n_total=state.smax, first=0, fill_dead=False)
        dist.all_gather_into_tensor(state.buf, state.local, group=group)
    return ShardedRecords(state.buf, S, world, pmax, dh.n_sel, state.block_bytes, state.fused), state
