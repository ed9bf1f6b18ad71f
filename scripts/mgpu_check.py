"""Multi-GPU check (run under torchrun, one rank per GPU): the mu sweep sharded over the ranks with the gather fused into the
sweep kernel (symmetric memory over NVLink) against the NCCL gather of the same compact blocks and against an unsharded
sweep; timings of both; the sharded solver / 2-D / grid entry points on small cases."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhmcanalysis_b200 import engine, parallel, synth  # noqa: E402
from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def log(*a):
    if rank == 0:
        print(*a, flush=True)


n = 1001
lnpi, mom = synth.two_peak_lnpi(n), synth.one_comp_moments(n)
hist = histogram.from_arrays(lnpi, mom, 1.0, [0.0], 10)
dh = hist.device_histogram(moments=("N", "N2"), device=dev)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
S = per * world
mu = np.linspace(-0.03, 0.03, S)
mu_dev = torch.from_numpy(mu).to(dev)

ref = None
if rank == 0:   # unsharded reference on one GPU (strided sample)
    sub = np.arange(0, S, 997)
    full = dh.sweep(mu_dev, pmax=4).host()       # plain records of the table-free product-form kernel: same integers, fe / avg to 1e-11
    ref = {k: (v[sub] if v is not None else None) for k, v in full.items()}
    del full

results = {}
for name, fused in (("fused NVLink stores", True), ("NCCL all_gather of compact blocks", False)):
    try:
        rec, state = parallel.sweep_sharded_compact(dh, mu_dev, pmax=4, fused=fused)
    except Exception as e:
        log("%s: unavailable: %r" % (name, e))
        continue
    torch.cuda.synchronize()
    h = rec.host()
    if rank == 0:
        P = ref["nphase"]
        assert np.array_equal(h["nphase"][sub], P) and np.array_equal(h["code"][sub], ref["code"]) and np.array_equal(h["safe"][sub], ref["safe"])
        live = np.arange(4)[None, :] < P[:, None]
        assert np.allclose(h["fe"][sub][live], ref["fe"][live], rtol=1e-11, atol=1e-12) and np.allclose(h["avg"][sub][live], ref["avg"][live], rtol=1e-11, atol=0)
        assert np.array_equal(h["bounds"][sub][live], ref["bounds"][live])
    # every rank holds the same gathered bytes for the live slots
    chk = torch.tensor([float(np.nansum(h["fe"])), float(h["nphase"].sum())], dtype=torch.float64, device=dev)
    if world > 1:
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        assert torch.equal(lo_, hi_), (lo_, hi_)
    # timing: barrier + sweep (+ gather) + barrier, CUDA events, max over ranks
    times = []
    for rep in range(12):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rec, state = parallel.sweep_sharded_compact(dh, mu_dev, pmax=4, state=state, fused=fused)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    ms = float(np.median(times[2:]))
    results[name] = ms
    log("%-36s fused=%s (%s): %.3f ms per %d-point sweep on %d GPUs = %.3e state points/s incl. gather" %
        (name, state.fused, state.why[:60], ms, S, world, S / ms * 1e3))

# records left sharded (gather=False): this rank's block equals its block of the gathered buffer
if results:
    g_rec, state = parallel.sweep_sharded_compact(dh, mu_dev, pmax=4, state=state)
    torch.cuda.synchronize()
    mine = {k: (v.clone() if v is not None else None) for k, v in g_rec.views(rank).items()}
    l_rec, state = parallel.sweep_sharded_compact(dh, mu_dev, pmax=4, state=state, gather=False)
    torch.cuda.synchronize()
    nloc = l_rec.sizes[rank]
    for k, v in l_rec.views(rank).items():
        if v is not None and k in ("status", "nphase"):
            assert torch.equal(v[:nloc], mine[k][:nloc]), k
    log("gather=False: this rank's block identical to its block of the gathered buffer (status, nphase)")

# compute only (local compact records, no gather), for the efficiency figure
lo, hi = parallel.shard_bounds(S, world, rank)
times = []
for rep in range(12):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dh.sweep_compact(mu_dev[lo:hi], pmax=4)
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
log("compute only (compact records, local): %.3f ms = %.3e state points/s" % (np.median(times[2:]), S / np.median(times[2:]) * 1e3))

# ---- the other sharded entry points, small cases against unsharded calls ---------------------------------------------
# config 4 (solver)
h4 = histogram.from_arrays(synth.two_peak_lnpi(801, scale=0.8), synth.one_comp_moments(801, max_order=3), 1.0, [0.0], 10)
betas = 1.0 / np.linspace(0.92, 1.03, 403)
mk4 = lambda: h4.device_histogram(beta=betas, order=2, moments=("N", "N2", "U"), device=dev)
r4 = parallel.find_phase_eq_sharded(mk4, 0.0, betas, lnz_tol=1e-10, continuation=False)
if rank == 0:
    one = mk4().find_phase_eq(np.zeros_like(betas), beta=betas, lnz_tol=1e-10, continuation=False).host()
    assert np.array_equal(one["code"], r4["code"]) and np.array_equal(one["mu_coex"], r4["mu_coex"]) and np.array_equal(one["nphase"], r4["nphase"])
    log("find_phase_eq_sharded ok: %d solves, converged %.3f" % (len(betas), r4["converged"].mean()))
# config 5 (2-D)
l2, b2 = synth.joint_2d(128, 96, cut=140)
a1, a2 = np.linspace(-0.05, 0.05, 1001), np.linspace(0.04, -0.04, 1001)
o1, o2 = np.arange(128.0), np.arange(96.0)
r5 = parallel.reweight_2d_sharded(l2, b2, o1, o2, a1, a2, device=dev)
if rank == 0:
    one5 = engine.reweight_2d(l2, b2, o1, o2, a1, a2, device=dev)
    assert np.allclose(r5, one5, rtol=1e-12, atol=0)
    log("reweight_2d_sharded ok: %d state points" % len(a1))
# config 3 (Taylor grid)
h3 = histogram.from_arrays(synth.two_peak_lnpi(301, scale=0.3), synth.two_comp_moments(301), 1.0, [-3.0, -2.5], 10)
bg, dg = np.linspace(0.97, 1.03, 37), np.linspace(0.4, 0.6, 29)
mk3 = lambda: h3.device_histogram(beta=bg, dmu=dg, order=2, moments=(), device=dev)
r3 = parallel.sweep_grid_sharded(mk3, -2.9, bg, dg)
if rank == 0:
    one3 = mk3().sweep(np.array([-2.9]), bg, dg, grid=True, pmax=4).host()
    assert np.array_equal(one3["nphase"], r3["nphase"]) and np.array_equal(one3["fe"][:, 0], r3["fe"][:, 0])
    log("sweep_grid_sharded ok: %d x %d grid" % (len(bg), len(dg)))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
log("mgpu_check done")
