"""CPU: the multi-process (N>1) host logic -- contiguous sharding of the state-point range and the final gather of the
packed result records -- exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fhmcanalysis_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for S in (0, 1, 7, 1000, 1000003):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(S, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == S
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.shard_sizes(S, world)


def _fake_records(lo, hi, pmax, nsel):
    """Deterministic stand-in for the per-state-point kernel outputs of the slice [lo, hi)."""
    s = torch.arange(lo, hi, dtype=torch.float64)
    S = hi - lo
    return {"lnnorm": s * 0.5, "fe": s[:, None] + torch.arange(pmax, dtype=torch.float64)[None, :],
            "avg": (s[:, None, None] * 3 + torch.arange(pmax * nsel, dtype=torch.float64).reshape(1, pmax, nsel)),
            "status": (s % 7).to(torch.int32), "nphase": (s % 3).to(torch.int32) + 1, "nmin": (s % 2).to(torch.int32) + 2,
            "bounds": torch.arange(S * pmax * 2, dtype=torch.int32).reshape(S, pmax, 2) + lo,
            "max_idx": torch.arange(S * pmax, dtype=torch.int32).reshape(S, pmax) + 2 * lo,
            "min_idx": torch.arange(S * (pmax + 1), dtype=torch.int32).reshape(S, pmax + 1) + 3 * lo}


def _soa16_bytes(S, pmax, nsel):
    return ((4 * S + 15) & ~15) + pmax * S * (8 * (1 + nsel) + 4)      # fhmc_pack_soa16_bytes (include/fhmc_b200.h)


def _fake_compact(lo, hi, pmax, nsel):
    s = np.arange(lo, hi)
    nph = (s % pmax + 1).astype(np.uint8)
    st = np.where(s % 11 == 0, 4, 0).astype(np.uint16) | np.where(s % 2 == 0, 0x100, 0).astype(np.uint16)
    fe = s[:, None] * 0.25 + np.arange(pmax)[None, :]
    avg = s[:, None, None] * 1.5 + np.arange(pmax * nsel).reshape(1, pmax, nsel)
    bounds = (s[:, None, None] % 100 + np.arange(pmax * 2).reshape(1, pmax, 2)).astype(np.int16)
    return {"status": st, "nphase": nph, "fe": fe, "avg": avg, "bounds": bounds}


def _pack_soa16(block, rec, smax, pmax, nsel):
    """NumPy writer of the fhmc_pack_phase_soa16 layout for smax records (only live phase slots are stored)."""
    from fhmcanalysis_b200.engine import soa16_views
    v = soa16_views(block, smax, pmax, nsel)
    m = len(rec["status"])
    v["status"][:m] = torch.from_numpy(rec["status"].view(np.int16))
    v["nphase"][:m] = torch.from_numpy(rec["nphase"])
    live = np.arange(pmax)[None, :] < np.where((rec["status"] & 0xFF) == 0, rec["nphase"], 0)[:, None]
    for p in range(pmax):
        rows = torch.from_numpy(np.where(live[:, p])[0])
        v["fe"][rows, p] = torch.from_numpy(rec["fe"][live[:, p], p])
        v["avg"][rows, p] = torch.from_numpy(rec["avg"][live[:, p], p])
        v["bounds"][rows, p] = torch.from_numpy(rec["bounds"][live[:, p], p])


def _worker(rank, world, port, S, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pmax, nsel = 4, 2
    lo, hi = parallel.shard_bounds(S, world, rank)
    f, i = parallel.pack_records(_fake_records(lo, hi, pmax, nsel))
    F, I = parallel.all_gather_records(f, i, S)
    out = parallel.unpack_records(F, I, pmax, nsel)
    ref = _fake_records(0, S, pmax, nsel)
    # bounds/max/min of the fake generator depend on the shard offset: rebuild the expectation shard by shard
    exp = {k: torch.cat([_fake_records(*parallel.shard_bounds(S, world, r), pmax, nsel)[k] for r in range(world)]) for k in ref}
    ok = all(torch.equal(out[k].reshape(exp[k].shape).to(exp[k].dtype), exp[k]) for k in exp)
    ok = ok and torch.equal(out["lnnorm"], ref["lnnorm"]) and torch.equal(out["fe"], ref["fe"])
    # fewer state points than ranks: the rank with the EMPTY shard still joins the collective (no deadlock, no error)
    lo1, hi1 = parallel.shard_bounds(1, world, rank)
    f1, i1 = parallel.pack_records(_fake_records(lo1, hi1, pmax, nsel))
    F1, I1 = parallel.all_gather_records(f1, i1, 1)
    ok = ok and F1.shape[0] == 1 and I1.shape[0] == 1 and torch.equal(F1[0], parallel.pack_records(_fake_records(0, 1, pmax, nsel))[0][0])
    # compact gather (NCCL-style path on CPU tensors): every rank packs its shard as a narrow phase-major block, one
    # all_gather_into_tensor of the blocks, ShardedRecords.host() reassembles and masks the phase slots that do not exist
    sizes = parallel.shard_sizes(S, world)
    smax = max(sizes)
    bb = (_soa16_bytes(smax, pmax, nsel) + 255) & ~255
    block = torch.full((bb,), 0x5A, dtype=torch.uint8)                    # junk where nothing is stored
    _pack_soa16(block, _fake_compact(lo, hi, pmax, nsel), smax, pmax, nsel)
    full = torch.empty(world * bb, dtype=torch.uint8)
    dist.all_gather_into_tensor(full, block)
    got = parallel.ShardedRecords(full, S, world, pmax, nsel, bb, False).host()
    exp = _fake_compact(0, S, pmax, nsel)
    live = np.arange(pmax)[None, :] < np.where((exp["status"] & 0xFF) == 0, exp["nphase"], 0)[:, None]
    ok = ok and np.array_equal(got["status"], exp["status"]) and np.array_equal(got["nphase"], exp["nphase"])
    ok = ok and np.array_equal(got["fe"][live], exp["fe"][live]) and np.all(np.isnan(got["fe"][~live]))
    ok = ok and np.array_equal(got["avg"][live], exp["avg"][live]) and np.array_equal(got["bounds"][live], exp["bounds"][live])
    ok = ok and np.all(got["bounds"][~live] == -1) and got["fe"].shape == (S, pmax)
    # full records trimmed to the live widths: the ranks agree on the largest phase / minima count (MAX all-reduce), only
    # that many per-phase columns are packed and gathered; _deliver pads them back to the capacity with NaN / -1
    pm8 = 8
    rec = _fake_records(lo, hi, pm8, nsel)
    rec["nphase"] = rec["nphase"].clamp(max=2 + rank)          # rank 1 owns the only three-phase records
    widths = parallel.live_widths(rec["nphase"], rec["nmin"], pm8)
    ok = ok and widths == (3, 3)
    ft, it = parallel.pack_records(rec, widths)
    ok = ok and ft.shape[1] == 1 + 3 + 3 * nsel and it.shape[1] == 3 + 6 + 3 + 3
    Ft, It = parallel.all_gather_records(ft, it, S)
    trimmed = parallel.unpack_records(Ft, It, pm8, nsel, widths)
    trimmed["widths"] = widths
    host = parallel._deliver(trimmed, True, pm8)
    exp8 = {k: torch.cat([_fake_records(*parallel.shard_bounds(S, world, r), pm8, nsel)[k] for r in range(world)]).numpy() for k in rec}
    ok = ok and host["fe"].shape == (S, pm8) and np.array_equal(host["fe"][:, :3], exp8["fe"][:, :3]) and np.all(np.isnan(host["fe"][:, 3:]))
    ok = ok and np.array_equal(host["avg"][:, :3], exp8["avg"][:, :3]) and np.array_equal(host["bounds"][:, :3], exp8["bounds"][:, :3])
    ok = ok and np.all(host["bounds"][:, 3:] == -1) and np.array_equal(host["max_idx"][:, :3], exp8["max_idx"][:, :3])
    ok = ok and host["min_idx"].shape == (S, pm8 + 1) and np.array_equal(host["min_idx"][:, :3], exp8["min_idx"][:, :3]) and np.all(host["min_idx"][:, 3:] == -1)
    ok = ok and np.array_equal(host["lnnorm"], exp8["lnnorm"]) and np.array_equal(host["status"], exp8["status"])
    # an empty shard joins the width agreement too
    e = parallel._empty_records(pm8, nsel, torch.device("cpu"))
    w1 = parallel.live_widths(rec["nphase"] if rank == 0 else e["nphase"], rec["nmin"] if rank == 0 else e["nmin"], pm8)
    ok = ok and w1 == (2, 3)
    q.put((rank, bool(ok), int(F.shape[0])))
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    S = 1001  # odd: shards of unequal size exercise the padding
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == S for r in res)
