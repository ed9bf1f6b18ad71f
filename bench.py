#!/usr/bin/env python
"""bench.py -- headline benchmark of the histogram-reweighting hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2|config3]

Workload (config.workload): BASELINE config 2 -- synthetic 1-component N_tot ln(PI), N_max = 1000 (1001 bins),
smooth = 10; a "step" is one pass of the fused sweep (reweight + normalise + phase split + per-phase lnZ +
<N>, <N^2>) over 10^6 state points mu in [-0.03, 0.03] PER GPU (weak scaling: state points are independent,
each rank owns a contiguous slice of the N*10^6-point sweep; no collective on the data path, one NCCL
all_gather of the packed results after the timed region, timed separately).

Timed numbers
  value      state points/s, inputs resident in HBM, CUDA events around each step on the launch stream,
             L2 flushed (256 MiB write) before every step, max over ranks.
  e2e        the same metric through the public batched API with HOST buffers: pinned-host mu -> device,
             kernel, packed results -> pinned host, every step.
  roofline   fp64-exp issue roofline: algorithmic exps = 1001 per state point, peak = the
             register-resident exp micro-benchmark run in this same process (DFMA peak beside it).
  cpu_baseline / --impl reference
             the compiled reference (oracle/_ref; else the C port) driven like its notebooks
             (fresh copy -> reweight -> thermo -> is_safe) on a bounded sample over all host cores.
"""
import argparse
import json
import os
import hashlib
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BINS = 1001
SMOOTH = 10
S_PER_GPU = 1000000
MU_LO, MU_HI = -0.03, 0.03
PMAX = 4
METRIC = "reweighted state points/sec (lnPI+thermo) at N_max=1000"
UNIT = "state points/s"
# warp instructions per state point of k_sweep_prod2<2,1> on this workload (ncu --set full, profiles/r01b_prod2_sweep_ncu_summary.txt)
FP64_INSTR_PER_POINT = 4355
INSTR_PER_POINT = 9653
PROFILE_SOURCE = "profiles/r01b_prod2_sweep_ncu_summary.txt"
E2E_FIELDS = ("status", "nphase", "bounds", "fe", "avg")   # what the e2e arm copies back to the host every step


def workload_arrays():
    from fhmcanalysis_b200 import synth
    lnpi = synth.two_peak_lnpi(N_BINS)
    mom = synth.one_comp_moments(N_BINS)
    return lnpi, mom


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def sample_indices(total):
    """Indices into the 10^6-point mu grid of the workload that the CPU leg evaluates: every stride-th point, so that the
    very same doubles are state points of the GPU arm's timed sweep (parity in the same run)."""
    stride = max(S_PER_GPU // max(int(total), 1), 1)
    return np.arange(0, S_PER_GPU, stride)[:total]


def _cpu_chunk(args):
    kind, idx, matched = args
    lnpi, mom = workload_arrays()
    mus = np.linspace(MU_LO, MU_HI, S_PER_GPU)[idx]
    m = len(mus)
    rec = {"nphase": np.zeros(m, np.int32), "safe": np.zeros(m, np.int8), "bounds": np.full((m, PMAX, 2), -1, np.int32),
           "max_idx": np.full((m, PMAX), -1, np.int32), "min_idx": np.full((m, PMAX + 1), -1, np.int32),
           "fe": np.full((m, PMAX), np.nan), "avg": np.full((m, PMAX, 2), np.nan)}
    t0 = time.perf_counter()
    if kind == "reference":
        from oracle import ref
        import copy
        base = ref.make_histogram(lnpi, mom, 1.0, [0.0], SMOOTH)
        for k, mu in enumerate(mus):
            h = copy.deepcopy(base)
            h.reweight(float(mu))
            h.thermo(props=not matched)
            rec["safe"][k] = h.is_safe()
            th = h.data["thermo"]
            P = min(len(th), PMAX)
            rec["nphase"][k] = len(th)
            M, mn = h.data["ln(PI)_maxima_idx"], h.data["ln(PI)_minima_idx"]
            rec["max_idx"][k, :min(len(M), PMAX)] = M[:PMAX]
            rec["min_idx"][k, :min(len(mn), PMAX + 1)] = mn[:PMAX + 1]
            for p in range(P):
                rec["bounds"][k, p] = th[p]["bound_idx"]
                rec["fe"][k, p] = th[p]["F.E./kT"]
                if not matched:
                    rec["avg"][k, p, 0] = th[p]["mom"][0, 1, 0, 0, 0]
                    rec["avg"][k, p, 1] = th[p]["mom"][0, 2, 0, 0, 0]
                else:   # <N>, <N^2> the way thermo() forms them (GH:530-541), for these two arrays only
                    l, r = th[p]["bound_idx"]
                    prob = np.exp(h.data["ln(PI)"][l:r])
                    sp = np.sum(prob)
                    rec["avg"][k, p, 0] = np.sum(prob * h.data["mom"][0, 1, 0, 0, 0, l:r]) / sp
                    rec["avg"][k, p, 1] = np.sum(prob * h.data["mom"][0, 2, 0, 0, 0, l:r]) / sp
    else:
        from oracle import fhmc_oracle as fo
        i = np.arange(N_BINS, dtype=float)
        sel = np.stack([i, i * i])
        for k, mu in enumerate(mus):
            r = fo.state_point(lnpi, np.arange(N_BINS), 1.0, 0.0, float(mu), SMOOTH, sel=sel, pmax=PMAX)
            P = min(r["nphase"], PMAX)
            rec["nphase"][k], rec["safe"][k] = r["nphase"], r["safe"]
            rec["max_idx"][k, :P] = r["max_idx"][:P]
            rec["min_idx"][k, :min(len(r["min_idx"]), PMAX + 1)] = r["min_idx"][:PMAX + 1]
            rec["bounds"][k, :P] = np.asarray(r["bounds"]).reshape(-1, 2)[:P]
            rec["fe"][k, :P] = r["fe"][:P]
            rec["avg"][k, :P] = r["avg"][:P, :2]
    return len(mus), time.perf_counter() - t0, idx, rec


def cpu_arm(sample_per_core, cores=None, matched=False, keep=None):
    """Time the CPU reference on `cores` processes; returns dict(value, cores, kind, sample).  keep: path of an .npz that
    receives the sample's outputs (indices into the 10^6-point grid + records) for the GPU arm's parity check."""
    import multiprocessing as mp
    from oracle import ref
    kind = "reference" if ref.available() else "port"
    if kind == "port":
        from oracle import fhmc_oracle
        fhmc_oracle.build()
    cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    total = sample_per_core * cores
    idx = sample_indices(total)
    chunks = [(kind, idx[c::cores], matched) for c in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, [(kind, idx[:2], matched)] * cores)  # warm-up: imports, page-in
        t0 = time.perf_counter()
        res = pool.map(_cpu_chunk, chunks)
        wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    if keep:
        order = np.argsort(np.concatenate([r[2] for r in res]))
        out = {"idx": np.concatenate([r[2] for r in res])[order]}
        for k in res[0][3]:
            out[k] = np.concatenate([r[3][k] for r in res])[order]
        np.savez(keep, **out)
    work = ("deepcopy->reweight->thermo(props=False)->is_safe + the two moment averages the GPU arm forms (matched work)" if matched
            else "deepcopy->reweight->thermo->is_safe (thermo averages all 27 moment arrays; the GPU arm averages 2: ~1.6x more CPU work per point)")
    return {"value": n / wall, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d of the 10^6 state points of the workload (every %d-th), %s driven as %s, "
                      "%d processes, %.1f s wall" % (n, max(S_PER_GPU // n, 1), "compiled reference (oracle/_ref)" if kind == "reference" else "C port (oracle/fhmc_oracle.c)", work, cores, wall),
            "wall_s": wall}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    per_core = 6000
    steps, warm = max(args.steps, 1), args.warmup
    vals = []
    last = None
    for k in range(steps + min(warm, 1)):
        last = cpu_arm(per_core)
        if k >= min(warm, 1):
            vals.append(last["value"])
    v = float(np.mean(vals))
    matched = cpu_arm(max(per_core // 3, 1), matched=True)   # same per-point work as the GPU arm (two moment averages)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * per_core * last["cores"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": "config2: synthetic 1-comp N_tot lnPI, N_max=1000, smooth=10, mu sweep in [-0.03,0.03]; bounded sample",
                       "state_points_per_step": per_core * last["cores"]},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
            "matched_work": {"value": matched["value"], "unit": UNIT, "sample": matched["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def parity_block(cpu_npz, gpu, tag, rtol=1e-10):
    """Compare GPU records (dict of arrays indexed by state point: nphase, safe, bounds, fe, avg [, max_idx, min_idx]) with
    the CPU leg's outputs at the same state points.  Integers must be identical, F.E./kT and the averages agree to rtol."""
    z = np.load(cpu_npz)
    idx = z["idx"]
    idx = idx[idx < len(gpu["nphase"])]
    m = len(idx)
    ref = {k: z[k][:m] for k in z.files if k != "idx"}
    P = ref["nphase"]
    bad = int(np.sum(gpu["nphase"][idx] != P)) + int(np.sum(gpu["safe"][idx].astype(bool) != ref["safe"].astype(bool)))
    live = np.arange(PMAX)[None, :] < np.minimum(P, PMAX)[:, None]
    bad += int(np.sum((gpu["bounds"][idx].astype(np.int64) != ref["bounds"])[live]))
    if "max_idx" in gpu:
        bad += int(np.sum((gpu["max_idx"][idx] != ref["max_idx"])[live]))
        live_m = np.arange(PMAX + 1)[None, :] < (ref["min_idx"] >= 0).sum(axis=1)[:, None]
        bad += int(np.sum((gpu["min_idx"][idx] != ref["min_idx"])[live_m]))
    rel = 0.0
    with np.errstate(all="ignore"):
        for a, b in ((gpu["fe"][idx], ref["fe"]), (gpu["avg"][idx][..., 0], ref["avg"][..., 0]), (gpu["avg"][idx][..., 1], ref["avg"][..., 1])):
            d = np.abs(a - b)[live] / np.abs(b[live])
            if d.size:
                rel = max(rel, float(np.nanmax(d)))
            bad += int(np.sum(np.isnan(a[live])))
    return {"checked": tag, "n": int(m), "int_mismatches": int(bad), "max_rel": rel, "rtol": rtol, "ok": bool(bad == 0 and rel <= rtol)}


def sha16(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.idx = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from fhmcanalysis_b200 import _lib, engine
    from fhmcanalysis_b200.moments.histogram.one_dim.ntot.gc_hist import histogram

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        guard_stdout()   # NCCL's version banner and any other library chatter on descriptor 1 go to stderr
        dist.init_process_group("nccl", device_id=dev)

    # CPU baseline first, on rank 0 at N=1 only (bounded sample), before the GPU is busy
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            keep_path = os.path.join(tempfile.mkdtemp(prefix="fhmc_bench_"), "cpu_sample.npz")
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker", "--keep", keep_path],
                                 capture_output=True, text=True, timeout=600)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])
            cpu["keep"] = keep_path
        except Exception as e:  # keep the GPU measurement even if the CPU leg breaks
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": "cpu baseline failed: %r" % (e,)}

    lnpi, mom = workload_arrays()
    hist = histogram.from_arrays(lnpi, mom, 1.0, [0.0], SMOOTH)
    moments = ("N", "N2")
    dh = hist.device_histogram(moments=moments, device=dev)
    S = args.points
    # this rank's slice of the global sweep
    mu_all_lo = MU_LO + (MU_HI - MU_LO) * rank / world
    mu_all_hi = MU_LO + (MU_HI - MU_LO) * (rank + 1) / world
    mu_host = torch.from_numpy(np.linspace(mu_all_lo, mu_all_hi, S, endpoint=(rank == world - 1))).pin_memory()
    mu_dev = mu_host.to(dev)
    states = dh.make_states(mu_dev)
    out = engine.SweepResult(S, PMAX, dh.n_sel, dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    peaks = engine.measure_peaks(dev)

    def step():
        dh.sweep(None, states=states, out=out, pmax=PMAX, lanes=args.lanes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()            # samples cover warm-up + timed region + e2e loop (same kernel under load throughout)
    t_w = time.perf_counter()
    n_warm = 0
    while n_warm < max(args.warmup, 3) or time.perf_counter() - t_w < 1.0:   # >= 1 s of load before timing
        flush.zero_()
        step()
        n_warm += 1
        if n_warm % 8 == 0:
            torch.cuda.synchronize(dev)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()          # L2 flush: inputs (8 MB of mu + 32 KB blob) are far smaller than the 126 MB L2
        ev[k][0].record()
        step()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - t_wall0
    timed_kernel = _lib.last_kernel()
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    t_ms = torch.tensor([float(np.sum(kern_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * S / (ms_per_step * 1e-3)

    # ---- e2e through the public batched API with host buffers ------------------------------------
    # the per-state-point outputs SURVEY 8(d) config 2 lists: nphase, bounds, is_safe (status bit), lnZ per phase,
    # <N>, <N^2> per phase.  (extrema index lists and the normalisation constant stay on the device.)
    # Only phase slots that exist cross PCIe (phase-major repack on the device, engine.sweep_host_compact).
    host_out = {"r": None}
    h2d = mu_host.numel() * 8 + dh.h2d_bytes

    def e2e_step():
        # public host-buffer entry point (one C-ABI call): chunked, double-buffered H2D(mu) -> kernel -> repack -> D2H(results)
        host_out["r"] = dh.sweep_host_compact(mu_host, pmax=PMAX, lanes=args.lanes, out=host_out["r"])

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * S / (float(e2e_ms.item()) * 1e-3)
    d2h = int(host_out["r"]["d2h_bytes"])
    e2e_launches = 2 * ((S + (1 << 17) - 1) >> 17)      # sweep + repack kernel per chunk
    clocks = sampler.stop()

    # ---- the one collective of the path: final gather of packed results (not in `value`) ----------
    gather_ms = None
    if world > 1:
        packed = torch.cat([out.lnnorm.view(-1), out.fe.view(-1), out.avg.view(-1)])
        gathered = torch.empty(world * packed.numel(), dtype=packed.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, packed)
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(gathered, packed)
        g1.record()
        g1.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms = float(gt.item())

    # sanity: the timed kernel really produced results
    hrec = out.host()
    h_status = hrec["status"]
    ok_frac = float(np.mean((h_status & 0xFF) == 0))
    fast_frac = float(np.mean((h_status & 0x1000) != 0))

    # ---- parity in the same run: the records the TIMED kernel left behind (and the e2e arm's host records) against the
    #      outputs of the cpu_baseline leg at the same state points -------------------------
    parity = None
    if rank == 0 and cpu is not None and cpu.get("keep") and os.path.exists(cpu["keep"]):
        try:
            parity = parity_block(cpu["keep"], hrec, "records of the timed kernel (%s) vs cpu_baseline outputs (%s), same mu" % (timed_kernel, cpu["kind"]))
            e2e_rec = {"nphase": host_out["r"]["nphase"].numpy().astype(np.int32),
                       "safe": (host_out["r"]["status"].numpy().astype(np.int64) & 0x100) != 0,
                       "bounds": host_out["r"]["bounds"].numpy(), "fe": host_out["r"]["fe"].numpy(), "avg": host_out["r"]["avg"].numpy()}
            pe = parity_block(cpu["keep"], e2e_rec, "e2e host records")
            parity["e2e"] = {k: pe[k] for k in ("n", "int_mismatches", "max_rel", "ok")}
        except Exception as e:
            parity = {"checked": False, "error": repr(e)}

    # ---- second half of the BASELINE metric: coexistence points/s (config 4: 10^4 temperatures, N_max = 2000, smooth 10,
    #      order-2 beta extrapolation, one batched find_phase_eq launch, guesses = mu_ref for every temperature) ----------
    coex = None
    if rank == 0:
        try:
            from fhmcanalysis_b200 import synth
            h4 = histogram.from_arrays(synth.two_peak_lnpi(2001, scale=2.0), synth.one_comp_moments(2001, max_order=3), 1.0, [0.0], SMOOTH)
            betas4 = 1.0 / np.linspace(0.90, 1.06, 10000)
            dh4 = h4.device_histogram(beta=betas4, order=2, moments=("N", "N2", "U"), device=dev)
            g4 = np.zeros_like(betas4)
            r4 = dh4.find_phase_eq(g4, beta=betas4, lnz_tol=1e-10, pmax=4)
            torch.cuda.synchronize(dev)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                r4 = dh4.find_phase_eq(g4, beta=betas4, lnz_tol=1e-10, pmax=4)
            c1.record()
            c1.synchronize()
            hr = r4.host()
            okc = hr["code"] == 0
            coex = {"value": 3 * len(betas4) / (c0.elapsed_time(c1) * 1e-3), "unit": "coexistence points/s", "solves": len(betas4),
                    "converged_fraction": float(np.mean(okc)), "mean_evaluations": float(np.mean(hr["iters"])),
                    "median_abs_dfe": float(np.median(np.abs(hr["dfe"][okc]))) if okc.any() else None,
                    "workload": "config4: N_max=2000, smooth=10, T in [0.90,1.06], order-2 beta extrapolation, all guesses = 0"}
        except Exception as e:  # secondary metric: never lose the headline line
            coex = {"value": None, "error": repr(e)}

    if rank == 0:
        exps = S * N_BINS / (np.mean(kern_ms) * 1e-3)   # this rank's kernel
        mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = None
        if os.path.exists(mp_path):
            try:
                hbm_peak = json.load(open(mp_path)).get("hbm_gbs")
            except Exception:
                hbm_peak = None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01b_sweep_dram_bytes.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        algo_bytes = S * (8 + out.nbytes() / S)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "gpu_launches": args.steps,
            "config": {"workload": "config2: synthetic 1-comp N_tot lnPI, N_max=1000 (1001 bins), smooth=10, 10^6-point mu sweep per GPU with thermo "
                                   "(<N>, <N^2>, per-phase lnZ, phase split, is_safe)",
                       "state_points_per_gpu": S, "bins": N_BINS, "smooth": SMOOTH, "pmax": PMAX, "moments": list(moments),
                       "lanes_per_point": args.lanes or "auto", "e2e_outputs": list(E2E_FIELDS), "e2e_path": "fhmc_sweep_host_compact (C ABI, host buffers): per 2^17-point chunk H2D(mu) -> sweep kernel -> k_pack_phase_soa16 (status i16, nphase u8, bounds i16; fe/avg f64) -> D2H of the live phase blocks; upload, compute and download streams", "e2e_gpu_launches_per_step": e2e_launches, "l2": "flushed (256 MiB memset) before every timed step",
                       "parallelism": "dp%d over state points, no data-path collective" % world,
                       "final_gather_ms": gather_ms, "ok_fraction": ok_frac, "fast_kernel_fraction": fast_frac, "wall_s_timed_region": wall},
            "clocks": clocks,
            "coexistence": coex,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            # Roofline of the dominant kernel: fp64 ISSUE.  frac = fp64-pipe instructions the kernel EXECUTES per second (count per
            # state point from the committed ncu capture of this build x measured state points/s) / the DFMA issue peak measured
            # in this process.  The algorithmic count of SURVEY 8(d) (1001 exp per state point against the measured exp-issue
            # peak) is kept beside it as algorithmic_frac: it exceeds 1 because the product form replaces most exps by FMAs.
            "roofline": {"bound": "fp64_issue", "kernel": timed_kernel,
                         "achieved": S * FP64_INSTR_PER_POINT / (np.mean(kern_ms) * 1e-3) / 1e9, "peak": peaks["dfma_per_s"] / 1e9,
                         "unit": "G fp64-pipe instr/s (per lane)",
                         "frac": S * FP64_INSTR_PER_POINT / (np.mean(kern_ms) * 1e-3) / peaks["dfma_per_s"], "traffic": traffic,
                         "peak_source": "k_bench_dfma, register-resident DFMA chains, measured in this process (no fp64 figure in MEASURED_PEAKS.json)",
                         "fp64_pipe_instr_per_state_point": FP64_INSTR_PER_POINT, "instr_per_state_point": INSTR_PER_POINT,
                         "instr_count_source": PROFILE_SOURCE,
                         "algorithmic_frac": exps / peaks["exp_per_s"], "algorithmic_exp_per_state_point": N_BINS,
                         "algorithmic_achieved_gexp_s": exps / 1e9, "exp_peak_gexp_s": peaks["exp_per_s"] / 1e9,
                         "hbm": {"algorithmic_bytes_per_launch": algo_bytes, "achieved_gbs": algo_bytes / (np.mean(kern_ms) * 1e-3) / 1e9,
                                 "peak_gbs": hbm_peak, "frac": (algo_bytes / (np.mean(kern_ms) * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_peak else "absent"}},
            "parity": parity,
            "inputs_sha256_16": {"lnpi": sha16(lnpi), "mom": sha16(mom), "mu_rank0": sha16(mu_host.numpy())},
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def emit(line):
    """The one JSON line of this run, on the process's ORIGINAL stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def guard_stdout():
    """stdout must carry exactly one JSON line, but libraries write to file descriptor 1 behind Python's back (NCCL prints
    its version banner there whenever NCCL_DEBUG is VERSION or higher).  Keep a private handle on the original stdout for
    emit() and point descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=S_PER_GPU, help="state points per GPU per step")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per state point (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--keep", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_baseline_worker:
        print(json.dumps(cpu_arm(12000, keep=args.keep)))
        return 0
    guard_stdout()   # from here on only emit() reaches the caller's stdout
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    return run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
