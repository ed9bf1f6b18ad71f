"""Kernel time of the dense compact mu sweep on tilt cells (k_sweep_cell) against the table walk, and the cost of building the cells."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fhmcanalysis_b200 import _lib, engine, synth

n = 1001
lnpi = synth.two_peak_lnpi(n)
N = np.arange(n, dtype=np.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for S in (100000, 1000000, 4000000):
    mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
    for cells in (False, True):
        dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
        dh.use_mu_cells = cells
        st = dh.make_states(mu)
        buf = torch.empty(int(_lib.load().fhmc_pack_soa16_bytes(S, 4, 2)), dtype=torch.uint8, device="cuda")
        fn = lambda: dh.sweep_compact(None, pmax=4, dst=buf, states=st, fill_dead=False)
        r = fn()
        torch.cuda.synchronize()
        med, best = timed(fn)
        frac = float(r["path"].double().mean()) if cells else 0.0
        print("S %8d cells %d  %-26s median %.1f us  min %.1f us -> %.3e points/s  cell fraction %.4f" % (S, cells, _lib.last_kernel(), 1e3 * med, 1e3 * best, S / (1e-3 * med), frac), flush=True)
        if cells:
            def rebuild():
                dh._cells_key = None
                dh.ensure_mu_cells(mu)
            med, best = timed(rebuild, 10)
            print("   cells build (range on the device + 4 kernels): median %.1f us  min %.1f us" % (1e3 * med, 1e3 * best))
            hdr = dh._mu_cells[(-dh._mu_cells.data_ptr()) % 256:][:64].cpu().numpy().view(np.int32)
            print("   header: n_pieces %d n_blocks %d truncated %d (piece_cap %d)" % (hdr[6], hdr[7], hdr[8], hdr[4]))
            def cold():
                dh._cells_key = None
                fn()
            med, best = timed(cold, 10)
            print("   cold step (cells rebuilt inside): median %.1f us  min %.1f us -> %.3e points/s" % (1e3 * med, 1e3 * best, S / (1e-3 * med)))

# host cost of one call (no synchronisation inside the loop) and back-to-back device time per call
import time
S = 1000000
mu = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
dh = engine.DeviceHistogram(lnpi, N, 1.0, 0.0, smooth=10, sel=["N", N * N])
st = dh.make_states(mu)
buf = torch.empty(int(_lib.load().fhmc_pack_soa16_bytes(S, 4, 2)), dtype=torch.uint8, device="cuda")
fn = lambda: dh.sweep_compact(None, pmax=4, dst=buf, states=st, fill_dead=False)
fn(); torch.cuda.synchronize()
for reps in (20, 200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("back to back x%d: host %.1f us per call, device %.1f us per call" % (reps, 1e6 * (t1 - t0) / reps, 1e3 * a.elapsed_time(b) / reps))

# the product's sharded call (what bench.py times), host cost per call
from fhmcanalysis_b200 import parallel
mu_all = torch.linspace(-0.03, 0.03, S, dtype=torch.float64, device="cuda")
hold = {"state": None}
def step():
    _, hold["state"] = parallel.sweep_sharded_compact(dh, mu_all, pmax=4, state=hold["state"], gather=False)
for _ in range(3):
    step()
torch.cuda.synchronize()
for reps in (200,):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        step()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("sweep_sharded_compact back to back x%d: host %.1f us per call, device %.1f us per call" % (reps, 1e6 * (t1 - t0) / reps, 1e3 * a.elapsed_time(b) / reps))
