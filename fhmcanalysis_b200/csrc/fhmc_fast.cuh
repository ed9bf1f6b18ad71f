// fhmc_fast.cuh -- the headline path: pure chemical-potential sweeps (no Taylor terms), one state
// point per thread, ONE pass over the bins.
//
// What makes one pass possible:
//  (1) the maximum of u_i = lnPI_i + s*N_i over i is attained on the upper concave envelope of the points
//      (N_i, lnPI_i); the host precomputes that hull once per histogram (two extra blob rows: edge slopes and
//      vertex indices) and each state point finds its vertex by a <= 11-step binary search, so the shift of the
//      max-shifted sums is known before the bins are touched;
//  (2) the phase boundaries are the windowed local minima (GH:329-330, 498-520); a sequential walk finds
//      them on the fly, so the per-phase sums can be flushed the moment a minimum is confirmed.  Bins are
//      handled four at a time: the sign bits of the four successive differences u_{k+1}-u_k (exact in sign)
//      say whether the block can contain a strict 1-neighbour extremum at all; only then (rare) are the exact
//      comparisons and the full +-smooth window test run, bin by bin.  Otherwise the four exp chains are
//      independent and interleave, which is what hides the fp64 pipe latency.
// Per bin: 2 fp64 ops for u (un-fused, bit-identical to GH:77), 1 difference, 10 for exp, 1 + NSEL accumulates.
// Everything that decides an index is afterwards validated exactly like the generic path (repair(), verify() on
// fl(u - c)); any state point that is not a plain "maxima and minima alternate, phases tile [0,n)" case, that
// overflows pmax or that contains a phase of negligible weight is re-run by the generic PointEval::run().
#pragma once
#include "fhmc_point.cuh"

namespace fhmc {

// out-of-line generic evaluation (own PointEval, so the hot loop's evaluator never has its address taken)
__device__ __noinline__ void run_generic_point(const SweepArgs &a, const double *sm, const double *s_tab, int lane,
                                               double mu1, long long sp)
{
    PointEval<1, false> pe(a, sm, lane, s_tab);
    pe.setup(mu1, a.d.beta_ref, a.d.dmu_ref);
    pe.run(sp);
}

template <int NSEL, bool SEL0N>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_mu_fast(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    const uint32_t blob_bytes = (uint32_t)a.d.n_rows * (uint32_t)a.d.n_pad * 8u;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + blob_bytes);
    double *s_tab = reinterpret_cast<double *>(smem_raw + blob_bytes + 16);
    stage_exp_table(s_tab);
    stage_blob(sm, a.blob, blob_bytes, bar);

    PointEval<1, false> pe(a, sm, threadIdx.x & 31, s_tab);
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax;
    const uint32_t row_bytes = (uint32_t)a.d.n_pad * 8u;
    const uint32_t s_lnpi = smem_u32(sm), s_n = s_lnpi + row_bytes, tab = pe.tab;
    // Interleave the rows this kernel walks into one packed array {lnPI_i, N_i, X_a(i), X_b(i), ...} (PK doubles per
    // bin, 16-byte aligned) so that a bin costs one or two LDS.128 with immediate offsets instead of one LDS.64 and
    // one address computation per row.  Built once per (persistent) CTA from the TMA-staged blob.
    constexpr int NX = NSEL - (SEL0N ? 1 : 0);          // quantities that need their own row
    constexpr int PK = 2 + NX + (NX & 1);               // doubles per packed bin (even)
    double *pk = s_tab + 64;
    for (int i = threadIdx.x; i < n; i += FHMC_CTA) {
        pk[i * PK + 0] = sm[i];
        pk[i * PK + 1] = sm[a.d.n_pad + i];
#pragma unroll
        for (int q = 0; q < NX; ++q) pk[i * PK + 2 + q] = sm[a.d.sel_row[q + (SEL0N ? 1 : 0)] * a.d.n_pad + i];
        if (NX & 1) pk[i * PK + 2 + NX] = 0.0;
    }
    __syncthreads();
    const uint32_t s_pk = smem_u32(pk);
    const uint32_t s_slope = s_lnpi + (uint32_t)a.d.hull_row * row_bytes, s_hidx = s_slope + row_bytes;
    const int H = a.d.hull_len;

    const ExpRegs ec = load_exp_regs();   // reduction / polynomial constants pinned in registers for the hot loop
    const long long S = a.st.n_states;
    for (long long sp = (long long)blockIdx.x * FHMC_CTA + threadIdx.x; sp < S; sp += (long long)gridDim.x * FHMC_CTA) {
        const double mu1 = a.st.mu1[(sp / a.st.mu1_div) % a.st.n_mu1];
        pe.setup(mu1, a.d.beta_ref, a.d.dmu_ref);
        const double s = pe.s;
        // ---- shift: hull vertex maximising lnPI + s*N ------------------------------------------
        int lo = 0, hi = H - 1;
        const double neg_s = -s;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (lds_f64(s_slope + 8u * mid) > neg_s) lo = mid + 1; else hi = mid;
        }
        const int i_max = (int)lds_f64(s_hidx + 8u * lo);
        const int Mq = shift_for_max(pe.U(i_max));

        int *maxl = a.out.max_idx + sp * pmax;
        int *minl = a.out.min_idx + sp * (pmax + 1);
        int *bl = a.out.bounds + sp * pmax * 2;
        int cntM = 0, cntm = 0, P = 0;
        bool bad = false;
        double Sacc = 0.0, Stot = 0.0, A[NSEL > 0 ? NSEL : 1];
#pragma unroll
        for (int q = 0; q < NSEL; ++q) A[q] = 0.0;

        struct Bin {
            double u, N, x[NX > 0 ? NX : 1];
        };
        auto load_bin = [&](int i, Bin &b) {
            const uint32_t addr = s_pk + (uint32_t)i * (uint32_t)(PK * 8);
            double l;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l), "=d"(b.N) : "r"(addr));
            b.u = __dadd_rn(l, __dmul_rn(s, b.N));   // un-fused, GH:77
#pragma unroll
            for (int q = 0; q < NX; q += 2) {
                double x0, x1;
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(addr + 16u + 8u * q));
                b.x[q] = x0;
                if (q + 1 < NX) b.x[q + 1] = x1;
            }
        };
        auto accumulate = [&](const Bin &b) {
            const double e = exp_scaled_r(b.u, Mq, tab, ec);
            Sacc += e;
            if (SEL0N) A[0] = fma(e, b.N, A[0]);
#pragma unroll
            for (int q = 0; q < NX; ++q) A[q + (SEL0N ? 1 : 0)] = fma(e, b.x[q], A[q + (SEL0N ? 1 : 0)]);
        };
        Bin b0;
        load_bin(0, b0);
        const double u0 = b0.u;
        auto flush = [&]() {
            if (P < pmax && Sacc >= 1e-280) {
                a.out.fe[sp * pmax + P] = -(add_shift(Mq, log(Sacc)) - u0);
#pragma unroll
                for (int q = 0; q < NSEL; ++q) a.out.avg[(sp * pmax + P) * NSEL + q] = A[q] / Sacc;
            } else {
                bad = true;
            }
            Stot += Sacc;
            Sacc = 0.0;
#pragma unroll
            for (int q = 0; q < NSEL; ++q) A[q] = 0.0;
            ++P;
        };
        // exact strict 1-neighbour test + window test of bin i (values xm, xc, xp), then its contribution
        auto slow_bin = [&](int i, double xm, const Bin &c, double xp) {
            const double xc = c.u;
            const bool is_max = (xc > xm) && (xc > xp), is_min = (xc < xm) && (xc < xp);
            if ((is_max || is_min) && pe.window_ok(i, xc, is_max, false, 0.0, 2)) {
                if (is_max) {
                    if (1 + cntM <= pmax - 1) maxl[1 + cntM] = i;
                    ++cntM;
                } else {
                    if (1 + cntm <= pmax) minl[1 + cntm] = i;
                    ++cntm;
                    flush();  // a minimum bin opens the phase to its right (GH:498-520)
                }
            }
            accumulate(c);
        };

        if (n >= 3) {
            accumulate(b0);
            Bin c;
            load_bin(1, c);
            double xm = u0;
            double dc = __dsub_rn(c.u, xm);  // sign(dc) is the exact order of (xm, xc)
            int i = 1;
#pragma unroll 2
            for (; i + 3 < last; i += 4) {   // bins i..i+3 are interior, i+4 <= last exists
                Bin b1, b2, b3, b4;
                load_bin(i + 1, b1);
                load_bin(i + 2, b2);
                load_bin(i + 3, b3);
                load_bin(i + 4, b4);
                const double d1 = __dsub_rn(b1.u, c.u), d2 = __dsub_rn(b2.u, b1.u), d3 = __dsub_rn(b3.u, b2.u), d4 = __dsub_rn(b4.u, b3.u);
                const int flip = (__double2hiint(dc) ^ __double2hiint(d1)) | (__double2hiint(d1) ^ __double2hiint(d2)) |
                                 (__double2hiint(d2) ^ __double2hiint(d3)) | (__double2hiint(d3) ^ __double2hiint(d4));
                if (flip < 0) {   // some pair of successive differences changes sign: look closely
                    slow_bin(i, xm, c, b1.u);
                    slow_bin(i + 1, c.u, b1, b2.u);
                    slow_bin(i + 2, b1.u, b2, b3.u);
                    slow_bin(i + 3, b2.u, b3, b4.u);
                } else {
                    accumulate(c);
                    accumulate(b1);
                    accumulate(b2);
                    accumulate(b3);
                }
                xm = b3.u;
                c = b4;
                dc = d4;
            }
            for (; i < last; ++i) {
                Bin nx;
                load_bin(i + 1, nx);
                slow_bin(i, xm, c, nx.u);
                xm = c.u;
                c = nx;
            }
            accumulate(c);
            flush();
        } else {
            bad = true;
        }

        // ---- validate with the exact rules of the generic path -----------------------------------
        bool done = false;
        unsigned flags = 0;
        if (!bad && !a.d.complete) {
            int nM = 0, nm = 0;
            bool part = false;
            const int rc = pe.repair(false, 0.0, cntM, cntm, 0.0, 0.0, maxl, minl, bl, nM, nm, flags, part);
            if (rc == FHMC_OK && part && nM == P) {
                pe.P = nM;
                pe.nmin = nm;
                const double c = add_shift(Mq, log(Stot));
                if (a.d.compare_raw || !pe.verify(maxl, minl, c)) {
                    const double xM = __dsub_rn(pe.U(maxl[nM - 1]), c), xl = __dsub_rn(pe.U(last), c);
                    if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
                    a.out.status[sp] = flags;
                    a.out.nphase[sp] = nM;
                    a.out.nmin[sp] = nm;
                    a.out.lnnorm[sp] = c;
                    done = true;
                }
            }
        }
        if (!done) run_generic_point(a, sm, s_tab, threadIdx.x & 31, mu1, sp);  // anything unusual: the generic evaluator redoes this state point
    }
}

}  // namespace fhmc
