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
    uint32_t s_sel[NSEL > 0 ? NSEL : 1];
#pragma unroll
    for (int q = 0; q < NSEL; ++q) s_sel[q] = s_lnpi + (uint32_t)a.d.sel_row[q] * row_bytes;
    const uint32_t s_slope = s_lnpi + (uint32_t)a.d.hull_row * row_bytes, s_hidx = s_slope + row_bytes;
    const int H = a.d.hull_len;

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

        auto accumulate = [&](double u, int i, double Ni) {
            const double e = exp_scaled(u, Mq, tab);
            Sacc += e;
#pragma unroll
            for (int q = 0; q < NSEL; ++q) {
                const double x = (SEL0N && q == 0) ? Ni : lds_f64(s_sel[q] + 8u * i);
                A[q] = fma(e, x, A[q]);
            }
        };
        auto load_u = [&](int i, double &Ni) {
            Ni = lds_f64(s_n + 8u * i);
            return __dadd_rn(lds_f64(s_lnpi + 8u * i), __dmul_rn(s, Ni));
        };
        double N0;
        const double u0 = load_u(0, N0);
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
        auto slow_bin = [&](int i, double xm, double xc, double xp, double Nc) {
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
            accumulate(xc, i, Nc);
        };

        if (n >= 3) {
            accumulate(u0, 0, N0);
            double Nc, xm = u0;
            double xc = load_u(1, Nc);
            double dc = __dsub_rn(xc, xm);  // sign(dc) is the exact order of (xm, xc)
            int i = 1;
            for (; i + 3 < last; i += 4) {   // bins i..i+3 are interior, i+4 <= last exists
                double N1, N2, N3, N4;
                const double x1 = load_u(i + 1, N1), x2 = load_u(i + 2, N2), x3 = load_u(i + 3, N3), x4 = load_u(i + 4, N4);
                const double d1 = __dsub_rn(x1, xc), d2 = __dsub_rn(x2, x1), d3 = __dsub_rn(x3, x2), d4 = __dsub_rn(x4, x3);
                const int flip = (__double2hiint(dc) ^ __double2hiint(d1)) | (__double2hiint(d1) ^ __double2hiint(d2)) |
                                 (__double2hiint(d2) ^ __double2hiint(d3)) | (__double2hiint(d3) ^ __double2hiint(d4));
                if (flip < 0) {   // some pair of successive differences changes sign: look closely
                    slow_bin(i, xm, xc, x1, Nc);
                    slow_bin(i + 1, xc, x1, x2, N1);
                    slow_bin(i + 2, x1, x2, x3, N2);
                    slow_bin(i + 3, x2, x3, x4, N3);
                } else {
                    accumulate(xc, i, Nc);
                    accumulate(x1, i + 1, N1);
                    accumulate(x2, i + 2, N2);
                    accumulate(x3, i + 3, N3);
                }
                xm = x3;
                xc = x4;
                Nc = N4;
                dc = d4;
            }
            for (; i < last; ++i) {
                double Np;
                const double xp = load_u(i + 1, Np);
                slow_bin(i, xm, xc, xp, Nc);
                xm = xc;
                xc = xp;
                Nc = Np;
            }
            accumulate(xc, last, Nc);
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
