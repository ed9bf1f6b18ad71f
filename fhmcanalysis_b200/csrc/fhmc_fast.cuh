// fhmc_fast.cuh -- the headline path: pure chemical-potential sweeps (no Taylor terms), one state
// point per thread, ONE pass over the bins.
//
// What makes one pass possible:
//  (1) the maximum of u_i = lnPI_i + s*N_i over i is attained on the upper concave envelope of the points
//      (N_i, lnPI_i); the host precomputes that hull once per histogram (two extra blob rows: edge slopes and
//      vertex indices) and each state point finds its vertex by a <= 11-step binary search, so the shift of the
//      max-shifted sums is known before the bins are touched;
//  (2) the phase boundaries are the windowed local minima (GH:329-330, 498-520); a sequential walk finds
//      them on the fly (strict 1-neighbour test per bin, full +-smooth window test only at the rare candidates,
//      evaluated by recomputing u from the broadcast shared-memory rows), so the per-phase sums can be flushed
//      the moment a minimum is confirmed.
// Per bin: 2 fp64 ops for u (un-fused, bit-identical to GH:77), 2 compares, 10 for exp, 1 + NSEL accumulates.
// Everything that decides an index is afterwards validated exactly like the generic path (repair(), verify() on
// fl(u - c)); any state point that is not a plain "maxima and minima alternate, phases tile [0,n)" case, that
// overflows pmax or that contains a phase of negligible weight is re-run by the generic PointEval::run().
#pragma once
#include "fhmc_point.cuh"

namespace fhmc {

template <int NSEL, bool SEL0N>
__global__ void __launch_bounds__(FHMC_CTA, 3) k_sweep_mu_fast(const __grid_constant__ SweepArgs a)
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
        int cntM = 0, cntm = 0, P = 0, left = 0;
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
        double N0 = lds_f64(s_n);
        const double u0 = __dadd_rn(lds_f64(s_lnpi), __dmul_rn(s, N0));
        auto flush = [&](int right) {
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
            left = right;
            ++P;
        };

        if (n >= 3) {
            accumulate(u0, 0, N0);
            double Nc = lds_f64(s_n + 8u);
            double xc = __dadd_rn(lds_f64(s_lnpi + 8u), __dmul_rn(s, Nc));
            bool gt_c = xc > u0, lt_c = xc < u0;
#pragma unroll 2
            for (int i = 1; i < last; ++i) {
                const double Np = lds_f64(s_n + 8u * (i + 1));
                const double xp = __dadd_rn(lds_f64(s_lnpi + 8u * (i + 1)), __dmul_rn(s, Np));
                const bool gt_p = xp > xc, lt_p = xp < xc;
                if ((gt_c && lt_p) || (lt_c && gt_p)) {  // strict 1-neighbour extremum: test the full window
                    if (pe.window_ok(i, xc, gt_c, false, 0.0, 2)) {
                        if (gt_c) {
                            if (1 + cntM <= pmax - 1) maxl[1 + cntM] = i;
                            ++cntM;
                        } else {
                            if (1 + cntm <= pmax) minl[1 + cntm] = i;
                            ++cntm;
                            flush(i);  // a minimum bin opens the phase to its right (GH:498-520)
                        }
                    }
                }
                accumulate(xc, i, Nc);
                xc = xp;
                Nc = Np;
                gt_c = gt_p;
                lt_c = lt_p;
            }
            accumulate(xc, last, Nc);
            flush(n);
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
        if (!done) pe.run(sp);  // anything unusual: the generic evaluator redoes this state point
    }
}

}  // namespace fhmc
