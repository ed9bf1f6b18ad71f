// fhmc_fast.cuh -- the throughput path for large sweeps: one state point per THREAD, one exp pass over the bins.
//
//   k_sweep_fast<NSEL, SEL0N, NC, NT>
//     NC = 0   pure chemical-potential sweep (the headline metric).  The shift of the max-shifted sums is known before
//              the bins are touched: max_i(lnPI_i + s*N_i) lies on the upper concave envelope of the points
//              (N_i, lnPI_i), precomputed by the host (two blob rows: edge slopes, vertex indices); <= 11-step binary search.
//     NC > 0   Taylor-extrapolated state points (beta and/or dmu_2 differ): lnPI' = lnPI + fl(s*N) + xi_0*N + sum_c xi_c*A_c
//              (term 0 is the dB*mu_1*N term, terms 1..NC-1 own a coefficient row).  No hull exists for a shift that is
//              nonlinear in the state variables, so a cheap max-only pre-pass (no exp) runs first.
//     NT       rows per averaged quantity (1 = no extrapolation of the quantity, 2/3 = first-order terms in dB [, dD]).
//
// Shared memory holds ONE packed, interleaved copy of the rows this kernel walks, {lnPI_i, N_i, A_1(i).., X..(i)}
// (PK doubles per bin, 16-byte aligned), so a bin costs one or a few LDS.128 with immediate offsets, and every lane
// of a warp reads the same bin (broadcast).  The packed copy is built once per persistent CTA: each needed blob row is
// staged by a 1-D TMA bulk copy (cp.async.bulk + mbarrier, SASS UBLKCP) into a one-row buffer and scattered.
//
// The walk: four bins per iteration; u_i = fl(lnPI_i + fl(s*N_i)) un-fused (bit-identical to GH:77) [+ fma terms in the
// same order as the generic evaluator]; the sign bits of the four successive differences (exact in sign) tell whether a
// strict 1-neighbour extremum can sit inside the block; only then (rare) the exact comparisons and the +-smooth window
// test of argrelextrema (GH:329-330) run, bin by bin, and a confirmed minimum flushes the running per-phase sums (a
// minimum bin opens the phase to its right, GH:498-520).  Otherwise four independent exp chains interleave, which is what
// hides the fp64 latency.  Per bin (NC = 0): 2 fp64 ops for u, 1 difference, 10 for exp, 1 + NSEL accumulates.
//
// Everything that decides an index is afterwards validated exactly like the generic path (repair(), verify() on
// fl(u - c)); any state point that is not a plain "maxima and minima alternate, phases tile [0,n)" case, that overflows
// pmax or that contains a phase of negligible weight is re-run by the generic evaluator (rows read from HBM/L2).
#pragma once
#include "fhmc_point.cuh"

#define FHMC_FAST_QUEUE 2048  // deferred-fallback queue entries per CTA (drained when fewer than one tile is free)

namespace fhmc {

template <bool TAYLOR>
__device__ __noinline__ void run_generic_point(const SweepArgs &a, const double *s_tab, int lane, double mu1, double beta,
                                               double dmu, long long sp)
{
    PointEval<1, TAYLOR> pe(a, a.blob, lane, s_tab);  // own evaluator: the hot loop's one never has its address taken
    pe.setup(mu1, beta, dmu);
    pe.run(sp);
}

// the same with all 32 lanes of the calling warp on ONE state point (queue drain: a queued point evaluated by a single lane
// kept the rest of its CTA waiting at the barrier for ~1 ms -- 29 % of the Taylor grid's kernel time)
template <bool TAYLOR>
__device__ __noinline__ void run_generic_point_warp(const SweepArgs &a, const double *s_tab, int lane, double mu1, double beta,
                                                    double dmu, long long sp)
{
    PointEval<32, TAYLOR> pe(a, a.blob, lane, s_tab);
    pe.setup(mu1, beta, dmu);
    pe.run(sp);
}

template <int NSEL, bool SEL0N, int NC, int NT, int REC = 0>
struct FastLayout {
    static constexpr int NX = NSEL - (SEL0N ? 1 : 0);   // quantities that need their own rows
    static constexpr int NCR = NC > 0 ? NC - 1 : 0;     // coefficient rows (term 0 re-uses the N row)
    static constexpr int ROWS = 2 + NCR + NX * NT;      // slots copied from blob rows
    static constexpr int RAW = ROWS + (REC == 1 ? 1 : 0);   // + the 4-bin ratio exp(lnPI_i - lnPI_{i-4}) of the recurrence
    static constexpr int PK = RAW + (RAW & 1);          // doubles per packed bin (even -> 16-byte aligned)
    static constexpr int XOFF = 2 + NCR;
    static constexpr int DOFF = ROWS;
    // REC == 2 (product form): per 4-bin block {key(Dmin), key(Dmax) | pad | P x4 | P*X_q x4 ...} doubles
    static constexpr int BW = 2 + 4 * (1 + NSEL);
    static constexpr int SEGB = 32;                     // blocks per anchor segment (128 bins)
    static constexpr int GRPB = 8;                      // blocks per key group (k_sweep_prod2: one range test clears 32 bins)
    static constexpr int QN = FHMC_FAST_QUEUE;          // deferred-fallback queue entries
    static constexpr int QTILES = 4;                    // tiles between two looks at the queue
    // (three CTAs per SM with a 512-entry queue and an 80-register cap measured 10 % slower for REC == 2)
};

// what the per-state-point walk needs to know about the CTA's packed copy
struct FastCtx {
    uint32_t s_pk, s_slope;   // shared-memory addresses: packed rows, hull edge slopes (NC == 0)
    double *s_tab;            // 2^(j/64) table, followed by the deferred-fallback queue of the sweep kernel
    const double *g_hidx;     // hull vertex bin indices (global)
    int H;                    // hull vertices
    uint32_t s_prod, s_anch;  // REC == 2: per-block product rows, per-segment anchors (max lnPI of the segment)
    uint32_t s_gkey;          // REC == 2: {min key(Dmin), max key(Dmax)} over each group of GRPB blocks
    double sdn_lim;           // REC == 2: largest |s dN| for which every segment can be walked in product form (see fast_prepare)
    double lmax;              // REC == 2: max |lnPI_i| (rounding margin of the extremum prefilter)
};

// order-preserving int32 key of the upper word of a double: key(a) < key(b) implies a < b
__device__ __forceinline__ int hi_key(double x)
{
    const int h = __double2hiint(x);
    return h ^ ((h >> 31) & 0x7fffffff);
}

// bytes of the packed rows + staging row + barrier + exp table + fallback queue (what every variant needs)
template <int PK, int QN = FHMC_FAST_QUEUE>
__host__ __device__ constexpr size_t fast_base_bytes(int n_pad)
{
    return (((size_t)n_pad * 8 * (PK + 1) + 16 + 512 + QN * 8 + 64) + 15) & ~(size_t)15;
}

// Build the CTA's packed copy (all threads of the CTA; ends with a barrier).
template <int NSEL, bool SEL0N, int NC, int NT, int REC>
__device__ __forceinline__ FastCtx fast_prepare(const SweepArgs &a, unsigned char *smem_raw)
{
    using LY = FastLayout<NSEL, SEL0N, NC, NT, REC>;
    constexpr int PK = LY::PK, XOFF = LY::XOFF;
    const int n = a.d.n, npad = a.d.n_pad;
    double *pk = reinterpret_cast<double *>(smem_raw);
    double *stage = pk + (size_t)npad * PK;
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage + npad);
    double *s_tab = reinterpret_cast<double *>(bar + 2);
    stage_exp_table(s_tab);
    // ---- build the packed copy: TMA-stage one blob row at a time, scatter it to its slot ----------------
    {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        uint32_t parity = 0;
        for (int slot = 0; slot < LY::ROWS; ++slot) {
            int row;
            if (slot < 2) row = slot;
            else if (slot < XOFF) row = a.d.coef_row[slot - 1];                      // terms 1..NC-1
            else row = a.d.sel_row[(slot - XOFF) / NT + (SEL0N ? 1 : 0)] + (slot - XOFF) % NT;
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, (uint32_t)npad * 8u);
                tma_bulk_g2s(stage, a.blob + (size_t)row * npad, (uint32_t)npad * 8u, bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
            for (int i = threadIdx.x; i < n; i += blockDim.x) pk[(size_t)i * PK + slot] = stage[i];
            __syncthreads();
        }
        if (REC == 1)   // G_i = exp(lnPI_i - lnPI_{i-4}): e_i = e_{i-4} * exp(4 s dN) * G_i along each of the four bin chains
            for (int i = threadIdx.x; i < n; i += blockDim.x)
                pk[(size_t)i * PK + LY::DOFF] = (i >= 4) ? exp(pk[(size_t)i * PK] - pk[(size_t)(i - 4) * PK]) : 1.0;
        if (LY::RAW & 1)
            for (int i = threadIdx.x; i < n; i += blockDim.x) pk[(size_t)i * PK + LY::RAW] = 0.0;
        if (NC == 0) {  // the one-row staging buffer is free now: keep the hull edge slopes in it for the binary search
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, (uint32_t)npad * 8u);
                tma_bulk_g2s(stage, a.blob + (size_t)a.d.hull_row * npad, (uint32_t)npad * 8u, bar);
            }
            mbar_wait(bar, parity);
        }
        __syncthreads();
    }
    FastCtx cx;
    cx.s_prod = cx.s_anch = cx.s_gkey = 0;
    cx.sdn_lim = 0.0;
    cx.lmax = 0.0;
    if (REC == 2) {
        // Product form: bins 1+4b .. 4+4b make block b (nb full blocks, bin 4+4b+... <= last - 0), SEGB blocks share the
        // anchor A_g = max lnPI over the segment.  P_i = exp(lnPI_i - A_g) <= 1 and P_i * X_q(i) are tabulated, so that
        // sum_i exp(lnPI_i + s N_i - shift) X_i over a block is a Horner polynomial in exp(s dN) times one running factor.
        // Each block also carries the range of tilts -s dN for which one of its bins can be a windowed extremum; only the
        // order-preserving keys of the upper words of that range are kept.
        const int nb = (n - 2) / 4;
        const int nseg = (nb + LY::SEGB - 1) / LY::SEGB;
        double *prod = reinterpret_cast<double *>(smem_raw + fast_base_bytes<PK, LY::QN>(npad));
        double *anch = prod + (size_t)nb * LY::BW;
        unsigned long long *s_lmax = reinterpret_cast<unsigned long long *>(anch + nseg);
        unsigned long long *s_rmax = s_lmax + 1;   // largest lnPI spread (max - min finite value) of a segment
        if (threadIdx.x == 0) { *s_lmax = 0ull; *s_rmax = 0ull; }
        __syncthreads();
        double lm = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) lm = fmax(lm, fabs(pk[(size_t)i * PK]));
        atomicMax(s_lmax, (unsigned long long)__double_as_longlong(lm));
        for (int g = threadIdx.x; g < nseg; g += blockDim.x) {
            const int i0 = 1 + 4 * LY::SEGB * g, i1 = min(1 + 4 * LY::SEGB * (g + 1), 1 + 4 * nb);
            double m = -CUDART_INF, lo = CUDART_INF;
            for (int i = i0; i < i1; ++i) {
                const double v = pk[(size_t)i * PK];
                m = fmax(m, v);
                if (v > -CUDART_INF) lo = fmin(lo, v);
            }
            anch[g] = m;
            if (lo < CUDART_INF) atomicMax(s_rmax, (unsigned long long)__double_as_longlong(m - lo));
        }
        __syncthreads();
        const double lmax_all = __longlong_as_double((long long)*s_lmax);
        const double dN0 = pk[PK + 1] - pk[1];
        const double slack = 4.0 * (1.8e-15 * (lmax_all + 4.5 * fmax(fabs(pk[1]), fabs(pk[(size_t)(n - 1) * PK + 1])) / dN0) + 1e-300)
                             + 1e-12 * fabs(dN0);
        for (int b = threadIdx.x; b < nb; b += blockDim.x) {
            const int i = 1 + 4 * b;
            const double A = anch[b / LY::SEGB];
            // For which tilts a = -s dN can a bin of this block be a windowed extremum?  u_j - u_k = lnPI_j - lnPI_k - a (j-k),
            // so bin j is a strict maximum over +-w iff max_d SR_d < a < min_d SL_d with the chord slopes
            // SL_d = (lnPI_j - lnPI_jl)/(j - jl), SR_d = (lnPI_jr - lnPI_j)/(jr - j) (jl, jr clipped like argrelextrema's
            // 'clip' mode), and a strict minimum iff max_d SL_d < a < min_d SR_d.  The block keeps the hull [dmin, dmax] of
            // its non-empty intervals (widened by `slack`, the largest rounding margin a chained state point can have).
            double dmin = CUDART_INF, dmax = -CUDART_INF;
            for (int j = i; j < i + 4; ++j) {
                const double lj = pk[(size_t)j * PK];
                double maxSL = -CUDART_INF, minSL = CUDART_INF, maxSR = -CUDART_INF, minSR = CUDART_INF;
                for (int d = 1; d <= a.d.smooth; ++d) {
                    const int jl = max(j - d, 0), jr = min(j + d, n - 1);
                    const double SL = (lj - pk[(size_t)jl * PK]) / (double)(j - jl);
                    const double SR = (pk[(size_t)jr * PK] - lj) / (double)(jr - j);
                    maxSL = fmax(maxSL, SL);
                    minSL = fmin(minSL, SL);
                    maxSR = fmax(maxSR, SR);
                    minSR = fmin(minSR, SR);
                }
                if (maxSR <= minSL + slack) { dmin = fmin(dmin, maxSR); dmax = fmax(dmax, minSL); }   // maximum possible
                if (maxSL <= minSR + slack) { dmin = fmin(dmin, maxSL); dmax = fmax(dmax, minSR); }   // minimum possible
            }
            double *pb = prod + (size_t)b * LY::BW;
            int *kb = reinterpret_cast<int *>(pb);
            kb[0] = hi_key(dmin);
            kb[1] = hi_key(dmax);
            pb[1] = 0.0;
            for (int k = 0; k < 4; ++k) {
                const double *row = pk + (size_t)(i + k) * PK;
                const double P = exp(row[0] - A);
                pb[2 + k] = P;
                if (SEL0N) pb[6 + k] = P * row[1];
                for (int q = 0; q < LY::NX; ++q) pb[2 + 4 * (1 + q + (SEL0N ? 1 : 0)) + k] = P * row[XOFF + q];
            }
        }
        __syncthreads();
        // hull of the tilt ranges of GRPB consecutive blocks: a group no state point of the thread can have an extremum in is
        // summed without looking at its blocks' keys
        int *gkey = reinterpret_cast<int *>(s_rmax + 1);
        for (int g = threadIdx.x; g < (nb + LY::GRPB - 1) / LY::GRPB; g += blockDim.x) {
            int lo = 0x7fffffff, hi = (int)0x80000000;
            for (int b = g * LY::GRPB; b < min(nb, (g + 1) * LY::GRPB); ++b) {
                const int *kb = reinterpret_cast<const int *>(prod + (size_t)b * LY::BW);
                lo = min(lo, kb[0]);
                hi = max(hi, kb[1]);
            }
            gkey[2 * g] = lo;
            gkey[2 * g + 1] = hi;
        }
        __syncthreads();
        cx.s_prod = smem_u32(prod);
        cx.s_anch = smem_u32(anch);
        cx.s_gkey = smem_u32(gkey);
        // P_i = exp(lnPI_i - A_g) and the running factor t r^k must both stay normal numbers wherever their product
        // matters: the lnPI spread of a segment plus the growth of t over its 128 bins has to fit the fp64 exponent
        // range, and exp(|s dN| * 128) itself must stay finite (4.5).  State points beyond the limit take true exps.
        cx.sdn_lim = fmin(4.5, (600.0 - __longlong_as_double((long long)*s_rmax)) / 128.0);
        cx.lmax = __longlong_as_double((long long)*s_lmax);
    }
    cx.s_slope = smem_u32(stage);
    cx.s_pk = smem_u32(pk);
    cx.s_tab = s_tab;
    cx.g_hidx = a.blob + (size_t)a.d.hull_row * npad + npad;
    cx.H = a.d.hull_len;
    return cx;
}

// One state point, walked by the calling thread; fe/avg/bounds/extrema/status go to record `sp`.  Returns false when the
// state point is not a plain case and must be re-run by the generic evaluator (nothing final has been written then).
template <int NSEL, bool SEL0N, int NC, int NT, int REC>
__device__ __forceinline__ bool fast_point(const SweepArgs &a, const FastCtx &cx, PointEval<1, (NC > 0) || (NT > 1)> &pe,
                                           const ExpRegs &ec, long long sp, double mu1, double beta, double dmu)
{
    using LY = FastLayout<NSEL, SEL0N, NC, NT, REC>;
    constexpr int NX = LY::NX, PK = LY::PK, XOFF = LY::XOFF;
    constexpr bool TAYLOR = (NC > 0) || (NT > 1);
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax;
    const uint32_t s_pk = cx.s_pk, s_slope = cx.s_slope, tab = pe.tab;
    const double *g_hidx = cx.g_hidx;
    const int H = cx.H;
    (void)TAYLOR; (void)XOFF; (void)s_slope; (void)g_hidx; (void)H;
        pe.setup(mu1, beta, dmu);
        const double s = pe.s;
        double xi[NC > 0 ? NC : 1], ts[NT];
#pragma unroll
        for (int c = 0; c < NC; ++c) xi[c] = pe.xi[c];
#pragma unroll
        for (int t = 0; t < NT; ++t) ts[t] = TAYLOR ? pe.ts[t] : 1.0;

        struct Bin {
            double u, N, x[NX > 0 ? NX : 1], g;
        };
        // u in the generic evaluator's order: fl(lnPI + fl(s*N)), then fma per Taylor term
        auto load_u = [&](int i, double &Ni) {
            const uint32_t addr = s_pk + (uint32_t)i * (uint32_t)(PK * 8);
            double l;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l), "=d"(Ni) : "r"(addr));
            double u = __dadd_rn(l, __dmul_rn(s, Ni));
            if (NC > 0) {
                u = fma(xi[0], Ni, u);
#pragma unroll
                for (int c = 1; c < NC; c += 2) {
                    double a0, a1;
                    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a0), "=d"(a1) : "r"(addr + 8u * (1 + c)));
                    u = fma(xi[c], a0, u);
                    if (c + 1 < NC) u = fma(xi[c + 1], a1, u);
                }
            }
            return u;
        };
        auto load_bin = [&](int i, Bin &b) {
            b.u = load_u(i, b.N);
            const uint32_t addr = s_pk + (uint32_t)i * (uint32_t)(PK * 8) + 8u * XOFF;
#pragma unroll
            for (int q = 0; q < NX; ++q) {
                double x = lds_f64(addr + 8u * (q * NT));
#pragma unroll
                for (int t = 1; t < NT; ++t) x = fma(ts[t], lds_f64(addr + 8u * (q * NT + t)), x);
                b.x[q] = x;
            }
            if (REC == 1) b.g = lds_f64(s_pk + (uint32_t)i * (uint32_t)(PK * 8) + 8u * LY::DOFF);
        };

        // ---- shift -------------------------------------------------------------------------------------
        int Mq;
        double approx_max = 0.0, seen_max = -CUDART_INF;
        if (NC == 0) {   // hull vertex maximising lnPI + s*N
            int lo = 0, hi = H - 1;
            const double neg_s = -s;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lds_f64(s_slope + 8u * mid) > neg_s) lo = mid + 1; else hi = mid;
            }
            double Nm;
            Mq = shift_for_max(load_u((int)g_hidx[lo], Nm));
        } else {
            // Max-only pre-pass over every 8th bin: the exponent shift need not sit on the exact maximum, only close
            // enough that nothing overflows.  The walk tracks the true maximum; a state point whose subsample missed
            // it by more than e^400 (pathologically spiky input) is handed to the generic evaluator.
            double m0 = -CUDART_INF, m1 = -CUDART_INF, Nd;
            int i = 0;
            for (; i + 8 < n; i += 16) {
                m0 = fmax(m0, load_u(i, Nd));
                m1 = fmax(m1, load_u(i + 8, Nd));
            }
            for (; i < n; i += 8) m0 = fmax(m0, load_u(i, Nd));
            m0 = fmax(m0, load_u(last, Nd));
            approx_max = fmax(m0, m1);
            Mq = shift_for_max(approx_max);
        }

        int *maxl = a.out.max_idx + sp * pmax;
        int *minl = a.out.min_idx + sp * (pmax + 1);
        int *bl = a.out.bounds + sp * pmax * 2;
        int cntM = 0, cntm = 0, P = 0;
        bool bad = false;
        unsigned rescue = 0;
        double Sacc = 0.0, Stot = 0.0, A[NSEL > 0 ? NSEL : 1];
#pragma unroll
        for (int q = 0; q < NSEL; ++q) A[q] = 0.0;

        auto add_term = [&](const Bin &b, double e) {
            Sacc += e;
            if (SEL0N) A[0] = fma(e, b.N, A[0]);
#pragma unroll
            for (int q = 0; q < NX; ++q) A[q + (SEL0N ? 1 : 0)] = fma(e, b.x[q], A[q + (SEL0N ? 1 : 0)]);
        };
        auto accumulate = [&](const Bin &b) {
            if (NC > 0) seen_max = fmax(seen_max, b.u);
            const double e = exp_scaled_r(b.u, Mq, tab, ec);
            add_term(b, e);
            return e;
        };
        Bin b0;
        load_bin(0, b0);
        const double u0 = b0.u;
        auto flush = [&]() {
            if (P < pmax && Sacc >= 1e-280) {
                a.out.fe[sp * pmax + P] = -(add_shift(Mq, log(Sacc)) - u0);
#pragma unroll
                for (int q = 0; q < NSEL; ++q) a.out.avg[(sp * pmax + P) * NSEL + q] = A[q] / Sacc;
            } else if (P < pmax && P < 32) {
                rescue |= 1u << P;   // phase too unlikely for the common shift: re-integrated about its own maximum below
            } else {   // (also: a negligible phase beyond the 32 the rescue mask can name -- left to the generic evaluator)
                bad = true;
            }
            Stot += Sacc;
            Sacc = 0.0;
#pragma unroll
            for (int q = 0; q < NSEL; ++q) A[q] = 0.0;
            ++P;
        };
        // remaining shifts d0..smooth of the argrelextrema test, on the packed shared-memory rows (the generic
        // evaluator's window_ok()/verify() would read the blob through global memory here)
        // `robust` (REC == 2): every comparison behind an accepted extremum was decided by more than vmargin, the most
        // that rounding fl(u - c) can move two values against each other -- the re-test on the normalised values
        // below then cannot differ and is skipped.
        bool robust = (REC == 2);
        double vmargin = 0.0;
        if (REC == 2) {
            const double Na = fmax(fabs(lds_f64(s_pk + 8u)), fabs(lds_f64(s_pk + (uint32_t)last * (uint32_t)(PK * 8) + 8u)));
            vmargin = 1.8e-15 * (cx.lmax + fabs(s) * Na) + 1e-14;   // >= 2^-51 (|u| + |c|) for every bin
        }
        auto window_fast = [&](int i, double xc, bool is_max, bool use_c, double cc, int d0) {
            double Nd;
            const double xg = is_max ? xc - vmargin : xc + vmargin;
            bool rb = true;
            for (int d = d0; d <= pe.w; ++d) {
                const int jl = (i - d < 0) ? 0 : i - d;
                const int jr = (i + d > last) ? last : i + d;
                double xl = load_u(jl, Nd), xr = load_u(jr, Nd);
                if (use_c) { xl = __dsub_rn(xl, cc); xr = __dsub_rn(xr, cc); }
                const bool ok = is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr);
                if (!ok) return false;
                if (REC == 2) rb &= is_max ? (xg > xl && xg > xr) : (xg < xl && xg < xr);
            }
            if (REC == 2 && !use_c) robust &= rb;
            return true;
        };
        // exact strict 1-neighbour test + window test of bin i (values xm, xc, xp), then its contribution
        auto test_bin = [&](int i, double xm, double xc, double xp) {
            const bool is_max = (xc > xm) && (xc > xp), is_min = (xc < xm) && (xc < xp);
            if ((is_max || is_min) && window_fast(i, xc, is_max, false, 0.0, 2)) {
                if (REC == 2) robust &= is_max ? (xc - vmargin > xm && xc - vmargin > xp) : (xc + vmargin < xm && xc + vmargin < xp);
                if (is_max) {
                    if (1 + cntM <= pmax - 1) maxl[1 + cntM] = i;
                    ++cntM;
                } else {
                    if (1 + cntm <= pmax) minl[1 + cntm] = i;
                    ++cntm;
                    flush();  // a minimum bin opens the phase to its right (GH:498-520)
                }
            }
        };
        auto slow_bin = [&](int i, double xm, const Bin &c, double xp) {
            test_bin(i, xm, c.u, xp);
            return accumulate(c);
        };
        // exp recurrence (REC): four chains, one per bin position in the block; e_i = (e_{i-4} * r4) * G_i.  Chains are
        // re-anchored with true exps every 16 blocks and in every block that takes the slow (exact-test) path, which
        // bounds the accumulated rounding at ~3e-15 relative.
        double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0, r4 = 1.0;
        int since_anchor = 16;
        double r1 = 1.0;
        if (REC) {
            const double sdn = s * (lds_f64(s_pk + (uint32_t)(PK * 8) + 8u) - lds_f64(s_pk + 8u));   // s dN
            const double t4 = 4.0 * sdn;
            if (!(fabs(t4) < 200.0)) bad = true;   // extreme tilt: leave it to the generic evaluator
            r4 = exp(t4);
            if (REC == 2) r1 = exp(sdn);
        }

        if (n >= 3 && !bad) {
            accumulate(b0);
            Bin c;
            load_bin(1, c);
            double xm = u0;
            double dc = __dsub_rn(c.u, xm);  // sign(dc) is the exact order of (xm, xc)
            int i = 1;
            if constexpr (REC == 2) {
                // ---- product form -------------------------------------------------------------------------------
                // A block's sums are Horner polynomials in r1 = exp(s dN) over the tabulated P_i (* X_i), scaled by the
                // running factor t = exp(A_g + s N_i - shift) of the block's first bin: 4 fp64 ops per block and summed
                // quantity, + 1 for t.  u itself is only formed in blocks that may hold an extremum: a block skips the
                // exact tests when every difference D_j + s dN that touches it is further from zero than the rounding
                // of fl(u_j - u_{j-1}) can reach (keys: upper words, conservative).
                const double sdn = s * (lds_f64(s_pk + (uint32_t)(PK * 8) + 8u) - lds_f64(s_pk + 8u));
                const double Na = fmax(fabs(lds_f64(s_pk + 8u)), fabs(lds_f64(s_pk + (uint32_t)last * (uint32_t)(PK * 8) + 8u)));
                const double margin = 1.8e-15 * (cx.lmax + fabs(s) * Na) + 1e-300;   // 8 * 2^-52 * (|lnPI| + |s N|)
                const bool chain_ok = fabs(sdn) < cx.sdn_lim;   // product form usable for this tilt (fast_prepare)
                // (the tabulated ranges assume |s dN| < 4.5 in their rounding slack: beyond it every block is examined)
                const int k_hi = chain_ok ? hi_key(-sdn + margin) : 0x7fffffff, k_lo = chain_ok ? hi_key(-sdn - margin) : (int)0x80000000;
                const int nb = (n - 2) / 4;
                const double r2 = r1 * r1, r8 = r4 * r4;
                constexpr uint32_t BWB = (uint32_t)(LY::BW * 8);
                // sums of one block from the tabulated products: Horner in r1, split in two halves (depth 2)
                auto fast_block = [&](uint32_t pb, double tb) {
                    double p0, p1, p2, p3;
                    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p0), "=d"(p1) : "r"(pb + 16u));
                    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p2), "=d"(p3) : "r"(pb + 32u));
                    Sacc = fma(fma(fma(p3, r1, p2), r2, fma(p1, r1, p0)), tb, Sacc);
#pragma unroll
                    for (int q = 0; q < NSEL; ++q) {
                        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p0), "=d"(p1) : "r"(pb + 48u + 32u * q));
                        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p2), "=d"(p3) : "r"(pb + 64u + 32u * q));
                        A[q] = fma(fma(fma(p3, r1, p2), r2, fma(p1, r1, p0)), tb, A[q]);
                    }
                };
                // exact tests on u, bin by bin (a confirmed minimum flushes the sums before its own term is added); the
                // terms themselves still come from the tabulated products: e_{i+k} = P_{i+k} t r1^k
                auto careful_block = [&](uint32_t pb, int ib, double tb) {
                    double Nd, uu[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) uu[k] = load_u(ib - 1 + k, Nd);
                    double tk = tb;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        test_bin(ib + k, uu[k], uu[k + 1], uu[k + 2]);
                        Sacc = fma(lds_f64(pb + 16u + 8u * k), tk, Sacc);
#pragma unroll
                        for (int q = 0; q < NSEL; ++q) A[q] = fma(lds_f64(pb + 48u + 32u * q + 8u * k), tk, A[q]);
                        tk *= r1;
                    }
                };
                auto flagged = [&](uint32_t pb) {   // an extremum may sit in the block at pb
                    int ka, kb;
                    asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka), "=r"(kb) : "r"(pb));
                    return !((ka > k_hi) | (kb < k_lo));
                };
                uint32_t pb = cx.s_prod;
                int b = 0;
                for (int g = 0; b < nb; ++g) {
                    const int bend = min(nb, b + LY::SEGB);
                    // the tabulated P_i of this segment are relative to A_g: t = exp(A_g + s N_i - shift) at its first bin
                    double Ni;
                    const double lA = lds_f64(cx.s_anch + 8u * (uint32_t)g);
                    asm("ld.shared.f64 %0, [%1];" : "=d"(Ni) : "r"(s_pk + (uint32_t)i * (uint32_t)(PK * 8) + 8u));
                    double t = exp_scaled_r(__dadd_rn(lA, __dmul_rn(s, Ni)), Mq, tab, ec);
                    // exp_scaled_r() clamps an underflowing result to [2^-1022, 2^-1020): anything that small is not usable.
                    if (!(t > 1e-300) || !chain_ok) {
                        // clamped (underflowed) or unusable factor: the segment takes one true exp per bin instead
                        for (; b < bend; ++b, i += 4, pb += BWB) {
                            if (flagged(pb)) {
                                double Nd;
                                Bin c0, b1, b2, b3;
                                const double um = load_u(i - 1, Nd);
                                load_bin(i, c0);
                                load_bin(i + 1, b1);
                                load_bin(i + 2, b2);
                                load_bin(i + 3, b3);
                                const double u4 = load_u(i + 4, Nd);
                                slow_bin(i, um, c0, b1.u);
                                slow_bin(i + 1, c0.u, b1, b2.u);
                                slow_bin(i + 2, b1.u, b2, b3.u);
                                slow_bin(i + 3, b2.u, b3, u4);
                            } else {
                                Bin c0;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    load_bin(i + k, c0);
                                    accumulate(c0);
                                }
                            }
                        }
                        continue;
                    }
                    // two blocks per iteration; the range keys of the next two are fetched while these are summed
                    int ka0, kb0, ka1, kb1;
                    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka0), "=r"(kb0) : "r"(pb));
                    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka1), "=r"(kb1) : "r"(pb + BWB));
                    for (; b + 1 < bend; b += 2, i += 8, pb += 2u * BWB) {
                        const bool fa = !((ka0 > k_hi) | (kb0 < k_lo)), fb = !((ka1 > k_hi) | (kb1 < k_lo));
                        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka0), "=r"(kb0) : "r"(pb + 2u * BWB));
                        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka1), "=r"(kb1) : "r"(pb + 3u * BWB));
                        const double t2 = t * r4;
                        if (fa | fb) {
                            if (fa) careful_block(pb, i, t); else fast_block(pb, t);
                            if (fb) careful_block(pb + BWB, i + 4, t2); else fast_block(pb + BWB, t2);
                        } else {
                            fast_block(pb, t);
                            fast_block(pb + BWB, t2);
                        }
                        t *= r8;
                    }
                    if (b < bend) {
                        if (flagged(pb)) careful_block(pb, i, t); else fast_block(pb, t);
                        ++b;
                        i += 4;
                        pb += BWB;
                    }
                }
                double Nd;
                xm = load_u(i - 1, Nd);
                load_bin(i, c);
            } else {
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
                    e0 = slow_bin(i, xm, c, b1.u);
                    e1 = slow_bin(i + 1, c.u, b1, b2.u);
                    e2 = slow_bin(i + 2, b1.u, b2, b3.u);
                    e3 = slow_bin(i + 3, b2.u, b3, b4.u);
                    since_anchor = 0;
                } else if (!REC || since_anchor >= 16) {
                    e0 = accumulate(c);
                    e1 = accumulate(b1);
                    e2 = accumulate(b2);
                    e3 = accumulate(b3);
                    since_anchor = 0;
                } else {
                    e0 = (e0 * r4) * c.g;
                    e1 = (e1 * r4) * b1.g;
                    e2 = (e2 * r4) * b2.g;
                    e3 = (e3 * r4) * b3.g;
                    add_term(c, e0);
                    add_term(b1, e1);
                    add_term(b2, e2);
                    add_term(b3, e3);
                }
                ++since_anchor;
                xm = b3.u;
                c = b4;
                dc = d4;
            }
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
            if (NC > 0 && !(seen_max - approx_max < 400.0)) bad = true;
        } else {
            bad = true;
        }

        // ---- validate with the exact rules of the generic path -----------------------------------
        bool done = false;
        unsigned flags = 0;
        if (!bad && !a.d.complete) {
            int nM = 0, nm = 0;
            bool part = false;
            const double c = add_shift(Mq, log(Stot));
            int rc;
            if (cntM == 0 && cntm == 0) {
                // No windowed extremum at all (monotone ln(PI), e.g. far from coexistence on a Taylor grid): the
                // reference takes the bins tied with the max / min of the NORMALISED array (GH:382-386).  One light
                // scan on fl(u - c) here; only genuine ties go to the generic evaluator.
                double vM = -CUDART_INF, vm = CUDART_INF, Nd;
                int cM = 0, cm = 0, pM = 0, pm = 0;
                robust = false;   // nothing was decided by the walk: keep the re-test of whatever repair() lists
                for (int j = 0; j < n; ++j) {
                    const double v = __dsub_rn(load_u(j, Nd), c);
                    if (v > vM) { vM = v; cM = 1; pM = j; } else if (v == vM) ++cM;
                    if (v < vm) { vm = v; cm = 1; pm = j; } else if (v == vm) ++cm;
                }
                rc = (cM == 1 && cm == 1) ? pe.repair(true, c, 0, 0, 0.0, 0.0, maxl, minl, bl, nM, nm, flags, part, 1, 1, pM, pm)
                                          : FHMC_NEED_SLOW;
            } else {
                rc = pe.repair(false, 0.0, cntM, cntm, 0.0, 0.0, maxl, minl, bl, nM, nm, flags, part);
            }
            if (rc == FHMC_OK && part && nM == P) {
                pe.P = nM;
                pe.nmin = nm;
                // re-test the detected interior extrema on the normalised values, as PointEval::verify()
                bool differs = false;
                if (!a.d.compare_raw && !robust) {
                    double Nd;
                    for (int k = 0; k < nM + nm && !differs; ++k) {
                        const bool is_max = k < nM;
                        const int idx = is_max ? maxl[k] : minl[k - nM];
                        if (idx > 0 && idx < last)
                            differs = !window_fast(idx, __dsub_rn(load_u(idx, Nd), c), is_max, true, c, 1);
                    }
                }
                if (!differs) {
                    double Nd;
                    // phases whose weight underflowed next to the global maximum (far from coexistence): integrate them
                    // about their own maximum, as PointEval::partition_sum_probe() does (status bit RESCUED)
                    for (int p = 0; rescue != 0 && p < nM; ++p) {
                        if (!((rescue >> p) & 1u)) continue;
                        const int left = bl[2 * p], right = bl[2 * p + 1];
                        double ml = -CUDART_INF;
                        for (int j = left; j < right; ++j) ml = fmax(ml, load_u(j, Nd));
                        const int Mp = shift_for_max(ml);
                        double Sp = 0.0, Ap[NSEL > 0 ? NSEL : 1];
#pragma unroll
                        for (int q = 0; q < NSEL; ++q) Ap[q] = 0.0;
                        for (int j = left; j < right; ++j) {
                            Bin b;
                            load_bin(j, b);
                            const double e = exp_scaled_r(b.u, Mp, tab, ec);
                            Sp += e;
                            if (SEL0N) Ap[0] = fma(e, b.N, Ap[0]);
#pragma unroll
                            for (int q = 0; q < NX; ++q) Ap[q + (SEL0N ? 1 : 0)] = fma(e, b.x[q], Ap[q + (SEL0N ? 1 : 0)]);
                        }
                        a.out.fe[sp * pmax + p] = -(add_shift(Mp, log(Sp)) - u0);
#pragma unroll
                        for (int q = 0; q < NSEL; ++q) a.out.avg[(sp * pmax + p) * NSEL + q] = Ap[q] / Sp;
                        flags |= FHMC_ST_RESCUED;
                    }
                    const double xM = __dsub_rn(load_u(maxl[nM - 1], Nd), c), xl = __dsub_rn(load_u(last, Nd), c);
                    if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
                    a.out.status[sp] = flags | FHMC_ST_FAST;
                    a.out.nphase[sp] = nM;
                    a.out.nmin[sp] = nm;
                    a.out.lnnorm[sp] = c;
                    done = true;
                }
            }
        }
        return done;
}

template <int NSEL, bool SEL0N, int NC, int NT, int REC = 0>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_fast(const __grid_constant__ SweepArgs a)
{
    static_assert(!REC || (NC == 0 && NT == 1), "the exp recurrence only exists for pure mu sweeps");
    constexpr bool TAYLOR = (NC > 0) || (NT > 1);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FastCtx cx = fast_prepare<NSEL, SEL0N, NC, NT, REC>(a, smem_raw);
    double *s_tab = cx.s_tab;
    PointEval<1, TAYLOR> pe(a, a.blob, threadIdx.x & 31, s_tab);   // rare paths (repair) read HBM/L2
    const ExpRegs ec = load_exp_regs();   // reduction / polynomial constants pinned in registers for the hot loop

    // Irregular state points are not re-run on the spot (one such lane would stall its whole warp for a full generic
    // evaluation): they are queued in shared memory and drained by all warps of the CTA, one queued point per warp.
    // The queue is inspected (one CTA barrier) only every QTILES-th tile.
    using LY = FastLayout<NSEL, SEL0N, NC, NT, REC>;
    long long *queue = reinterpret_cast<long long *>(s_tab + 64);
    int *q_count = reinterpret_cast<int *>(queue + LY::QN);
    if (threadIdx.x == 0) *q_count = 0;
    __syncthreads();
    auto drain = [&]() {
        const int cnt = *q_count;
        for (int k = threadIdx.x >> 5; k < cnt; k += FHMC_CTA / 32) {   // one queued state point per WARP
            const long long qs = queue[k];
            const double qm = a.st.mu1[(qs / a.st.mu1_div) % a.st.n_mu1];
            const double qb = (TAYLOR && a.st.beta) ? a.st.beta[(qs / a.st.beta_div) % a.st.n_beta] : a.d.beta_ref;
            const double qd = (TAYLOR && a.st.dmu) ? a.st.dmu[(qs / a.st.dmu_div) % a.st.n_dmu] : a.d.dmu_ref;
            run_generic_point_warp<TAYLOR>(a, s_tab, threadIdx.x & 31, qm, qb, qd, qs);
        }
        __syncthreads();
        if (threadIdx.x == 0) *q_count = 0;
        __syncthreads();
    };
    int tile_no = 0;

    const long long S = a.st.n_states;
    for (long long base = (long long)blockIdx.x * FHMC_CTA; base < S; base += (long long)gridDim.x * FHMC_CTA) {
      const long long sp = base + threadIdx.x;
      if (sp < S) {
        const double mu1 = a.st.mu1[(sp / a.st.mu1_div) % a.st.n_mu1];
        const double beta = (TAYLOR && a.st.beta) ? a.st.beta[(sp / a.st.beta_div) % a.st.n_beta] : a.d.beta_ref;
        const double dmu = (TAYLOR && a.st.dmu) ? a.st.dmu[(sp / a.st.dmu_div) % a.st.n_dmu] : a.d.dmu_ref;
        if (!fast_point<NSEL, SEL0N, NC, NT, REC>(a, cx, pe, ec, sp, mu1, beta, dmu))
            queue[atomicAdd(q_count, 1)] = sp;  // anything unusual: defer to the generic evaluator
      }
      if ((++tile_no % LY::QTILES) == 0) {
          __syncthreads();
          if (*q_count > LY::QN - LY::QTILES * FHMC_CTA) drain();   // uniform across the CTA (read after the barrier)
      }
    }
    __syncthreads();
    drain();
}

// shared memory the fast kernel needs for this histogram
template <int NSEL, bool SEL0N, int NC, int NT, int REC = 0>
static size_t fast_smem_bytes(int n_pad)
{
    using LY = FastLayout<NSEL, SEL0N, NC, NT, REC>;
    size_t b = fast_base_bytes<LY::PK, LY::QN>(n_pad);
    if (REC == 2) {
        const size_t nb = (size_t)(n_pad / 4 + 1);
        b += (nb + 3) * LY::BW * 8 + (nb / LY::SEGB + 2) * 8 + 32;   // (+3 blocks: the key prefetch reads ahead)
        b += (nb / LY::GRPB + 2) * 8;                                  // group keys
    }
    return b;
}

}  // namespace fhmc

// ---------------------------------------------------------------------------------------------------------
// host side: launch helpers shared by fhmc_b200.cu (NC = 0 instantiations) and fhmc_fast_taylor.cu (NC > 0)
// ---------------------------------------------------------------------------------------------------------
namespace fhmc {

// returns 0 ok, 1 error, -1 "does not fit / not applicable" (caller falls back to the generic kernel)
template <int NSEL, bool SEL0N, int NC, int NT, int REC = 0>
static int launch_fast(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const size_t smem = fast_smem_bytes<NSEL, SEL0N, NC, NT, REC>(args.d.n_pad);
    if (smem > (size_t)smem_optin) return -1;
    auto kern = k_sweep_fast<NSEL, SEL0N, NC, NT, REC>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FHMC_CTA, smem), "occupancy query")) return 1;
    if (occ < 1) return -1;
    const long long ntiles = (args.st.n_states + FHMC_CTA - 1) / FHMC_CTA;
    long long grid = (long long)sm_count * occ;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, FHMC_CTA, smem, stream>>>(args);
    note_kernel(REC == 2 ? "k_sweep_fast<prod>" : (REC == 1 ? "k_sweep_fast<rec>" : (NC > 0 || NT > 1 ? "k_sweep_fast<taylor>" : "k_sweep_fast")));
    return check_cuda(cudaGetLastError(), "k_sweep_fast launch");
}

// Taylor-extrapolated sweeps (fhmc_fast_taylor.cu)
int launch_fast_taylor(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream);
// pure mu sweeps with the exp recurrence (fhmc_fast_rec.cu)
int launch_fast_mu_rec(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream);
// pure mu sweeps, product form of the recurrence (fhmc_fast_prod.cu)
int launch_fast_mu_prod(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream);

}  // namespace fhmc
