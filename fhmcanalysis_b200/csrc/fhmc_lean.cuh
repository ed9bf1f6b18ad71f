// fhmc_lean.cuh -- one state point evaluated by ONE WARP, lean form (K1+K3+K2 for the batched coexistence solver, K4, and for
// warp-per-point sweeps).
//
// Same arithmetic, bit for bit, as PointEval<32, TAYLOR> (fhmc_point.cuh): u_i = fl(lnPI_i + fl(s N_i)) then one fma per
// Taylor term in descriptor order; comparisons that decide an index are made on u and re-tested on fl(u - c); every
// irregular case (no maxima / no minima with ties, phases that do not tile [0, n), a phase whose weight underflows, more
// candidates than the scratch holds) returns false and the caller re-runs the state point with PointEval.  What differs
// is how the work is laid out (r01b profile of the group evaluator: 150 K warp instructions per solve, 7 % of them fp64;
// IMAD/BRA/ISETP/LD of descriptor-driven row addressing and a loop over FHMC_MAX_TERMS with a break per term):
//   * Taylor / quantity term counts are template parameters, row byte offsets live in registers, rows are read with
//     LDS from the staged blob (not through a generic pointer);
//   * a lane owns a CONTIGUOUS chunk of C bins (C odd: conflict-free shared-memory banks), so the neighbours of a bin are
//     the lane's own previous values -- no shuffles in the detection pass; strict 1-neighbour extrema are recorded in a
//     64-bit mask per lane and laid out in bin order by one warp scan afterwards;
//   * the +-smooth window test runs one lane per candidate for the first shifts (noise candidates die there), then one
//     lane per shift for the survivors;
//   * phase sums: one pass, a lane hands its partial sums to a per-warp slot when it crosses a phase boundary (exactly one
//     lane crosses each boundary: no atomics, deterministic), a butterfly per phase closes it.
#pragma once
#include "fhmc_point.cuh"

#define FHMC_LEAN_CAND 160  // candidate 1-neighbour extrema per evaluation held in scratch (more: PointEval fallback)
#define FHMC_LEAN_PMAX 8    // phases per state point the lean path handles (more: PointEval fallback)
#define FHMC_LEAN_D1 4      // window shifts 2..D1 are tested by the candidate's own lane

namespace fhmc {

// diagnostic counters (fhmc_lean_stats): [0] evaluations finished by the lean path, then the reasons for handing a state point
// to PointEval: [1] too many candidates, [2] repair needs the normalised array / reference raises / phases do not tile,
// [3] phase count, [4] a phase underflowed, [5] re-test on fl(u - c) failed, [6] monotone ln(PI) with tied arg max / arg min,
// [7] monotone ln(PI) whose repair is not the single phase [0, n)
static __device__ unsigned long long g_lean_stats[8];   // (one copy per translation unit: no relocatable device code)
#define FHMC_LEAN_STAT(k) do { if (lane == 0) atomicAdd(&g_lean_stats[k], 1ull); } while (0)
// -DFHMC_LEAN_PROFILE: cycles per section of run() accumulated in g_lean_prof (sections: 0 pass A, 1 candidate list, 2 window
// shifts 2..D1, 3 remaining shifts, 4 repair, 5 pass B, 6 phase reductions, 7 re-test + is_safe); read with fhmc_lean_stats(out, .)
// through out[8..15]
#ifdef FHMC_LEAN_PROFILE
static __device__ unsigned long long g_lean_prof[16];
#define FHMC_LEAN_T0 long long t_prof = clock64();
#define FHMC_LEAN_TICK(k) do { const long long t_now = clock64(); if (lane == 0) atomicAdd(&g_lean_prof[k], (unsigned long long)(t_now - t_prof)); t_prof = t_now; } while (0)
#else
#define FHMC_LEAN_T0
#define FHMC_LEAN_TICK(k)
#endif

struct LeanScratch {  // per warp, shared memory
    double slot[FHMC_LEAN_PMAX][1 + FHMC_MAX_SEL];
    // the record of the evaluation in flight (same fields as fhmc_sweep_out; copied to the caller's arrays by commit()):
    // an iterative caller (K4) evaluates several state points into the same record and only the last one is kept, and
    // every list access in between stays in shared memory instead of making a round trip to L2
    double fe[FHMC_LEAN_PMAX], avg[FHMC_LEAN_PMAX * FHMC_MAX_SEL], lnnorm;
    int maxl[FHMC_LEAN_PMAX + 1], minl[FHMC_LEAN_PMAX + 2], bl[2 * FHMC_LEAN_PMAX];
    unsigned status;
    int nphase, nmin;
    int cand[FHMC_LEAN_CAND];
};

template <int NC, int NSEL, int NT>
struct LeanEval {
    static constexpr int NCA = NC > 0 ? NC : 1, NSA = NSEL > 0 ? NSEL : 1;
    const SweepArgs &a;
    LeanScratch *const ws;
    const int lane;
    const uint32_t sb, tab;  // shared-memory byte address of blob row 0 / of the 2^(j/64) table
    const int n, last, w, pmax;
    const uint32_t npad8;
    int C, lo, hi;
    uint32_t crow8[NCA], srow8[NSA];
    double s, xi[NCA], ts[NT];
    int P, nmin;
    double m, c;

    __device__ LeanEval(const SweepArgs &args, const double *blob_smem, LeanScratch *scratch, int lane_, const double *exp_table)
        : a(args), ws(scratch), lane(lane_), sb(smem_u32(blob_smem)), tab(smem_u32(exp_table)), n(args.d.n), last(args.d.n - 1),
          w(args.d.smooth), pmax(args.d.pmax), npad8((uint32_t)args.d.n_pad * 8u)
    {
        C = ((n + 31) / 32) | 1;
        lo = lane * C;
        hi = min(n, lo + C);
#pragma unroll
        for (int t = 0; t < NC; ++t) crow8[t] = (uint32_t)args.d.coef_row[t] * npad8;
#pragma unroll
        for (int q = 0; q < NSEL; ++q) srow8[q] = (uint32_t)args.d.sel_row[q] * npad8;
    }

    // what the lean path can take on at all (checked once per launch on the host as well)
    __device__ __forceinline__ bool usable() const { return C <= 64 && pmax <= FHMC_LEAN_PMAX && n >= 3 && !a.d.complete; }

    __device__ __forceinline__ void setup(double mu1, double beta, double dmu)
    {
        s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);  // GH:77, evaluated left to right
        const double dB = beta - a.d.beta_ref, dD = dmu - a.d.dmu_ref;
#pragma unroll
        for (int t = 0; t < NC; ++t) xi[t] = monomial(a.d.coef_kind[t], dB, dD, mu1);
#pragma unroll
        for (int t = 0; t < NT; ++t) ts[t] = (t == 0) ? 1.0 : monomial(a.d.sel_kind[t], dB, dD, mu1);
    }

    // u_i, in the evaluation order of PointEval::U
    __device__ __forceinline__ double U(int i) const
    {
        const uint32_t ai = sb + 8u * (uint32_t)i;
        double u = __dadd_rn(lds_f64(ai), __dmul_rn(s, lds_f64(ai + npad8)));
#pragma unroll
        for (int t = 0; t < NC; ++t) u = fma(xi[t], lds_f64(ai + crow8[t]), u);
        return u;
    }
    __device__ __forceinline__ double X(int q, int i) const
    {
        const uint32_t ai = sb + 8u * (uint32_t)i + srow8[q];
        double x = lds_f64(ai);
#pragma unroll
        for (int t = 1; t < NT; ++t) x = fma(ts[t], lds_f64(ai + (uint32_t)t * npad8), x);
        return x;
    }
    __device__ __forceinline__ double wsum(double v) const { return group_sum<32>(v, 0xffffffffu); }

    // No windowed extremum at all (monotone ln(PI): far from coexistence, e.g. while the solver looks for the two-phase
    // window).  The reference takes the bins tied with the maximum / minimum of the NORMALISED array (GH:382-386), so c comes
    // first (one exp pass over [0, n), which is also the single phase's sum), then one scan on fl(u - c).  Unique arg max and
    // arg min only; ties go to PointEval.  Mirrors the slow path of PointEval::run (status bit SLOW_PATH).
    __device__ __noinline__ bool run_monotone(unsigned &status_out)
    {
        int *const maxl = ws->maxl, *const minl = ws->minl, *const bl = ws->bl;
        double S = 0.0, A[NSA], mn = CUDART_INF;
#pragma unroll
        for (int q = 0; q < NSEL; ++q) A[q] = 0.0;
        for (int i = lo; i < hi; ++i) {
            const double u = U(i);
            mn = fmin(mn, u);
            const double e = exp_nonpos(u - m, tab);
            S += e;
#pragma unroll
            for (int q = 0; q < NSEL; ++q) A[q] = fma(e, X(q, i), A[q]);
        }
        const double Sg = wsum(S), umin = group_min<32>(mn, 0xffffffffu);
        double Ag[NSA];
#pragma unroll
        for (int q = 0; q < NSEL; ++q) Ag[q] = wsum(A[q]);
        if (Sg < 1e-280) { FHMC_LEAN_STAT(4); return false; }
        c = m + log(Sg);
        const double vmax = __dsub_rn(m, c), vmin = __dsub_rn(umin, c);
        int cM = 0, cm = 0, pM = 0x7fffffff, pm = 0x7fffffff;
        for (int i = lo; i < hi; ++i) {
            const double v = __dsub_rn(U(i), c);
            if (v == vmax) { ++cM; pM = min(pM, i); }
            if (v == vmin) { ++cm; pm = min(pm, i); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cM += __shfl_xor_sync(0xffffffffu, cM, o);
            cm += __shfl_xor_sync(0xffffffffu, cm, o);
            pM = min(pM, __shfl_xor_sync(0xffffffffu, pM, o));
            pm = min(pm, __shfl_xor_sync(0xffffffffu, pm, o));
        }
        if (cM != 1 || cm != 1) { FHMC_LEAN_STAT(6); return false; }
        int packed = 0, nm_l = 0;
        unsigned flags = 0;
        if (lane == 0) {
            PointEval<1, false> pe(a, a.blob, 0, nullptr);   // with a unique arg max / arg min repair() reads no array value
            int nM, nm;
            bool part;
            const int rc = pe.repair(true, c, 0, 0, m, umin, maxl, minl, bl, nM, nm, flags, part, 1, 1, pM, pm);
            packed = rc | (part ? 0x100 : 0) | (nM << 9);
            nm_l = nm;
        }
        __syncwarp();
        packed = __shfl_sync(0xffffffffu, packed, 0);
        nmin = __shfl_sync(0xffffffffu, nm_l, 0);
        P = packed >> 9;
        if ((packed & 0xff) != FHMC_OK || !(packed & 0x100) || P != 1 || bl[0] != 0 || bl[1] != n) { FHMC_LEAN_STAT(7); return false; }
        const double u0 = U(0);
        flags = FHMC_ST_SLOW_PATH | FHMC_ST_LEAN;
        const double xM = __dsub_rn(U(maxl[0]), c), xl = __dsub_rn(U(last), c);
        if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
        status_out = flags;
        FHMC_LEAN_STAT(0);
        if (lane == 0) {
            ws->fe[0] = -(c - u0);
#pragma unroll
            for (int q = 0; q < NSEL; ++q) ws->avg[q] = Ag[q] / Sg;
            ws->status = flags;
            ws->nphase = 1;
            ws->nmin = nmin;
            ws->lnnorm = c;
        }
        __syncwarp();
        return true;
    }

    // copy the record of the last evaluation from the scratch to record `rec` of the caller's arrays (all lanes)
    __device__ void commit(long long rec) const
    {
        const int Pn = ws->nphase, nm = ws->nmin;
        for (int k = lane; k < Pn; k += 32) {
            a.out.fe[rec * pmax + k] = ws->fe[k];
            a.out.max_idx[rec * pmax + k] = ws->maxl[k];
            a.out.bounds[(rec * pmax + k) * 2] = ws->bl[2 * k];
            a.out.bounds[(rec * pmax + k) * 2 + 1] = ws->bl[2 * k + 1];
        }
        for (int k = lane; k < Pn * NSEL; k += 32) a.out.avg[rec * pmax * NSEL + k] = ws->avg[k];
        for (int k = lane; k < nm; k += 32) a.out.min_idx[rec * (pmax + 1) + k] = ws->minl[k];
        if (lane == 0) {
            a.out.status[rec] = ws->status;
            a.out.nphase[rec] = Pn;
            a.out.nmin[rec] = nm;
            a.out.lnnorm[rec] = ws->lnnorm;
        }
        __syncwarp();
    }

    // One state point -> the record in the scratch (commit() copies it out).  Returns false when PointEval must take it.
    __device__ bool run(unsigned &status_out)
    {
        int *const maxl = ws->maxl, *const minl = ws->minl, *const bl = ws->bl;
        FHMC_LEAN_T0
        // ---- pass A: max of u and the candidate extrema of this lane's chunk (bit j of the mask = bin lo + j).  A bin is a
        // candidate when the successive differences d = u_{i+1} - u_i and d' = u_i - u_{i-1} differ in sign bit: every strict
        // 1-neighbour extremum is one (the sign of a rounded difference is exact); the few extra ones (a difference of exactly
        // zero) are thrown out by the exact comparisons of the window test below.  The maximum of u sits at an end point or at
        // a candidate (first bin after a plateau included), so only candidates touch the running maximum.
        unsigned mlo = 0u, mhi = 0u;
        double mx = -CUDART_INF;
        if (lo < n) {
            double uc = U(lo);
            int hprev = 0;                       // bin 0: "rising" on its left, so that a falling start offers u_0 to the maximum
            if (lo > 0) hprev = __double2hiint(__dsub_rn(uc, U(lo - 1)));
            const int jend = min(hi, last) - lo;   // bins lo + j < last have a right neighbour
            auto bin = [&](int j, unsigned &mask, unsigned bit) {
                const double up = U(lo + j + 1);
                const int h = __double2hiint(__dsub_rn(up, uc));
                if ((h ^ hprev) < 0) {
                    mask |= bit;
                    mx = fmax(mx, uc);
                }
                hprev = h;
                uc = up;
            };
            const int j32 = min(jend, 32);
            unsigned bit = 1u;
            int j = 0;
            for (; j < j32; ++j, bit <<= 1) bin(j, mlo, bit);
            bit = 1u;
            for (; j < jend; ++j, bit <<= 1) bin(j, mhi, bit);
            if (hi == n) mx = fmax(mx, uc);      // uc = u_last here
            if (lo == 0) mlo &= ~1u;             // bin 0 has no left neighbour: never an extremum (argrelextrema, mode 'clip')
        }
        m = group_max<32>(mx, 0xffffffffu);
        FHMC_LEAN_TICK(0);
        // candidates in bin order: exclusive scan of the per-lane counts
        const int cnt_l = __popc(mlo) + __popc(mhi);
        int incl = cnt_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total > FHMC_LEAN_CAND) { FHMC_LEAN_STAT(1); return false; }
        {
            int off = incl - cnt_l;
            while (mlo) {
                const int j = __ffs((int)mlo) - 1;
                mlo &= mlo - 1u;
                ws->cand[off++] = lo + j;
            }
            while (mhi) {
                const int j = __ffs((int)mhi) - 1;
                mhi &= mhi - 1u;
                ws->cand[off++] = lo + 32 + j;
            }
        }
        __syncwarp();
        FHMC_LEAN_TICK(1);
        // ---- window test, shifts 2..D1: one lane per candidate; survivors compacted in place (order kept) ---------------------
        int nsurv = 0;
        const unsigned below = (1u << lane) - 1u;
        const int d1 = min(w, FHMC_LEAN_D1);
        for (int base = 0; base < total; base += 32) {
            const int k = base + lane;
            bool ok = false;
            int code = 0;
            if (k < total) {
                const int i = ws->cand[k];
                const double xc = U(i), xm = U(i - 1), xp = U(i + 1);
                const bool is_max = xc > xm;
                code = (i << 1) | (is_max ? 1 : 0);
                ok = is_max ? (xc > xp) : (xc < xm && xc < xp);   // strict 1-neighbour extremum (exact comparisons)
                for (int d = 2; d <= d1 && ok; ++d) {
                    const double xl = U(max(i - d, 0)), xr = U(min(i + d, last));
                    ok = is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr);
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, ok);
            __syncwarp();
            if (ok) ws->cand[nsurv + __popc(b & below)] = code;
            nsurv += __popc(b);
            __syncwarp();
        }
        FHMC_LEAN_TICK(2);
        // ---- shifts D1+1..w of the survivors: one lane per shift; confirmed extrema go to the lists at position 1 + k ---------
        int cntM = 0, cntm = 0;
        for (int k = 0; k < nsurv; ++k) {
            const int code = ws->cand[k];
            const int i = code >> 1;
            const bool is_max = code & 1;
            bool ok = true;
            if (w > d1) {
                const double xc = U(i);
                for (int d = d1 + 1 + lane; d <= w; d += 32) {
                    const double xl = U(max(i - d, 0)), xr = U(min(i + d, last));
                    ok = ok && (is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr));
                }
                ok = __all_sync(0xffffffffu, ok);
            }
            if (ok) {
                if (is_max) {
                    if (lane == 0 && 1 + cntM <= pmax - 1) maxl[1 + cntM] = i;
                    ++cntM;
                } else {
                    if (lane == 0 && 1 + cntm <= pmax) minl[1 + cntm] = i;
                    ++cntm;
                }
            }
        }
        FHMC_LEAN_TICK(3);
        // ---- repair / validation / bounds (leader lane; GH:333-415, 498-520) ----------------------------------------------------
        int packed = 0, nm_l = 0;
        unsigned flags = 0;
        if (lane == 0) {
            PointEval<1, false> pe(a, a.blob, 0, nullptr);   // repair() without c touches the lists only
            int nM, nm;
            bool part;
            const int rc = pe.repair(false, 0.0, cntM, cntm, 0.0, 0.0, maxl, minl, bl, nM, nm, flags, part);
            packed = (rc == FHMC_NEED_SLOW) ? -1 : (rc | (part ? 0x100 : 0) | (nM << 9));
            nm_l = nm;
        }
        __syncwarp();
        packed = __shfl_sync(0xffffffffu, packed, 0);
        nmin = __shfl_sync(0xffffffffu, nm_l, 0);
        if (packed < 0 && cntM == 0 && cntm == 0) return run_monotone(status_out);
        if (packed < 0 || (packed & 0xff) != FHMC_OK || !(packed & 0x100)) { FHMC_LEAN_STAT(2); return false; }
        P = packed >> 9;
        if (P < 1 || P > FHMC_LEAN_PMAX) { FHMC_LEAN_STAT(3); return false; }
        FHMC_LEAN_TICK(4);
        // ---- pass B: per-phase sums about the common shift m ------------------------------------------------------------------
        for (int k = lane; k < P * (1 + FHMC_MAX_SEL); k += 32) (&ws->slot[0][0])[k] = 0.0;
        __syncwarp();
        double S = 0.0, A[NSA];
#pragma unroll
        for (int q = 0; q < NSEL; ++q) A[q] = 0.0;
        int cur = 0;
        if (lo < n) {
            while (cur < P - 1 && bl[2 * cur + 1] <= lo) ++cur;
            int i = lo;
            while (i < hi) {
                const int nb = (cur < P - 1) ? bl[2 * cur + 1] : 0x7fffffff;   // right bound of the phase bin i is in
                const int end = min(hi, nb);
                double S1 = 0.0, A1[NSA];
#pragma unroll
                for (int q = 0; q < NSEL; ++q) A1[q] = 0.0;
                for (; i + 3 < end; i += 4) {   // four independent exp chains in flight (fp64 latency ~8 cycles, 4 warps per scheduler)
                    const double e0 = exp_nonpos(U(i) - m, tab), e1 = exp_nonpos(U(i + 1) - m, tab);
                    const double e2 = exp_nonpos(U(i + 2) - m, tab), e3 = exp_nonpos(U(i + 3) - m, tab);
                    S += e0;
                    S1 += e1;
                    S += e2;
                    S1 += e3;
#pragma unroll
                    for (int q = 0; q < NSEL; ++q) {
                        A[q] = fma(e0, X(q, i), A[q]);
                        A1[q] = fma(e1, X(q, i + 1), A1[q]);
                        A[q] = fma(e2, X(q, i + 2), A[q]);
                        A1[q] = fma(e3, X(q, i + 3), A1[q]);
                    }
                }
                for (; i + 1 < end; i += 2) {
                    const double e0 = exp_nonpos(U(i) - m, tab), e1 = exp_nonpos(U(i + 1) - m, tab);
                    S += e0;
                    S1 += e1;
#pragma unroll
                    for (int q = 0; q < NSEL; ++q) {
                        A[q] = fma(e0, X(q, i), A[q]);
                        A1[q] = fma(e1, X(q, i + 1), A1[q]);
                    }
                }
                if (i < end) {
                    const double e0 = exp_nonpos(U(i) - m, tab);
                    S += e0;
#pragma unroll
                    for (int q = 0; q < NSEL; ++q) A[q] = fma(e0, X(q, i), A[q]);
                    ++i;
                }
                S += S1;
#pragma unroll
                for (int q = 0; q < NSEL; ++q) A[q] += A1[q];
                if (i < hi) {   // i == nb: this lane holds the boundary -- hand the closed phase's partial sums over
                    ws->slot[cur][0] = S;
#pragma unroll
                    for (int q = 0; q < NSEL; ++q) ws->slot[cur][1 + q] = A[q];
                    S = 0.0;
#pragma unroll
                    for (int q = 0; q < NSEL; ++q) A[q] = 0.0;
                    ++cur;
                }
            }
        } else {
            cur = -1;
        }
        __syncwarp();
        FHMC_LEAN_TICK(5);
        const double u0 = U(0);
        double Stot = 0.0;
        bool under = false;
        for (int p = 0; p < P; ++p) {
            const bool mine = (cur == p);
            const double Sg = wsum(mine ? S : 0.0) + ws->slot[p][0];
            double Ag[NSA];
#pragma unroll
            for (int q = 0; q < NSEL; ++q) Ag[q] = wsum(mine ? A[q] : 0.0) + ws->slot[p][1 + q];
            if (Sg < 1e-280) under = true;
            if (lane == 0) {
                ws->fe[p] = -((m + log(Sg)) - u0);
#pragma unroll
                for (int q = 0; q < NSEL; ++q) ws->avg[p * NSEL + q] = Ag[q] / Sg;
            }
            Stot += Sg;
        }
        if (under) { FHMC_LEAN_STAT(4); return false; }   // a phase too unlikely for the common shift: PointEval integrates it about its own maximum
        c = m + log(Stot);
        FHMC_LEAN_TICK(6);
        // ---- re-test the interior extrema on the normalised values fl(u - c), one lane per (extremum, shift) -----------------
        if (!a.d.compare_raw) {
            bool bad = false;
            const int ne = P + nmin;
            for (int t = lane; t < ne * w; t += 32) {
                const int k = t / w, d = t % w + 1;
                const bool is_max = k < P;
                const int idx = is_max ? maxl[k] : minl[k - P];
                if (idx > 0 && idx < last) {
                    const double xc = __dsub_rn(U(idx), c);
                    const double xl = __dsub_rn(U(max(idx - d, 0)), c), xr = __dsub_rn(U(min(idx + d, last)), c);
                    if (!(is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr))) bad = true;
                }
            }
            if (__any_sync(0xffffffffu, bad)) { FHMC_LEAN_STAT(5); return false; }
        }
        const double xM = __dsub_rn(U(maxl[P - 1]), c), xl = __dsub_rn(U(last), c);
        if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
        flags = __shfl_sync(0xffffffffu, flags, 0) | (flags & FHMC_ST_SAFE) | FHMC_ST_LEAN;
        status_out = flags;
        FHMC_LEAN_STAT(0);
        FHMC_LEAN_TICK(7);
        if (lane == 0) {
            ws->status = flags;
            ws->nphase = P;
            ws->nmin = nmin;
            ws->lnnorm = c;
        }
        __syncwarp();
        return true;
    }
};

}  // namespace fhmc
