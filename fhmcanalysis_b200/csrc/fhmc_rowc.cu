// fhmc_rowc.cu -- Taylor-extrapolated GRID sweeps (temp_dmu_extrap_multi-style beta x dmu_2 grids, reference
// gc_hist.pyx:813-1239) with dmu_2 as the fastest axis: "row-combined" form of the one-thread-per-state-point kernel.
//
// Along one grid row (mu_1 and beta fixed) the extrapolated histogram is a quadratic in dD = dmu_2 - dmu_2,ref:
//     lnPI'(N) = L(N) + dD * C1(N) + (dD^2 / 2) * C2(N)
//     L  = fl(lnPI + fl(s N)) + sum over the terms without dD  (dB mu_1 N, dB A_b, dB^2/2 A_bb, ...)
//     C1 = A_d + dB * A_bd,     C2 = A_dd
// so a CTA combines the coefficient rows ONCE per row of the grid (n bins x <= 8 terms, against n_dmu x n bins of
// walking) and a bin then costs 2 fp64 instructions for u instead of 8 and three 8-byte broadcast loads instead of four
// 16-byte ones.  Everything else is the walk of fhmc_fast.cuh (one exp pass, sign-of-difference prefilter, exact windowed
// tests, repair(), re-test on the normalised values) with three changes that the profile of the Taylor kernel asked for:
//   * two state points per thread: every row load serves both, their independent exp chains interleave;
//   * a lane-replicated 2^(j/64) table (16 copies, lane l reads copy l mod 16): the per-lane table look-up of exp is
//     conflict-free (it cost ~6 shared-memory wavefronts per warp and bin, as much as the row loads);
//   * no CTA-wide fallback queue: a state point that is not a plain case is re-run on the spot by its whole warp with the
//     general evaluator -- on the SAME combined rows, so both paths see bit-identical u -- and nothing synchronises the
//     warps of a CTA between two row builds.
// The exponent shift comes from a subsampled maximum (every 8th bin); a term may exceed it by up to 2^900, a larger
// miss saturates (the exponent offset is clamped) and sends the state point to the general evaluator.
#include "fhmc_fast.cuh"

namespace fhmc {

struct RowcPlan {
    long long n_run;      // state points per run: (mu_1, beta) are constant inside a run, the dmu index is the position in it
    long long n_runs;
    int chunk;            // state points per work item (a multiple of 512)
    int chunks_per_run;
};

struct RowcCtx {
    uint32_t sL;          // shared-memory address of row L; C2 at +oC2, C1 at +oC1
    uint32_t oC1, oC2;
    uint32_t tabL;        // this lane's copy of the 2^(j/64) table: entry j at tabL + (j << 7)
    const double *rows;   // the same rows / the plain table for the general evaluator
    const double *tab64;
};

// exp_scaled_r() with the lane-replicated table and an upper clamp of the exponent offset (see the file header)
__device__ __forceinline__ double exp_rowc(double u, int Mq, uint32_t tabL, const ExpRegs &c)
{
    const double kd0 = fma(u, c.inv, FHMC_EXP_MAGIC);
    const int k = __double2loint(kd0);
    const double kd = kd0 - FHMC_EXP_MAGIC;
    double r = fma(kd, c.nhi, u);
    r = fma(kd, c.nlo, r);
    double p = fma(r, c.c5, c.c4);
    p = fma(p, r, c.c3);
    p = fma(p, r, c.c2);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double T = lds_f64(tabL + ((k & 63) << 7));
    const double v = T * p;
    const int q = min(max((k >> 6) - Mq, -1022), 1000);
    const int hi = __double2hiint(v) + (q << 20);
    return __hiloint2double(hi, __double2loint(v));
}

// per-state-point state of the walk (kept in registers: never pass its address to a non-inlined function)
struct RowcPt {
    double dD, q2, Sacc, Stot, u0, xm, uc, dc;
    long long sp;
    int Mq, P, cntM, cntm;
    unsigned rescue;
    bool bad;
};

template <bool HC2>
__device__ __forceinline__ double rowc_u(const RowcCtx &cx, int i, double dD, double q2)
{
    const uint32_t addr = cx.sL + 8u * (uint32_t)i;
    const double t = fma(dD, lds_f64(addr + cx.oC1), lds_f64(addr));
    return HC2 ? fma(q2, lds_f64(addr + cx.oC2), t) : t;
}

// The general evaluator on the combined rows (all 32 lanes of the calling warp on one state point): row 0 = L with s = 0,
// row 1 = C2 (doubles as the "N" row: fl(0 * C2) = 0), row 2 = C1; u = fma(dD^2/2, C2, fma(dD, C1, L)) as in the walk.
__device__ __noinline__ void rowc_generic(const SweepArgs &a, const double *rows, const double *tab64, int lane, double dmu,
                                          long long sp, int hc2)
{
    SweepArgs b = a;
    b.d.n_rows = 3;
    b.d.n_coef = hc2 ? 2 : 1;
    b.d.coef_row[0] = 2;
    b.d.coef_kind[0] = FHMC_M_DD;
    b.d.coef_row[1] = 1;
    b.d.coef_kind[1] = FHMC_M_DD2;
    b.d.n_term = 1;
    PointEval<32, true> pe(b, rows, lane, tab64);
    pe.setup(a.d.mu1_ref, a.d.beta_ref, dmu);
    pe.run(sp);
}

// Validation and the record of one walked state point (exact rules of the general path: repair(), re-test of the extrema
// on fl(u - c), rescue of underflowed phases, is_safe).  false: not a plain case, re-run it with the general evaluator.
template <bool HC2>
__device__ __noinline__ bool rowc_finish(const SweepArgs &a, const RowcCtx cx, int lane, long long sp, double dD, double q2, int Mq,
                                         int P, int cntM, int cntm, unsigned rescue, double Stot, double u0)
{
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax, w = a.d.smooth;
    PointEval<1, true> pe(a, a.blob, lane, cx.tab64);   // repair() only: it reads no histogram value on these paths
    int *maxl = a.out.max_idx + sp * pmax;
    int *minl = a.out.min_idx + sp * (pmax + 1);
    int *bl = a.out.bounds + sp * pmax * 2;
    if (!(Stot < 0x1p900) || !(Stot > 0.0)) return false;   // the subsampled shift missed the maximum by too much
    const double c = add_shift(Mq, log(Stot));
    auto window_c = [&](int i, double xc, bool is_max) {   // all shifts 1..w on the normalised values
        for (int d = 1; d <= w; ++d) {
            const int jl = (i - d < 0) ? 0 : i - d;
            const int jr = (i + d > last) ? last : i + d;
            const double xl = __dsub_rn(rowc_u<HC2>(cx, jl, dD, q2), c), xr = __dsub_rn(rowc_u<HC2>(cx, jr, dD, q2), c);
            if (!(is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr))) return false;
        }
        return true;
    };
    int nM = 0, nm = 0, rc;
    unsigned flags = 0;
    bool part = false;
    if (cntM == 0 && cntm == 0) {
        // no windowed extremum at all (monotone ln(PI) far from coexistence): the reference takes the bins tied with the
        // max / min of the NORMALISED array (GH:382-386); genuine ties go to the general evaluator
        double vM = -CUDART_INF, vm = CUDART_INF;
        int cM = 0, cm = 0, pM = 0, pm = 0;
        for (int j = 0; j < n; ++j) {
            const double v = __dsub_rn(rowc_u<HC2>(cx, j, dD, q2), c);
            if (v > vM) { vM = v; cM = 1; pM = j; } else if (v == vM) ++cM;
            if (v < vm) { vm = v; cm = 1; pm = j; } else if (v == vm) ++cm;
        }
        rc = (cM == 1 && cm == 1) ? pe.repair(true, c, 0, 0, 0.0, 0.0, maxl, minl, bl, nM, nm, flags, part, 1, 1, pM, pm)
                                  : FHMC_NEED_SLOW;
    } else {
        rc = pe.repair(false, 0.0, cntM, cntm, 0.0, 0.0, maxl, minl, bl, nM, nm, flags, part);
    }
    if (!(rc == FHMC_OK && part && nM == P)) return false;
    if (!a.d.compare_raw) {
        for (int k = 0; k < nM + nm; ++k) {
            const bool is_max = k < nM;
            const int idx = is_max ? maxl[k] : minl[k - nM];
            if (idx > 0 && idx < last && !window_c(idx, __dsub_rn(rowc_u<HC2>(cx, idx, dD, q2), c), is_max)) return false;
        }
    }
    // phases whose weight underflowed next to the global maximum: integrate them about their own maximum
    for (int p = 0; rescue != 0 && p < nM; ++p) {
        if (!((rescue >> p) & 1u)) continue;
        const int left = bl[2 * p], right = bl[2 * p + 1];
        double ml = -CUDART_INF;
        for (int j = left; j < right; ++j) ml = fmax(ml, rowc_u<HC2>(cx, j, dD, q2));
        const int Mp = shift_for_max(ml);
        double Sp = 0.0;
        for (int j = left; j < right; ++j) Sp += exp_scaled(rowc_u<HC2>(cx, j, dD, q2), Mp, smem_u32(cx.tab64));
        a.out.fe[sp * pmax + p] = -(add_shift(Mp, log(Sp)) - u0);
        flags |= FHMC_ST_RESCUED;
    }
    const double xM = __dsub_rn(rowc_u<HC2>(cx, maxl[nM - 1], dD, q2), c), xl = __dsub_rn(rowc_u<HC2>(cx, last, dD, q2), c);
    if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
    a.out.status[sp] = flags | FHMC_ST_FAST;
    a.out.nphase[sp] = nM;
    a.out.nmin[sp] = nm;
    a.out.lnnorm[sp] = c;
    return true;
}

// The walk of one (TWO = false) or two state points of the calling thread over the combined rows.
template <bool HC2, bool TWO>
__device__ __forceinline__ void rowc_walk(const SweepArgs &a, const RowcCtx &cx, const ExpRegs &ec, RowcPt &A, RowcPt &B)
{
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax, w = a.d.smooth;
    const uint32_t sL = cx.sL, oC1 = cx.oC1, oC2 = cx.oC2, tabL = cx.tabL;
    auto uof = [&](const RowcPt &p, double L, double c1, double c2) {
        const double t = fma(p.dD, c1, L);
        return HC2 ? fma(p.q2, c2, t) : t;
    };
    auto load_u = [&](const RowcPt &p, int i) { return rowc_u<HC2>(cx, i, p.dD, p.q2); };
    // ---- shift: maximum over every 8th bin (+ the last one) ---------------------------------------------------------
    {
        double mA = -CUDART_INF, mB = -CUDART_INF;
        for (int i = 0; i < n; i += 8) {
            const uint32_t addr = sL + 8u * (uint32_t)i;
            const double L = lds_f64(addr), c1 = lds_f64(addr + oC1), c2 = HC2 ? lds_f64(addr + oC2) : 0.0;
            mA = fmax(mA, uof(A, L, c1, c2));
            if (TWO) mB = fmax(mB, uof(B, L, c1, c2));
        }
        mA = fmax(mA, load_u(A, last));
        A.Mq = shift_for_max(mA);
        if (TWO) {
            mB = fmax(mB, load_u(B, last));
            B.Mq = shift_for_max(mB);
        }
    }
    auto flush = [&](RowcPt &p) {
        if (p.P < pmax && p.Sacc >= 1e-280) a.out.fe[p.sp * pmax + p.P] = -(add_shift(p.Mq, log(p.Sacc)) - p.u0);
        else if (p.P < pmax && p.P < 32) p.rescue |= 1u << p.P;   // re-integrated about its own maximum in rowc_finish()
        else p.bad = true;
        p.Stot += p.Sacc;
        p.Sacc = 0.0;
        ++p.P;
    };
    auto window = [&](const RowcPt &p, int i, double xc, bool is_max) {   // shifts 2..w (shift 1 already passed)
        for (int d = 2; d <= w; ++d) {
            const int jl = (i - d < 0) ? 0 : i - d;
            const int jr = (i + d > last) ? last : i + d;
            const double xl = load_u(p, jl), xr = load_u(p, jr);
            if (!(is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr))) return false;
        }
        return true;
    };
    auto test_bin = [&](RowcPt &p, int i, double xm, double xc, double xp) {
        const bool is_max = (xc > xm) && (xc > xp), is_min = (xc < xm) && (xc < xp);
        if ((is_max || is_min) && window(p, i, xc, is_max)) {
            if (is_max) {
                if (1 + p.cntM <= pmax - 1) a.out.max_idx[p.sp * pmax + 1 + p.cntM] = i;
                ++p.cntM;
            } else {
                if (1 + p.cntm <= pmax) a.out.min_idx[p.sp * (pmax + 1) + 1 + p.cntm] = i;
                ++p.cntm;
                flush(p);   // a minimum bin opens the phase to its right (GH:498-520)
            }
        }
    };
    auto ex = [&](const RowcPt &p, double u) { return exp_rowc(u, p.Mq, tabL, ec); };
    // one block of four bins i..i+3 of one state point: u of the bins i+1..i+4 given, p.xm / p.uc / p.dc carried
    auto fast_block = [&](RowcPt &p, double u1, double u2, double u3, double u4, double d4) {
        const double e0 = ex(p, p.uc), e1 = ex(p, u1), e2 = ex(p, u2), e3 = ex(p, u3);
        p.Sacc += (e0 + e1) + (e2 + e3);
        p.xm = u3;
        p.uc = u4;
        p.dc = d4;
    };
    auto slow_block = [&](RowcPt &p, int i, double u1, double u2, double u3, double u4, double d4) {
        test_bin(p, i, p.xm, p.uc, u1);
        p.Sacc += ex(p, p.uc);
        test_bin(p, i + 1, p.uc, u1, u2);
        p.Sacc += ex(p, u1);
        test_bin(p, i + 2, u1, u2, u3);
        p.Sacc += ex(p, u2);
        test_bin(p, i + 3, u2, u3, u4);
        p.Sacc += ex(p, u3);
        p.xm = u3;
        p.uc = u4;
        p.dc = d4;
    };
    auto start = [&](RowcPt &p) {
        p.u0 = load_u(p, 0);
        p.Sacc = ex(p, p.u0);
        p.uc = load_u(p, 1);
        p.xm = p.u0;
        p.dc = __dsub_rn(p.uc, p.xm);   // sign(dc) is the exact order of (xm, uc)
    };
    if (n < 3) {
        A.bad = true;
        if (TWO) B.bad = true;
        return;
    }
    start(A);
    if (TWO) start(B);
    int i = 1;
    uint32_t addr = sL + 16u;   // bin i + 1
#pragma unroll 1
    for (; i + 3 < last; i += 4, addr += 32u) {   // bins i..i+3 are interior, i+4 <= last exists
        double L[4], c1[4], c2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            L[k] = lds_f64(addr + 8u * k);
            c1[k] = lds_f64(addr + oC1 + 8u * k);
            c2[k] = HC2 ? lds_f64(addr + oC2 + 8u * k) : 0.0;
        }
        double ua[4], ub[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ua[k] = uof(A, L[k], c1[k], c2[k]);
            if (TWO) ub[k] = uof(B, L[k], c1[k], c2[k]);
        }
        const double a1 = __dsub_rn(ua[0], A.uc), a2 = __dsub_rn(ua[1], ua[0]), a3 = __dsub_rn(ua[2], ua[1]), a4 = __dsub_rn(ua[3], ua[2]);
        const int flipA = (__double2hiint(A.dc) ^ __double2hiint(a1)) | (__double2hiint(a1) ^ __double2hiint(a2)) |
                          (__double2hiint(a2) ^ __double2hiint(a3)) | (__double2hiint(a3) ^ __double2hiint(a4));
        if (TWO) {
            const double b1 = __dsub_rn(ub[0], B.uc), b2 = __dsub_rn(ub[1], ub[0]), b3 = __dsub_rn(ub[2], ub[1]), b4 = __dsub_rn(ub[3], ub[2]);
            const int flipB = (__double2hiint(B.dc) ^ __double2hiint(b1)) | (__double2hiint(b1) ^ __double2hiint(b2)) |
                              (__double2hiint(b2) ^ __double2hiint(b3)) | (__double2hiint(b3) ^ __double2hiint(b4));
            if ((flipA | flipB) < 0) {   // some pair of successive differences changes sign: look closely (per point)
                if (flipA < 0) slow_block(A, i, ua[0], ua[1], ua[2], ua[3], a4); else fast_block(A, ua[0], ua[1], ua[2], ua[3], a4);
                if (flipB < 0) slow_block(B, i, ub[0], ub[1], ub[2], ub[3], b4); else fast_block(B, ub[0], ub[1], ub[2], ub[3], b4);
            } else {
                fast_block(A, ua[0], ua[1], ua[2], ua[3], a4);
                fast_block(B, ub[0], ub[1], ub[2], ub[3], b4);
            }
        } else {
            if (flipA < 0) slow_block(A, i, ua[0], ua[1], ua[2], ua[3], a4); else fast_block(A, ua[0], ua[1], ua[2], ua[3], a4);
        }
    }
    auto tail = [&](RowcPt &p) {
        for (int j = i; j < last; ++j) {
            const double un = load_u(p, j + 1);
            test_bin(p, j, p.xm, p.uc, un);
            p.Sacc += ex(p, p.uc);
            p.xm = p.uc;
            p.uc = un;
        }
        p.Sacc += ex(p, p.uc);
        flush(p);
    };
    tail(A);
    if (TWO) tail(B);
}

__device__ __forceinline__ void rowc_init(RowcPt &p, long long sp, double dmu, double dmu_ref)
{
    p.sp = sp;
    p.dD = dmu - dmu_ref;
    p.q2 = monomial(FHMC_M_DD2, 0.0, p.dD, 0.0);
    p.Sacc = p.Stot = 0.0;
    p.P = p.cntM = p.cntm = 0;
    p.rescue = 0;
    p.bad = false;
    p.Mq = 0;
    p.u0 = p.xm = p.uc = p.dc = 0.0;
}

template <bool HC2>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_rowc(const __grid_constant__ SweepArgs a, const __grid_constant__ RowcPlan pl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = a.d.n, npad = a.d.n_pad, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *rows = reinterpret_cast<double *>(smem_raw);   // L | C2 | C1
    double *tab64 = rows + 3 * (size_t)npad;
    double *tabR = tab64 + 64;                             // [64][16]: entry j of copy c at (j * 16 + c)
    stage_exp_table(tab64);
    for (int j = threadIdx.x; j < 64 * 16; j += blockDim.x) tabR[j] = c_exp.tab[j >> 4];
    const ExpRegs ec = load_exp_regs();
    RowcCtx cx;
    cx.sL = smem_u32(rows);
    cx.oC2 = (uint32_t)npad * 8u;
    cx.oC1 = (uint32_t)npad * 16u;
    cx.tabL = smem_u32(tabR) + ((uint32_t)(lane & 15) << 3);
    cx.rows = rows;
    cx.tab64 = tab64;
    const long long items = pl.n_runs * pl.chunks_per_run;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const long long run = it / pl.chunks_per_run;
        const int ch = (int)(it % pl.chunks_per_run);
        const long long sp0 = run * pl.n_run;
        const double mu1 = a.st.mu1[(sp0 / a.st.mu1_div) % a.st.n_mu1];
        const double beta = a.st.beta ? a.st.beta[(sp0 / a.st.beta_div) % a.st.n_beta] : a.d.beta_ref;
        __syncthreads();   // every warp is done with the previous rows
        {
            // combine the coefficient rows of this (mu_1, beta): terms in descriptor order, each into the row of its power of dD
            const double s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);   // GH:77, evaluated left to right
            const double dB = beta - a.d.beta_ref;
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                double L = 0.0, C1 = 0.0, C2 = 0.0;
                if (i < n) {
                    L = __dadd_rn(a.blob[i], __dmul_rn(s, a.blob[npad + i]));
                    for (int t = 0; t < a.d.n_coef; ++t) {
                        const int kind = a.d.coef_kind[t];
                        const double v = a.blob[(size_t)a.d.coef_row[t] * npad + i];
                        if (kind == FHMC_M_DD) C1 = fma(1.0, v, C1);
                        else if (kind == FHMC_M_DBDD) C1 = fma(dB, v, C1);
                        else if (kind == FHMC_M_DD2) C2 = fma(1.0, v, C2);
                        else L = fma(monomial(kind, dB, 0.0, mu1), v, L);
                    }
                }
                rows[i] = L;
                rows[npad + i] = C2;
                rows[2 * npad + i] = C1;
            }
        }
        __syncthreads();
        const long long j0 = (long long)ch * pl.chunk, j1 = min(j0 + (long long)pl.chunk, pl.n_run);
        for (long long t = j0 + warp * 64; t < j1; t += (FHMC_CTA / 32) * 64) {
            const long long jA = t + lane, jB = t + 32 + lane;
            bool okA = true, okB = true;
            if (t + 64 <= j1) {
                RowcPt A, B;
                rowc_init(A, sp0 + jA, a.st.dmu[jA], a.d.dmu_ref);
                rowc_init(B, sp0 + jB, a.st.dmu[jB], a.d.dmu_ref);
                rowc_walk<HC2, true>(a, cx, ec, A, B);
                okA = !A.bad && rowc_finish<HC2>(a, cx, lane, A.sp, A.dD, A.q2, A.Mq, A.P, A.cntM, A.cntm, A.rescue, A.Stot, A.u0);
                okB = !B.bad && rowc_finish<HC2>(a, cx, lane, B.sp, B.dD, B.q2, B.Mq, B.P, B.cntM, B.cntm, B.rescue, B.Stot, B.u0);
            } else {   // last, partly filled tile of the run: one state point at a time
                for (int h = 0; h < 2; ++h) {
                    const long long j = h ? jB : jA;
                    if (j < j1) {
                        RowcPt A;
                        rowc_init(A, sp0 + j, a.st.dmu[j], a.d.dmu_ref);
                        rowc_walk<HC2, false>(a, cx, ec, A, A);
                        const bool ok = !A.bad && rowc_finish<HC2>(a, cx, lane, A.sp, A.dD, A.q2, A.Mq, A.P, A.cntM, A.cntm, A.rescue, A.Stot, A.u0);
                        if (h) okB = ok; else okA = ok;
                    }
                }
            }
            // anything unusual: the whole warp re-runs it with the general evaluator, on the combined rows
            __syncwarp();
            unsigned fA = __ballot_sync(0xffffffffu, !okA), fB = __ballot_sync(0xffffffffu, !okB);
            while (fA | fB) {
                const bool second = fA == 0;
                unsigned &f = second ? fB : fA;
                const int src = __ffs(f) - 1;
                f &= f - 1;
                const long long j = t + (second ? 32 : 0) + src;
                rowc_generic(a, rows, tab64, lane, a.st.dmu[j], sp0 + j, HC2 ? 1 : 0);
                __syncwarp();
            }
        }
    }
}

// returns 0 ok, 1 error, -1 "not applicable" (the caller falls back to the flat Taylor kernel)
int launch_rowc(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = args.d;
    const fhmc_states &st = args.st;
    if (d.n_sel != 0 || d.n_coef < 1 || d.complete || d.n < 3) return -1;
    if (!st.dmu || st.dmu_div != 1 || st.n_dmu < 512 || st.n_states % st.n_dmu) return -1;
    const long long n_run = st.n_dmu;
    if (!(st.n_mu1 == 1 || st.mu1_div % n_run == 0)) return -1;
    if (st.beta && !(st.n_beta == 1 || st.beta_div % n_run == 0)) return -1;
    bool has_dd = false, hc2 = false;
    for (int t = 0; t < d.n_coef; ++t) {
        switch (d.coef_kind[t]) {
        case FHMC_M_DD: case FHMC_M_DBDD: has_dd = true; break;
        case FHMC_M_DD2: hc2 = true; break;
        case FHMC_M_DB: case FHMC_M_DB2: case FHMC_M_DB3: case FHMC_M_DB_MU1: case FHMC_M_ONE: break;
        default: return -1;
        }
    }
    if (!has_dd && !hc2) return -1;
    const size_t smem = ((size_t)3 * d.n_pad + 64 + 64 * 16) * 8;
    if (smem > (size_t)smem_optin) return -1;
    auto kern = hc2 ? k_sweep_rowc<true> : k_sweep_rowc<false>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FHMC_CTA, smem), "occupancy query")) return 1;
    if (occ < 1) return -1;
    RowcPlan pl;
    pl.n_run = n_run;
    pl.n_runs = st.n_states / n_run;
    const long long ctas = (long long)sm_count * occ;
    int chunk = 4096;   // state points per work item: as large as leaves >= 8 items per CTA, not below one tile per warp
    while (chunk > 512 && pl.n_runs * ((n_run + chunk - 1) / chunk) < 8 * ctas) chunk >>= 1;
    pl.chunk = chunk;
    pl.chunks_per_run = (int)((n_run + chunk - 1) / chunk);
    const long long items = pl.n_runs * pl.chunks_per_run;
    const long long grid = items < ctas ? items : ctas;
    kern<<<(unsigned)grid, FHMC_CTA, smem, stream>>>(args, pl);
    note_kernel("k_sweep_rowc");
    return check_cuda(cudaGetLastError(), "k_sweep_rowc launch");
}

}  // namespace fhmc
