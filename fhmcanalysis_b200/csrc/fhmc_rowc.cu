// fhmc_rowc.cu -- Taylor-extrapolated GRID sweeps (temp_dmu_extrap_multi-style beta x dmu_2 grids, reference
// gc_hist.pyx:813-1239) with dmu_2 as the fastest axis: "row-combined" form of the one-thread-per-state-point kernel.
//
// Along one grid row (mu_1 and beta fixed) the extrapolated histogram is a quadratic in dD = dmu_2 - dmu_2,ref:
//     lnPI'(N) = L(N) + dD * C1(N) + (dD^2 / 2) * C2(N)
//     L  = fl(lnPI + fl(s N)) + sum over the terms without dD  (dB mu_1 N, dB A_b, dB^2/2 A_bb, ...)
//     C1 = A_d + dB * A_bd,     C2 = A_dd
// so the coefficient rows are combined ONCE per row of the grid (n bins x <= 8 terms, against n_dmu x n bins of walking)
// and a bin then costs 2 fp64 instructions for u instead of 8 and three 8-byte broadcast loads instead of four 16-byte
// ones.  Everything else is the walk of fhmc_fast.cuh (one exp pass, sign-of-difference prefilter, exact windowed tests,
// repair(), re-test on the normalised values) with the changes that the profile of the Taylor kernel asked for:
//   * two state points per thread: every row load serves both, their independent exp chains interleave;
//   * a lane-replicated 2^(j/256) table (16 copies, lane l reads copy l mod 16): the per-lane table look-up of exp is
//     conflict-free (it cost ~6 shared-memory wavefronts per warp and bin, as much as the row loads), and the larger
//     table pays for one polynomial degree (9 fp64-pipe instructions per exp);
//   * no CTA-wide fallback queue: a state point that is not a plain case is re-run on the spot by its whole warp with the
//     general evaluator -- on the SAME combined rows, so both paths see bit-identical u;
//   * no CTA barrier per grid row: every warp combines the rows of its next grid row itself (all warps store identical
//     values) into one of two alternating row buffers, guarded by release counters in shared memory.
// The exponent shift comes from a subsampled maximum (every 8th bin); a term may exceed it by up to 2^900, a larger
// miss saturates (the exponent offset is clamped) and sends the state point to the general evaluator.
#include "fhmc_fast.cuh"
#include "fhmc_exp256.cuh"

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
    uint32_t tabL;        // this lane's copy of the 2^(j/256) table: entry j at tabL + (j << 7)
    const double *rows;   // the same rows / the plain table for the general evaluator
    const double *tab64;
};

// exp(u - Mq ln2) for the walk: k = rint(u * 256/ln2), two-constant Cody-Waite, |r| <= ln2/512, degree-4 polynomial
// (remainder 3.8e-17), 2^(j/256) from the lane-replicated table, exponent by integer add with the offset clamped on both
// sides (see the file header): 9 fp64-pipe instructions, one less than exp_scaled_r().
struct ExpRegs256 {
    double inv, nhi, nlo, c4, c3, c2;
};
__device__ __forceinline__ ExpRegs256 load_exp_regs256()
{
    ExpRegs256 r = {FHMC_EXP256_INV, -FHMC_EXP256_HI, -FHMC_EXP256_LO, 1.0 / 24.0, 1.0 / 6.0, 0.5};
    asm volatile("" : "+d"(r.inv), "+d"(r.nhi), "+d"(r.nlo), "+d"(r.c4), "+d"(r.c3), "+d"(r.c2));
    return r;
}
__device__ __forceinline__ double exp_rowc(double u, int Mq, uint32_t tabL, const ExpRegs256 &c)
{
    const double kd0 = fma(u, c.inv, FHMC_EXP_MAGIC);
    const int k = __double2loint(kd0);
    const double kd = kd0 - FHMC_EXP_MAGIC;
    double r = fma(kd, c.nhi, u);
    r = fma(kd, c.nlo, r);
    double p = fma(r, c.c4, c.c3);
    p = fma(p, r, c.c2);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double T = lds_f64(tabL + ((k & 255) << 7));
    const double v = T * p;
    const int q = min(max((k >> 8) - Mq, -1022), 1000);
    const int hi = __double2hiint(v) + (q << 20);
    return __hiloint2double(hi, __double2loint(v));
}
// per-state-point state of the walk (kept in registers: never pass its address to a non-inlined function)
struct RowcPt {
    double dD, q2, Sacc, Stot, u0, xm, uc, dc;
    long long sp;
    int Mq, P, cntM, cntm;
    unsigned rescue;
    bool bad, live;   // !live: a lane beyond the end of its run walks a copy of the last state point and writes nothing
    bool wig;         // some pair of successive differences of u changed sign (false: u is monotone, all differences share a sign bit)
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
                                         int P, int cntM, int cntm, unsigned rescue, double Stot, double u0, int mono)
{
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax, w = a.d.smooth;
    PointEval<1, true> pe(a, a.blob, lane, cx.tab64);   // repair() only: it reads no histogram value on these paths
    int *maxl = a.out.max_idx + sp * pmax;
    int *minl = a.out.min_idx + sp * (pmax + 1);
    int *bl = a.out.bounds + sp * pmax * 2;
    if (!(Stot < 0x1p900) || !(Stot > 0x1p-900)) return false;   // the shift (subsampled maximum / carried over) was too far off
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
        if (mono != 0) {
            // the walk saw no sign change between successive differences: u, and with it fl(u - c), is monotone (mono = 1:
            // non-decreasing, 2: strictly decreasing), so the maximum and the minimum sit at the two ends and can only be
            // tied with their neighbours there
            const double v0 = __dsub_rn(rowc_u<HC2>(cx, 0, dD, q2), c), v1 = __dsub_rn(rowc_u<HC2>(cx, 1, dD, q2), c);
            const double vl = __dsub_rn(rowc_u<HC2>(cx, last, dD, q2), c), vk = __dsub_rn(rowc_u<HC2>(cx, last - 1, dD, q2), c);
            const bool up = mono == 1;
            cM = (up ? vl > vk : v0 > v1) ? 1 : 2;
            cm = (up ? v0 < v1 : vl < vk) ? 1 : 2;
            pM = up ? last : 0;
            pm = up ? 0 : last;
        } else {
            for (int j = 0; j < n; ++j) {
                const double v = __dsub_rn(rowc_u<HC2>(cx, j, dD, q2), c);
                if (v > vM) { vM = v; cM = 1; pM = j; } else if (v == vM) ++cM;
                if (v < vm) { vm = v; cm = 1; pm = j; } else if (v == vm) ++cm;
            }
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

// F.E./kT of a phase from its max-shifted sum (out of line: log() is ~60 instructions and sits on the rare flush path)
__device__ __noinline__ double rowc_fe(double Sacc, int Mq, double u0) { return -(add_shift(Mq, log(Sacc)) - u0); }

// The walk of the two state points of the calling thread over the combined rows.
template <bool HC2, int NB>
__device__ __forceinline__ void rowc_walk(const SweepArgs &a, const RowcCtx &cx, const ExpRegs256 &ec, RowcPt &A, RowcPt &B, bool have_shift)
{
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax, w = a.d.smooth;
    const uint32_t sL = cx.sL, oC1 = cx.oC1, oC2 = cx.oC2, tabL = cx.tabL;
    auto uof = [&](const RowcPt &p, double L, double c1, double c2) {
        const double t = fma(p.dD, c1, L);
        return HC2 ? fma(p.q2, c2, t) : t;
    };
    auto load_u = [&](const RowcPt &p, int i) { return rowc_u<HC2>(cx, i, p.dD, p.q2); };
    // ---- shift: maximum over every 8th bin (+ the last one), unless the caller carried one over from the thread's previous
    // state points (512 positions earlier in the same run: ln Z moves by far less than the 2^+-900 a shift may be off)
    if (!have_shift) {
        double mA = -CUDART_INF, mB = -CUDART_INF;
        for (int i = 0; i < n; i += 8) {
            const uint32_t addr = sL + 8u * (uint32_t)i;
            const double L = lds_f64(addr), c1 = lds_f64(addr + oC1), c2 = HC2 ? lds_f64(addr + oC2) : 0.0;
            const double va = uof(A, L, c1, c2), vb = uof(B, L, c1, c2);
            mA = va > mA ? va : mA;
            mB = vb > mB ? vb : mB;
        }
        A.Mq = shift_for_max(fmax(mA, load_u(A, last)));
        B.Mq = shift_for_max(fmax(mB, load_u(B, last)));
    }
    auto flush = [&](RowcPt &p) {
        if (p.P < pmax && p.Sacc >= 1e-280) { if (p.live) a.out.fe[p.sp * pmax + p.P] = rowc_fe(p.Sacc, p.Mq, p.u0); }
        else if (p.P < pmax && p.P < 32) p.rescue |= 1u << p.P;   // re-integrated about its own maximum in rowc_finish()
        else p.bad = true;
        p.Stot += p.Sacc;
        p.Sacc = 0.0;
        ++p.P;
    };
    // bin i passed the strict 1-neighbour test: remaining shifts 2..w of argrelextrema, then the list entry; a confirmed
    // minimum flushes the running sums (a minimum bin opens the phase to its right, GH:498-520)
    auto confirm = [&](RowcPt &p, int i, double xc, bool is_max) {
        for (int d = 2; d <= w; ++d) {
            const int jl = (i - d < 0) ? 0 : i - d;
            const int jr = (i + d > last) ? last : i + d;
            const double xl = load_u(p, jl), xr = load_u(p, jr);
            if (!(is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr))) return;
        }
        if (is_max) {
            if (p.live && 1 + p.cntM <= pmax - 1) a.out.max_idx[p.sp * pmax + 1 + p.cntM] = i;
            ++p.cntM;
        } else {
            if (p.live && 1 + p.cntm <= pmax) a.out.min_idx[p.sp * (pmax + 1) + 1 + p.cntm] = i;
            ++p.cntm;
            flush(p);
        }
    };
    auto ex = [&](const RowcPt &p, double u) { return exp_rowc(u, p.Mq, tabL, ec); };
    // one block of NB bins i..i+NB-1 of one state point: un[] = u of the bins i+1..i+NB, p.xm / p.uc / p.dc carried
    auto fast_block = [&](RowcPt &p, const double (&un)[NB], double dlast) {
        double e[NB];
        e[0] = ex(p, p.uc);
#pragma unroll
        for (int k = 1; k < NB; ++k) e[k] = ex(p, un[k - 1]);
        p.Sacc += NB == 4 ? (e[0] + e[1]) + (e[2] + e[NB - 1]) : e[0] + e[1];
        p.xm = un[NB - 2];
        p.uc = un[NB - 1];
        p.dc = dlast;
    };
    // a block in which successive differences change sign: exact strict 1-neighbour tests; only a bin that passes them
    // (rare) takes the window test, in bin order, each bin's own term added after its test
    auto slow_block = [&](RowcPt &p, int i, const double (&un)[NB], double dlast) {
        p.wig = true;
        double x[NB + 2], e[NB];
        x[0] = p.xm;
        x[1] = p.uc;
#pragma unroll
        for (int k = 0; k < NB; ++k) x[k + 2] = un[k];
#pragma unroll
        for (int k = 0; k < NB; ++k) e[k] = ex(p, x[k + 1]);
        unsigned cM = 0, cm = 0;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            cM |= (x[k + 1] > x[k] && x[k + 1] > x[k + 2]) ? (1u << k) : 0u;
            cm |= (x[k + 1] < x[k] && x[k + 1] < x[k + 2]) ? (1u << k) : 0u;
        }
        if ((cM | cm) == 0) {
            p.Sacc += NB == 4 ? (e[0] + e[1]) + (e[2] + e[NB - 1]) : e[0] + e[1];
        } else {
            double xs[NB], es[NB];   // indexed by the loop counter below: local memory, on this rare path only
#pragma unroll
            for (int k = 0; k < NB; ++k) { xs[k] = x[k + 1]; es[k] = e[k]; }
#pragma unroll 1
            for (int k = 0; k < NB; ++k) {
                if (((cM | cm) >> k) & 1u) confirm(p, i + k, xs[k], ((cM >> k) & 1u) != 0);
                p.Sacc += es[k];
            }
        }
        p.xm = un[NB - 2];
        p.uc = un[NB - 1];
        p.dc = dlast;
    };
    auto start = [&](RowcPt &p) {
        p.u0 = load_u(p, 0);
        p.Sacc = ex(p, p.u0);
        p.uc = load_u(p, 1);
        p.xm = p.u0;
        p.dc = __dsub_rn(p.uc, p.xm);   // sign(dc) is the exact order of (xm, uc)
    };
    if (n < 3) {
        A.bad = B.bad = true;
        return;
    }
    start(A);
    start(B);
    int i = 1;
    uint32_t addr = sL + 16u;   // bin i + 1
    // sign bits of the NB + 1 successive differences around a block: any change means a strict 1-neighbour extremum may sit in it
    auto flips = [&](const RowcPt &p, const double (&un)[NB], double &dlast) {
        double d[NB];
        d[0] = __dsub_rn(un[0], p.uc);
#pragma unroll
        for (int k = 1; k < NB; ++k) d[k] = __dsub_rn(un[k], un[k - 1]);
        int f = __double2hiint(p.dc) ^ __double2hiint(d[0]);
#pragma unroll
        for (int k = 1; k < NB; ++k) f |= __double2hiint(d[k - 1]) ^ __double2hiint(d[k]);
        dlast = d[NB - 1];
        return f;
    };
#pragma unroll 1
    for (; i + NB - 1 < last; i += NB, addr += 8u * NB) {   // bins i..i+NB-1 are interior, i+NB <= last exists
        double L[NB], c1[NB], c2[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            L[k] = lds_f64(addr + 8u * k);
            c1[k] = lds_f64(addr + oC1 + 8u * k);
            c2[k] = HC2 ? lds_f64(addr + oC2 + 8u * k) : 0.0;
        }
        double ua[NB], ub[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            ua[k] = uof(A, L[k], c1[k], c2[k]);
            ub[k] = uof(B, L[k], c1[k], c2[k]);
        }
        double da, db;
        const int flipA = flips(A, ua, da), flipB = flips(B, ub, db);
        if ((flipA | flipB) < 0) {   // look closely (per point)
            if (flipA < 0) slow_block(A, i, ua, da); else fast_block(A, ua, da);
            if (flipB < 0) slow_block(B, i, ub, db); else fast_block(B, ub, db);
        } else {
            fast_block(A, ua, da);
            fast_block(B, ub, db);
        }
    }
    auto tail = [&](RowcPt &p) {
#pragma unroll 1
        for (int j = i; j < last; ++j) {
            const double un = load_u(p, j + 1), xc = p.uc, dn = __dsub_rn(un, xc);
            if ((__double2hiint(dn) ^ __double2hiint(p.dc)) < 0) p.wig = true;
            p.dc = dn;
            const bool is_max = (xc > p.xm) && (xc > un), is_min = (xc < p.xm) && (xc < un);
            if (is_max || is_min) confirm(p, j, xc, is_max);
            p.Sacc += ex(p, xc);
            p.xm = xc;
            p.uc = un;
        }
        p.Sacc += ex(p, p.uc);
        flush(p);
    };
    tail(A);
    tail(B);
}

__device__ __forceinline__ void rowc_init(RowcPt &p, long long sp, double dmu, double dmu_ref, bool live, int Mq)
{
    p.sp = sp;
    p.live = live;
    p.wig = false;
    p.dD = dmu - dmu_ref;
    p.q2 = monomial(FHMC_M_DD2, 0.0, p.dD, 0.0);
    p.Sacc = p.Stot = 0.0;
    p.P = p.cntM = p.cntm = 0;
    p.rescue = 0;
    p.bad = false;
    p.Mq = Mq;
    p.u0 = p.xm = p.uc = p.dc = 0.0;
}

// bytes of shared memory: two row buffers (L | C2 | C1 each), the plain 2^(j/64) table of the general evaluator, the
// lane-replicated 2^(j/256) table, two release counters
static size_t rowc_smem_bytes(int n_pad) { return ((size_t)6 * n_pad + 64 + 256 * 16) * 8 + 16; }

template <bool HC2>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_rowc(const __grid_constant__ SweepArgs a, const __grid_constant__ RowcPlan pl)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = a.d.n, npad = a.d.n_pad, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *rows0 = reinterpret_cast<double *>(smem_raw);   // two buffers of L | C2 | C1
    double *tab64 = rows0 + 6 * (size_t)npad;
    double *tabR = tab64 + 64;                              // [256][16]: entry j of copy c at (j * 16 + c)
    unsigned *released = reinterpret_cast<unsigned *>(tabR + 256 * 16);   // per buffer: warps that are done with a use of it
    stage_exp_table(tab64);
    for (int j = threadIdx.x; j < 256 * 16; j += blockDim.x) tabR[j] = c_exp256[j >> 4];
    if (threadIdx.x < 2) released[threadIdx.x] = 0u;
    __syncthreads();   // the only CTA-wide barrier of the kernel
    const ExpRegs256 ec = load_exp_regs256();
    RowcCtx cx;
    cx.oC2 = (uint32_t)npad * 8u;
    cx.oC1 = (uint32_t)npad * 16u;
    cx.tabL = smem_u32(tabR) + ((uint32_t)(lane & 15) << 3);
    cx.tab64 = tab64;
    cx.rows = rows0;
    cx.sL = smem_u32(rows0);
    const long long items = pl.n_runs * pl.chunks_per_run;
    long long cur_run = -1;
    unsigned epoch = 0;   // row builds so far; build e uses buffer e & 1
    // contiguous items per CTA: the chunks of one run follow each other, so its rows are combined once
    const long long it0 = items * blockIdx.x / gridDim.x, it1 = items * (blockIdx.x + 1) / gridDim.x;
    for (long long it = it0; it < it1; ++it) {
        const long long run = it / pl.chunks_per_run;
        const int ch = (int)(it % pl.chunks_per_run);
        const long long sp0 = run * pl.n_run;
        if (run != cur_run) {
            // Every warp combines the coefficient rows of this (mu_1, beta) itself (all warps store identical values), so no
            // warp waits for another one here; the two row buffers alternate, and a buffer is overwritten only after all
            // warps of the CTA have released its previous use, two builds back (counter in shared memory).
            __syncwarp();
            if (cur_run >= 0 && lane == 0) {
                __threadfence_block();
                atomicAdd(&released[(epoch - 1) & 1u], 1u);
            }
            cur_run = run;
            const unsigned b = epoch & 1u, need = (epoch >> 1) * (FHMC_CTA / 32);
            if (lane == 0)
                while (*reinterpret_cast<volatile unsigned *>(&released[b]) < need) { }
            __syncwarp();
            double *rows = rows0 + (size_t)b * 3 * npad;
            const double mu1 = a.st.mu1[(sp0 / a.st.mu1_div) % a.st.n_mu1];
            const double beta = a.st.beta ? a.st.beta[(sp0 / a.st.beta_div) % a.st.n_beta] : a.d.beta_ref;
            // terms in descriptor order, each into the row of its power of dD
            const double s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);   // GH:77, evaluated left to right
            const double dB = beta - a.d.beta_ref;
            for (int i = lane; i < npad; i += 32) {
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
            __syncwarp();
            cx.rows = rows;
            cx.sL = smem_u32(rows);
            ++epoch;
        }
        const long long j0 = (long long)ch * pl.chunk, j1 = min(j0 + (long long)pl.chunk, pl.n_run);
        int MqA = 0, MqB = 0;
        bool carry = false;   // MqA / MqB: ln Z / ln 2 of the thread's previous two state points of this item
        for (long long t = j0 + warp * 64; t < j1; t += (FHMC_CTA / 32) * 64) {
            // a lane beyond the end of the run walks the run's last state point again without writing anything
            const bool liveA = t + lane < j1, liveB = t + 32 + lane < j1;
            const long long jA = liveA ? t + lane : j1 - 1, jB = liveB ? t + 32 + lane : j1 - 1;
            RowcPt A, B;
            rowc_init(A, sp0 + jA, a.st.dmu[jA], a.d.dmu_ref, liveA, MqA);
            rowc_init(B, sp0 + jB, a.st.dmu[jB], a.d.dmu_ref, liveB, MqB);
            rowc_walk<HC2, 4>(a, cx, ec, A, B, carry);
            const int monoA = A.wig ? 0 : (__double2hiint(A.dc) < 0 ? 2 : 1), monoB = B.wig ? 0 : (__double2hiint(B.dc) < 0 ? 2 : 1);
            const bool okA = !liveA || !A.bad && rowc_finish<HC2>(a, cx, lane, A.sp, A.dD, A.q2, A.Mq, A.P, A.cntM, A.cntm, A.rescue, A.Stot, A.u0, monoA);
            const bool okB = !liveB || !B.bad && rowc_finish<HC2>(a, cx, lane, B.sp, B.dD, B.q2, B.Mq, B.P, B.cntM, B.cntm, B.rescue, B.Stot, B.u0, monoB);
            // next tile of this warp: shift = exponent of the sums just formed (both walks plain and in range), else a fresh pre-pass
            carry = !A.bad && !B.bad && A.Stot < 0x1p900 && A.Stot > 0x1p-900 && B.Stot < 0x1p900 && B.Stot > 0x1p-900;
            MqA = A.Mq + ((__double2hiint(A.Stot) >> 20) & 0x7ff) - 1022;
            MqB = B.Mq + ((__double2hiint(B.Stot) >> 20) & 0x7ff) - 1022;
            // anything unusual: the whole warp re-runs it with the general evaluator, on the combined rows
            __syncwarp();
            unsigned fA = __ballot_sync(0xffffffffu, !okA), fB = __ballot_sync(0xffffffffu, !okB);
            while (fA | fB) {
                const bool second = fA == 0;
                unsigned &f = second ? fB : fA;
                const int src = __ffs(f) - 1;
                f &= f - 1;
                const long long j = t + (second ? 32 : 0) + src;
                rowc_generic(a, cx.rows, tab64, lane, a.st.dmu[j], sp0 + j, HC2 ? 1 : 0);
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
    const size_t smem = rowc_smem_bytes(d.n_pad);
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
