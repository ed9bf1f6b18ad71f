// fhmc_prod.cuh -- pure-mu sweeps in product form, TWO state points per thread (k_sweep_prod2).
//
// Same arithmetic as the REC == 2 walk of k_sweep_fast (fhmc_fast.cuh): exp(lnPI_i + s N_i - shift) = P_i t_i with
// P_i = exp(lnPI_i - A_seg) tabulated per CTA and t_i geometric in the bin index, so a 4-bin block costs one Horner
// polynomial in exp(s dN) per summed quantity.  Every lane of a warp reads the same table entry (a broadcast LDS.128 costs
// two shared-memory wavefronts for 16 bytes), and with one state point per thread the SM-wide shared-memory pipe, not the
// fp64 pipe, is the limiter (ncu: 69 % vs 41 %, profiles/r01b_prod_sweep_ncu_summary.txt).  Here a thread carries two
// state points through the blocks together, so each table entry is loaded once for both: half the shared-memory traffic
// per state point, and the two points' independent Horner chains give the instruction-level parallelism that two blocks
// per iteration gave before.
//
// The two walks share their control flow wherever both are plain (no block of either flagged by the tilt-range keys, both
// segment anchors usable); anything else is done per point with the single-point building blocks, which follow the
// single-point kernel line by line.  Prologue (hull search, shift) and epilogue (tail bins, repair(), verification,
// rescue of underflowed phases, is_safe, record) run per point.
#pragma once
#include "fhmc_fast.cuh"

namespace fhmc {

// COMPACT: records leave in the phase-major narrow form of fhmc_pack_phase_soa16, written by the walk itself to every
// destination of SweepArgs::c (this GPU's buffer and, for a gather fused into the sweep, the peers' buffers over NVLink):
// no fhmc_sweep_out arrays, no repack kernel.  The extrema / bounds lists live in a per-thread local array meanwhile.
template <int NSEL, bool SEL0N, bool COMPACT = false>
struct ProdWalk {
    using LY = FastLayout<NSEL, SEL0N, 0, 1, 2>;
    static constexpr int NX = LY::NX, PK = LY::PK, XOFF = LY::XOFF;
    static constexpr uint32_t BWB = (uint32_t)(LY::BW * 8);
    static constexpr int NA = NSEL > 0 ? NSEL : 1;

    struct Bin {
        double u, N, x[NX > 0 ? NX : 1];
    };
    // one state point in flight
    // (kept small on purpose: two of these plus one block of table entries must fit the 128-register budget of two CTAs
    // per SM without the hot loop re-deriving exp(s dN)^2 every iteration -- u_0 and the decision margin are recomputed
    // where the rare paths need them, the three flags share a word)
    struct PS {
        double s, Sacc, Stot, A[NA], t, r1, r4;
        long long sp;
        int Mq, k_hi, k_lo, cntM, cntm, P;
        unsigned rescue, fl;
        int *lst;   // COMPACT: this state point's lists {max_idx[PM], min_idx[PM + 1], bounds[2 PM]} (local memory)
    };
    static constexpr int PM = FHMC_COMPACT_PMAX, LST = 4 * FHMC_COMPACT_PMAX + 1;
    static constexpr bool USES_LST = COMPACT;
    static constexpr unsigned F_BAD = 1u, F_ROBUST = 2u, F_CHAIN = 4u;
    __device__ __forceinline__ double u0(const PS &p) const
    {
        double Nd;
        return load_u(p, 0, Nd);
    }
    // more than rounding of fl(u - c) can move two values against each other: >= 2^-51 (|u| + |c|) for every bin
    __device__ __forceinline__ double vmargin(const PS &p) const
    {
        const double Na = fmax(fabs(lds_f64(s_pk + 8u)), fabs(lds_f64(s_pk + (uint32_t)last * (uint32_t)(PK * 8) + 8u)));
        return 1.8e-15 * (cx.lmax + fabs(p.s) * Na) + 1e-300 + 1e-14;
    }

    const SweepArgs &a;
    const FastCtx &cx;
    PointEval<1, false> *const pe_;   // prologue / epilogue only (setup, repair); null inside slow_range_call()
    const int n, last, pmax, smooth;
    const uint32_t s_pk, tab;

    __device__ ProdWalk(const SweepArgs &a_, const FastCtx &cx_, PointEval<1, false> *pe, int smooth_, uint32_t tab_)
        : a(a_), cx(cx_), pe_(pe), n(a_.d.n), last(a_.d.n - 1), pmax(a_.d.pmax), smooth(smooth_), s_pk(cx_.s_pk), tab(tab_)
    {
    }

    __device__ __forceinline__ int *maxl(const PS &p) const { return COMPACT ? p.lst : a.out.max_idx + p.sp * pmax; }
    __device__ __forceinline__ int *minl(const PS &p) const { return COMPACT ? p.lst + PM : a.out.min_idx + p.sp * (pmax + 1); }
    __device__ __forceinline__ int *bl(const PS &p) const { return COMPACT ? p.lst + 2 * PM + 1 : a.out.bounds + p.sp * pmax * 2; }

    // ---- compact records: layout of fhmc_pack_phase_soa16 for c.n_total records, record c.first + sp ------------------------
    __device__ __forceinline__ double *cF(int d, int ph, long long sp) const
    {
        unsigned char *base = a.c.dst[d] + ((4 * a.c.n_total + 15) & ~15ll);
        return reinterpret_cast<double *>(base) + ((long long)ph * a.c.n_total + a.c.first + sp) * (1 + NSEL);
    }
    __device__ __forceinline__ short2 *cB(int d, int ph, long long sp) const
    {
        unsigned char *base = a.c.dst[d] + ((4 * a.c.n_total + 15) & ~15ll) + (long long)pmax * a.c.n_total * (1 + NSEL) * 8;
        return reinterpret_cast<short2 *>(base) + (long long)ph * a.c.n_total + a.c.first + sp;
    }
    // F.E./kT and the averages of phase ph of state point sp
    __device__ __forceinline__ void put_phase(long long sp, int ph, double fe, double S, const double *A) const
    {
        if (COMPACT) {
            double v[NA];
#pragma unroll
            for (int q = 0; q < NSEL; ++q) v[q] = A[q] / S;
            for (int d = 0; d < a.c.n_dst; ++d) {
                double *f = cF(d, ph, sp);
                f[0] = fe;
#pragma unroll
                for (int q = 0; q < NSEL; ++q) f[1 + q] = v[q];
            }
        } else {
            a.out.fe[sp * pmax + ph] = fe;
#pragma unroll
            for (int q = 0; q < NSEL; ++q) a.out.avg[(sp * pmax + ph) * NSEL + q] = A[q] / S;
        }
    }

    // u_i = fl(lnPI_i + fl(s N_i)), bit-identical to GH:77
    __device__ __forceinline__ double load_u(const PS &p, int i, double &Ni) const
    {
        double l;
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l), "=d"(Ni) : "r"(s_pk + (uint32_t)i * (uint32_t)(PK * 8)));
        return __dadd_rn(l, __dmul_rn(p.s, Ni));
    }
    __device__ __forceinline__ void load_bin(const PS &p, int i, Bin &b) const
    {
        b.u = load_u(p, i, b.N);
        const uint32_t addr = s_pk + (uint32_t)i * (uint32_t)(PK * 8) + 8u * XOFF;
#pragma unroll
        for (int q = 0; q < NX; ++q) b.x[q] = lds_f64(addr + 8u * q);
    }
    __device__ __forceinline__ void accumulate(PS &p, const Bin &b) const
    {
        const double e = exp_scaled(b.u, p.Mq, tab);
        p.Sacc += e;
        if (SEL0N) p.A[0] = fma(e, b.N, p.A[0]);
#pragma unroll
        for (int q = 0; q < NX; ++q) p.A[q + (SEL0N ? 1 : 0)] = fma(e, b.x[q], p.A[q + (SEL0N ? 1 : 0)]);
    }
    // close the running phase (a minimum bin opens the phase to its right, GH:498-520)
    __device__ __forceinline__ void flush(PS &p) const
    {
        if (p.P < pmax && p.Sacc >= 1e-280) {
            put_phase(p.sp, p.P, -(add_shift(p.Mq, log(p.Sacc)) - u0(p)), p.Sacc, p.A);
        } else if (p.P < pmax && p.P < 32) {
            p.rescue |= 1u << p.P;   // phase too unlikely for the common shift: re-integrated about its own maximum in finish()
        } else {   // (also: a negligible phase beyond the 32 the rescue mask can name -- left to the generic evaluator)
            p.fl |= F_BAD;
        }
        p.Stot += p.Sacc;
        p.Sacc = 0.0;
#pragma unroll
        for (int q = 0; q < NSEL; ++q) p.A[q] = 0.0;
        ++p.P;
    }
    // shifts d0..smooth of the argrelextrema test (GH:329-330) of bin i; `track`: fold the decision margins into p.robust
    __device__ __forceinline__ bool window(PS &p, int i, double xc, bool is_max, bool use_c, double cc, int d0, bool track) const
    {
        double Nd;
        const double vm = vmargin(p);
        const double xg = is_max ? xc - vm : xc + vm;
        bool rb = true;
        for (int d = d0; d <= smooth; ++d) {
            const int jl = (i - d < 0) ? 0 : i - d;
            const int jr = (i + d > last) ? last : i + d;
            double xl = load_u(p, jl, Nd), xr = load_u(p, jr, Nd);
            if (use_c) { xl = __dsub_rn(xl, cc); xr = __dsub_rn(xr, cc); }
            const bool ok = is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr);
            if (!ok) return false;
            rb &= is_max ? (xg > xl && xg > xr) : (xg < xl && xg < xr);
        }
        if (track && !rb) p.fl &= ~F_ROBUST;
        return true;
    }
    // exact strict 1-neighbour test + window test of bin i (values xm, xc, xp); bookkeeping of a confirmed extremum
    __device__ __forceinline__ void test_bin(PS &p, int i, double xm, double xc, double xp) const
    {
        const bool is_max = (xc > xm) && (xc > xp), is_min = (xc < xm) && (xc < xp);
        if ((is_max || is_min) && window(p, i, xc, is_max, false, 0.0, 2, true)) {
            const double vm = vmargin(p);
            if (!(is_max ? (xc - vm > xm && xc - vm > xp) : (xc + vm < xm && xc + vm < xp))) p.fl &= ~F_ROBUST;
            if (is_max) {
                if (1 + p.cntM <= pmax - 1) maxl(p)[1 + p.cntM] = i;
                ++p.cntM;
            } else {
                if (1 + p.cntm <= pmax) minl(p)[1 + p.cntm] = i;
                ++p.cntm;
                flush(p);
            }
        }
    }
    __device__ __forceinline__ bool flagged(const PS &p, int ka, int kb) const { return !((ka > p.k_hi) | (kb < p.k_lo)); }

    // ---- single-point building blocks of the walk ---------------------------------------------------------------
    // one block whose table entries are in registers: tb = {P0..P3, (P X_q)0..3 ...}
    __device__ __forceinline__ void fast_block_regs(PS &p, const double (&tb)[4 * (1 + NA)], double r2) const
    {
        p.Sacc = fma(fma(fma(tb[3], p.r1, tb[2]), r2, fma(tb[1], p.r1, tb[0])), p.t, p.Sacc);
#pragma unroll
        for (int q = 0; q < NSEL; ++q)
            p.A[q] = fma(fma(fma(tb[7 + 4 * q], p.r1, tb[6 + 4 * q]), r2, fma(tb[5 + 4 * q], p.r1, tb[4 + 4 * q])), p.t, p.A[q]);
    }
    __device__ __forceinline__ void load_block(uint32_t pb, double (&tb)[4 * (1 + NA)]) const
    {
#pragma unroll
        for (int v = 0; v < 2 * (1 + NSEL); ++v)
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tb[2 * v]), "=d"(tb[2 * v + 1]) : "r"(pb + 16u + 16u * v));
    }
    // exact tests on u, bin by bin (a confirmed minimum flushes the sums before its own term is added); the terms still come
    // from the tabulated products: e_{i+k} = P_{i+k} t r1^k
    __device__ __forceinline__ void careful_block(PS &p, uint32_t pb, int ib) const
    {
        double Nd, uu[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) uu[k] = load_u(p, ib - 1 + k, Nd);
        double tk = p.t;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            test_bin(p, ib + k, uu[k], uu[k + 1], uu[k + 2]);
            p.Sacc = fma(lds_f64(pb + 16u + 8u * k), tk, p.Sacc);
#pragma unroll
            for (int q = 0; q < NSEL; ++q) p.A[q] = fma(lds_f64(pb + 48u + 32u * q + 8u * k), tk, p.A[q]);
            tk *= p.r1;
        }
    }
    // one block with one true exp per bin (segment whose factor t is clamped or unusable)
    __device__ __forceinline__ void exact_block(PS &p, uint32_t pb, int ib, bool flag) const
    {
        if (flag) {
            double Nd;
            Bin c0, b1, b2, b3;
            const double um = load_u(p, ib - 1, Nd);
            load_bin(p, ib, c0);
            load_bin(p, ib + 1, b1);
            load_bin(p, ib + 2, b2);
            load_bin(p, ib + 3, b3);
            const double u4 = load_u(p, ib + 4, Nd);
            test_bin(p, ib, um, c0.u, b1.u);
            accumulate(p, c0);
            test_bin(p, ib + 1, c0.u, b1.u, b2.u);
            accumulate(p, b1);
            test_bin(p, ib + 2, b1.u, b2.u, b3.u);
            accumulate(p, b2);
            test_bin(p, ib + 3, b2.u, b3.u, u4);
            accumulate(p, b3);
        } else {
            Bin c0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                load_bin(p, ib + k, c0);
                accumulate(p, c0);
            }
        }
    }
    // t = exp(A_g + s N_i - shift) at the first bin of a segment; false when clamped / chain switched off
    __device__ __forceinline__ bool anchor(PS &p, int g, int i) const
    {
        double Ni;
        const double lA = lds_f64(cx.s_anch + 8u * (uint32_t)g);
        asm("ld.shared.f64 %0, [%1];" : "=d"(Ni) : "r"(s_pk + (uint32_t)i * (uint32_t)(PK * 8) + 8u));
        p.t = exp_scaled(__dadd_rn(lA, __dmul_rn(p.s, Ni)), p.Mq, tab);
        // exp_scaled() clamps an underflowing result to [2^-1022, 2^-1020): anything that small is not a usable factor.
        return (p.t > 1e-300) && (p.fl & F_CHAIN);
    }
    // blocks [b, bend) of one segment for one point (usable: chained products; else true exps)
    __device__ __forceinline__ void segment_single(PS &p, bool usable, int b, int bend, int i, uint32_t pb) const
    {
        const double r2 = p.r1 * p.r1;
        for (; b < bend; ++b, i += 4, pb += BWB) {
            int ka, kb;
            asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka), "=r"(kb) : "r"(pb));
            const bool f = flagged(p, ka, kb);
            if (!usable) {
                exact_block(p, pb, i, f);
                continue;
            }
            if (f) {
                careful_block(p, pb, i);
            } else {
                double tb[4 * (1 + NA)];
                load_block(pb, tb);
                fast_block_regs(p, tb, r2);
            }
            p.t *= p.r4;
        }
    }

    // ---- prologue -----------------------------------------------------------------------------------------------
    __device__ __forceinline__ void init(PS &p, long long sp, double mu1) const
    {
        PointEval<1, false> &pe = *pe_;
        pe.setup(mu1, a.d.beta_ref, a.d.dmu_ref);
        p.s = pe.s;
        p.sp = sp;
        p.cntM = p.cntm = p.P = 0;
        p.rescue = 0;
        p.fl = F_ROBUST | (n < 3 ? F_BAD : 0u);
        p.Sacc = p.Stot = 0.0;
        p.t = 0.0;
#pragma unroll
        for (int q = 0; q < NA; ++q) p.A[q] = 0.0;
        // hull vertex maximising lnPI + s N
        int lo = 0, hi = cx.H - 1;
        const double neg_s = -p.s;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (lds_f64(cx.s_slope + 8u * mid) > neg_s) lo = mid + 1; else hi = mid;
        }
        double Nm;
        p.Mq = shift_for_max(load_u(p, (int)cx.g_hidx[lo], Nm));
        const double sdn = p.s * (lds_f64(s_pk + (uint32_t)(PK * 8) + 8u) - lds_f64(s_pk + 8u));   // s dN
        if (!(fabs(4.0 * sdn) < 200.0)) p.fl |= F_BAD;   // extreme tilt: leave it to the generic evaluator
        p.r1 = exp(sdn);
        p.r4 = exp(4.0 * sdn);
        const double Na = fmax(fabs(lds_f64(s_pk + 8u)), fabs(lds_f64(s_pk + (uint32_t)last * (uint32_t)(PK * 8) + 8u)));
        const double margin = 1.8e-15 * (cx.lmax + fabs(p.s) * Na) + 1e-300;   // 8 * 2^-52 * (|lnPI| + |s N|)
        const bool chain_ok = fabs(sdn) < cx.sdn_lim;   // product form usable for this tilt (fast_prepare); else every block is examined
        if (chain_ok) p.fl |= F_CHAIN;
        p.k_hi = chain_ok ? hi_key(-sdn + margin) : 0x7fffffff;
        p.k_lo = chain_ok ? hi_key(-sdn - margin) : (int)0x80000000;
        Bin b0;
        load_bin(p, 0, b0);
        if (!(p.fl & F_BAD)) accumulate(p, b0);
    }

    // ---- the walk over the full blocks (bins 1 .. 4 nb) -------------------------------------------------------
    __device__ __forceinline__ void walk1(PS &p) const
    {
        const int nb = (n - 2) / 4;
        uint32_t pb = cx.s_prod;
        int b = 0, i = 1;
        for (int g = 0; b < nb; ++g) {
            const int bend = min(nb, b + LY::SEGB);
            segment_single(p, anchor(p, g, i), b, bend, i, pb);
            i += 4 * (bend - b);
            pb += BWB * (uint32_t)(bend - b);
            b = bend;
        }
    }
    // The hot loop.  A group of GRPB blocks whose key hull clears both state points is summed with no test at all
    // (6 LDS.128 + 24 DFMA + 2 DMUL per block for the two points); a group that may hold an extremum of either point is
    // walked point by point by slow_group() (out of line: the hot loop stays a few hundred bytes of straight code).
    __device__ __forceinline__ void walk2(PS &p0, PS &p1) const
    {
        constexpr int GB = LY::GRPB;
        const int nb = (n - 2) / 4;
        const double r2a = p0.r1 * p0.r1, r2b = p1.r1 * p1.r1;
        uint32_t pb = cx.s_prod, pg = cx.s_gkey;
        int b = 0, i = 1;
        for (int g = 0; b < nb; ++g) {
            const int bend = min(nb, b + LY::SEGB);
            const bool ua = anchor(p0, g, i), ub = anchor(p1, g, i);
            if (!(ua & ub)) {   // rare: walk this segment point by point
                slow_range(p0, ua, b, bend, i, pb);
                slow_range(p1, ub, b, bend, i, pb);
                i += 4 * (bend - b);
                pb += BWB * (uint32_t)(bend - b);
                pg += 8u * (uint32_t)((bend - b + GB - 1) / GB);
                b = bend;
                continue;
            }
            while (b < bend) {
                const int ge = min(bend, b + GB);
                int ka, kb;
                asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(ka), "=r"(kb) : "r"(pg));
                pg += 8u;
                if (flagged(p0, ka, kb) | flagged(p1, ka, kb)) {
                    // block by block with the blocks' own keys; a flagged block runs the exact tests for its point
                    int ib = i;
                    for (int k = b; k < ge; ++k, ib += 4, pb += BWB) {
                        int kba, kbb;
                        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(kba), "=r"(kbb) : "r"(pb));
                        const bool fa = flagged(p0, kba, kbb), fb = flagged(p1, kba, kbb);
                        double tb[4 * (1 + NA)];
                        load_block(pb, tb);
                        if (fa) careful_block(p0, pb, ib); else fast_block_regs(p0, tb, r2a);
                        if (fb) careful_block(p1, pb, ib); else fast_block_regs(p1, tb, r2b);
                        p0.t *= p0.r4;
                        p1.t *= p1.r4;
                    }
                } else if (ge - b == GB) {
#pragma unroll 2
                    for (int k = 0; k < GB; ++k, pb += BWB) {
                        double tb[4 * (1 + NA)];
                        load_block(pb, tb);
                        fast_block_regs(p0, tb, r2a);
                        fast_block_regs(p1, tb, r2b);
                        p0.t *= p0.r4;
                        p1.t *= p1.r4;
                    }
                } else {
                    for (int k = b; k < ge; ++k, pb += BWB) {
                        double tb[4 * (1 + NA)];
                        load_block(pb, tb);
                        fast_block_regs(p0, tb, r2a);
                        fast_block_regs(p1, tb, r2b);
                        p0.t *= p0.r4;
                        p1.t *= p1.r4;
                    }
                }
                i += 4 * (ge - b);
                b = ge;
            }
        }
    }
    // blocks [b, bend) for one point with per-block key tests, exact tests in flagged blocks (segment_single), out of line.
    // The state travels through a copy so that the caller's PS never has its address taken.
    __device__ __forceinline__ void slow_range(PS &p, bool usable, int b, int bend, int i, uint32_t pb) const
    {
        PS q = p;
        slow_range_call(a, cx, smooth, tab, &q, usable, b, bend, i, pb);
        p = q;
    }
    static __device__ __noinline__ void slow_range_call(const SweepArgs &a_, FastCtx cx_, int smooth_, uint32_t tab_, PS *q, bool usable,
                                                        int b, int bend, int i, uint32_t pb)
    {
        const ProdWalk w(a_, cx_, nullptr, smooth_, tab_);
        PS p = *q;
        w.segment_single(p, usable, b, bend, i, pb);
        *q = p;
    }

    // ---- epilogue: tail bins, validation with the exact rules of the generic path, record ------------------------
    // Returns false when the state point is not a plain case and must be re-run by the generic evaluator.
    __device__ __forceinline__ bool finish(PS &p) const
    {
        if (p.fl & F_BAD) return false;
        PointEval<1, false> &pe = *pe_;
        pe.s = p.s;
        {
            int i = 1 + 4 * ((n - 2) / 4);
            double Nd;
            double xm = load_u(p, i - 1, Nd);
            Bin c;
            load_bin(p, i, c);
            for (; i < last; ++i) {
                Bin nx;
                load_bin(p, i + 1, nx);
                test_bin(p, i, xm, c.u, nx.u);
                accumulate(p, c);
                xm = c.u;
                c = nx;
            }
            accumulate(p, c);
            flush(p);
        }
        if ((p.fl & F_BAD) || a.d.complete) return false;
        int *const ml = maxl(p), *const mn = minl(p), *const bb = bl(p);
        const long long sp = p.sp;
        int nM = 0, nm = 0;
        bool part = false;
        unsigned flags = 0;
        const double c = add_shift(p.Mq, log(p.Stot));
        int rc;
        if (p.cntM == 0 && p.cntm == 0) {
            // No windowed extremum at all (monotone ln(PI)): the reference takes the bins tied with the max / min of the
            // NORMALISED array (GH:382-386).  One light scan on fl(u - c); only genuine ties go to the generic evaluator.
            double vM = -CUDART_INF, vm = CUDART_INF, Nd;
            int cM = 0, cm = 0, pM = 0, pm = 0;
            p.fl &= ~F_ROBUST;   // nothing was decided by the walk: keep the re-test of whatever repair() lists
            for (int j = 0; j < n; ++j) {
                const double v = __dsub_rn(load_u(p, j, Nd), c);
                if (v > vM) { vM = v; cM = 1; pM = j; } else if (v == vM) ++cM;
                if (v < vm) { vm = v; cm = 1; pm = j; } else if (v == vm) ++cm;
            }
            rc = (cM == 1 && cm == 1) ? pe.repair(true, c, 0, 0, 0.0, 0.0, ml, mn, bb, nM, nm, flags, part, 1, 1, pM, pm)
                                      : FHMC_NEED_SLOW;
        } else {
            rc = pe.repair(false, 0.0, p.cntM, p.cntm, 0.0, 0.0, ml, mn, bb, nM, nm, flags, part);
        }
        if (!(rc == FHMC_OK && part && nM == p.P)) return false;
        pe.P = nM;
        pe.nmin = nm;
        // re-test the detected interior extrema on the normalised values, as PointEval::verify() -- unless every
        // comparison behind them was decided by more than rounding can move (p.robust)
        if (!a.d.compare_raw && !(p.fl & F_ROBUST)) {
            double Nd;
            for (int k = 0; k < nM + nm; ++k) {
                const bool is_max = k < nM;
                const int idx = is_max ? ml[k] : mn[k - nM];
                if (idx > 0 && idx < last && !window(p, idx, __dsub_rn(load_u(p, idx, Nd), c), is_max, true, c, 1, false))
                    return false;
            }
        }
        double Nd;
        // phases whose weight underflowed next to the global maximum: integrate them about their own maximum, as
        // PointEval::partition_sum_probe() does (status bit RESCUED)
        for (int ph = 0; p.rescue != 0 && ph < nM; ++ph) {
            if (!((p.rescue >> ph) & 1u)) continue;
            const int left = bb[2 * ph], right = bb[2 * ph + 1];
            double mlx = -CUDART_INF;
            for (int j = left; j < right; ++j) mlx = fmax(mlx, load_u(p, j, Nd));
            const int Mp = shift_for_max(mlx);
            double Sp = 0.0, Ap[NA];
#pragma unroll
            for (int q = 0; q < NA; ++q) Ap[q] = 0.0;
            for (int j = left; j < right; ++j) {
                Bin b;
                load_bin(p, j, b);
                const double e = exp_scaled(b.u, Mp, tab);
                Sp += e;
                if (SEL0N) Ap[0] = fma(e, b.N, Ap[0]);
#pragma unroll
                for (int q = 0; q < NX; ++q) Ap[q + (SEL0N ? 1 : 0)] = fma(e, b.x[q], Ap[q + (SEL0N ? 1 : 0)]);
            }
            put_phase(sp, ph, -(add_shift(Mp, log(Sp)) - u0(p)), Sp, Ap);
            flags |= FHMC_ST_RESCUED;
        }
        const double xM = __dsub_rn(load_u(p, ml[nM - 1], Nd), c), xl = __dsub_rn(load_u(p, last, Nd), c);
        if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
        if (COMPACT) {
            put_compact_tail(sp, flags | FHMC_ST_FAST, nM, bb);
        } else {
            a.out.status[sp] = flags | FHMC_ST_FAST;
            a.out.nphase[sp] = nM;
            a.out.nmin[sp] = nm;
            a.out.lnnorm[sp] = c;
        }
        return true;
    }

    // head + bounds (+ NaN / -1 in the phase slots that do not exist) of a compact record; fe / avg of the live phases are
    // already in place (put_phase)
    __device__ __forceinline__ void put_compact_tail(long long sp, unsigned status, int nM, const int *bb) const
    {
        put_compact_tail_fill(sp, status, nM, bb, a.c.fill_dead != 0);
    }
    __device__ __forceinline__ void put_compact_tail_fill(long long sp, unsigned status, int nM, const int *bb, bool fill) const
    {
        const int Pe = ((status & FHMC_ST_CODE_MASK) == FHMC_OK) ? min(max(nM, 0), pmax) : 0;
        const uchar4 h = make_uchar4((unsigned char)(status & 0xFFu), (unsigned char)((status >> 8) & 0xFFu),
                                     (unsigned char)min(max(nM, 0), 255), 0);
        for (int d = 0; d < a.c.n_dst; ++d) {
            reinterpret_cast<uchar4 *>(a.c.dst[d])[a.c.first + sp] = h;
            for (int ph = 0; ph < Pe; ++ph) *cB(d, ph, sp) = make_short2((short)bb[2 * ph], (short)bb[2 * ph + 1]);
            if (fill)
                for (int ph = Pe; ph < pmax; ++ph) {
                    double *f = cF(d, ph, sp);
#pragma unroll
                    for (int q = 0; q <= NSEL; ++q) f[q] = CUDART_NAN;
                    *cB(d, ph, sp) = make_short2(-1, -1);
                }
        }
    }
};

// Two state points per thread: thread t of tile T owns state points T*512 + t and T*512 + 256 + t.  W = ProdWalk or the
// table-driven TabWalk (fhmc_tab.cuh): same tiles, same deferred queue, same drain.
template <int NSEL, bool COMPACT, class W>
__device__ __forceinline__ void prod2_tiles(const SweepArgs &a, const W &w, double *s_tab)
{
    using LY = typename W::LY;
    // deferred fallback queue, as in k_sweep_fast
    long long *queue = reinterpret_cast<long long *>(s_tab + 64);
    int *q_count = reinterpret_cast<int *>(queue + LY::QN);
    if (threadIdx.x == 0) *q_count = 0;
    __syncthreads();
    int top = 0;   // COMPACT: largest phase count this thread has written
    auto drain = [&]() {
        const int cnt = *q_count;
        for (int k = threadIdx.x >> 5; k < cnt; k += FHMC_CTA / 32) {   // one queued state point per warp, as in k_sweep_fast
            const long long qs = queue[k];
            const double qm = a.st.mu1[(qs / a.st.mu1_div) % a.st.n_mu1];
            if (!COMPACT) {
                run_generic_point_warp<false>(a, s_tab, threadIdx.x & 31, qm, a.d.beta_ref, a.d.dmu_ref, qs);
            } else {
                // general evaluator into this warp's scratch record, then the compact record from it
                const long long slot = (long long)blockIdx.x * (FHMC_CTA / 32) + (threadIdx.x >> 5);
                run_generic_point_warp<false>(a, s_tab, threadIdx.x & 31, qm, a.d.beta_ref, a.d.dmu_ref, slot);
                __syncwarp();
                const unsigned st_ = a.out.status[slot];
                const int nM = a.out.nphase[slot];
                const int lane = threadIdx.x & 31;
                const int Pe = ((st_ & FHMC_ST_CODE_MASK) == FHMC_OK) ? min(max(nM, 0), a.d.pmax) : 0;
                if (lane < Pe)
                    w.put_phase(qs, lane, a.out.fe[slot * a.d.pmax + lane], 1.0, a.out.avg + (slot * a.d.pmax + lane) * NSEL);
                if (lane == 0) {
                    // (a point that went through the walk first may have left fe / avg in slots the final record does not
                    // have: always blank the dead slots of a re-evaluated point)
                    w.put_compact_tail_fill(qs, st_, nM, a.out.bounds + slot * a.d.pmax * 2, true);
                }
                top = max(top, Pe);
                __syncwarp();
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) *q_count = 0;
        __syncthreads();
    };
    int tile_no = 0;
    const long long S = a.st.n_states;
    int lst0[W::USES_LST ? W::LST : 1], lst1[W::USES_LST ? W::LST : 1];
    for (long long base = (long long)blockIdx.x * (2 * FHMC_CTA); base < S; base += (long long)gridDim.x * (2 * FHMC_CTA)) {
        const long long sp0 = base + threadIdx.x, sp1 = sp0 + FHMC_CTA;
        if (sp0 < S) {
            typename W::PS p0, p1;
            p0.lst = lst0;
            p1.lst = lst1;
            w.init(p0, sp0, a.st.mu1[(sp0 / a.st.mu1_div) % a.st.n_mu1]);
            bool ok0, ok1 = true;
            if (sp1 < S) {
                w.init(p1, sp1, a.st.mu1[(sp1 / a.st.mu1_div) % a.st.n_mu1]);
                if (!((p0.fl | p1.fl) & W::F_BAD)) {
                    w.walk2(p0, p1);
                } else {
                    if (!(p0.fl & W::F_BAD)) w.walk1(p0);
                    if (!(p1.fl & W::F_BAD)) w.walk1(p1);
                }
                ok0 = w.finish(p0);
                ok1 = w.finish(p1);
                if (COMPACT && ok1) top = max(top, p1.P);
            } else {
                if (!(p0.fl & W::F_BAD)) w.walk1(p0);
                ok0 = w.finish(p0);
            }
            if (COMPACT && ok0) top = max(top, p0.P);
            if (!ok0) queue[atomicAdd(q_count, 1)] = sp0;   // anything unusual: defer to the generic evaluator
            if (!ok1) queue[atomicAdd(q_count, 1)] = sp1;
        }
        if ((++tile_no & 1) == 0) {   // a tile can queue 2 * FHMC_CTA entries
            __syncthreads();
            if (*q_count > LY::QN - 4 * FHMC_CTA) drain();   // uniform across the CTA (read after the barrier)
        }
    }
    __syncthreads();
    drain();
    if (COMPACT && a.c.max_nphase) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
        if ((threadIdx.x & 31) == 0 && top > 0) atomicMax(a.c.max_nphase, top);
    }
}

template <int NSEL, bool SEL0N, bool COMPACT = false>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_prod2(const __grid_constant__ SweepArgs a)
{
    using W = ProdWalk<NSEL, SEL0N, COMPACT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FastCtx cx = fast_prepare<NSEL, SEL0N, 0, 1, 2>(a, smem_raw);
    double *s_tab = cx.s_tab;
    PointEval<1, false> pe(a, a.blob, threadIdx.x & 31, s_tab);   // rare paths (repair) read HBM/L2
    const W w(a, cx, &pe, a.d.smooth, pe.tab);
    prod2_tiles<NSEL, COMPACT, W>(a, w, s_tab);
}

template <int NSEL, bool SEL0N, bool COMPACT = false>
static int launch_prod2(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream, int *grid_out = nullptr, bool dry = false)
{
    const size_t smem = fast_smem_bytes<NSEL, SEL0N, 0, 1, 2>(args.d.n_pad);
    if (smem > (size_t)smem_optin) return -1;
    auto kern = k_sweep_prod2<NSEL, SEL0N, COMPACT>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FHMC_CTA, smem), "occupancy query")) return 1;
    if (occ < 1) return -1;
    const long long ntiles = (args.st.n_states + 2 * FHMC_CTA - 1) / (2 * FHMC_CTA);
    long long grid = (long long)sm_count * occ;
    if (grid_out) *grid_out = (int)grid;   // (upper bound: what a scratch record per resident warp has to cover)
    if (dry) return 0;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, FHMC_CTA, smem, stream>>>(args);
    note_kernel(COMPACT ? "k_sweep_prod2<compact>" : "k_sweep_prod2");
    return check_cuda(cudaGetLastError(), "k_sweep_prod2 launch");
}

}  // namespace fhmc
