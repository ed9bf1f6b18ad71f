// fhmc_tab.cuh -- pure-mu sweeps on PRECOMPUTED per-histogram tables (k_sweep_prod2<..., TAB = true>).
//
// For a pure chemical-potential sweep on uniformly spaced N the reweighted array is u_i = lnPI_i + s N_i, and with the tilt
// a = -s dN (per bin)   u_j - u_k = (j - k) (S_jk - a),  S_jk = (lnPI_j - lnPI_k) / (j - k)  (a chord slope of the histogram).
// Whether bin j is a strict extremum over its +-smooth window (scipy argrelextrema 'clip', GH:329-330) therefore depends on
// the tilt only through an interval test:  j is a maximum iff max_d SR_d < a < min_d SL_d, a minimum iff
// max_d SL_d < a < min_d SR_d (SL_d / SR_d: chord slopes to the clipped neighbours at distance d).  The complete outcome of
// relextrema() + the phase bounds of thermo() (GH:317-415, 498-520) is thus PIECEWISE CONSTANT in the tilt and changes
// only at the interval endpoints of the bins -- a few per bin, known from the histogram alone.
//
// fhmc_mu_tables_build() (once per histogram) sorts those endpoints (plus the edge slopes of the upper hull, at which the
// position of max_i u_i moves) and evaluates ONE representative state point inside every elementary interval with the
// general evaluator (PointEval<32>::run, the parity-pinned path), keeping {number of phases, phase bounds, last maximum,
// hull vertex} per interval.  The sweep kernel then does, per state point: one binary search over the sorted endpoints,
// a margin test (the tilt must be further from both neighbouring endpoints than rounding of fl(lnPI + fl(s N)) - c can
// move a comparison: then the reference's own comparisons come out as in exact arithmetic, i.e. as at the
// representative), and a walk over the bins that only SUMS: product-form blocks (fhmc_prod.cuh) with the per-phase sums
// flushed at the known boundary bins.  No extremum test, no repair(), no re-test on the normalised values in the hot
// kernel.  A state point that fails the margin test, lands in an interval whose representative was not a plain case
// (the reference raises, gap filling, ties, more phases than the record holds) or shows an underflowing anchor goes to
// the general evaluator through the deferred queue, exactly like the non-plain points of k_sweep_prod2.
//
// The product tables themselves (packed rows, P_i = exp(lnPI_i - A_seg), P_i X_q(i), anchors) are built once as well:
// the build runs fast_prepare() in one CTA and saves its shared-memory image; every sweep CTA stages that image with two
// TMA bulk copies instead of recomputing it (r01b: ~40 us per launch and CTA, 105 us per 2^17-point chunk against 64 us).
#pragma once
#include <stdlib.h>
#include <string.h>

#include "fhmc_prod.cuh"

namespace fhmc {

#define FHMC_TAB_MAGIC 0x46484d54u   // 'FHMT'
#define FHMC_TAB_REC_I16 32          // int16 words per interval record (64 bytes)
// record words
#define FHMC_TR_VALID 0
#define FHMC_TR_NPHASE 1
#define FHMC_TR_HIDX 2     // bin of max_i u_i (hull vertex): the shift of the sums
#define FHMC_TR_LASTMAX 3  // last entry of the maxima list (is_safe, GH:586-591)
#define FHMC_TR_CNTM 4     // windowed maxima / minima found (what PointEval::repair() tests against the caller's pmax)
#define FHMC_TR_CNTMIN 5
#define FHMC_TR_NMIN 6     // entries of the minima list
#define FHMC_TR_BOUNDS 8   // nphase x {left, right}

struct MuTabHeader {   // 256 bytes at the start of the tables buffer; the static part is written by the build's first kernel
    unsigned magic;
    int n, n_pad, smooth, n_sel, sel0n, sel_row[2];
    int ep_cap;        // capacity of the endpoint / record arrays
    int n_ep;          // finite endpoints (device)
    int bad;           // tables unusable: non-finite ln(PI) (device)
    int regA_bytes, regB_bytes, regB_off;   // shared-memory image: region A at offset 0, region B at regB_off
    int pad0;
    long long off_imgA, off_imgB, off_iv, off_raw, off_ep, off_rec, off_mu, off_scratch;
    double sdn_lim, lmax, dN, Na;
    double pad1[8];
};
static_assert(sizeof(MuTabHeader) <= 256, "header must fit its slot");

// rounding margin of a state point's comparisons, in tilt units (see the file header): 4 x vmargin of fhmc_prod.cuh
__device__ __forceinline__ double tab_margin(double lmax, double s_abs, double Na) { return 4.0 * (1.8e-15 * (lmax + s_abs * Na) + 1e-14); }

// ---------------------------------------------------------------------------------------------------------------------
// build, kernel 1: shared-memory image by fast_prepare() itself + the tilt intervals of every bin + hull edge slopes
// ---------------------------------------------------------------------------------------------------------------------
template <int NSEL, bool SEL0N>
__global__ void __launch_bounds__(FHMC_CTA, 1) k_tab_image(const __grid_constant__ SweepArgs a, unsigned char *tables, MuTabHeader h0)
{
    using LY = FastLayout<NSEL, SEL0N, 0, 1, 2>;
    constexpr int PK = LY::PK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FastCtx cx = fast_prepare<NSEL, SEL0N, 0, 1, 2>(a, smem_raw);
    MuTabHeader *h = reinterpret_cast<MuTabHeader *>(tables);
    const int n = a.d.n, npad = a.d.n_pad;
    const double *pk = reinterpret_cast<const double *>(smem_raw);
    const double dN = pk[PK + 1] - pk[1];
    const double Na = fmax(fabs(pk[1]), fabs(pk[(size_t)(n - 1) * PK + 1]));
    if (threadIdx.x == 0) {
        h0.sdn_lim = cx.sdn_lim;
        h0.lmax = cx.lmax;
        h0.dN = dN;
        h0.Na = Na;
        h0.n_ep = 0;
        h0.bad = 0;
        *h = h0;
    }
    __syncthreads();
    // image: region A = packed rows + hull-slope row, region B = product blocks .. group keys
    {
        const uint4 *srcA = reinterpret_cast<const uint4 *>(smem_raw);
        uint4 *dstA = reinterpret_cast<uint4 *>(tables + h0.off_imgA);
        for (int i = threadIdx.x; i < h0.regA_bytes / 16; i += blockDim.x) dstA[i] = srcA[i];
        const uint4 *srcB = reinterpret_cast<const uint4 *>(smem_raw + h0.regB_off);
        uint4 *dstB = reinterpret_cast<uint4 *>(tables + h0.off_imgB);
        for (int i = threadIdx.x; i < h0.regB_bytes / 16; i += blockDim.x) dstB[i] = srcB[i];
    }
    // per-bin tilt intervals {max: lo, hi; min: lo, hi} (+inf, +inf when the bin can never be that kind of extremum), with
    // the slack of fast_prepare(); bins 0 and n-1 are never strict extrema in 'clip' mode
    double *iv = reinterpret_cast<double *>(tables + h0.off_iv);
    double *raw = reinterpret_cast<double *>(tables + h0.off_raw);
    const double slack = 4.0 * (1.8e-15 * (cx.lmax + 4.5 * Na / dN) + 1e-300) + 1e-12 * fabs(dN);
    bool bad = !(dN > 0.0);   // (N ascending: the hull search and the tilt sign convention assume it)
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double lj = pk[(size_t)j * PK];
        if (!(fabs(lj) < CUDART_INF)) bad = true;
        double o[4] = {CUDART_INF, CUDART_INF, CUDART_INF, CUDART_INF};
        if (j > 0 && j < n - 1) {
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
            if (maxSR <= minSL + slack) { o[0] = maxSR; o[1] = minSL; }
            if (maxSL <= minSR + slack) { o[2] = maxSL; o[3] = minSR; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { iv[4 * (size_t)j + k] = o[k]; raw[4 * (size_t)j + k] = o[k]; }
    }
    // hull edge slopes (per N) -> per bin; the rest of the raw array is +inf
    const double *slope = reinterpret_cast<const double *>(smem_raw) + (size_t)npad * PK;
    for (int k = threadIdx.x; k < h0.ep_cap - 4 * n; k += blockDim.x)
        raw[4 * (size_t)n + k] = (k < a.d.hull_len - 1) ? slope[k] * dN : CUDART_INF;
    if (bad) h->bad = 1;
}

// build, kernel 2: rank sort of the raw endpoints (ties broken by position); +inf entries end up behind the finite ones
static __global__ void __launch_bounds__(256) k_tab_rank(unsigned char *tables)
{
    MuTabHeader *h = reinterpret_cast<MuTabHeader *>(tables);
    const int E = h->ep_cap;
    const double *raw = reinterpret_cast<const double *>(tables + h->off_raw);
    double *ep = reinterpret_cast<double *>(tables + h->off_ep);
    __shared__ double tile[1024];
    const int me = blockIdx.x * blockDim.x + threadIdx.x;
    const double v = (me < E) ? raw[me] : CUDART_INF;
    int rank = 0;
    for (int base = 0; base < E; base += 1024) {
        __syncthreads();
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) tile[k] = (base + k < E) ? raw[base + k] : CUDART_INF;
        __syncthreads();
        const int lim = min(1024, E - base);
        for (int k = 0; k < lim; ++k) {
            const double w = tile[k];
            rank += (w < v) || (w == v && base + k < me);
        }
    }
    if (me < E) {
        ep[rank] = v;
        if (v < CUDART_INF) atomicAdd(&h->n_ep, 1);
    }
}

// build, kernel 3: one representative mu per elementary interval (k = number of endpoints <= tilt)
static __global__ void __launch_bounds__(256) k_tab_mu(unsigned char *tables, double mu1_ref, double beta_ref)
{
    const MuTabHeader *h = reinterpret_cast<const MuTabHeader *>(tables);
    const double *ep = reinterpret_cast<const double *>(tables + h->off_ep);
    double *mu = reinterpret_cast<double *>(tables + h->off_mu);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > h->ep_cap) return;
    const int ne = h->n_ep, kk = min(k, ne);
    double arep;
    if (ne == 0) arep = 0.0;
    else if (kk == 0) arep = ep[0] - 1e-3;
    else if (kk == ne) arep = ep[ne - 1] + 1e-3;
    else arep = 0.5 * (ep[kk - 1] + ep[kk]);
    mu[k] = mu1_ref + (-arep / h->dN) / beta_ref;
}

// build, kernel 4: the general evaluator's records of the representatives -> interval records
static __global__ void __launch_bounds__(128) k_tab_records(unsigned char *tables, fhmc_sweep_out out, int pmax, double mu1_ref, double beta_ref,
                                                     const double *hull_slope, const double *hull_idx, int hull_len)
{
    const MuTabHeader *h = reinterpret_cast<const MuTabHeader *>(tables);
    const double *ep = reinterpret_cast<const double *>(tables + h->off_ep);
    const double *iv = reinterpret_cast<const double *>(tables + h->off_iv);
    const double *mu = reinterpret_cast<const double *>(tables + h->off_mu);
    short *rec = reinterpret_cast<short *>(tables + h->off_rec);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > h->ep_cap) return;
    short *r = rec + (size_t)k * FHMC_TAB_REC_I16;
    for (int q = 0; q < FHMC_TAB_REC_I16; ++q) r[q] = 0;
    const int ne = h->n_ep, n = h->n, last = n - 1;
    if (k > ne || h->bad) return;
    const double s = __dmul_rn(__dsub_rn(mu[k], mu1_ref), beta_ref);
    const double av = -(s * h->dN);
    const double dl = tab_margin(h->lmax, fabs(s), h->Na);
    const double lo = k > 0 ? ep[k - 1] : -CUDART_INF, hi = k < ne ? ep[k] : CUDART_INF;
    if (!(av - lo > dl && hi - av > dl)) return;   // interval narrower than the margins
    const unsigned st = out.status[k];
    const int P = out.nphase[k], nm = out.nmin[k];
    if ((st & FHMC_ST_CODE_MASK) != FHMC_OK || (st & FHMC_ST_GAP_FILL) || P < 1 || P > FHMC_COMPACT_PMAX || P > pmax) return;
    const int *bl = out.bounds + (size_t)k * pmax * 2, *ml = out.max_idx + (size_t)k * pmax, *mn = out.min_idx + (size_t)k * (pmax + 1);
    // phases tile [0, n) with no empty phase
    int prev = 0;
    for (int p = 0; p < P; ++p) {
        if (bl[2 * p] != prev || bl[2 * p + 1] <= bl[2 * p]) return;
        prev = bl[2 * p + 1];
    }
    if (prev != n) return;
    // cross-check with the interval tests: the windowed extrema at this tilt are exactly the interior entries of the lists
    int cM = 0, cm = 0;
    for (int j = 1; j < last; ++j) {
        cM += (iv[4 * (size_t)j] < av) && (av < iv[4 * (size_t)j + 1]);
        cm += (iv[4 * (size_t)j + 2] < av) && (av < iv[4 * (size_t)j + 3]);
    }
    int lM = 0, lm = 0;
    for (int q = 0; q < P; ++q) {
        const int j = ml[q];
        if (j > 0 && j < last) { if (!((iv[4 * (size_t)j] < av) && (av < iv[4 * (size_t)j + 1]))) return; ++lM; }
    }
    for (int q = 0; q < nm; ++q) {
        const int j = mn[q];
        if (j > 0 && j < last) { if (!((iv[4 * (size_t)j + 2] < av) && (av < iv[4 * (size_t)j + 3]))) return; ++lm; }
    }
    if (cM + cm > 0) {
        // extrema found by the window test; the branches that need values of the normalised array (gap filling) were
        // refused above, a re-detection on the normalised array is not expected this far from every endpoint
        if (lM != cM || lm != cm || (st & FHMC_ST_SLOW_PATH)) return;
        if (!((cM > 0 && cm > 0) || (cM == 1 && cm == 0) || (cM == 0 && cm == 1))) return;
    } else {
        // monotone at this tilt (GH:382-386): unique maximum at one edge, unique minimum at the other, one phase
        if (P != 1 || nm != 1 || lM != 0 || lm != 0) return;
        if (!((ml[0] == 0 && mn[0] == last) || (ml[0] == last && mn[0] == 0))) return;
    }
    // hull vertex of this tilt, as ProdWalk::init() finds it
    int hl = 0, hh = hull_len - 1;
    const double neg_s = -s;
    while (hl < hh) {
        const int mid = (hl + hh) >> 1;
        if (hull_slope[mid] > neg_s) hl = mid + 1; else hh = mid;
    }
    r[FHMC_TR_NPHASE] = (short)P;
    r[FHMC_TR_HIDX] = (short)(int)hull_idx[hl];
    r[FHMC_TR_LASTMAX] = (short)ml[P - 1];
    r[FHMC_TR_CNTM] = (short)cM;
    r[FHMC_TR_CNTMIN] = (short)cm;
    r[FHMC_TR_NMIN] = (short)nm;
    for (int p = 0; p < P; ++p) {
        r[FHMC_TR_BOUNDS + 2 * p] = (short)bl[2 * p];
        r[FHMC_TR_BOUNDS + 2 * p + 1] = (short)bl[2 * p + 1];
    }
    r[FHMC_TR_VALID] = 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// sweep side
// ---------------------------------------------------------------------------------------------------------------------
struct TabCtx {
    const double *ep;
    const short *rec;
    int n_ep;
    double dN, Na;
};

// Stage the saved image (all threads of the CTA; ends with a barrier) and rebuild the FastCtx fast_prepare() would return.
template <int NSEL, bool SEL0N>
__device__ __forceinline__ FastCtx tab_prepare(const SweepArgs &a, unsigned char *smem_raw, TabCtx &tc, bool &usable)
{
    using LY = FastLayout<NSEL, SEL0N, 0, 1, 2>;
    constexpr int PK = LY::PK;
    const unsigned char *tables = static_cast<const unsigned char *>(a.d.mu_tables);
    const MuTabHeader *h = reinterpret_cast<const MuTabHeader *>(tables);
    const int npad = a.d.n_pad;
    double *pk = reinterpret_cast<double *>(smem_raw);
    double *stage = pk + (size_t)npad * PK;
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage + npad);
    double *s_tab = reinterpret_cast<double *>(bar + 2);
    stage_exp_table(s_tab);
    const int regA = h->regA_bytes, regB = h->regB_bytes, offB = h->regB_off;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(regA + regB));
        const uint32_t chunk = 32768;
        for (uint32_t off = 0; off < (uint32_t)regA; off += chunk)
            tma_bulk_g2s(smem_raw + off, tables + h->off_imgA + off, min(chunk, (uint32_t)regA - off), bar);
        for (uint32_t off = 0; off < (uint32_t)regB; off += chunk)
            tma_bulk_g2s(smem_raw + offB + off, tables + h->off_imgB + off, min(chunk, (uint32_t)regB - off), bar);
    }
    mbar_wait(bar, 0);
    const int n = a.d.n, nb = (n - 2) / 4, nseg = (nb + LY::SEGB - 1) / LY::SEGB;
    double *prod = reinterpret_cast<double *>(smem_raw + fast_base_bytes<PK, LY::QN>(npad));
    double *anch = prod + (size_t)nb * LY::BW;
    FastCtx cx;
    cx.s_pk = smem_u32(pk);
    cx.s_slope = smem_u32(stage);
    cx.s_tab = s_tab;
    cx.g_hidx = a.blob + (size_t)a.d.hull_row * npad + npad;
    cx.H = a.d.hull_len;
    cx.s_prod = smem_u32(prod);
    cx.s_anch = smem_u32(anch);
    cx.s_gkey = smem_u32(anch + nseg + 2);
    cx.sdn_lim = h->sdn_lim;
    cx.lmax = h->lmax;
    // exp(A_{g+1} - A_g): the running factor crosses a segment border by one multiplication (TabWalk::next_anchor); 0 where
    // the ratio leaves the fp64 range.  Kept where fast_prepare() keeps the group keys, which this kernel never reads.
    {
        double *ratio = anch + nseg + 2;
        for (int g = threadIdx.x; g < nseg; g += blockDim.x) {
            const double e = (g + 1 < nseg) ? exp(anch[g + 1] - anch[g]) : 0.0;
            ratio[g] = (e > 1e-280 && e < 1e280) ? e : 0.0;
        }
        __syncthreads();
    }
    tc.ep = reinterpret_cast<const double *>(tables + h->off_ep);
    tc.rec = reinterpret_cast<const short *>(tables + h->off_rec);
    tc.n_ep = h->n_ep;
    tc.dN = h->dN;
    tc.Na = h->Na;
    // tables built for another histogram shape: every state point goes to the general evaluator (correct, slow)
    usable = h->magic == FHMC_TAB_MAGIC && !h->bad && h->n == n && h->n_pad == npad && h->smooth == a.d.smooth && h->n_sel == NSEL &&
             h->sel0n == (SEL0N ? 1 : 0) && (NSEL < 1 || h->sel_row[0] == a.d.sel_row[0]) && (NSEL < 2 || h->sel_row[1] == a.d.sel_row[1]) &&
             offB == (int)fast_base_bytes<PK, LY::QN>(npad);
    return cx;
}

template <int NSEL, bool SEL0N>
struct TabWalk : ProdWalk<NSEL, SEL0N, true> {
    using B = ProdWalk<NSEL, SEL0N, true>;
    using LY = typename B::LY;
    using Bin = typename B::Bin;
    static constexpr int NA = B::NA;
    static constexpr uint32_t BWB = B::BWB;
    static constexpr bool USES_LST = false;
    struct PS : B::PS {
        int Bn;             // next boundary bin: the running phase ends before it (n: none left)
        int nph;            // phases of the record
        int ivl;            // elementary tilt interval of this state point
        const short *rec;
    };
    const TabCtx &tc;
    const bool tables_ok;

    __device__ TabWalk(const SweepArgs &a_, const FastCtx &cx_, PointEval<1, false> *pe, int smooth_, uint32_t tab_, const TabCtx &tc_, bool ok)
        : B(a_, cx_, pe, smooth_, tab_), tc(tc_), tables_ok(ok)
    {
    }

    // ---- prologue: interval lookup instead of hull search + key ranges ------------------------------------------------
    // hint: interval index of a neighbouring state point (-1: none); consecutive state points of a sweep mostly share it
    __device__ __forceinline__ void init(PS &p, long long sp, double mu1, int hint = -1) const
    {
        p.s = __dmul_rn(__dsub_rn(mu1, this->a.d.mu1_ref), this->a.d.beta_ref);   // GH:77, evaluated left to right
        p.sp = sp;
        p.cntM = p.cntm = p.P = 0;
        p.rescue = 0;
        p.fl = (this->n < 3 || !tables_ok) ? B::F_BAD : 0u;
        p.Sacc = p.Stot = 0.0;
        p.t = 0.0;
#pragma unroll
        for (int q = 0; q < NA; ++q) p.A[q] = 0.0;
        const double sdn = p.s * tc.dN;
        if (!(fabs(4.0 * sdn) < 200.0)) p.fl |= B::F_BAD;   // extreme tilt: leave it to the general evaluator
        if (fabs(sdn) < this->cx.sdn_lim) p.fl |= B::F_CHAIN;
        const double av = -sdn;
        int lo = 0, hi = tc.n_ep;
        if (hint >= 0) {   // the hinted interval holds the tilt: no search
            const double h_lo = hint > 0 ? __ldg(tc.ep + hint - 1) : -CUDART_INF, h_hi = hint < tc.n_ep ? __ldg(tc.ep + hint) : CUDART_INF;
            if (h_lo <= av && av < h_hi) lo = hi = hint;
        }
        while (lo < hi) {   // number of endpoints <= tilt
            const int mid = (lo + hi) >> 1;
            if (__ldg(tc.ep + mid) <= av) lo = mid + 1; else hi = mid;
        }
        p.ivl = lo;
        const double dl = tab_margin(this->cx.lmax, fabs(p.s), tc.Na);
        const double e_lo = lo > 0 ? __ldg(tc.ep + lo - 1) : -CUDART_INF, e_hi = lo < tc.n_ep ? __ldg(tc.ep + lo) : CUDART_INF;
        p.rec = tc.rec + (size_t)lo * FHMC_TAB_REC_I16;
        const int4 head = __ldg(reinterpret_cast<const int4 *>(p.rec));   // {valid, nphase}, {hidx, lastmax}, {cntM, cntm}, {nmin, -}
        p.nph = head.x >> 16;
        // (the capacity rules of PointEval::repair() for the caller's pmax: such a state point is a capacity error)
        const bool cap = p.nph > this->pmax || (head.z & 0xffff) > this->pmax - 1 || (head.z >> 16) > this->pmax || (head.w & 0xffff) > this->pmax + 1;
        if (!(av - e_lo > dl && e_hi - av > dl) || (head.x & 0xffff) != 1 || cap) p.fl |= B::F_BAD;
        p.r1 = exp(sdn);
        p.r4 = (p.r1 * p.r1) * (p.r1 * p.r1);   // (3 ulp instead of 1: <= 1e-14 over the 32 blocks between two anchors)
        p.Bn = this->n;
        if (!(p.fl & B::F_BAD)) {
            double Nm;
            p.Mq = shift_for_max(this->load_u(p, head.y & 0xffff, Nm));
            p.Bn = next_boundary(p);
            Bin b0;
            this->load_bin(p, 0, b0);
            this->accumulate(p, b0);
        }
    }
    // t at the first bin of segment g: carried over from the end of segment g - 1 (t has been advanced by r4 per block, so
    // only the anchors' ratio is missing) when that segment was walked in product form and t is a comfortable normal
    // number, else a true exp like ProdWalk::anchor().  Rounding: one more multiplication per 128 bins.
    __device__ __forceinline__ bool next_anchor(PS &p, int g, int i, bool prev_usable) const
    {
        if (g > 0 && prev_usable) {
            const double e = lds_f64(this->cx.s_gkey + 8u * (uint32_t)(g - 1));
            const double t = p.t * e;
            if (t > 1e-250 && t < 1e250) {
                p.t = t;
                return true;
            }
        }
        return this->anchor(p, g, i);
    }
    // right end of the running phase p.P (the last phase ends at n and is closed by finish())
    __device__ __forceinline__ int next_boundary(const PS &p) const
    {
        return (p.P + 1 < p.nph) ? (int)__ldg(p.rec + FHMC_TR_BOUNDS + 2 * p.P + 1) : 0x7fffffff;
    }
    // block that holds the next boundary bin (blocks cover bins 1 .. 4 nb); beyond the blocks: a huge index
    __device__ __forceinline__ int boundary_block(const PS &p) const { return (p.Bn - 1) >> 2; }
    __device__ __forceinline__ void cross(PS &p, int bin) const
    {
        if (bin == p.Bn) {
            this->flush(p);
            p.Bn = next_boundary(p);
        }
    }
    // the block holding a boundary: term by term, e_{i+k} = P_{i+k} t r1^k, the running sums closed at the boundary bin
    __device__ __forceinline__ void split_block(PS &p, uint32_t pb, int ib) const
    {
        double tk = p.t;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cross(p, ib + k);
            p.Sacc = fma(lds_f64(pb + 16u + 8u * k), tk, p.Sacc);
#pragma unroll
            for (int q = 0; q < NSEL; ++q) p.A[q] = fma(lds_f64(pb + 48u + 32u * q + 8u * k), tk, p.A[q]);
            tk *= p.r1;
        }
    }
    // one block with a true exp per bin (segment whose factor t is clamped or unusable)
    __device__ __forceinline__ void exact_block(PS &p, int ib) const
    {
        Bin c0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cross(p, ib + k);
            this->load_bin(p, ib + k, c0);
            this->accumulate(p, c0);
        }
    }
    // blocks [b, bend) of one segment for one point
    __device__ __forceinline__ void segment_single(PS &p, bool usable, int b, int bend, int i, uint32_t pb) const
    {
        const double r2 = p.r1 * p.r1;
        for (; b < bend; ++b, i += 4, pb += BWB) {
            if (!usable) {
                exact_block(p, i);
                continue;
            }
            if (b == boundary_block(p)) {
                split_block(p, pb, i);
            } else {
                double tb[4 * (1 + NA)];
                this->load_block(pb, tb);
                this->fast_block_regs(p, tb, r2);
            }
            p.t *= p.r4;
        }
    }
    __device__ __forceinline__ void walk1(PS &p) const
    {
        const int nb = (this->n - 2) / 4;
        uint32_t pb = this->cx.s_prod;
        int b = 0, i = 1;
        bool u = false;
        for (int g = 0; b < nb; ++g) {
            const int bend = min(nb, b + LY::SEGB);
            u = next_anchor(p, g, i, u);
            segment_single(p, u, b, bend, i, pb);
            i += 4 * (bend - b);
            pb += BWB * (uint32_t)(bend - b);
            b = bend;
        }
    }
    // The hot loop: between two boundary blocks nothing but 6 LDS.128 + 24 DFMA + 2 DMUL per block for the two points.
    __device__ __forceinline__ void walk2(PS &p0, PS &p1) const
    {
        const int nb = (this->n - 2) / 4;
        const double r2a = p0.r1 * p0.r1, r2b = p1.r1 * p1.r1;
        uint32_t pb = this->cx.s_prod;
        int b = 0, i = 1;
        bool ua = false, ub = false;
        for (int g = 0; b < nb; ++g) {
            const int bend = min(nb, b + LY::SEGB);
            ua = next_anchor(p0, g, i, ua);
            ub = next_anchor(p1, g, i, ub);
            if (!(ua & ub)) {   // rare: walk this segment point by point
                segment_single(p0, ua, b, bend, i, pb);
                segment_single(p1, ub, b, bend, i, pb);
                i += 4 * (bend - b);
                pb += BWB * (uint32_t)(bend - b);
                b = bend;
                continue;
            }
            while (b < bend) {
                const int stop = min(bend, min(boundary_block(p0), boundary_block(p1)));
                int cnt = stop - b;
                i += 4 * cnt;
                b = stop;
                // software pipeline: the table entries of a block are fetched while the previous block is summed (an iteration
                // that starts with its own loads waits out the shared-memory latency first: 360 cycles per two blocks at two
                // warps per scheduler against 104 cycles of fp64 pipe).  The last prefetch of a run reads at most two blocks
                // past it: still inside the product rows' allocation (fast_smem_bytes keeps three spare blocks).
                if (cnt > 0) {
                    double ta[4 * (1 + NA)], tc2[4 * (1 + NA)];
                    this->load_block(pb, ta);
#pragma unroll 1
                    for (; cnt >= 2; cnt -= 2, pb += 2 * BWB) {
                        this->load_block(pb + BWB, tc2);
                        this->fast_block_regs(p0, ta, r2a);
                        this->fast_block_regs(p1, ta, r2b);
                        p0.t *= p0.r4;
                        p1.t *= p1.r4;
                        this->load_block(pb + 2 * BWB, ta);
                        this->fast_block_regs(p0, tc2, r2a);
                        this->fast_block_regs(p1, tc2, r2b);
                        p0.t *= p0.r4;
                        p1.t *= p1.r4;
                    }
                    if (cnt > 0) {
                        this->fast_block_regs(p0, ta, r2a);
                        this->fast_block_regs(p1, ta, r2b);
                        p0.t *= p0.r4;
                        p1.t *= p1.r4;
                        pb += BWB;
                    }
                }
                if (b < bend) {   // the block at `b` holds a boundary of at least one of the two points
                    double ta[4 * (1 + NA)];
                    this->load_block(pb, ta);
                    if (b == boundary_block(p0)) split_block(p0, pb, i); else this->fast_block_regs(p0, ta, r2a);
                    if (b == boundary_block(p1)) split_block(p1, pb, i); else this->fast_block_regs(p1, ta, r2b);
                    p0.t *= p0.r4;
                    p1.t *= p1.r4;
                    pb += BWB;
                    i += 4;
                    ++b;
                }
            }
        }
    }

    // ---- epilogue: tail bins, last flush, record -----------------------------------------------------------------------
    __device__ __forceinline__ bool finish(PS &p) const
    {
        if (p.fl & B::F_BAD) return false;
        const int n = this->n, last = this->last;
        {
            Bin c;
            for (int i = 1 + 4 * ((n - 2) / 4); i <= last; ++i) {
                cross(p, i);
                this->load_bin(p, i, c);
                this->accumulate(p, c);
            }
            this->flush(p);
        }
        if ((p.fl & B::F_BAD) || p.P != p.nph) return false;
        const int nM = p.nph;
        int bb[2 * FHMC_COMPACT_PMAX];
        {
            const int4 lo4 = __ldg(reinterpret_cast<const int4 *>(p.rec + FHMC_TR_BOUNDS)), hi4 = __ldg(reinterpret_cast<const int4 *>(p.rec + FHMC_TR_BOUNDS) + 1);
            const int wds[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
            for (int q = 0; q < FHMC_COMPACT_PMAX; ++q) { bb[2 * q] = wds[q] & 0xffff; bb[2 * q + 1] = (wds[q] >> 16) & 0xffff; }
        }
        unsigned flags = 0;
        const double c = add_shift(p.Mq, log(p.Stot));
        double Nd;
        // phases whose weight underflowed next to the global maximum: integrate them about their own maximum (as ProdWalk)
        for (int ph = 0; p.rescue != 0 && ph < nM; ++ph) {
            if (!((p.rescue >> ph) & 1u)) continue;
            const int left = bb[2 * ph], right = bb[2 * ph + 1];
            double mlx = -CUDART_INF;
            for (int j = left; j < right; ++j) mlx = fmax(mlx, this->load_u(p, j, Nd));
            const int Mp = shift_for_max(mlx);
            double Sp = 0.0, Ap[NA];
#pragma unroll
            for (int q = 0; q < NA; ++q) Ap[q] = 0.0;
            for (int j = left; j < right; ++j) {
                Bin b;
                this->load_bin(p, j, b);
                const double e = exp_scaled(b.u, Mp, this->tab);
                Sp += e;
                if (SEL0N) Ap[0] = fma(e, b.N, Ap[0]);
#pragma unroll
                for (int q = 0; q < B::NX; ++q) Ap[q + (SEL0N ? 1 : 0)] = fma(e, b.x[q], Ap[q + (SEL0N ? 1 : 0)]);
            }
            this->put_phase(p.sp, ph, -(add_shift(Mp, log(Sp)) - this->u0(p)), Sp, Ap);
            flags |= FHMC_ST_RESCUED;
        }
        const int lastmax = (int)__ldg(p.rec + FHMC_TR_LASTMAX);
        const double xM = __dsub_rn(this->load_u(p, lastmax, Nd), c), xl = __dsub_rn(this->load_u(p, last, Nd), c);
        if (!(__dsub_rn(xM, xl) < this->a.d.cutoff)) flags |= FHMC_ST_SAFE;
        this->put_compact_tail(p.sp, flags | FHMC_ST_FAST, nM, bb);
        return true;
    }
};

#ifdef FHMC_TAB_PROFILE
__device__ unsigned long long g_tab_prof[8];   // cycles (lane 0 of every warp): init, walk, finish, drain; [4] warp tiles
#define TAB_PROF_T(k) { const long long t_ = clock64(); if (lane == 0) prof[k] += t_ - tp; tp = t_; }
#else
#define TAB_PROF_T(k)
#endif

// Warp-granular tiles: a warp owns 64 consecutive state points (lane l: points l and 32 + l of the warp tile), warp tiles are
// dealt round-robin over all warps of the grid, and NOTHING synchronises the warps of a CTA after the tables are staged: the
// per-warp fallback queue (at most the 64 points of the tile) is drained by the warp itself.  With CTA-wide tiles and a
// barrier every second tile the eight warps of a CTA ran their prologues, product loops and epilogues in lockstep -- fp64
// pipe saturated in the loop phase (math-pipe throttle) and idle in the latency-bound phases (r2f capture: 55 % active).
// Independent warps drift apart (a warp that reaches the loop first finds the pipe free and gets further ahead), so one
// warp's epilogue overlaps the others' loops.
// IDX: the state points are list[0 .. *count) of SweepArgs::c (what k_sweep_cell left over) instead of 0 .. n_states - 1.
template <int NSEL, class W, bool IDX = false>
__device__ __forceinline__ void tab2_warp_tiles(const SweepArgs &a, const W &w, double *s_tab)
{
    using LY = typename W::LY;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    long long *queue = reinterpret_cast<long long *>(s_tab + 64) + wib * 64;   // this warp's 64 entries of the queue area
    static_assert(LY::QN >= (FHMC_CTA / 32) * 64, "queue area too small for the per-warp queues");
    const long long S = IDX ? min((long long)*reinterpret_cast<volatile int *>(a.c.ix_count), a.st.n_states) : a.st.n_states;
    const long long slot = (long long)blockIdx.x * (FHMC_CTA / 32) + wib;          // scratch record of this warp
    const long long nwarps = (long long)gridDim.x * (FHMC_CTA / 32);
    int top = 0;
#ifdef FHMC_TAB_STAGGER
    // seed the phase offsets of the four warps that share a scheduler: warp pairs (w, w + 4) of the two resident CTAs
    {
        const int g = ((wib >> 2) & 1) * 2 + (int)((blockIdx.x / 148u) & 1u);
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)g * FHMC_TAB_STAGGER) __nanosleep(2000);
    }
#endif
    // warp w of CTA c takes warp tiles (w * gridDim + c) + k * nwarps: the tiles of one round go to different SMs first
#ifdef FHMC_TAB_PROFILE
    long long prof[5] = {0, 0, 0, 0, 0}, tp = clock64();
#endif
    for (long long wt = (long long)wib * gridDim.x + blockIdx.x; wt * 64 < S; wt += nwarps) {
        const long long k0 = wt * 64 + lane, k1 = k0 + 32;
        const long long sp0 = (IDX && k0 < S) ? a.c.ix_list[k0] : k0, sp1 = (IDX && k1 < S) ? a.c.ix_list[k1] : k1;
        bool ok0 = true, ok1 = true;
        // (an index list can only name state points of this sweep; anything else is a stale entry and is skipped)
        const bool v0 = k0 < S && (!IDX || (sp0 >= 0 && sp0 < a.st.n_states)), v1 = k1 < S && (!IDX || (sp1 >= 0 && sp1 < a.st.n_states));
        if (v0) {
            typename W::PS p0, p1;
            w.init(p0, sp0, a.st.mu1[(sp0 / a.st.mu1_div) % a.st.n_mu1]);
            if (v1) {
                w.init(p1, sp1, a.st.mu1[(sp1 / a.st.mu1_div) % a.st.n_mu1], p0.ivl);
                TAB_PROF_T(0)
                if (!((p0.fl | p1.fl) & W::F_BAD)) {
                    w.walk2(p0, p1);
                } else {
                    if (!(p0.fl & W::F_BAD)) w.walk1(p0);
                    if (!(p1.fl & W::F_BAD)) w.walk1(p1);
                }
                TAB_PROF_T(1)
                ok0 = w.finish(p0);
                ok1 = w.finish(p1);
                TAB_PROF_T(2)
                if (ok1) top = max(top, p1.P);
            } else {
                if (!(p0.fl & W::F_BAD)) w.walk1(p0);
                ok0 = w.finish(p0);
            }
            if (ok0) top = max(top, p0.P);
        }
#ifdef FHMC_TAB_PROFILE
        if (lane == 0) prof[4] += 1;
#endif
        // anything unusual: the whole warp evaluates it with the general evaluator, point by point
        const unsigned m0 = __ballot_sync(0xffffffffu, !ok0), m1 = __ballot_sync(0xffffffffu, !ok1);
        if (m0 | m1) {
            if (!ok0) queue[__popc(m0 & ((1u << lane) - 1u))] = sp0;
            if (!ok1) queue[__popc(m0) + __popc(m1 & ((1u << lane) - 1u))] = sp1;
            __syncwarp();
            const int cnt = __popc(m0) + __popc(m1);
            for (int k = 0; k < cnt; ++k) {
                const long long qs = queue[k];
                const double qm = a.st.mu1[(qs / a.st.mu1_div) % a.st.n_mu1];
                run_generic_point_warp<false>(a, s_tab, lane, qm, a.d.beta_ref, a.d.dmu_ref, slot);
                __syncwarp();
                const unsigned st_ = a.out.status[slot];
                const int nM = a.out.nphase[slot];
                const int Pe = ((st_ & FHMC_ST_CODE_MASK) == FHMC_OK) ? min(max(nM, 0), a.d.pmax) : 0;
                if (lane < Pe)
                    w.put_phase(qs, lane, a.out.fe[slot * a.d.pmax + lane], 1.0, a.out.avg + (slot * a.d.pmax + lane) * NSEL);
                // (a point that went through the walk first may have left fe / avg in slots the final record does not have:
                // always blank the dead slots of a re-evaluated point)
                if (lane == 0) w.put_compact_tail_fill(qs, st_, nM, a.out.bounds + slot * a.d.pmax * 2, true);
                top = max(top, Pe);
                __syncwarp();
            }
        }
    }
    if (a.c.max_nphase) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
        if (lane == 0 && top > 0) atomicMax(a.c.max_nphase, top);
    }
#ifdef FHMC_TAB_PROFILE
    if (lane == 0)
        for (int k = 0; k < 5; ++k) atomicAdd(&g_tab_prof[k], (unsigned long long)prof[k]);
#endif
}

template <int NSEL, bool SEL0N>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_tab2(const __grid_constant__ SweepArgs a)
{
    using W = TabWalk<NSEL, SEL0N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TabCtx tc;
    bool ok;
    const FastCtx cx = tab_prepare<NSEL, SEL0N>(a, smem_raw, tc, ok);
    double *s_tab = cx.s_tab;
    const W w(a, cx, nullptr, a.d.smooth, smem_u32(s_tab), tc, ok);
    tab2_warp_tiles<NSEL, W>(a, w, s_tab);
}

// the same over an index list (the leftovers of k_sweep_cell): nothing staged when the list is empty
template <int NSEL, bool SEL0N>
__global__ void __launch_bounds__(FHMC_CTA, 2) k_sweep_tab2_idx(const __grid_constant__ SweepArgs a)
{
    using W = TabWalk<NSEL, SEL0N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // launched behind k_sweep_cell with programmatic stream serialisation: wait for that grid (and its list) to be complete
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long cnt = min((long long)*reinterpret_cast<volatile int *>(a.c.ix_count), a.st.n_states);
    if ((long long)blockIdx.x * 64 < cnt) {   // this CTA's first warp tile is tile blockIdx.x (uniform per CTA)
        TabCtx tc;
        bool ok;
        const FastCtx cx = tab_prepare<NSEL, SEL0N>(a, smem_raw, tc, ok);
        double *s_tab = cx.s_tab;
        const W w(a, cx, nullptr, a.d.smooth, smem_u32(s_tab), tc, ok);
        tab2_warp_tiles<NSEL, W, true>(a, w, s_tab);
    }
    // the last CTA through resets the counter and the ticket for the next sweep (every CTA has read the count by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.c.ix_count + 1, 1) == (int)gridDim.x - 1) {
            a.c.ix_count[0] = 0;
            a.c.ix_count[1] = 0;
            __threadfence();
        }
    }
}

template <int NSEL, bool SEL0N>
static int launch_tab2_idx(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const size_t smem = fast_smem_bytes<NSEL, SEL0N, 0, 1, 2>(args.d.n_pad);
    if (smem > (size_t)smem_optin) return -1;
    auto kern = k_sweep_tab2_idx<NSEL, SEL0N>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FHMC_CTA, smem), "occupancy query")) return 1;
    if (occ < 1) return -1;
    const long long ntiles = (args.st.n_states + 63) / 64;   // (the list cannot be longer than the sweep)
    long long grid = (long long)sm_count * occ;
    if (grid > ntiles) grid = ntiles;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(FHMC_CTA);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, args), "k_sweep_tab2_idx launch");
}

template <int NSEL, bool SEL0N>
static int launch_tab2(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream, int *grid_out, bool dry)
{
    const size_t smem = fast_smem_bytes<NSEL, SEL0N, 0, 1, 2>(args.d.n_pad);
    if (smem > (size_t)smem_optin) return -1;
    auto kern = k_sweep_tab2<NSEL, SEL0N>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FHMC_CTA, smem), "occupancy query")) return 1;
    if (occ < 1) return -1;
    const long long ntiles = (args.st.n_states + 2 * FHMC_CTA - 1) / (2 * FHMC_CTA);
    long long grid = (long long)sm_count * occ;
    if (grid_out) *grid_out = (int)grid;
    if (dry) return 0;
    if (grid > ntiles) grid = ntiles;
    if (const char *e = getenv("FHMC_TAB_GRID")) { const long long g = atoll(e); if (g > 0 && g < grid) grid = g; }   // (probe: fewer resident warps)
    kern<<<(unsigned)grid, FHMC_CTA, smem, stream>>>(args);
    note_kernel("k_sweep_tab2<compact>");
    return check_cuda(cudaGetLastError(), "k_sweep_tab2 launch");
}

}  // namespace fhmc
