// fhmc_cell.cu -- dense pure-mu sweeps on TILT CELLS: moment expansions of the per-phase sums (k_sweep_cell).
//
// The table walk (fhmc_tab.cuh) already knows, per elementary tilt interval, the phases [left_p, right_p) of every state point
// in it; what it still does per state point is the sum over the bins,  S_p(s) = sum_{i in p} exp(lnPI_i + s N_i)  and the
// weighted sums  A_pq(s) = sum_i exp(..) X_q(i)  (reweight() + the per-phase averages of thermo(), GH:71-78, 498-554): ~3.8 fp64
// instructions per bin and state point.  With the phase fixed these are entire functions of s.  About a cell centre s_c,
// with N_i = N_c + dN (i - c) on uniformly spaced N,
//      S_p(s_c + d) = exp(M_p + d N_c) * sum_k y^k C_k,    y = d dN R,   C_k = sum_i w_i x_i^k / k!,
//      w_i = exp(lnPI_i + s_c N_i - M_p),   x_i = (i - c) / R in [-1, 1]   (c, R: centre and half width of the phase in bins)
// and the same with w_i X_q(i) for the averages.  All w_i >= 0, so the series is dominated term by term by exp(|y|) sum_i w_i:
// cut after K = 8 terms with |y| <= 0.05 the relative truncation error is below 0.05^8 / 8! * e^0.1 = 1.1e-15.  A cell is
// therefore at most 0.1 / R wide in tilt units (2e-4 for a 1001-bin one-phase histogram): a few hundred cells over the mu
// range of a dense sweep, each costing one pass over the bins to build (k_cell_coef: a CTA per cell) -- against one pass
// PER STATE POINT for the walk.  The coefficients are stored divided by C_0, so the sum is C_0 (1 + eps) with |eps| <= e^0.05 - 1
// and its logarithm ln C_0 + log1p(eps) by a 12-term series: a state point costs, per phase, (1 + n_sel) degree-7 polynomials,
// that series and one reciprocal, whatever the histogram length -- what is left is its record traffic (8 bytes of mu in, 60 bytes
// out per two-phase state point).
//
// What stays exactly as in the table walk: integers (phase count, bounds) come from the interval record, i.e. from the general
// evaluator at the interval's representative; the rounding-margin test of the interval lookup is folded into a safe range per cell
// (k_cell_pieces); the capacity rules for the caller's pmax; is_safe() (decided from u_lastmax - u_last against the cutoff, with a
// margin; GH:586-591); the RESCUED diagnostic bit.  Whatever the cells do not cover (state points outside the range they were built for, intervals
// without a valid record, failed margin tests, is_safe closer to its cutoff than rounding) is appended to an index list and
// walked by k_sweep_tab2_idx -- the parity-pinned table walk with its own fallback to the general evaluator -- in the same
// call.
#include <stddef.h>
#include <string.h>

#include "fhmc_tab.cuh"

namespace fhmc {

#define FHMC_CELL_MAGIC 0x4648434cu   // 'FHCL'
#define FHMC_CELL_K 8                 // terms of the expansion
#define FHMC_CELL_YMAX 0.05           // |y| at a cell's edge

struct CellHeader {   // 256 bytes at the start of the cells buffer
    unsigned magic;
    int n_sel, blk;                       // doubles per phase block
    int iv_cap, piece_cap, block_cap;
    int n_pieces, n_blocks;               // (device)
    int truncated;                        // capacity reached: the upper intervals of the range are not covered (device)
    int grid_n;                           // cells of the lookup grid over [a_lo, a_hi]
    int q_count, q_ticket;                // sweeps: entries of the leftover list / CTAs of the indexed walk that are through; both
                                          // are zero between two sweeps (the last CTA of the walk resets them)
    long long off_iv, off_piece, off_coef, off_pstart, off_grid;
    double a_lo, a_hi;                    // covered tilt range (device)
    double inv_g;                         // grid cells per unit of tilt (device)
    unsigned long long kmin, kmax;        // order-preserving keys of min / max mu of the sweep (k_cell_range; device)
    int sel_row[2], n, smooth;            // what the coefficients were built for
    double pad1[12];
};
static_assert(sizeof(CellHeader) <= 256, "header must fit its slot");

struct CellIv {     // per elementary interval (build only)
    int first, m;   // pieces first .. first + m - 1 (m == 0: not covered)
    int bfirst, nph;
    double a0, inv_w;
};
// One cell.  [safe_lo, safe_hi]: the tilts of the cell that are further from both ends of its elementary interval than the
// rounding margin of the table walk (tab_margin at the largest |s| of the cell, doubled) -- the margin test of a state point is
// one range check.  The words the sweep needs from the interval record travel with the cell.
struct CellPiece {
    double s_c, safe_lo, safe_hi;
    int block, ivl;
    short nph, lastmax, hidx, cntM, cntm, nmin, pad0, pad1;
    double lM, NM, lH, NH;   // ln(PI) and N at the last maximum (is_safe) and at the hull vertex (the shift of the RESCUED bit)
};
static_assert(sizeof(CellIv) == 32 && sizeof(CellPiece) == 80, "cell table strides");

// centre and half width (bins) of the phase [left, right)
__device__ __forceinline__ void cell_geom(int left, int right, double &c, double &R)
{
    c = 0.5 * (double)(left + right - 1);
    R = fmax(0.5 * (double)(right - 1 - left), 0.5);
}

__host__ __device__ constexpr int cell_blk(int nsel) { return 4 + FHMC_CELL_K * (1 + nsel); }

// ---------------------------------------------------------------------------------------------------------------------
// build 0 (device-side range): min / max of the finite mu of a sweep, as order-preserving integer keys
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cell_key(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double cell_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__global__ void __launch_bounds__(256) k_cell_range(const double *mu, long long n, unsigned char *cells)
{
    CellHeader *h = reinterpret_cast<CellHeader *>(cells);
    unsigned long long lo = ~0ull, hi = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = __ldg(mu + i);
        if (fabs(x) < CUDART_INF) {
            const unsigned long long k = cell_key(x);
            lo = min(lo, k);
            hi = max(hi, k);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ unsigned long long s_lo[8], s_hi[8];
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        if (lo <= hi) {
            atomicMin(&h->kmin, lo);
            atomicMax(&h->kmax, hi);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// build 1 (one CTA): pieces per interval for the tilt range of [mu_lo, mu_hi] (dev_range: of the keys k_cell_range left in
// the header), exclusive scans
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_cell_plan(const unsigned char *tables, unsigned char *cells, CellHeader h0, double mu_lo, double mu_hi,
                                                    double mu1_ref, double beta_ref, int dev_range)
{
    bool have_range = true;
    if (dev_range) {
        const CellHeader *hr = reinterpret_cast<const CellHeader *>(cells);
        have_range = hr->kmin <= hr->kmax;
        mu_lo = have_range ? cell_unkey(hr->kmin) : 0.0;
        mu_hi = have_range ? cell_unkey(hr->kmax) : 0.0;
    }
    __syncthreads();   // (every thread has read the keys before thread 0 rewrites the header)
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const double *ep = reinterpret_cast<const double *>(tables + th->off_ep);
    const short *rec = reinterpret_cast<const short *>(tables + th->off_rec);
    CellHeader *h = reinterpret_cast<CellHeader *>(cells);
    CellIv *iv = reinterpret_cast<CellIv *>(cells + h0.off_iv);
    const int ne = th->n_ep, niv = ne + 1;
    // tilt of a state point: a = -s dN, s = fl(fl(mu - mu_ref) beta)
    const double t0 = -(__dmul_rn(__dsub_rn(mu_lo, mu1_ref), beta_ref) * th->dN), t1 = -(__dmul_rn(__dsub_rn(mu_hi, mu1_ref), beta_ref) * th->dN);
    double a_lo = fmin(t0, t1), a_hi = fmax(t0, t1);
    const double pad = 1e-9 * fmax(1.0, fmax(fabs(a_lo), fabs(a_hi)));
    a_lo -= pad;
    a_hi += pad;
    const bool usable = th->magic == FHMC_TAB_MAGIC && !th->bad && niv <= h0.iv_cap && have_range;
    const int per = (niv + (int)blockDim.x - 1) / (int)blockDim.x;
    const int k_lo = (int)threadIdx.x * per, k_hi = min(niv, k_lo + per);
    int np = 0, nb = 0;
    for (int k = k_lo; k < k_hi; ++k) {
        const short *r = rec + (size_t)k * FHMC_TAB_REC_I16;
        const double lo = fmax(k > 0 ? ep[k - 1] : -CUDART_INF, a_lo), hi = fmin(k < ne ? ep[k] : CUDART_INF, a_hi);
        CellIv c;
        c.first = c.m = c.bfirst = c.nph = 0;
        c.a0 = lo;
        c.inv_w = 0.0;
        if (usable && r[FHMC_TR_VALID] == 1 && hi > lo) {
            const int P = r[FHMC_TR_NPHASE];
            double Rmax = 0.5;
            for (int p = 0; p < P; ++p) {
                double cc, R;
                cell_geom(r[FHMC_TR_BOUNDS + 2 * p], r[FHMC_TR_BOUNDS + 2 * p + 1], cc, R);
                Rmax = fmax(Rmax, R);
            }
            const double hmax = 2.0 * FHMC_CELL_YMAX / Rmax;
            const double mm = ceil((hi - lo) / hmax);
            if (mm < 16384.0) {
                c.m = max(1, (int)mm);
                c.nph = P;
                c.inv_w = (double)c.m / (hi - lo);
            }
        }
        iv[k] = c;
        np += c.m;
        nb += c.m * c.nph;
    }
    // exclusive scan of (np, nb) over the threads
    __shared__ int s_p[1024], s_b[1024];
    s_p[threadIdx.x] = np;
    s_b[threadIdx.x] = nb;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
        const int vp = threadIdx.x >= o ? s_p[threadIdx.x - o] : 0, vb = threadIdx.x >= o ? s_b[threadIdx.x - o] : 0;
        __syncthreads();
        s_p[threadIdx.x] += vp;
        s_b[threadIdx.x] += vb;
        __syncthreads();
    }
    int op = s_p[threadIdx.x] - np, ob = s_b[threadIdx.x] - nb;
    __shared__ int s_endp, s_endb, s_trunc;
    if (threadIdx.x == 0) s_endp = s_endb = s_trunc = 0;
    __syncthreads();
    for (int k = k_lo; k < k_hi; ++k) {
        CellIv c = iv[k];
        if (c.m > 0 && (op + c.m > h0.piece_cap || ob + c.m * c.nph > h0.block_cap)) {   // out of room: this interval is left to the walk
            op += c.m;
            ob += c.m * c.nph;
            c.m = 0;
            s_trunc = 1;
        } else if (c.m > 0) {
            c.first = op;
            c.bfirst = ob;
            op += c.m;
            ob += c.m * c.nph;
            atomicMax(&s_endp, op);
            atomicMax(&s_endb, ob);
        }
        iv[k] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        h->n_pieces = s_endp;
        h->n_blocks = s_endb;
        h->truncated = s_trunc;
        h->a_lo = a_lo;
        h->a_hi = a_hi;
        h->inv_g = (double)h0.grid_n / (a_hi - a_lo);
    }
}

// build 2: the pieces of every covered interval (a thread per interval), in tilt order; pstart[] = their lower ends
__global__ void __launch_bounds__(256) k_cell_pieces(const unsigned char *tables, unsigned char *cells, const double *blob, int n_pad)
{
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const CellHeader *h = reinterpret_cast<const CellHeader *>(cells);
    const double *ep = reinterpret_cast<const double *>(tables + th->off_ep);
    const short *rec = reinterpret_cast<const short *>(tables + th->off_rec);
    const CellIv *iv = reinterpret_cast<const CellIv *>(cells + h->off_iv);
    CellPiece *pc = reinterpret_cast<CellPiece *>(cells + h->off_piece);
    double *pstart = reinterpret_cast<double *>(cells + h->off_pstart);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > th->n_ep) return;
    const CellIv c = iv[k];
    if (c.m < 1) return;
    const short *r = rec + (size_t)k * FHMC_TAB_REC_I16;
    const double e_lo = k > 0 ? ep[k - 1] : -CUDART_INF, e_hi = k < th->n_ep ? ep[k] : CUDART_INF;
    const double w = 1.0 / c.inv_w;
    for (int j = 0; j < c.m; ++j) {
        CellPiece p;
        const double p_lo = c.a0 + (double)j * w, p_hi = c.a0 + (double)(j + 1) * w;
        const double a_c = c.a0 + ((double)j + 0.5) * w;
        p.s_c = -a_c / th->dN;
        const double s_abs = fmax(fabs(p_lo), fabs(p_hi)) / th->dN;
        const double dl = 2.0 * tab_margin(th->lmax, s_abs, th->Na);
        p.safe_lo = fmax(p_lo, e_lo + dl);
        p.safe_hi = fmin(p_hi, e_hi - dl);
        p.block = c.bfirst + j * c.nph;
        p.ivl = k;
        p.nph = (short)c.nph;
        p.lastmax = r[FHMC_TR_LASTMAX];
        p.hidx = r[FHMC_TR_HIDX];
        p.cntM = r[FHMC_TR_CNTM];
        p.cntm = r[FHMC_TR_CNTMIN];
        p.nmin = r[FHMC_TR_NMIN];
        p.pad0 = p.pad1 = 0;
        p.lM = blob[p.lastmax];
        p.NM = blob[n_pad + p.lastmax];
        p.lH = blob[p.hidx];
        p.NH = blob[n_pad + p.hidx];
        pc[c.first + j] = p;
        pstart[c.first + j] = p_lo;
    }
}

// build 2b: lookup grid over [a_lo, a_hi]: gfirst[g] = last piece whose lower end is <= the lower edge of grid cell g (0 if none)
__global__ void __launch_bounds__(256) k_cell_grid(unsigned char *cells)
{
    const CellHeader *h = reinterpret_cast<const CellHeader *>(cells);
    const double *pstart = reinterpret_cast<const double *>(cells + h->off_pstart);
    int *gfirst = reinterpret_cast<int *>(cells + h->off_grid);
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > h->grid_n) return;
    const double edge = h->a_lo + (double)g / h->inv_g;
    int lo = 0, hi = h->n_pieces;   // number of pieces with pstart <= edge
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pstart[mid] <= edge) lo = mid + 1; else hi = mid;
    }
    gfirst[g] = max(lo - 1, 0);
}

// build 3: the expansion coefficients, a CTA (8 warps) per cell: one pass over the bins of every phase
template <int NSEL>
__global__ void __launch_bounds__(256) k_cell_coef(const unsigned char *tables, unsigned char *cells, const double *blob, int n_pad, int row0, int row1)
{
    constexpr int K = FHMC_CELL_K, BLK = cell_blk(NSEL), NV = (1 + NSEL) * K;
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const short *rec = reinterpret_cast<const short *>(tables + th->off_rec);
    const CellHeader *h = reinterpret_cast<const CellHeader *>(cells);
    const CellPiece *pc = reinterpret_cast<const CellPiece *>(cells + h->off_piece);
    double *coef = reinterpret_cast<double *>(cells + h->off_coef);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const double *lnpi = blob, *Nrow = blob + n_pad;
    const double *xr[2] = {blob + (size_t)row0 * n_pad, blob + (size_t)row1 * n_pad};
    const double N0 = Nrow[0], dN = th->dN;
    __shared__ double s_red[8][NV + 1];
    for (int pi = blockIdx.x; pi < h->n_pieces; pi += gridDim.x) {
        const CellPiece p = pc[pi];
        const short *r = rec + (size_t)p.ivl * FHMC_TAB_REC_I16;
        const int P = r[FHMC_TR_NPHASE];
        for (int ph = 0; ph < P; ++ph) {
            const int left = r[FHMC_TR_BOUNDS + 2 * ph], right = r[FHMC_TR_BOUNDS + 2 * ph + 1];
            double c, R;
            cell_geom(left, right, c, R);
            double M = -CUDART_INF;
            for (int i = left + (int)threadIdx.x; i < right; i += 256) M = fmax(M, lnpi[i] + p.s_c * Nrow[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
            __syncthreads();   // (s_red of the previous phase has been read)
            if (lane == 0) s_red[wib][NV] = M;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < 8; ++w) M = fmax(M, s_red[w][NV]);
            double acc[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) acc[q] = 0.0;
            const double invR = 1.0 / R;
            for (int i = left + (int)threadIdx.x; i < right; i += 256) {
                const double w = exp(lnpi[i] + p.s_c * Nrow[i] - M);
                const double x = ((double)i - c) * invR;
                double wq[1 + NSEL];
                wq[0] = w;
#pragma unroll
                for (int q = 0; q < NSEL; ++q) wq[1 + q] = w * xr[q][i];
                double t = 1.0;   // x^k / k!
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int q = 0; q <= NSEL; ++q) acc[q * K + k] = fma(wq[q], t, acc[q * K + k]);
                    t *= x * (1.0 / (double)(k + 1));
                }
            }
#pragma unroll
            for (int q = 0; q < NV; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            }
#pragma unroll
            for (int q = 0; q < NV; ++q)
                if (lane == (q & 31)) s_red[wib][q] = acc[q];
            __syncthreads();
            if (wib == 0 && lane < NV) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += s_red[w][lane];
                // normalised by C_0 = sum_i w_i: the sum is C_0 (1 + eps) with |eps| <= e^|y| - 1, its log ln C_0 + log1p(eps) by a
                // short series (no log() per state point), and the averages are ratios of the normalised polynomials
                const double C0 = __shfl_sync((1u << NV) - 1u, v, 0);
                double *b = coef + (size_t)(p.block + ph) * BLK;
                b[4 + lane] = v / C0;
                if (lane == 0) {
                    b[0] = M + log(C0);   // ln S_p at the cell centre
                    b[1] = N0 + c * dN;   // N at the phase centre
                    b[2] = dN * R;        // y = d * b[2]
                    reinterpret_cast<int *>(b + 3)[0] = (left & 0xffff) | (right << 16);   // {left, right} as two int16: the record's bounds word
                    reinterpret_cast<int *>(b + 3)[1] = 0;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// sweep: a thread per state point
// ---------------------------------------------------------------------------------------------------------------------
// resident CTAs per SM: the per-lane-store kernel at 80 registers (3 CTAs, no spills on its hot path, room for the hint), the
// transposing one at 64 (4 CTAs) -- measured: 100 vs 111 us per 4x10^6 state points for the former, 115 vs 118 us for the latter
#ifndef FHMC_CELL_MINB
#define FHMC_CELL_MINB 3
#endif
#ifndef FHMC_CELL_T_MINB
#define FHMC_CELL_T_MINB 4
#endif

// what every thread of a sweep kernel reads from the two table headers and the descriptor
struct CellCtx {
    const CellPiece *cpc;
    const double *coef, *pstart;
    const int *gfirst;
    int last, pmax, grid_n;
    double dN, Na, lmax, a_lo, a_hi, inv_g, l0, N_0, l_last, N_last;
    bool usable;   // the cells belong to this descriptor and hold at least one cell
};

template <int NSEL>
__device__ __forceinline__ CellCtx cell_ctx(const SweepArgs &a)
{
    const unsigned char *tables = static_cast<const unsigned char *>(a.d.mu_tables);
    const unsigned char *cells = static_cast<const unsigned char *>(a.d.mu_cells);
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const CellHeader *ch = reinterpret_cast<const CellHeader *>(cells);
    const int n = a.d.n;
    const double *lnpi = a.blob, *Nrow = a.blob + a.d.n_pad;
    CellCtx c;
    c.cpc = reinterpret_cast<const CellPiece *>(cells + ch->off_piece);
    c.coef = reinterpret_cast<const double *>(cells + ch->off_coef);
    c.pstart = reinterpret_cast<const double *>(cells + ch->off_pstart);
    c.gfirst = reinterpret_cast<const int *>(cells + ch->off_grid);
    c.last = n - 1;
    c.pmax = a.d.pmax;
    c.grid_n = ch->grid_n;
    c.dN = th->dN; c.Na = th->Na; c.lmax = th->lmax;
    c.a_lo = ch->a_lo; c.a_hi = ch->a_hi; c.inv_g = ch->inv_g;
    c.usable = th->magic == FHMC_TAB_MAGIC && !th->bad && th->n == n && th->smooth == a.d.smooth && ch->magic == FHMC_CELL_MAGIC &&
               ch->n_sel == NSEL && th->n_sel == NSEL && (NSEL < 1 || (th->sel_row[0] == a.d.sel_row[0] && ch->sel_row[0] == a.d.sel_row[0])) &&
               (NSEL < 2 || (th->sel_row[1] == a.d.sel_row[1] && ch->sel_row[1] == a.d.sel_row[1])) && ch->n == n && ch->smooth == a.d.smooth &&
               n >= 3 && ch->n_pieces > 0;
    c.l0 = lnpi[0]; c.N_0 = Nrow[0]; c.l_last = lnpi[c.last]; c.N_last = Nrow[c.last];
    return c;
}

// log1p(x) = x + x^2 (c[0] + c[1] x + ... + c[11] x^11),  c[k] = (-1)^(k+1) / (k + 2)
__constant__ double c_log1p[12] = {-1.0 / 2.0, 1.0 / 3.0, -1.0 / 4.0, 1.0 / 5.0, -1.0 / 6.0, 1.0 / 7.0, -1.0 / 8.0, 1.0 / 9.0, -1.0 / 10.0, 1.0 / 11.0,
                                   -1.0 / 12.0, 1.0 / 13.0};

// one state point after its lookup
struct CellPoint {
    double d, u0, resc;   // s - s_c; fl(lnPI_0 + fl(s N_0)); threshold of the RESCUED bit on ln S_p - u_0
    const double *b;      // coefficient block of phase 0
    int nph;
    unsigned flags;
};

// The cell of the state point this thread evaluated last: consecutive rounds of a thread are 256 state points apart and mostly fall
// into the same cell, whose safe range is then checked from registers -- no lookup chain (grid -> cell starts -> cell), and the
// coefficient loads do not wait for the cell's record.  lo > hi: no hint.
struct CellHint {
    const CellPiece *cp;
    const double *b;
    double lo, hi, s_c;
    int nph;
};
__device__ __forceinline__ CellHint cell_no_hint(const CellCtx &c)
{
    CellHint h;
    h.cp = c.cpc;
    h.b = c.coef;
    h.lo = CUDART_INF;
    h.hi = -CUDART_INF;
    h.s_c = 0.0;
    h.nph = 0;
    return h;
}

// Lookup and the per-state-point tests.  Returns false when the state point is left to the table walk (outside the cells, margin test,
// capacity rules); `border`: is_safe is closer to its cutoff than rounding -- the caller leaves such a state point to the walk as well,
// but may evaluate its phases first (the walk overwrites them), so that nothing waits for the words is_safe needs.  act == false (an
// idle lane of a warp that stays convergent) returns false at once.  Requires c.usable.
// HINT = false: the hint is only scratch (every state point is looked up).
template <int NSEL, bool HINT>
__device__ __forceinline__ bool cell_point(const SweepArgs &a, const CellCtx &c, double mu1, bool act, CellPoint &p, CellHint &h, bool &border)
{
    border = false;
    if (!act) return false;
    const double s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);   // GH:77, evaluated left to right
    const double sdn = s * c.dN, av = -sdn;
    if (!(fabs(4.0 * sdn) < 200.0)) return false;
    if (!HINT || !(av >= h.lo && av <= h.hi)) {
        if (!(av >= c.a_lo && av <= c.a_hi)) return false;
        // the cell of this tilt: lookup grid, then a bisection over the few cells that start inside the grid cell
        const int g = min((int)((av - c.a_lo) * c.inv_g), c.grid_n - 1);
        int lo = __ldg(c.gfirst + g), hi = __ldg(c.gfirst + g + 1);
        while (lo < hi) {   // last cell with pstart <= tilt, in [lo, hi]
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(c.pstart + mid) <= av) lo = mid; else hi = mid - 1;
        }
        const CellPiece *cp = c.cpc + lo;
        const double2 w0 = __ldg(reinterpret_cast<const double2 *>(cp));   // {s_c, safe_lo}
        const int4 w1 = __ldg(reinterpret_cast<const int4 *>(cp) + 1);     // {safe_hi lo, safe_hi hi, block, ivl}
        const int4 w2 = __ldg(reinterpret_cast<const int4 *>(cp) + 2);     // {nph | lastmax, hidx | cntM, cntm | nmin, -}
        const double safe_hi = __hiloint2double(w1.y, w1.x);
        const int nph = w2.x & 0xffff, pmax = c.pmax;
        // margin test of the table walk (the tilt is further from both interval ends than rounding can move a comparison), folded
        // into the cell's safe range; the capacity rules of PointEval::repair() for the caller's pmax (a capacity error otherwise)
        if (!(av >= w0.y && av <= safe_hi) ||
            nph > pmax || ((w2.y >> 16) & 0xffff) > pmax - 1 || (w2.z & 0xffff) > pmax || ((w2.z >> 16) & 0xffff) > pmax + 1)
            return false;
        h.cp = cp;
        h.b = c.coef + (size_t)w1.z * cell_blk(NSEL);
        h.lo = w0.y;
        h.hi = safe_hi;
        h.s_c = w0.x;
        h.nph = nph;
    }
    const double2 w3 = __ldg(reinterpret_cast<const double2 *>(h.cp) + 3);   // {lnPI, N} at the last maximum
    const double2 w4 = __ldg(reinterpret_cast<const double2 *>(h.cp) + 4);   // {lnPI, N} at the hull vertex
    // is_safe (GH:586-591): fl(fl(u_M - c) - fl(u_last - c)) < cutoff, decided without c when it is not a rounding matter (a last
    // maximum AT the last bin gives u_M == u_last bit for bit, D = 0, as in the reference)
    const double dl = tab_margin(c.lmax, fabs(s), c.Na);
    const double D = __dadd_rn(w3.x, __dmul_rn(s, w3.y)) - __dadd_rn(c.l_last, __dmul_rn(s, c.N_last));
    border = fabs(D - a.d.cutoff) <= 2.0 * dl;
    if (!HINT && border) return false;
    p.flags = FHMC_ST_FAST | (!(D < a.d.cutoff) ? FHMC_ST_SAFE : 0u);
    p.d = s - h.s_c;
    p.u0 = __dadd_rn(c.l0, __dmul_rn(s, c.N_0));
    // (diagnostic bit of the walk: a phase whose sum underflows next to the global maximum's shift -- ln S_p - u_0 below this)
    p.resc = (double)shift_for_max(__dadd_rn(w4.x, __dmul_rn(s, w4.y))) * 0.6931471805599453 - 644.7236 - p.u0;
    p.b = h.b;
    p.nph = h.nph;
    return true;
}

// One phase of one state point from its coefficient block b: v[0] = F.E./kT, v[1 ..] = the averages, bword = {left, right}.
// Returns false when the state point is outside the range of the expansion (never for a cell found by cell_point(), up to rounding).
template <int NSEL>
__device__ __forceinline__ bool cell_phase(const double *b, CellPoint &p, double (&v)[1 + NSEL], int &bword)
{
    constexpr int K = FHMC_CELL_K;
    const double2 g0 = __ldg(reinterpret_cast<const double2 *>(b)), g1 = __ldg(reinterpret_cast<const double2 *>(b) + 1);
    const double y = p.d * g1.x;
    bword = __double2loint(g1.y);
    double P[1 + NSEL];
#pragma unroll
    for (int q = 0; q <= NSEL; ++q) {
        const double2 *cq = reinterpret_cast<const double2 *>(b + 4 + q * K);
        const double2 k01 = __ldg(cq), k23 = __ldg(cq + 1), k45 = __ldg(cq + 2), k67 = __ldg(cq + 3);
        double t = fma(k67.y, y, k67.x);
        t = fma(t, y, k45.y);
        t = fma(t, y, k45.x);
        t = fma(t, y, k23.y);
        t = fma(t, y, k23.x);
        t = fma(t, y, k01.y);
        P[q] = q == 0 ? t * y : fma(t, y, k01.x);   // (quantity 0: eps = P_0 / C_0 - 1, its constant term is 1)
    }
    const double eps = P[0];
    if (!(fabs(y) <= FHMC_CELL_YMAX * (1.0 + 1e-6)) || !(fabs(eps) < 0.06)) return false;   // (|eps| <= e^0.05 - 1 by construction)
    // log1p(eps), |eps| < 0.06: alternating series through eps^13 / 13 (next term < 1e-17 relative); the coefficients are
    // constant-bank operands of the DFMAs (as literals each costs two UMOV per use)
    double l1 = fma(eps, c_log1p[11], c_log1p[10]);
#pragma unroll
    for (int k = 9; k >= 0; --k) l1 = fma(l1, eps, c_log1p[k]);
    l1 = fma(l1 * eps, eps, eps);
    const double lnS = (g0.x - p.u0) + fma(p.d, g0.y, l1);   // ln S_p - u_0
    if (lnS < p.resc) p.flags |= FHMC_ST_RESCUED;
    const double inv = 1.0 / (1.0 + eps);
    v[0] = -lnS;
#pragma unroll
    for (int q = 0; q < NSEL; ++q) v[1 + q] = P[1 + q] * inv;
    return true;
}

// head of a finished record (+ NaN / -1 in the dead phase slots when the caller asked for it) in every destination
template <int NSEL>
__device__ __forceinline__ void cell_head(const SweepArgs &a, long long rix, unsigned flags, int nph, long long fbase, long long bbase)
{
    const long long cN = a.c.n_total;
    const uchar4 hd = make_uchar4((unsigned char)(flags & 0xFFu), (unsigned char)((flags >> 8) & 0xFFu), (unsigned char)nph, 1);   // (byte 3: written by the tilt cells -- diagnostic)
    for (int dd = 0; dd < a.c.n_dst; ++dd) {
        reinterpret_cast<uchar4 *>(a.c.dst[dd])[rix] = hd;
        if (a.c.fill_dead)
            for (int ph = nph; ph < a.d.pmax; ++ph) {
                double *f = reinterpret_cast<double *>(a.c.dst[dd] + fbase) + ((long long)ph * cN + rix) * (1 + NSEL);
#pragma unroll
                for (int q = 0; q <= NSEL; ++q) f[q] = CUDART_NAN;
                reinterpret_cast<int *>(a.c.dst[dd] + bbase)[(long long)ph * cN + rix] = -1;
            }
    }
}

// leftovers of a warp: appended to the index list of the table walk (one atomic per warp; the list holds n_states + 32 entries)
__device__ __forceinline__ void cell_leftovers(const SweepArgs &a, bool left, long long sp, int lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, left);
    if (m) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(a.c.ix_count, __popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (left && pos >= 0 && pos <= a.st.n_states) a.c.ix_list[pos + __popc(m & ((1u << lane) - 1u))] = sp;
    }
}

// the contiguous run of state points of this CTA, and the mu of one of them (flat lists: no 64-bit division per state point)
struct CellRun {
    long long first, end;
    bool flat;
};
__device__ __forceinline__ CellRun cell_run(const SweepArgs &a)
{
    // Blocked partition: neighbouring state points of a sweep share their cell, so after the first round the coefficient blocks a
    // warp needs are in this SM's L1 (with a grid-stride loop every round of every CTA lands in another cell and waits for L2).
    const long long S = a.st.n_states;
    const long long run = (((S + gridDim.x - 1) / gridDim.x) + 255) & ~255ll;
    CellRun r;
    r.first = (long long)blockIdx.x * run;
    r.end = r.first + run < S ? r.first + run : S;
    r.flat = a.st.mu1_div == 1 && a.st.n_mu1 >= S;
    return r;
}
__device__ __forceinline__ double cell_mu(const SweepArgs &a, const CellRun &r, long long q)
{
    if (q >= r.end) return 0.0;
    return r.flat ? __ldg(a.st.mu1 + q) : a.st.mu1[(q / a.st.mu1_div) % a.st.n_mu1];
}

template <int NSEL>
__global__ void __launch_bounds__(256, FHMC_CELL_MINB) k_sweep_cell(const __grid_constant__ SweepArgs a)
{
    constexpr int BLK = cell_blk(NSEL), NF = 1 + NSEL;
    // the indexed walk behind this kernel is launched with programmatic stream serialisation: its CTAs may be placed as soon as
    // SMs drain here (it waits for this grid's completion itself before it reads the leftover list)
    asm volatile("griddepcontrol.launch_dependents;");
    const CellCtx c = cell_ctx<NSEL>(a);
    const int lane = threadIdx.x & 31;
    const long long cN = a.c.n_total;
    const long long fbase = (4 * cN + 15) & ~15ll, bbase = fbase + (long long)c.pmax * cN * NF * 8;
    const CellRun r = cell_run(a);
    int top = 0;
    double mu_next = cell_mu(a, r, r.first + threadIdx.x);   // the next round's mu is fetched a round ahead
    CellHint hint = cell_no_hint(c);
    for (long long base = r.first; base < r.end; base += 256) {
        const long long sp = base + threadIdx.x;
        const double mu1 = mu_next;
        mu_next = cell_mu(a, r, sp + 256);
        bool done = sp >= r.end;
        if (!done && c.usable) {
            CellPoint p;
            bool border;
            #ifndef FHMC_CELL_PLAIN_HINT
#define FHMC_CELL_PLAIN_HINT true
#endif
            if (cell_point<NSEL, FHMC_CELL_PLAIN_HINT>(a, c, mu1, true, p, hint, border)) {
                const long long rix = a.c.first + sp;
                // first destination: running pointers over the phase blocks (the other destinations of a fused gather are
                // addressed per phase)
                double *f0 = reinterpret_cast<double *>(a.c.dst[0] + fbase) + rix * NF;
                int *b0 = reinterpret_cast<int *>(a.c.dst[0] + bbase) + rix;
                const double *b = p.b;
                bool good = true;
                // (a state point that gives up after its first phases has left them in the record: the table walk that takes it
                // over finds the same phases in the same interval record and overwrites every one of them)
#pragma unroll
                for (int ph = 0; ph < FHMC_COMPACT_PMAX; ++ph) {
                    if (ph >= p.nph) break;
                    double v[NF];
                    int bword;
                    if (!cell_phase<NSEL>(b, p, v, bword)) { good = false; break; }
#pragma unroll
                    for (int q = 0; q < NF; ++q) f0[q] = v[q];
                    *b0 = bword;
                    for (int dd = 1; dd < a.c.n_dst; ++dd) {
                        double *f = reinterpret_cast<double *>(a.c.dst[dd] + fbase) + ((long long)ph * cN + rix) * NF;
#pragma unroll
                        for (int q = 0; q < NF; ++q) f[q] = v[q];
                        reinterpret_cast<int *>(a.c.dst[dd] + bbase)[(long long)ph * cN + rix] = bword;
                    }
                    f0 += cN * NF;
                    b0 += cN;
                    b += BLK;
                }
                if (good && !border) {
                    cell_head<NSEL>(a, rix, p.flags, p.nph, fbase, bbase);
                    top = max(top, p.nph);
                    done = true;
                }
            }
        }
        cell_leftovers(a, !done, sp, lane);
    }
    if (a.c.max_nphase) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
        if (lane == 0 && top > 0) atomicMax(a.c.max_nphase, top);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same sweep with the phase loop uniform across the warp and the fp64 fields of a phase leaving through a shared-memory
// transpose: the 32 state points of a warp fill 32 (1 + NSEL) consecutive doubles of the phase block, written as 16-byte
// chunks (full 32-byte sectors) instead of (1 + NSEL) 8-byte stores per lane at a stride of 8 (1 + NSEL) bytes -- a third of the
// store sectors at NSEL = 2, which is what bounds a gather fused into the sweep (every record crosses NVLink to every peer).
// ---------------------------------------------------------------------------------------------------------------------
template <int NSEL>
__global__ void __launch_bounds__(256, FHMC_CELL_T_MINB) k_sweep_cell_t(const __grid_constant__ SweepArgs a)
{
    constexpr int BLK = cell_blk(NSEL), NF = 1 + NSEL, CHUNKS = 16 * NF, ROUNDS = (CHUNKS + 31) / 32;
    asm volatile("griddepcontrol.launch_dependents;");
    __shared__ __align__(16) double s_slab[8][32 * NF];
    const CellCtx c = cell_ctx<NSEL>(a);
    const int lane = threadIdx.x & 31;
    double *slab = s_slab[threadIdx.x >> 5];
    const long long cN = a.c.n_total;
    const long long fbase = (4 * cN + 15) & ~15ll, bbase = fbase + (long long)c.pmax * cN * NF * 8;
    const CellRun r = cell_run(a);
    int top = 0;
    if (!c.usable) {   // cells that do not belong to this descriptor: every state point is left to the table walk
        for (long long base = r.first; base < r.end; base += 256) cell_leftovers(a, base + threadIdx.x < r.end, base + threadIdx.x, lane);
        return;
    }
    double mu_next = cell_mu(a, r, r.first + threadIdx.x);
    CellHint hint = cell_no_hint(c);
    for (long long base = r.first; base < r.end; base += 256) {
        const long long sp = base + threadIdx.x;
        const double mu1 = mu_next;
        mu_next = cell_mu(a, r, sp + 256);
        const bool mine = sp < r.end;
        CellPoint p;
        bool border;
        bool act = cell_point<NSEL, true>(a, c, mu1, mine, p, hint, border);
        const long long rix = a.c.first + sp, rix0 = rix - lane;
        const int nph_l = act ? p.nph : 0;
        const int nmax = __reduce_max_sync(0xffffffffu, nph_l);
        const double *b = act ? p.b : c.coef;
        for (int ph = 0; ph < nmax; ++ph, b += BLK) {   // (uniform across the warp)
            bool on = act && ph < nph_l;
            double v[NF];
#pragma unroll
            for (int q = 0; q < NF; ++q) v[q] = 0.0;
            int bword = 0;
            if (on && !cell_phase<NSEL>(b, p, v, bword)) act = on = false;
            const long long prow = (long long)ph * cN;
            // all 32 state points of the warp have this phase and the warp's slice of the block starts on a 16-byte boundary:
            // transpose through shared memory and write full 16-byte chunks; else every lane writes its own fields
            const bool whole = __all_sync(0xffffffffu, on) && ((NF & 1) == 0 || (((prow + rix0) & 1) == 0));
            if (whole) {
#pragma unroll
                for (int q = 0; q < NF; ++q) slab[NF * lane + q] = v[q];
                __syncwarp();
                double2 ch[ROUNDS];
#pragma unroll
                for (int k = 0; k < ROUNDS; ++k)
                    ch[k] = (lane + 32 * k < CHUNKS) ? reinterpret_cast<const double2 *>(slab)[lane + 32 * k] : make_double2(0.0, 0.0);
                __syncwarp();
                for (int dd = 0; dd < a.c.n_dst; ++dd) {
                    double2 *o = reinterpret_cast<double2 *>(reinterpret_cast<double *>(a.c.dst[dd] + fbase) + (prow + rix0) * NF);
#pragma unroll
                    for (int k = 0; k < ROUNDS; ++k)
                        if (lane + 32 * k < CHUNKS) o[lane + 32 * k] = ch[k];
                }
            } else if (on) {
                for (int dd = 0; dd < a.c.n_dst; ++dd) {
                    double *f = reinterpret_cast<double *>(a.c.dst[dd] + fbase) + (prow + rix) * NF;
#pragma unroll
                    for (int q = 0; q < NF; ++q) f[q] = v[q];
                }
            }
            if (on)
                for (int dd = 0; dd < a.c.n_dst; ++dd) reinterpret_cast<int *>(a.c.dst[dd] + bbase)[prow + rix] = bword;
        }
        // (a state point that gave up after its first phases has left them in the record: the table walk that takes it over finds
        // the same phases in the same interval record and overwrites every one of them)
        act = act && !border;
        if (act) {
            cell_head<NSEL>(a, rix, p.flags, p.nph, fbase, bbase);
            top = max(top, p.nph);
        }
        cell_leftovers(a, mine && !act, sp, lane);
    }
    if (a.c.max_nphase) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
        if (lane == 0 && top > 0) atomicMax(a.c.max_nphase, top);
    }
}

struct CellLayout {
    CellHeader h;
    size_t total;
};

static size_t upc(size_t x, size_t a) { return (x + a - 1) / a * a; }

static CellLayout cell_layout(const fhmc_hist_desc &d, int extra_pieces)
{
    CellLayout L;
    memset(&L, 0, sizeof(L));
    CellHeader &h = L.h;
    h.magic = FHMC_CELL_MAGIC;
    h.n_sel = d.n_sel;
    h.sel_row[0] = d.n_sel > 0 ? d.sel_row[0] : 0;
    h.sel_row[1] = d.n_sel > 1 ? d.sel_row[1] : 0;
    h.n = d.n;
    h.smooth = d.smooth;
    h.blk = cell_blk(d.n_sel);
    h.iv_cap = (int)upc((size_t)4 * d.n + d.hull_len, 256) + 1;   // (= MuTabHeader::ep_cap + 1)
    h.piece_cap = h.iv_cap + (extra_pieces > 0 ? extra_pieces : 0);
    h.block_cap = 3 * h.piece_cap;
    size_t off = 256;
    auto take = [&](size_t bytes) { const size_t o = off; off = upc(off + bytes, 256); return (long long)o; };
    h.off_iv = take((size_t)h.iv_cap * sizeof(CellIv));
    h.off_piece = take((size_t)h.piece_cap * sizeof(CellPiece));
    h.off_coef = take((size_t)h.block_cap * h.blk * 8);
    h.off_pstart = take((size_t)(h.piece_cap + 1) * 8);
    int g = 1024;
    while (g < 4 * h.piece_cap && g < 65536) g <<= 1;
    h.grid_n = g;
    h.off_grid = take((size_t)(g + 2) * 4);
    L.total = off;
    return L;
}

// cells + index list of the leftovers: returns 0 ok, 1 error, -1 not applicable
int launch_cell_compact(const SweepArgs &args_in, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = args_in.d;
    if (!d.mu_tables || !d.mu_cells || d.pmax > FHMC_COMPACT_PMAX || d.n_sel > 2 || !args_in.c.ix_list) return -1;
    // the list's counter lives in the cells buffer (zero between sweeps: fhmc_mu_cells_build zeroes it, the last CTA of the
    // indexed walk resets it) -- sweeps that share a cells buffer must therefore be ordered on one stream
    SweepArgs args = args_in;
    args.c.ix_count = reinterpret_cast<int *>(static_cast<unsigned char *>(const_cast<void *>(d.mu_cells)) + offsetof(CellHeader, q_count));
    const long long S = args.st.n_states;
    // persistent grid: exactly the CTAs that are resident at once (a partial last wave would idle a third of the SMs)
    // the transposing variant when the records also go to NVLink peers (FHMC_CELL_T = 1 / 0 forces it on / off)
    static int t_mode = -1;
    if (t_mode < 0) { const char *e = getenv("FHMC_CELL_T"); t_mode = e ? (atoi(e) ? 1 : 0) : 2; }
    const bool tr = t_mode == 1 || (t_mode == 2 && args.c.n_dst > 1);
    static int occ_cache[2][3] = {{0, 0, 0}, {0, 0, 0}};
    const int q = d.n_sel;
    if (occ_cache[tr][q] == 0) {
        int occ = 0;
        cudaError_t e;
        if (tr)
            e = q == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_cell_t<0>, 256, 0)
              : q == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_cell_t<1>, 256, 0)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_cell_t<2>, 256, 0);
        else
            e = q == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_cell<0>, 256, 0)
              : q == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_cell<1>, 256, 0)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_cell<2>, 256, 0);
        if (check_cuda(e, "occupancy query")) return 1;
        occ_cache[tr][q] = occ > 0 ? occ : 1;
    }
    long long grid = (S + 255) / 256;
    if (grid > (long long)sm_count * occ_cache[tr][q]) grid = (long long)sm_count * occ_cache[tr][q];
    if (const char *e = getenv("FHMC_CELL_GRID")) { const long long g = atoll(e); if (g > 0) grid = g; }   // (probe)
    if (tr) {
        switch (d.n_sel) {
        case 0: k_sweep_cell_t<0><<<(unsigned)grid, 256, 0, stream>>>(args); break;
        case 1: k_sweep_cell_t<1><<<(unsigned)grid, 256, 0, stream>>>(args); break;
        default: k_sweep_cell_t<2><<<(unsigned)grid, 256, 0, stream>>>(args); break;
        }
    } else {
        switch (d.n_sel) {
        case 0: k_sweep_cell<0><<<(unsigned)grid, 256, 0, stream>>>(args); break;
        case 1: k_sweep_cell<1><<<(unsigned)grid, 256, 0, stream>>>(args); break;
        default: k_sweep_cell<2><<<(unsigned)grid, 256, 0, stream>>>(args); break;
        }
    }
    if (check_cuda(cudaGetLastError(), "k_sweep_cell launch")) return 1;
    const bool s0n = d.n_sel > 0 && d.sel_row[0] == 1;
    int rc;
    switch (d.n_sel) {
    case 0: rc = launch_tab2_idx<0, false>(args, sm_count, smem_optin, stream); break;
    case 1: rc = s0n ? launch_tab2_idx<1, true>(args, sm_count, smem_optin, stream) : launch_tab2_idx<1, false>(args, sm_count, smem_optin, stream); break;
    default: rc = s0n ? launch_tab2_idx<2, true>(args, sm_count, smem_optin, stream) : launch_tab2_idx<2, false>(args, sm_count, smem_optin, stream); break;
    }
    if (rc < 0) { set_error("k_sweep_tab2_idx: histogram too large for shared memory"); return 1; }
    if (rc) return rc;
    note_kernel("k_sweep_cell<compact>");
    return 0;
}

}  // namespace fhmc

using namespace fhmc;

extern "C" {

size_t fhmc_mu_cells_bytes(const fhmc_hist_desc *desc, int extra_pieces)
{
    if (!desc || fhmc_mu_tables_bytes(desc) == 0) return 0;
    return cell_layout(*desc, extra_pieces).total;
}

static int cells_build_impl(const fhmc_hist_desc *desc, const double *blob, void *cells, size_t cells_bytes, int extra_pieces, double mu_lo,
                            double mu_hi, const double *mu_dev, long long n_mu, void *stream)
{
    if (!desc || !blob || !cells) { set_error("null argument"); return 1; }
    if (!desc->mu_tables || fhmc_mu_tables_bytes(desc) == 0) { set_error("mu cells need the mu tables of the same descriptor"); return 2; }
    if (!mu_dev && !(mu_lo <= mu_hi)) { set_error("mu cells: empty or non-finite range"); return 1; }
    if ((uintptr_t)cells & 255) { set_error("cells must be a 256-byte aligned device pointer"); return 1; }
    CellLayout L = cell_layout(*desc, extra_pieces);
    if (cells_bytes < L.total) { set_error("cells buffer too small: need fhmc_mu_cells_bytes() bytes"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *cb = static_cast<unsigned char *>(cells);
    const unsigned char *tb = static_cast<const unsigned char *>(desc->mu_tables);
    L.h.kmin = ~0ull;
    L.h.kmax = 0ull;
    if (check_cuda(cudaMemcpyAsync(cb, &L.h, sizeof(CellHeader), cudaMemcpyHostToDevice, s), "cudaMemcpyAsync")) return 1;
    if (mu_dev) {
        long long g = (n_mu + 1023) / 1024;
        k_cell_range<<<(unsigned)(g < 1 ? 1 : (g > 592 ? 592 : g)), 256, 0, s>>>(mu_dev, n_mu, cb);
    }
    k_cell_plan<<<1, 1024, 0, s>>>(tb, cb, L.h, mu_lo, mu_hi, desc->mu1_ref, desc->beta_ref, mu_dev ? 1 : 0);
    k_cell_pieces<<<(L.h.iv_cap + 255) / 256, 256, 0, s>>>(tb, cb, blob, desc->n_pad);
    k_cell_grid<<<(L.h.grid_n + 1 + 255) / 256, 256, 0, s>>>(cb);
    const int r0 = desc->n_sel > 0 ? desc->sel_row[0] : 0, r1 = desc->n_sel > 1 ? desc->sel_row[1] : 0;
    const int grid = 148 * 4;
    switch (desc->n_sel) {
    case 0: k_cell_coef<0><<<grid, 256, 0, s>>>(tb, cb, blob, desc->n_pad, r0, r1); break;
    case 1: k_cell_coef<1><<<grid, 256, 0, s>>>(tb, cb, blob, desc->n_pad, r0, r1); break;
    default: k_cell_coef<2><<<grid, 256, 0, s>>>(tb, cb, blob, desc->n_pad, r0, r1); break;
    }
    return check_cuda(cudaGetLastError(), "mu cells build launch");
}

int fhmc_mu_cells_build(const fhmc_hist_desc *desc, const double *blob, void *cells, size_t cells_bytes, int extra_pieces, double mu_lo,
                        double mu_hi, void *stream)
{
    return cells_build_impl(desc, blob, cells, cells_bytes, extra_pieces, mu_lo, mu_hi, nullptr, 0, stream);
}

int fhmc_mu_cells_build_for(const fhmc_hist_desc *desc, const double *blob, void *cells, size_t cells_bytes, int extra_pieces,
                            const double *mu_dev, long long n_mu, void *stream)
{
    if (!mu_dev || n_mu < 1) { set_error("mu cells: no state points"); return 1; }
    return cells_build_impl(desc, blob, cells, cells_bytes, extra_pieces, 0.0, 0.0, mu_dev, n_mu, stream);
}

}  // extern "C"
