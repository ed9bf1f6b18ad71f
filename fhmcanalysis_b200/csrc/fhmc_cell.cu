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
// range of a dense sweep, each costing one pass over the bins to build (k_cell_coef: a warp per cell) -- against one pass
// PER STATE POINT for the walk.  A state point then costs, per phase, (1 + n_sel) degree-7 polynomials, one log and one
// reciprocal, whatever the histogram length: the sweep is bound by its record traffic (8 bytes of mu in, 60 bytes out per
// two-phase state point), not by the fp64 pipe.
//
// What stays exactly as in the table walk: the interval lookup and its rounding-margin test (integers -- phase count and
// bounds -- come from the interval record, i.e. from the general evaluator at the interval's representative), the capacity
// rules for the caller's pmax, is_safe() (decided from u_lastmax - u_last against the cutoff, with a margin; GH:586-591), the
// RESCUED diagnostic bit.  Whatever the cells do not cover (state points outside the range they were built for, intervals
// without a valid record, failed margin tests, is_safe closer to its cutoff than rounding) is appended to an index list and
// walked by k_sweep_tab2_idx -- the parity-pinned table walk with its own fallback to the general evaluator -- in the same
// call.
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
    int pad0;
    long long off_iv, off_piece, off_coef;
    double a_lo, a_hi;                    // covered tilt range (device)
    double pad1[20];
};
static_assert(sizeof(CellHeader) <= 256, "header must fit its slot");

struct CellIv {     // per elementary interval
    int first, m;   // pieces first .. first + m - 1 (m == 0: not covered)
    int bfirst, nph;
    double a0, inv_w;
};
struct CellPiece {
    double s_c;
    int block, ivl;
};
static_assert(sizeof(CellIv) == 32 && sizeof(CellPiece) == 16, "cell table strides");

// centre and half width (bins) of the phase [left, right)
__device__ __forceinline__ void cell_geom(int left, int right, double &c, double &R)
{
    c = 0.5 * (double)(left + right - 1);
    R = fmax(0.5 * (double)(right - 1 - left), 0.5);
}

__host__ __device__ constexpr int cell_blk(int nsel) { return 4 + FHMC_CELL_K * (1 + nsel); }

// ---------------------------------------------------------------------------------------------------------------------
// build 1 (one CTA): pieces per interval for the tilt range of [mu_lo, mu_hi], exclusive scans
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_cell_plan(const unsigned char *tables, unsigned char *cells, CellHeader h0, double mu_lo, double mu_hi,
                                                    double mu1_ref, double beta_ref)
{
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
    const bool usable = th->magic == FHMC_TAB_MAGIC && !th->bad && niv <= h0.iv_cap;
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
    }
}

// build 2: the pieces of every covered interval (a thread per interval)
__global__ void __launch_bounds__(256) k_cell_pieces(const unsigned char *tables, unsigned char *cells)
{
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const CellHeader *h = reinterpret_cast<const CellHeader *>(cells);
    const CellIv *iv = reinterpret_cast<const CellIv *>(cells + h->off_iv);
    CellPiece *pc = reinterpret_cast<CellPiece *>(cells + h->off_piece);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > th->n_ep) return;
    const CellIv c = iv[k];
    for (int j = 0; j < c.m; ++j) {
        CellPiece p;
        const double a_c = c.a0 + ((double)j + 0.5) / c.inv_w;
        p.s_c = -a_c / th->dN;
        p.block = c.bfirst + j * c.nph;
        p.ivl = k;
        pc[c.first + j] = p;
    }
}

// build 3: the expansion coefficients, a warp per piece
template <int NSEL>
__global__ void __launch_bounds__(256) k_cell_coef(const unsigned char *tables, unsigned char *cells, const double *blob, int n_pad, int row0, int row1)
{
    constexpr int K = FHMC_CELL_K, BLK = cell_blk(NSEL);
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const short *rec = reinterpret_cast<const short *>(tables + th->off_rec);
    const CellHeader *h = reinterpret_cast<const CellHeader *>(cells);
    const CellPiece *pc = reinterpret_cast<const CellPiece *>(cells + h->off_piece);
    double *coef = reinterpret_cast<double *>(cells + h->off_coef);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const double *lnpi = blob, *Nrow = blob + n_pad;
    const double *xr[2] = {blob + (size_t)row0 * n_pad, blob + (size_t)row1 * n_pad};
    const double N0 = Nrow[0], dN = th->dN;
    for (int pi = warp; pi < h->n_pieces; pi += nwarp) {
        const CellPiece p = pc[pi];
        const short *r = rec + (size_t)p.ivl * FHMC_TAB_REC_I16;
        const int P = r[FHMC_TR_NPHASE];
        for (int ph = 0; ph < P; ++ph) {
            const int left = r[FHMC_TR_BOUNDS + 2 * ph], right = r[FHMC_TR_BOUNDS + 2 * ph + 1];
            double c, R;
            cell_geom(left, right, c, R);
            double M = -CUDART_INF;
            for (int i = left + lane; i < right; i += 32) M = fmax(M, lnpi[i] + p.s_c * Nrow[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
            double acc[(1 + NSEL) * K];
#pragma unroll
            for (int q = 0; q < (1 + NSEL) * K; ++q) acc[q] = 0.0;
            const double invR = 1.0 / R;
            for (int i = left + lane; i < right; i += 32) {
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
            for (int q = 0; q < (1 + NSEL) * K; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            }
            double *b = coef + (size_t)(p.block + ph) * BLK;
            if (lane == 0) {
                b[0] = M;
                b[1] = N0 + c * dN;   // N at the phase centre
                b[2] = dN * R;        // y = d * b[2]
                b[3] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < (1 + NSEL) * K; ++q)
                if (lane == (q & 31)) b[4 + q] = acc[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// sweep: a thread per state point
// ---------------------------------------------------------------------------------------------------------------------
template <int NSEL>
__global__ void __launch_bounds__(256) k_sweep_cell(const __grid_constant__ SweepArgs a)
{
    constexpr int K = FHMC_CELL_K, BLK = cell_blk(NSEL);
    const unsigned char *tables = static_cast<const unsigned char *>(a.d.mu_tables);
    const unsigned char *cells = static_cast<const unsigned char *>(a.d.mu_cells);
    const MuTabHeader *th = reinterpret_cast<const MuTabHeader *>(tables);
    const CellHeader *ch = reinterpret_cast<const CellHeader *>(cells);
    const double *ep = reinterpret_cast<const double *>(tables + th->off_ep);
    const short *rec = reinterpret_cast<const short *>(tables + th->off_rec);
    const CellIv *civ = reinterpret_cast<const CellIv *>(cells + ch->off_iv);
    const CellPiece *cpc = reinterpret_cast<const CellPiece *>(cells + ch->off_piece);
    const double *coef = reinterpret_cast<const double *>(cells + ch->off_coef);
    const int n = a.d.n, last = n - 1, pmax = a.d.pmax, ne = th->n_ep;
    const double dN = th->dN, Na = th->Na, lmax = th->lmax;
    const double *lnpi = a.blob, *Nrow = a.blob + a.d.n_pad;
    const bool usable = th->magic == FHMC_TAB_MAGIC && !th->bad && th->n == n && th->smooth == a.d.smooth && ch->magic == FHMC_CELL_MAGIC &&
                        ch->n_sel == NSEL && th->n_sel == NSEL && (NSEL < 1 || th->sel_row[0] == a.d.sel_row[0]) &&
                        (NSEL < 2 || th->sel_row[1] == a.d.sel_row[1]) && n >= 3;
    const double l0 = lnpi[0], N_0 = Nrow[0], l_last = lnpi[last], N_last = Nrow[last];
    const int lane = threadIdx.x & 31;
    const long long S = a.st.n_states;
    const long long cN = a.c.n_total;
    int top = 0;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < S; base += (long long)gridDim.x * blockDim.x) {
        const long long sp = base + threadIdx.x;
        bool done = true;
        if (sp < S) {
            done = false;
            const double mu1 = a.st.mu1[(sp / a.st.mu1_div) % a.st.n_mu1];
            const double s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);   // GH:77, evaluated left to right
            const double sdn = s * dN, av = -sdn;
            do {
                if (!usable || !(fabs(4.0 * sdn) < 200.0)) break;
                int lo = 0, hi = ne;
                while (lo < hi) {   // number of endpoints <= tilt
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(ep + mid) <= av) lo = mid + 1; else hi = mid;
                }
                const double dl = tab_margin(lmax, fabs(s), Na);
                const double e_lo = lo > 0 ? __ldg(ep + lo - 1) : -CUDART_INF, e_hi = lo < ne ? __ldg(ep + lo) : CUDART_INF;
                const short *r = rec + (size_t)lo * FHMC_TAB_REC_I16;
                const int4 head = __ldg(reinterpret_cast<const int4 *>(r));   // {valid, nphase}, {hidx, lastmax}, {cntM, cntm}, {nmin, -}
                const int nph = head.x >> 16;
                // (the capacity rules of PointEval::repair() for the caller's pmax: such a state point is a capacity error)
                const bool cap = nph > pmax || (head.z & 0xffff) > pmax - 1 || (head.z >> 16) > pmax || (head.w & 0xffff) > pmax + 1;
                if (!(av - e_lo > dl && e_hi - av > dl) || (head.x & 0xffff) != 1 || cap) break;
                const int4 c0 = __ldg(reinterpret_cast<const int4 *>(civ + lo));           // {first, m, bfirst, nph}
                const double2 c1 = __ldg(reinterpret_cast<const double2 *>(civ + lo) + 1);   // {a0, inv_w}
                if (c0.y < 1 || c0.w != nph) break;
                const double fj = (av - c1.x) * c1.y;
                if (!(fj >= 0.0 && fj <= (double)c0.y)) break;   // outside the range the cells were built for
                const int j = min((int)fj, c0.y - 1);
                const int4 pw = __ldg(reinterpret_cast<const int4 *>(cpc + c0.x + j));   // {s_c lo, s_c hi, block, ivl}
                const double s_c = __hiloint2double(pw.y, pw.x);
                const double d = s - s_c;
                // is_safe (GH:586-591): fl(fl(u_M - c) - fl(u_last - c)) < cutoff, decided without c when it is not a rounding matter
                const int hidx = head.y & 0xffff, lastmax = head.y >> 16;
                const double u_last = __dadd_rn(l_last, __dmul_rn(s, N_last));
                unsigned flags = FHMC_ST_FAST;
                if (lastmax != last) {
                    const double uM = __dadd_rn(__ldg(lnpi + lastmax), __dmul_rn(s, __ldg(Nrow + lastmax)));
                    const double D = uM - u_last;
                    if (fabs(D - a.d.cutoff) <= 2.0 * dl) break;
                    if (!(D < a.d.cutoff)) flags |= FHMC_ST_SAFE;
                } else if (!(0.0 < a.d.cutoff)) {
                    flags |= FHMC_ST_SAFE;
                }
                const double u0 = __dadd_rn(l0, __dmul_rn(s, N_0));
                const int Mq = shift_for_max(__dadd_rn(__ldg(lnpi + hidx), __dmul_rn(s, __ldg(Nrow + hidx))));
                const double *b = coef + (size_t)pw.z * BLK;
                // bounds of the phases, as the interval record holds them ({left, right} as two int16 per word)
                int wds[8];
                {
                    const int4 lo4 = __ldg(reinterpret_cast<const int4 *>(r + FHMC_TR_BOUNDS));
                    wds[0] = lo4.x; wds[1] = lo4.y; wds[2] = lo4.z; wds[3] = lo4.w;
                    wds[4] = wds[5] = wds[6] = wds[7] = 0;
                    if (nph > 4) {
                        const int4 hi4 = __ldg(reinterpret_cast<const int4 *>(r + FHMC_TR_BOUNDS) + 1);
                        wds[4] = hi4.x; wds[5] = hi4.y; wds[6] = hi4.z; wds[7] = hi4.w;
                    }
                }
                const long long fbase = (4 * cN + 15) & ~15ll, bbase = fbase + (long long)pmax * cN * (1 + NSEL) * 8, rix = a.c.first + sp;
                bool good = true;
                // (a state point that gives up after its first phases has left them in the record: the table walk that takes it
                // over finds the same phases in the same interval record and overwrites every one of them)
#pragma unroll
                for (int ph = 0; ph < FHMC_COMPACT_PMAX; ++ph) {
                    if (ph >= nph) break;
                    const double2 g0 = __ldg(reinterpret_cast<const double2 *>(b)), g1 = __ldg(reinterpret_cast<const double2 *>(b) + 1);
                    const double y = d * g1.x;
                    if (!(fabs(y) <= FHMC_CELL_YMAX * (1.0 + 1e-6))) { good = false; break; }
                    double P[1 + NSEL];
#pragma unroll
                    for (int q = 0; q <= NSEL; ++q) {
                        const double2 *cq = reinterpret_cast<const double2 *>(b + 4 + q * K);
                        const double2 k01 = __ldg(cq), k23 = __ldg(cq + 1), k45 = __ldg(cq + 2), k67 = __ldg(cq + 3);
                        double v = fma(k67.y, y, k67.x);
                        v = fma(v, y, k45.y);
                        v = fma(v, y, k45.x);
                        v = fma(v, y, k23.y);
                        v = fma(v, y, k23.x);
                        v = fma(v, y, k01.y);
                        v = fma(v, y, k01.x);
                        P[q] = v;
                    }
                    if (!(P[0] > 0.0)) { good = false; break; }
                    const double lnS = (g0.x - u0) + fma(d, g0.y, log(P[0]));   // ln S_p - u_0
                    // (diagnostic bit of the walk: this phase's sum underflows next to the global maximum's shift)
                    if (lnS + u0 - (double)Mq * 0.6931471805599453 < -644.7236) flags |= FHMC_ST_RESCUED;
                    const double inv = 1.0 / P[0];
                    for (int dd = 0; dd < a.c.n_dst; ++dd) {
                        double *f = reinterpret_cast<double *>(a.c.dst[dd] + fbase) + ((long long)ph * cN + rix) * (1 + NSEL);
                        f[0] = -lnS;
#pragma unroll
                        for (int q = 0; q < NSEL; ++q) f[1 + q] = P[1 + q] * inv;
                        reinterpret_cast<int *>(a.c.dst[dd] + bbase)[(long long)ph * cN + rix] = wds[ph];
                    }
                    b += BLK;
                }
                if (!good) break;
                const uchar4 hd = make_uchar4((unsigned char)(flags & 0xFFu), (unsigned char)((flags >> 8) & 0xFFu), (unsigned char)nph, 1);   // (byte 3: written by the tilt cells -- diagnostic)
                for (int dd = 0; dd < a.c.n_dst; ++dd) {
                    reinterpret_cast<uchar4 *>(a.c.dst[dd])[rix] = hd;
                    if (a.c.fill_dead)
                        for (int ph = nph; ph < pmax; ++ph) {
                            double *f = reinterpret_cast<double *>(a.c.dst[dd] + fbase) + ((long long)ph * cN + rix) * (1 + NSEL);
#pragma unroll
                            for (int q = 0; q <= NSEL; ++q) f[q] = CUDART_NAN;
                            reinterpret_cast<int *>(a.c.dst[dd] + bbase)[(long long)ph * cN + rix] = -1;
                        }
                }
                top = max(top, nph);
                done = true;
            } while (false);
        }
        // leftovers: appended to the index list of the table walk (one atomic per warp)
        const unsigned m = __ballot_sync(0xffffffffu, !done);
        if (m) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(a.c.ix_count, __popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (!done) a.c.ix_list[pos + __popc(m & ((1u << lane) - 1u))] = sp;
        }
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
    h.blk = cell_blk(d.n_sel);
    h.iv_cap = (int)upc((size_t)4 * d.n + d.hull_len, 256) + 1;   // (= MuTabHeader::ep_cap + 1)
    h.piece_cap = h.iv_cap + (extra_pieces > 0 ? extra_pieces : 0);
    h.block_cap = 3 * h.piece_cap;
    size_t off = 256;
    auto take = [&](size_t bytes) { const size_t o = off; off = upc(off + bytes, 256); return (long long)o; };
    h.off_iv = take((size_t)h.iv_cap * sizeof(CellIv));
    h.off_piece = take((size_t)h.piece_cap * sizeof(CellPiece));
    h.off_coef = take((size_t)h.block_cap * h.blk * 8);
    L.total = off;
    return L;
}

// cells + index list of the leftovers: returns 0 ok, 1 error, -1 not applicable
int launch_cell_compact(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = args.d;
    if (!d.mu_tables || !d.mu_cells || d.pmax > FHMC_COMPACT_PMAX || d.n_sel > 2 || !args.c.ix_list || !args.c.ix_count) return -1;
    if (check_cuda(cudaMemsetAsync(args.c.ix_count, 0, sizeof(int), stream), "cudaMemsetAsync")) return 1;
    const long long S = args.st.n_states;
    long long grid = (S + 255) / 256;
    if (grid > (long long)sm_count * 8) grid = (long long)sm_count * 8;
    switch (d.n_sel) {
    case 0: k_sweep_cell<0><<<(unsigned)grid, 256, 0, stream>>>(args); break;
    case 1: k_sweep_cell<1><<<(unsigned)grid, 256, 0, stream>>>(args); break;
    default: k_sweep_cell<2><<<(unsigned)grid, 256, 0, stream>>>(args); break;
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

int fhmc_mu_cells_build(const fhmc_hist_desc *desc, const double *blob, void *cells, size_t cells_bytes, int extra_pieces, double mu_lo,
                        double mu_hi, void *stream)
{
    if (!desc || !blob || !cells) { set_error("null argument"); return 1; }
    if (!desc->mu_tables || fhmc_mu_tables_bytes(desc) == 0) { set_error("mu cells need the mu tables of the same descriptor"); return 2; }
    if (!(mu_lo <= mu_hi)) { set_error("mu cells: empty or non-finite range"); return 1; }
    if ((uintptr_t)cells & 255) { set_error("cells must be a 256-byte aligned device pointer"); return 1; }
    const CellLayout L = cell_layout(*desc, extra_pieces);
    if (cells_bytes < L.total) { set_error("cells buffer too small: need fhmc_mu_cells_bytes() bytes"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *cb = static_cast<unsigned char *>(cells);
    const unsigned char *tb = static_cast<const unsigned char *>(desc->mu_tables);
    if (check_cuda(cudaMemcpyAsync(cb, &L.h, sizeof(CellHeader), cudaMemcpyHostToDevice, s), "cudaMemcpyAsync")) return 1;
    k_cell_plan<<<1, 1024, 0, s>>>(tb, cb, L.h, mu_lo, mu_hi, desc->mu1_ref, desc->beta_ref);
    k_cell_pieces<<<(L.h.iv_cap + 255) / 256, 256, 0, s>>>(tb, cb);
    const int r0 = desc->n_sel > 0 ? desc->sel_row[0] : 0, r1 = desc->n_sel > 1 ? desc->sel_row[1] : 0;
    const int grid = 148 * 4;
    switch (desc->n_sel) {
    case 0: k_cell_coef<0><<<grid, 256, 0, s>>>(tb, cb, blob, desc->n_pad, r0, r1); break;
    case 1: k_cell_coef<1><<<grid, 256, 0, s>>>(tb, cb, blob, desc->n_pad, r0, r1); break;
    default: k_cell_coef<2><<<grid, 256, 0, s>>>(tb, cb, blob, desc->n_pad, r0, r1); break;
    }
    return check_cuda(cudaGetLastError(), "mu cells build launch");
}

}  // extern "C"
