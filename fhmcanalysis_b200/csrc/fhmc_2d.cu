// fhmc_2d.cu -- K5: reweighting of a two-dimensional joint histogram lnPI(op1, op2).
//
// The reference has the CONTAINER (two_dim/joint_hist.pyx:201-247: dense lnPI[n1][n2] padded with
// -inf, bounds_idx[n1][2]) but no reweighting of it; the closest reference semantics are the ragged
// 2-D log-sum-exp / averages of pore_hist (two_dim/h_ntot/pore_hist.pyx:57-80, 154-184).  For each
// state point s:  v_ij = lnPI_ij + a1[s]*op1_i + a2[s]*op2_j  over j in [lo_i, hi_i),
//     out[s] = ( ln sum exp v,  <op1>,  <op2>,  <prop_q> ... ).
//
// Mapping: a CTA owns a chunk of consecutive rows, staged once in shared memory (one TMA bulk copy
// when 16-byte aligned), and 128 state points, one per thread.  All lanes walk the same bins, so
// every shared-memory read is a broadcast and the 2 MiB surface is read from L2/HBM once per
// (row chunk, 128 state points).  Chunk-local (max, sums) partials are merged by a second kernel.
#include "fhmc_common.cuh"

namespace fhmc {

#define FHMC_2D_CTA 128
#define FHMC_2D_MAXPROP 2

struct Rw2dArgs {
    const double *lnpi;
    const int *bounds;
    const double *op1, *op2, *props;
    const double *a1, *a2;
    double *ws, *out;
    long long n_states;
    int n1, n2, n_prop, rows, n_chunks;
};

template <int NPROP>
__global__ void __launch_bounds__(FHMC_2D_CTA) k_rw2d_partial(const __grid_constant__ Rw2dArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_tab[64];
    stage_exp_table(s_tab);
    const uint32_t tab = smem_u32(s_tab);
    const int n2 = a.n2;
    const int r0 = blockIdx.x * a.rows;
    const int nr = min(a.rows, a.n1 - r0);
    double *s_lnpi = reinterpret_cast<double *>(smem_raw);
    double *s_prop = s_lnpi + (size_t)a.rows * n2;
    double *s_op2 = s_prop + (size_t)NPROP * a.rows * n2;
    int *s_bounds = reinterpret_cast<int *>(s_op2 + n2);
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_bounds + 2 * a.rows + (2 * a.rows & 1 ? 1 : 0));
    // ---- stage the row chunk -------------------------------------------------------------
    const size_t chunk_bytes = (size_t)nr * n2 * 8;
    const double *g_lnpi = a.lnpi + (size_t)r0 * n2;
    bool tma_ok = (((uintptr_t)g_lnpi | chunk_bytes) & 15) == 0;
    for (int q = 0; q < NPROP; ++q) tma_ok = tma_ok && ((((uintptr_t)(a.props + ((size_t)q * a.n1 + r0) * n2)) & 15) == 0);
    // the property chunks land rows * n2 doubles apart in shared memory: bulk copies need 16-byte aligned destinations too
    if (NPROP > 0) tma_ok = tma_ok && ((((size_t)a.rows * n2 * 8) & 15) == 0);
    if (tma_ok) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (uint32_t)(chunk_bytes * (1 + NPROP)));
            const uint32_t step = 32768;
            for (int q = 0; q <= NPROP; ++q) {
                const char *src = (q == 0) ? (const char *)g_lnpi : (const char *)(a.props + ((size_t)(q - 1) * a.n1 + r0) * n2);
                char *dst = (q == 0) ? (char *)s_lnpi : (char *)(s_prop + (size_t)(q - 1) * a.rows * n2);
                for (size_t off = 0; off < chunk_bytes; off += step)
                    tma_bulk_g2s(dst + off, src + off, (uint32_t)min((size_t)step, chunk_bytes - off), bar);
            }
        }
    } else {
        for (size_t k = threadIdx.x; k < (size_t)nr * n2; k += FHMC_2D_CTA) {
            s_lnpi[k] = g_lnpi[k];
            for (int q = 0; q < NPROP; ++q) s_prop[(size_t)q * a.rows * n2 + k] = a.props[((size_t)q * a.n1 + r0) * n2 + k];
        }
    }
    for (int k = threadIdx.x; k < n2; k += FHMC_2D_CTA) s_op2[k] = a.op2[k];
    for (int k = threadIdx.x; k < 2 * nr; k += FHMC_2D_CTA) s_bounds[k] = a.bounds[2 * r0 + k];
    __syncthreads();
    if (tma_ok) mbar_wait(bar, 0);

    const long long sp = (long long)blockIdx.y * FHMC_2D_CTA + threadIdx.x;
    if (sp >= a.n_states) return;
    const double a1 = a.a1[sp], a2 = a.a2[sp];
    // ---- pass 1: chunk maximum -----------------------------------------------------------------
    double m = -CUDART_INF;
    for (int i = 0; i < nr; ++i) {
        const double ri = a1 * a.op1[r0 + i];
        const double *row = s_lnpi + (size_t)i * n2;
        const int lo = s_bounds[2 * i], hi = s_bounds[2 * i + 1];
        double m0 = -CUDART_INF, m1 = -CUDART_INF;
        int j = lo;
        for (; j + 1 < hi; j += 2) {
            m0 = fmax(m0, fma(a2, s_op2[j], row[j] + ri));
            m1 = fmax(m1, fma(a2, s_op2[j + 1], row[j + 1] + ri));
        }
        if (j < hi) m0 = fmax(m0, fma(a2, s_op2[j], row[j] + ri));
        m = fmax(m, fmax(m0, m1));
    }
    // ---- pass 2: shifted sums -------------------------------------------------------------------
    double S = 0.0, S1 = 0.0, S2 = 0.0, Sp[FHMC_2D_MAXPROP] = {0.0, 0.0};
    if (m > -CUDART_INF) {
        for (int i = 0; i < nr; ++i) {
            const double o1 = a.op1[r0 + i];
            const double ri = a1 * o1 - m;
            const double *row = s_lnpi + (size_t)i * n2;
            const int lo = s_bounds[2 * i], hi = s_bounds[2 * i + 1];
            double Sa = 0.0, Sb = 0.0, S2a = 0.0, S2b = 0.0;
            int j = lo;
            for (; j + 1 < hi; j += 2) {
                // (a bin of -inf inside the support -- zero probability -- is clamped to a term of ~1e-308)
                const double ea = exp_nonpos(fmax(fma(a2, s_op2[j], row[j] + ri), -800.0), tab);
                const double eb = exp_nonpos(fmax(fma(a2, s_op2[j + 1], row[j + 1] + ri), -800.0), tab);
                Sa += ea;
                Sb += eb;
                S2a = fma(ea, s_op2[j], S2a);
                S2b = fma(eb, s_op2[j + 1], S2b);
#pragma unroll
                for (int q = 0; q < NPROP; ++q) {
                    const double *pr = s_prop + ((size_t)q * a.rows + i) * n2;
                    Sp[q] = fma(ea, pr[j], Sp[q]);
                    Sp[q] = fma(eb, pr[j + 1], Sp[q]);
                }
            }
            if (j < hi) {
                const double ea = exp_nonpos(fmax(fma(a2, s_op2[j], row[j] + ri), -800.0), tab);
                Sa += ea;
                S2a = fma(ea, s_op2[j], S2a);
#pragma unroll
                for (int q = 0; q < NPROP; ++q) Sp[q] = fma(ea, s_prop[((size_t)q * a.rows + i) * n2 + j], Sp[q]);
            }
            const double Srow = Sa + Sb;
            S += Srow;
            S1 = fma(Srow, o1, S1);
            S2 += S2a + S2b;
        }
    }
    double *w = a.ws + ((size_t)sp * a.n_chunks + blockIdx.x) * (4 + NPROP);
    w[0] = m; w[1] = S; w[2] = S1; w[3] = S2;
#pragma unroll
    for (int q = 0; q < NPROP; ++q) w[4 + q] = Sp[q];
}

__global__ void __launch_bounds__(256) k_rw2d_merge(const __grid_constant__ Rw2dArgs a)
{
    const long long sp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sp >= a.n_states) return;
    const int rec = 4 + a.n_prop;
    const double *w = a.ws + (size_t)sp * a.n_chunks * rec;
    double m = -CUDART_INF;
    for (int k = 0; k < a.n_chunks; ++k) m = fmax(m, w[(size_t)k * rec]);
    double S = 0.0, S1 = 0.0, S2 = 0.0, Sp[FHMC_2D_MAXPROP] = {0.0, 0.0};
    for (int k = 0; k < a.n_chunks; ++k) {
        const double *r = w + (size_t)k * rec;
        if (!(r[0] > -CUDART_INF)) continue;
        const double f = exp(r[0] - m);
        S = fma(f, r[1], S);
        S1 = fma(f, r[2], S1);
        S2 = fma(f, r[3], S2);
        for (int q = 0; q < a.n_prop; ++q) Sp[q] = fma(f, r[4 + q], Sp[q]);
    }
    double *o = a.out + (size_t)sp * (3 + a.n_prop);
    o[0] = m + log(S);
    o[1] = S1 / S;
    o[2] = S2 / S;
    for (int q = 0; q < a.n_prop; ++q) o[3 + q] = Sp[q] / S;
}

// ---------------------------------------------------------------------------------------------------------------
// Product form (op2 uniformly spaced, moderate tilts; fhmc_reweight_2d_prod).  Along a row
//     exp(lnPI_ij + a1 op1_i + a2 op2_j - m) = P_ij * t_ij,  P_ij = exp(lnPI_ij - A_ig) tabulated (A_ig = max of the row's
// 128-bin segment g), t_ij = exp(A_ig + a1 op1_i + a2 op2_j - m) geometric in j, so the sums of a 4-bin block are Horner
// polynomials in exp(a2 d2): 4 FMAs per block and summed quantity (sum, op2 moment, properties) instead of an exp per bin,
// with one true exp per segment.  No maximum pass: any upper bound m of v over the row chunk keeps the terms <= 1, and
// max_i (rowmax_i + a1 op1_i) + max_j a2 op2_j is one.  Table entries are warp-uniform (broadcast LDS.128, two
// shared-memory wavefronts each), so a thread carries TWO state points per table load.
// ---------------------------------------------------------------------------------------------------------------
#define FHMC_2D_SEGB 32   // blocks per anchor segment
#define FHMC_2DP_CTA 256  // threads per CTA of the product-form kernel (two state points each)

struct Rw2dProdArgs {
    const double *lnpi;
    const int *bounds;
    const double *op1, *op2, *props;
    const double *a1, *a2;
    double *tab;       // [n1][nblk][4*(2+n_prop)] product tables
    double *anch;      // [n1][nseg] segment anchors (0 for a segment without finite bins), then rowmax[n1]
    double *ws, *out;
    long long n_states;
    int n1, n2, n_prop, rows, n_chunks, nblk, nseg;
};

__global__ void __launch_bounds__(128) k_rw2d_tables(const __grid_constant__ Rw2dProdArgs a)
{
    const int i = blockIdx.x;
    const int NQ = 2 + a.n_prop;
    const double *row = a.lnpi + (size_t)i * a.n2;
    const int lo = a.bounds[2 * i], hi = a.bounds[2 * i + 1];
    __shared__ double s_anch[64];
    __shared__ double s_w[4];
    double rmax = -CUDART_INF;
    for (int g = threadIdx.x; g < a.nseg; g += blockDim.x) {
        double m = -CUDART_INF;
        const int j0 = max(lo, g * 4 * FHMC_2D_SEGB), j1 = min(min(hi, a.n2), (g + 1) * 4 * FHMC_2D_SEGB);
        for (int j = j0; j < j1; ++j) m = fmax(m, row[j]);
        if (g < 64) s_anch[g] = m;
        a.anch[(size_t)i * a.nseg + g] = (m > -CUDART_INF) ? m : 0.0;
        rmax = fmax(rmax, m);
    }
    rmax = group_max<32>(rmax, 0xffffffffu);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = rmax;
    __syncthreads();
    if (threadIdx.x == 0) a.anch[(size_t)a.n1 * a.nseg + i] = fmax(fmax(s_w[0], s_w[1]), fmax(s_w[2], s_w[3]));
    double *trow = a.tab + (size_t)i * a.nblk * 4 * NQ;
    for (int j = threadIdx.x; j < a.nblk * 4; j += blockDim.x) {
        const int b = j >> 2, k = j & 3, g = b / FHMC_2D_SEGB;
        double P = 0.0;
        if (j >= lo && j < hi && j < a.n2) {
            const double A = (g < 64) ? s_anch[g] : a.anch[(size_t)i * a.nseg + g];   // (a finite x implies a finite anchor)
            const double x = row[j];
            if (x > -CUDART_INF) P = exp(x - A);
        }
        double *tb = trow + (size_t)b * 4 * NQ;
        tb[k] = P;
        tb[4 + k] = (P > 0.0) ? P * a.op2[j] : 0.0;
        for (int q = 0; q < a.n_prop; ++q) tb[8 + 4 * q + k] = (P > 0.0) ? P * a.props[((size_t)q * a.n1 + i) * a.n2 + j] : 0.0;
    }
}

template <int NPROP>
__global__ void __launch_bounds__(FHMC_2DP_CTA) k_rw2d_prod(const __grid_constant__ Rw2dProdArgs a)
{
    constexpr int NQ = 2 + NPROP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_tab[64];
    __shared__ uint64_t bar;
    stage_exp_table(s_tab);
    const uint32_t tab = smem_u32(s_tab);
    const int r0 = blockIdx.x * a.rows;
    const int nr = min(a.rows, a.n1 - r0);
    const size_t row_d = (size_t)a.nblk * 4 * NQ;   // doubles per table row
    double *s_t = reinterpret_cast<double *>(smem_raw);
    double *s_anch = s_t + (size_t)a.rows * row_d;
    double *s_op2 = s_anch + (size_t)a.rows * a.nseg;
    double *s_rmax = s_op2 + a.n2;
    int *s_bounds = reinterpret_cast<int *>(s_rmax + a.rows);
    // ---- stage the chunk's tables with one TMA bulk copy stream (contiguous rows, 32-byte multiples) ----------------
    const size_t chunk_bytes = (size_t)nr * row_d * 8;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)chunk_bytes);
        const char *src = reinterpret_cast<const char *>(a.tab + (size_t)r0 * row_d);
        const uint32_t step = 32768;
        for (size_t off = 0; off < chunk_bytes; off += step)
            tma_bulk_g2s(reinterpret_cast<char *>(s_t) + off, src + off, (uint32_t)min((size_t)step, chunk_bytes - off), &bar);
    }
    for (int k = threadIdx.x; k < nr * a.nseg; k += FHMC_2DP_CTA) s_anch[k] = a.anch[(size_t)r0 * a.nseg + k];
    for (int k = threadIdx.x; k < a.n2; k += FHMC_2DP_CTA) s_op2[k] = a.op2[k];
    for (int k = threadIdx.x; k < nr; k += FHMC_2DP_CTA) s_rmax[k] = a.anch[(size_t)a.n1 * a.nseg + r0 + k];
    for (int k = threadIdx.x; k < 2 * nr; k += FHMC_2DP_CTA) s_bounds[k] = a.bounds[2 * r0 + k];
    __syncthreads();
    mbar_wait(&bar, 0);

    const long long sp0 = ((long long)blockIdx.y * FHMC_2DP_CTA + threadIdx.x) * 2;   // this thread: state points sp0, sp0 + 1
    if (sp0 >= a.n_states) return;
    const bool two = sp0 + 1 < a.n_states;
    double a1[2], a2[2], r1[2], r2[2], r4[2], m[2], S[2], S1[2], S2[2], Sp[2][NPROP > 0 ? NPROP : 1];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const long long sp = (k == 0 || two) ? sp0 + k : sp0;
        a1[k] = a.a1[sp];
        a2[k] = a.a2[sp];
        const double d2 = (a.n2 > 1) ? s_op2[1] - s_op2[0] : 0.0;
        r1[k] = exp(a2[k] * d2);
        r2[k] = r1[k] * r1[k];
        r4[k] = r2[k] * r2[k];
        double mm = -CUDART_INF;
        for (int i = 0; i < nr; ++i) mm = fmax(mm, s_rmax[i] + a1[k] * a.op1[r0 + i]);
        m[k] = mm + fmax(a2[k] * s_op2[0], a2[k] * s_op2[a.n2 - 1]);
        S[k] = S1[k] = S2[k] = 0.0;
#pragma unroll
        for (int q = 0; q < NPROP; ++q) Sp[k][q] = 0.0;
    }
    if (m[0] > -CUDART_INF || m[1] > -CUDART_INF) {
        for (int i = 0; i < nr; ++i) {
            const int lo = s_bounds[2 * i], hi = min(s_bounds[2 * i + 1], a.n2);
            if (hi <= lo || !(s_rmax[i] > -CUDART_INF)) continue;
            const double o1 = a.op1[r0 + i];
            const uint32_t rowaddr = smem_u32(s_t + (size_t)i * row_d);
            const int b_end = (hi + 3) >> 2;
            double rowS[2] = {0.0, 0.0};
            for (int b = lo >> 2; b < b_end;) {
                const int g = b / FHMC_2D_SEGB;
                const int bseg = min(b_end, (g + 1) * FHMC_2D_SEGB);
                const double A = s_anch[i * a.nseg + g];
                double t[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) t[k] = exp_nonpos(fma(a2[k], s_op2[4 * b], fma(a1[k], o1, A)) - m[k], tab);
                uint32_t pb = rowaddr + (uint32_t)b * (uint32_t)(32 * NQ);
                for (; b < bseg; ++b, pb += 32u * NQ) {
                    double tb[4 * NQ];
#pragma unroll
                    for (int v = 0; v < 2 * NQ; ++v)
                        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tb[2 * v]), "=d"(tb[2 * v + 1]) : "r"(pb + 16u * v));
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        rowS[k] = fma(fma(fma(tb[3], r1[k], tb[2]), r2[k], fma(tb[1], r1[k], tb[0])), t[k], rowS[k]);
                        S2[k] = fma(fma(fma(tb[7], r1[k], tb[6]), r2[k], fma(tb[5], r1[k], tb[4])), t[k], S2[k]);
#pragma unroll
                        for (int q = 0; q < NPROP; ++q)
                            Sp[k][q] = fma(fma(fma(tb[11 + 4 * q], r1[k], tb[10 + 4 * q]), r2[k], fma(tb[9 + 4 * q], r1[k], tb[8 + 4 * q])), t[k], Sp[k][q]);
                        t[k] *= r4[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                S[k] += rowS[k];
                S1[k] = fma(rowS[k], o1, S1[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
        double *w = a.ws + ((size_t)(sp0 + k) * a.n_chunks + blockIdx.x) * (4 + NPROP);
        const bool any = S[k] > 0.0;
        w[0] = any ? m[k] : -CUDART_INF;
        w[1] = S[k]; w[2] = S1[k]; w[3] = S2[k];
#pragma unroll
        for (int q = 0; q < NPROP; ++q) w[4 + q] = Sp[k][q];
    }
}

static int plan_rows_prod(int n1, int n2, int n_prop, int smem_budget)
{
    const long long nblk = (n2 + 3) / 4, nseg = (nblk + FHMC_2D_SEGB - 1) / FHMC_2D_SEGB;
    const long long per_row = nblk * 32 * (2 + n_prop) + nseg * 8 + 8 + 8;
    long long rows = ((long long)smem_budget - (long long)n2 * 8 - 64) / per_row;
    if (rows > n1) rows = n1;
    return (int)rows;   // < 1: a row does not fit
}

static int plan_rows(int n1, int n2, int n_prop)
{
    long long per_row = (long long)n2 * 8 * (1 + n_prop);
    long long rows = 49152 / per_row;
    if (rows < 1) rows = 1;
    if (rows > n1) rows = n1;
    return (int)rows;
}

}  // namespace fhmc

using namespace fhmc;

extern "C" size_t fhmc_reweight_2d_workspace(int n1, int n2, int n_prop, long long n_states)
{
    if (n1 < 1 || n2 < 1 || n_prop < 0 || n_prop > FHMC_2D_MAXPROP || n_states < 0) return 0;
    const int rows = plan_rows(n1, n2, n_prop);
    const int n_chunks = (n1 + rows - 1) / rows;
    return (size_t)n_states * n_chunks * (4 + n_prop) * sizeof(double);
}

extern "C" int fhmc_reweight_2d(const double *lnpi, const int *bounds, int n1, int n2, const double *op1, const double *op2,
                                const double *props, int n_prop, const double *a1, const double *a2, long long n_states,
                                double *out, double *workspace, size_t workspace_bytes, void *stream)
{
    if (!lnpi || !bounds || !op1 || !op2 || !a1 || !a2 || !out || n1 < 1 || n2 < 1 || n_states < 0) { set_error("bad arguments"); return 1; }
    if (n_prop < 0 || n_prop > FHMC_2D_MAXPROP || (n_prop > 0 && !props)) { set_error("n_prop must be in [0,%d]", FHMC_2D_MAXPROP); return 1; }
    if (n_states == 0) return 0;
    const size_t need = fhmc_reweight_2d_workspace(n1, n2, n_prop, n_states);
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes", need); return 1; }
    Rw2dArgs a;
    a.lnpi = lnpi; a.bounds = bounds; a.op1 = op1; a.op2 = op2; a.props = props; a.a1 = a1; a.a2 = a2;
    a.ws = workspace; a.out = out; a.n_states = n_states; a.n1 = n1; a.n2 = n2; a.n_prop = n_prop;
    a.rows = plan_rows(n1, n2, n_prop);
    a.n_chunks = (n1 + a.rows - 1) / a.rows;
    const size_t smem = (size_t)a.rows * n2 * 8 * (1 + n_prop) + (size_t)n2 * 8 + (size_t)(2 * a.rows + 2) * 4 + 16;
    int dev = 0, smem_optin = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (check_cuda(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), "device attribute")) return 1;
    if (smem > (size_t)smem_optin) { set_error("a single row of the joint histogram does not fit in shared memory"); return 1; }
    const long long tiles = (n_states + FHMC_2D_CTA - 1) / FHMC_2D_CTA;
    if (tiles > 65535) { set_error("too many state points per call (max %d)", 65535 * FHMC_2D_CTA); return 1; }
    dim3 grid(a.n_chunks, (unsigned)tiles);
    cudaStream_t s = (cudaStream_t)stream;
#define FHMC_LAUNCH_2D(NP)                                                                                              \
    do {                                                                                                                \
        if (check_cuda(cudaFuncSetAttribute(k_rw2d_partial<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                       "cudaFuncSetAttribute")) return 1;                                                               \
        k_rw2d_partial<NP><<<grid, FHMC_2D_CTA, smem, s>>>(a);                                                          \
    } while (0)
    if (n_prop == 0) FHMC_LAUNCH_2D(0);
    else if (n_prop == 1) FHMC_LAUNCH_2D(1);
    else FHMC_LAUNCH_2D(2);
#undef FHMC_LAUNCH_2D
    if (check_cuda(cudaGetLastError(), "k_rw2d_partial launch")) return 1;
    note_kernel("k_rw2d_partial");
    k_rw2d_merge<<<(unsigned)((n_states + 255) / 256), 256, 0, s>>>(a);
    return check_cuda(cudaGetLastError(), "k_rw2d_merge launch");
}

// ---------------------------------------------------------------------------------------------------------------
// product-form entry point: same outputs as fhmc_reweight_2d; the caller promises uniformly spaced op2 and
// max_s |a2[s]| * (op2[n2-1] - op2[0]) < 300 (otherwise use fhmc_reweight_2d).
// workspace = tables + anchors + per-chunk partials (fhmc_reweight_2d_prod_workspace bytes).
// ---------------------------------------------------------------------------------------------------------------
#define FHMC_2D_PROD_SMEM (96 * 1024)

static size_t prod_table_doubles(int n1, int n2, int n_prop)
{
    const size_t nblk = (size_t)(n2 + 3) / 4, nseg = (nblk + FHMC_2D_SEGB - 1) / FHMC_2D_SEGB;
    return (size_t)n1 * nblk * 4 * (2 + n_prop) + (size_t)n1 * nseg + (size_t)n1 + 4;
}

extern "C" size_t fhmc_reweight_2d_prod_workspace(int n1, int n2, int n_prop, long long n_states)
{
    if (n1 < 1 || n2 < 1 || n_prop < 0 || n_prop > FHMC_2D_MAXPROP || n_states < 0) return 0;
    const int rows = plan_rows_prod(n1, n2, n_prop, FHMC_2D_PROD_SMEM);
    if (rows < 1) return 0;
    const int n_chunks = (n1 + rows - 1) / rows;
    return (prod_table_doubles(n1, n2, n_prop) + (size_t)n_states * n_chunks * (4 + n_prop)) * sizeof(double);
}

extern "C" int fhmc_reweight_2d_prod(const double *lnpi, const int *bounds, int n1, int n2, const double *op1, const double *op2,
                                     const double *props, int n_prop, const double *a1, const double *a2, long long n_states,
                                     double *out, double *workspace, size_t workspace_bytes, void *stream)
{
    if (!lnpi || !bounds || !op1 || !op2 || !a1 || !a2 || !out || n1 < 1 || n2 < 1 || n_states < 0) { set_error("bad arguments"); return 1; }
    if (n_prop < 0 || n_prop > FHMC_2D_MAXPROP || (n_prop > 0 && !props)) { set_error("n_prop must be in [0,%d]", FHMC_2D_MAXPROP); return 1; }
    if (n_states == 0) return 0;
    const size_t need = fhmc_reweight_2d_prod_workspace(n1, n2, n_prop, n_states);
    if (need == 0) { set_error("a single table row of the joint histogram does not fit in shared memory"); return 1; }
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes", need); return 1; }
    Rw2dProdArgs a;
    a.lnpi = lnpi; a.bounds = bounds; a.op1 = op1; a.op2 = op2; a.props = props; a.a1 = a1; a.a2 = a2; a.out = out;
    a.n_states = n_states; a.n1 = n1; a.n2 = n2; a.n_prop = n_prop;
    a.nblk = (n2 + 3) / 4;
    a.nseg = (a.nblk + FHMC_2D_SEGB - 1) / FHMC_2D_SEGB;
    a.rows = plan_rows_prod(n1, n2, n_prop, FHMC_2D_PROD_SMEM);
    a.n_chunks = (n1 + a.rows - 1) / a.rows;
    a.tab = workspace;
    a.anch = workspace + (size_t)n1 * a.nblk * 4 * (2 + n_prop);
    a.ws = workspace + prod_table_doubles(n1, n2, n_prop);
    const long long tiles = (n_states + 2 * FHMC_2DP_CTA - 1) / (2 * FHMC_2DP_CTA);
    if (tiles > 65535) { set_error("too many state points per call (max %d)", 65535 * 2 * FHMC_2DP_CTA); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    k_rw2d_tables<<<n1, 128, 0, s>>>(a);
    if (check_cuda(cudaGetLastError(), "k_rw2d_tables launch")) return 1;
    const size_t row_d = (size_t)a.nblk * 4 * (2 + n_prop);
    const size_t smem = ((size_t)a.rows * row_d + (size_t)a.rows * a.nseg + (size_t)n2 + (size_t)a.rows) * 8 + (size_t)(2 * a.rows + 2) * 4 + 16;
    dim3 grid(a.n_chunks, (unsigned)tiles);
#define FHMC_LAUNCH_2DP(NP)                                                                                            \
    do {                                                                                                               \
        if (check_cuda(cudaFuncSetAttribute(k_rw2d_prod<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),    \
                       "cudaFuncSetAttribute")) return 1;                                                              \
        k_rw2d_prod<NP><<<grid, FHMC_2DP_CTA, smem, s>>>(a);                                                            \
    } while (0)
    if (n_prop == 0) FHMC_LAUNCH_2DP(0);
    else if (n_prop == 1) FHMC_LAUNCH_2DP(1);
    else FHMC_LAUNCH_2DP(2);
#undef FHMC_LAUNCH_2DP
    if (check_cuda(cudaGetLastError(), "k_rw2d_prod launch")) return 1;
    note_kernel("k_rw2d_prod");
    Rw2dArgs mga;
    mga.lnpi = lnpi; mga.bounds = bounds; mga.op1 = op1; mga.op2 = op2; mga.props = props; mga.a1 = a1; mga.a2 = a2;
    mga.ws = a.ws; mga.out = out; mga.n_states = n_states; mga.n1 = n1; mga.n2 = n2; mga.n_prop = n_prop;
    mga.rows = a.rows; mga.n_chunks = a.n_chunks;
    k_rw2d_merge<<<(unsigned)((n_states + 255) / 256), 256, 0, s>>>(mga);
    return check_cuda(cudaGetLastError(), "k_rw2d_merge launch");
}
