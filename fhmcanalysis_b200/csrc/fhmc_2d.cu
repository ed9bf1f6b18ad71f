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
                const double ea = exp_nonpos(fma(a2, s_op2[j], row[j] + ri), tab);
                const double eb = exp_nonpos(fma(a2, s_op2[j + 1], row[j + 1] + ri), tab);
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
                const double ea = exp_nonpos(fma(a2, s_op2[j], row[j] + ri), tab);
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
    k_rw2d_merge<<<(unsigned)((n_states + 255) / 256), 256, 0, s>>>(a);
    return check_cuda(cudaGetLastError(), "k_rw2d_merge launch");
}
