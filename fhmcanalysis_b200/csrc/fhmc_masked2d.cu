// fhmc_masked2d.cu -- one-shot log-sum-exp and masked averages over a 2-D surface lnPI(h, N):
//   pore_hist.normalize  (two_dim/h_ntot/pore_hist.pyx:57-80, 147-152): lnPI -= ln sum_{i, j <= edge_i} exp lnPI_ij
//   pore_hist.thermo     (pore_hist.pyx:154-184): averages of every property matrix over the bins a mask selects,
//                        and the positions of the largest lnPI inside the mask.
// Both are single passes over n1*n2 fp64 values (+ one matrix per property): HBM-bound streaming reductions.
//   k_m2d_max     grid-stride, coalesced 8-byte loads, CTA tree reduction, one atomicMax on an order-preserving key
//   k_m2d_sums    same walk with the shift known: sum exp(x - M), sum exp(x - M) * prop_q; per-CTA partials (deterministic),
//                 and the flat indices of the bins equal to M
//   k_m2d_final   one CTA folds the partials in a fixed order -> out
// Algorithmic bytes per call: 8 n1 n2 (2 + n_prop) + 2 n1 n2 (mask, read twice).
#include "fhmc_common.cuh"

namespace fhmc {

#define FHMC_M2D_CTA 256
#define FHMC_M2D_MAXPROP 8

struct M2dArgs {
    const double *lnpi;
    const unsigned char *mask;   // nullable
    const int *edge;             // nullable: row i keeps j <= edge[i]
    const double *props;         // [n_prop][n1*n2]
    double *ws;                  // [0] key of the maximum (as u64), [1] peak counter (as u64), then partials[grid][1+n_prop]
    double *out;                 // [0] ln sum exp, [1] max, [2+q] <prop_q>
    long long *peak;             // [0] count, [1..cap] flat indices
    long long total;
    int n2, n_prop, peak_cap, grid;
};

__device__ __forceinline__ unsigned long long f64_key(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ bool m2d_selected(const M2dArgs &a, long long idx)
{
    if (a.mask && !a.mask[idx]) return false;
    if (a.edge) {
        const long long i = idx / a.n2;
        if ((int)(idx - i * a.n2) > a.edge[i]) return false;
    }
    return true;
}

__global__ void __launch_bounds__(FHMC_M2D_CTA) k_m2d_max(const __grid_constant__ M2dArgs a)
{
    __shared__ double s_red[FHMC_M2D_CTA / 32];
    double m = -CUDART_INF;
    for (long long idx = (long long)blockIdx.x * FHMC_M2D_CTA + threadIdx.x; idx < a.total; idx += (long long)gridDim.x * FHMC_M2D_CTA)
        if (m2d_selected(a, idx)) m = fmax(m, a.lnpi[idx]);
    m = group_max<32>(m, 0xffffffffu);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < FHMC_M2D_CTA / 32; ++w) m = fmax(m, s_red[w]);
        atomicMax(reinterpret_cast<unsigned long long *>(a.ws), f64_key(m));
    }
}

__global__ void __launch_bounds__(FHMC_M2D_CTA) k_m2d_sums(const __grid_constant__ M2dArgs a)
{
    __shared__ double s_tab[64];
    __shared__ double s_red[FHMC_M2D_CTA / 32][1 + FHMC_M2D_MAXPROP];
    stage_exp_table(s_tab);
    __syncthreads();
    const uint32_t tab = smem_u32(s_tab);
    const double M = f64_unkey(*reinterpret_cast<const unsigned long long *>(a.ws));
    double acc[1 + FHMC_M2D_MAXPROP];
#pragma unroll
    for (int q = 0; q <= FHMC_M2D_MAXPROP; ++q) acc[q] = 0.0;
    if (M > -CUDART_INF) {
        for (long long idx = (long long)blockIdx.x * FHMC_M2D_CTA + threadIdx.x; idx < a.total; idx += (long long)gridDim.x * FHMC_M2D_CTA) {
            if (!m2d_selected(a, idx)) continue;
            const double x = a.lnpi[idx];
            if (x == M) {
                const unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long *>(a.ws) + 1, 1ull);
                if (k < (unsigned long long)a.peak_cap) a.peak[1 + k] = idx;
            }
            const double e = (x > -CUDART_INF) ? exp_nonpos(x - M, tab) : 0.0;
            acc[0] += e;
#pragma unroll
            for (int q = 0; q < FHMC_M2D_MAXPROP; ++q)
                if (q < a.n_prop && e > 0.0) acc[1 + q] = fma(e, a.props[(size_t)q * a.total + idx], acc[1 + q]);
        }
    }
#pragma unroll
    for (int q = 0; q <= FHMC_M2D_MAXPROP; ++q) {
        const double v = group_sum<32>(acc[q], 0xffffffffu);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x <= a.n_prop) {
        double v = 0.0;
        for (int w = 0; w < FHMC_M2D_CTA / 32; ++w) v += s_red[w][threadIdx.x];
        a.ws[2 + (size_t)blockIdx.x * (1 + a.n_prop) + threadIdx.x] = v;
    }
}

__global__ void k_m2d_final(const __grid_constant__ M2dArgs a)
{
    const int q = threadIdx.x;   // one warp: lane q folds column q of the partials, in a fixed order
    __shared__ double s_S;
    double v = 0.0;
    if (q <= a.n_prop)
        for (int b = 0; b < a.grid; ++b) v += a.ws[2 + (size_t)b * (1 + a.n_prop) + q];
    if (q == 0) s_S = v;
    __syncthreads();
    const double M = f64_unkey(*reinterpret_cast<const unsigned long long *>(a.ws));
    if (q == 0) {
        a.out[0] = (M > -CUDART_INF) ? M + log(v) : -CUDART_INF;
        a.out[1] = M;
        a.peak[0] = (long long)*(reinterpret_cast<const unsigned long long *>(a.ws) + 1);
    } else if (q <= a.n_prop) {
        a.out[1 + q] = v / s_S;
    }
}

// lnpi_out = lnpi - shift (pore_hist.normalize writes the shifted surface back, PH:80)
__global__ void __launch_bounds__(FHMC_M2D_CTA) k_m2d_shift(const double *lnpi, const double *out, double *dst, long long total)
{
    const double c = out[0];
    for (long long idx = (long long)blockIdx.x * FHMC_M2D_CTA + threadIdx.x; idx < total; idx += (long long)gridDim.x * FHMC_M2D_CTA)
        dst[idx] = lnpi[idx] - c;
}

static int m2d_grid(long long total)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long g = (total + FHMC_M2D_CTA * 4 - 1) / (FHMC_M2D_CTA * 4);
    if (g > (long long)sms * 8) g = (long long)sms * 8;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace fhmc

using namespace fhmc;

extern "C" size_t fhmc_masked_lse_2d_workspace(int n1, int n2, int n_prop)
{
    if (n1 < 1 || n2 < 1 || n_prop < 0 || n_prop > FHMC_M2D_MAXPROP) return 0;
    return (size_t)(2 + (size_t)m2d_grid((long long)n1 * n2) * (1 + n_prop)) * sizeof(double);
}

extern "C" int fhmc_masked_lse_2d(const double *lnpi, const unsigned char *mask, const int *edge, int n1, int n2,
                                  const double *props, int n_prop, double *out, long long *peak, int peak_cap,
                                  double *lnpi_shifted, double *workspace, size_t workspace_bytes, void *stream)
{
    if (!lnpi || !out || !peak || n1 < 1 || n2 < 1 || peak_cap < 0) { set_error("bad arguments"); return 1; }
    if (n_prop < 0 || n_prop > FHMC_M2D_MAXPROP || (n_prop > 0 && !props)) { set_error("n_prop must be in [0,%d]", FHMC_M2D_MAXPROP); return 1; }
    const size_t need = fhmc_masked_lse_2d_workspace(n1, n2, n_prop);
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes", need); return 1; }
    M2dArgs a;
    a.lnpi = lnpi; a.mask = mask; a.edge = edge; a.props = props; a.ws = workspace; a.out = out; a.peak = peak;
    a.total = (long long)n1 * n2; a.n2 = n2; a.n_prop = n_prop; a.peak_cap = peak_cap;
    a.grid = m2d_grid(a.total);
    cudaStream_t s = (cudaStream_t)stream;
    if (check_cuda(cudaMemsetAsync(workspace, 0, 16, s), "cudaMemsetAsync")) return 1;   // key 0 = below every double
    k_m2d_max<<<a.grid, FHMC_M2D_CTA, 0, s>>>(a);
    k_m2d_sums<<<a.grid, FHMC_M2D_CTA, 0, s>>>(a);
    k_m2d_final<<<1, 32, 0, s>>>(a);
    if (lnpi_shifted) k_m2d_shift<<<a.grid, FHMC_M2D_CTA, 0, s>>>(lnpi, out, lnpi_shifted, a.total);
    return check_cuda(cudaGetLastError(), "masked 2-D reduction launch");
}
