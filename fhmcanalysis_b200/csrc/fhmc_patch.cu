// fhmc_patch.cu -- batched window-patching shift solve.
//
// Reference: moments/win_patch/fhmc_patch.pyx:640-709 (identical copies in chkpt_patch.pyx:638 and feasst_patch.pyx:384):
// patch_window_pair() minimises window_patch_error(x) = sum_i ((a_i + x) - b_i)^2 over the overlap of two windows with
// scipy.optimize.fmin and returns (x*, error(x*) / len).  The objective is a parabola in x: x* = mean(b - a) in closed form
// and error(x*) = sum (a_i + x* - b_i)^2.  One warp per window pair: coalesced 8-byte loads of both slices, shuffle
// reductions, two passes (mean first, squared residuals about it second) so that the error does not lose digits to
// cancellation.  HBM-bound: 2 * 8 bytes per overlapping bin, read twice (second pass from L2).
#include "fhmc_common.cuh"

namespace fhmc {

__global__ void __launch_bounds__(256) k_patch_shifts(const double *__restrict__ a, const double *__restrict__ b,
                                                      const long long *__restrict__ offsets, int n_pairs,
                                                      double *__restrict__ shift, double *__restrict__ err2)
{
    const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const long long lo = offsets[w], hi = offsets[w + 1], len = hi - lo;
    double s = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long long i = lo + lane;
    for (; i + 96 < hi; i += 128) {   // eight independent loads in flight per lane
        s += b[i] - a[i];
        s1 += b[i + 32] - a[i + 32];
        s2 += b[i + 64] - a[i + 64];
        s3 += b[i + 96] - a[i + 96];
    }
    for (; i < hi; i += 32) s += b[i] - a[i];
    s = group_sum<32>((s + s1) + (s2 + s3), 0xffffffffu);
    const double x = (len > 0) ? s / (double)len : CUDART_NAN;
    double e = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    i = lo + lane;
    for (; i + 96 < hi; i += 128) {
        const double r0 = (a[i] + x) - b[i], r1 = (a[i + 32] + x) - b[i + 32];   // as window_patch_error (fhmc_patch.pyx:663)
        const double r2 = (a[i + 64] + x) - b[i + 64], r3 = (a[i + 96] + x) - b[i + 96];
        e = fma(r0, r0, e);
        e1 = fma(r1, r1, e1);
        e2 = fma(r2, r2, e2);
        e3 = fma(r3, r3, e3);
    }
    for (; i < hi; i += 32) {
        const double r = (a[i] + x) - b[i];
        e = fma(r, r, e);
    }
    e = group_sum<32>((e + e1) + (e2 + e3), 0xffffffffu);
    if (lane == 0) {
        shift[w] = x;
        err2[w] = (len > 0) ? e / (double)len : CUDART_NAN;
    }
}

}  // namespace fhmc

using namespace fhmc;

extern "C" int fhmc_patch_shifts(const double *a, const double *b, const long long *offsets, int n_pairs, double *shift,
                                 double *err2, void *stream)
{
    if (!a || !b || !offsets || !shift || !err2 || n_pairs < 0) { set_error("bad arguments"); return 1; }
    if (n_pairs == 0) return 0;
    const int warps_per_cta = 8;
    k_patch_shifts<<<(n_pairs + warps_per_cta - 1) / warps_per_cta, 32 * warps_per_cta, 0, (cudaStream_t)stream>>>(a, b, offsets, n_pairs, shift, err2);
    return check_cuda(cudaGetLastError(), "k_patch_shifts launch");
}
