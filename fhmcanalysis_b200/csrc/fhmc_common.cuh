// fhmc_common.cuh -- device helpers shared by the sm_100a kernels of libfhmc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

#include "fhmc_b200.h"

#define FHMC_CTA 256  // threads per CTA of the 1-D kernels

namespace fhmc {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

// ---------------------------------------------------------------------------------------------
// fp64 exp for arguments t <= 0 (the only kind the max-shifted sums produce).
//   k = rint(t*log2 e) by the 1.5*2^52 trick, r = t - k ln2 (two-constant Cody-Waite),
//   exp(r) by a degree-13 Taylor polynomial (|r| <= 0.3466 -> remainder 4e-18), 2^k by integer add to
//   the exponent field (no fp64 op).  Inputs below -707 return 0: the reference sums them as
//   (sub)normal numbers < 1e-307 next to a leading term of 1 (np.seterr(under='ignore'), GH:29).
//   17 fp64-pipe instructions, no special-case branch, max error < 1 ulp.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double exp_nonpos(double t)
{
    const double magic = 6755399441055744.0;
    const double kd0 = fma(t, 1.4426950408889634074, magic);
    const int k = __double2loint(kd0);
    const double kd = kd0 - magic;
    double r = fma(kd, -6.93147180369123816490e-01, t);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int hi = __double2hiint(p) + (k << 20);
    const double res = __hiloint2double(hi, __double2loint(p));
    return (t < -707.0) ? 0.0 : res;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) used to stage a histogram blob
// in shared memory once per CTA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Stage `bytes` (multiple of 16) from global `src` (16-byte aligned) into shared `dst`.
// Called by every thread of the CTA; returns when the data is visible to all of them.
__device__ __forceinline__ void stage_blob(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        const uint32_t chunk = 32768;
        for (uint32_t off = 0; off < bytes; off += chunk) {
            const uint32_t nb = (bytes - off < chunk) ? (bytes - off) : chunk;
            tma_bulk_g2s((char *)dst + off, (const char *)src + off, nb, bar);
        }
    }
    mbar_wait(bar, 0);
}

// ---------------------------------------------------------------------------------------------
// sub-warp group reductions (G lanes cooperate on one state point); xor butterflies give every
// lane the same, order-deterministic result.
// ---------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned member)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(member, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v, unsigned member)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(member, v, o));
    return v;
}
template <int G>
__device__ __forceinline__ double group_min(double v, unsigned member)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(member, v, o));
    return v;
}

__device__ __forceinline__ double monomial(int kind, double dB, double dD, double mu1)
{
    switch (kind) {
    case FHMC_M_DB: return dB;
    case FHMC_M_DD: return dD;
    case FHMC_M_DB2: return 0.5 * dB * dB;
    case FHMC_M_DBDD: return dB * dD;
    case FHMC_M_DD2: return 0.5 * dD * dD;
    case FHMC_M_DB3: return dB * dB * dB * (1.0 / 6.0);
    case FHMC_M_DB_MU1: return dB * mu1;
    default: return 1.0;
    }
}

}  // namespace fhmc
