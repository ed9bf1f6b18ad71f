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
void note_kernel(const char *name);   // diagnostic: name of the sweep / solver kernel this thread launched last (fhmc_last_kernel)

// ---------------------------------------------------------------------------------------------
// fp64 exp for the max-shifted sums:   exp_scaled(u, Mq) = exp(u - Mq*ln2),  u - Mq*ln2 <~ 0.
//   k = rint(u * 64/ln2) (1.5*2^52 trick), r = u - k*ln2/64 (two-constant Cody-Waite, |r| <= ln2/128),
//   exp(r) by a degree-5 Taylor polynomial (remainder 3.5e-17), 2^(j/64) (j = k & 63) from a 64-entry
//   shared-memory table, 2^((k>>6) - Mq) by an integer add to the exponent field (no fp64 op).
//   10 fp64-pipe instructions (the register-resident roofline micro-benchmark times exactly this
//   function), no special-case branch, max error ~2 ulp.  Results below 2^-1021 return 0: the reference
//   sums them as (sub)normal numbers < 1e-307 next to a leading term of ~1 (np.seterr(under='ignore'), GH:29).
// The polynomial/reduction constants live in __constant__ memory so that DFMA takes them as
// constant-bank operands (no per-use UMOV/IMAD.MOV materialisation).
// ---------------------------------------------------------------------------------------------
struct ExpConst {
    double inv, hi, lo, c5, c4, c3, c2, ln2_hi, ln2_lo, inv_ln2;
    double tab[64];
};
__constant__ ExpConst c_exp = {
    92.33248261689366,        // 64/ln2
    0.01083042469326756,      // ln2/64, upper 32 bits (k*hi exact for |k| < 2^20)
    2.9815858269852933e-12,   // ln2/64 - hi
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
    0.6931471803691238, 1.9082149292705877e-10, 1.4426950408889634074,
    {1, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.1023825833078409, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.2021567314527031, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.2553807570246911, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.3396675240533029,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.5590044002378369, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.6457554781539649, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.7186192981224779, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.9784560263879509}};

#define FHMC_EXP_MAGIC 6755399441055744.0

// stage the 2^(j/64) table into shared memory (call from all threads, then __syncthreads)
__device__ __forceinline__ void stage_exp_table(double *s_tab)
{
    for (int j = threadIdx.x; j < 64; j += blockDim.x) s_tab[j] = c_exp.tab[j];
}

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ double exp_scaled(double u, int Mq, uint32_t tab_addr)
{
    const double kd0 = fma(u, c_exp.inv, FHMC_EXP_MAGIC);
    const int k = __double2loint(kd0);
    const double kd = kd0 - FHMC_EXP_MAGIC;
    double r = fma(kd, -c_exp.hi, u);
    r = fma(kd, -c_exp.lo, r);
    double p = fma(r, c_exp.c5, c_exp.c4);
    p = fma(p, r, c_exp.c3);
    p = fma(p, r, c_exp.c2);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double T = lds_f64(tab_addr + ((k & 63) << 3));
    const double v = T * p;
    // v in [1, 4): its biased exponent is 1023 or 1024, so any q >= -1022 keeps the result a normal number.
    // Terms further down are clamped to ~2^-1022 (< 4.5e-308) instead of being flushed to zero: one integer max
    // instead of a compare and two selects; next to a leading term >= 0.5 the difference is < 1e-304 relative.
    const int q = max((k >> 6) - Mq, -1022);
    const int hi = __double2hiint(v) + (q << 20);
    return __hiloint2double(hi, __double2loint(v));
}

// Same function with the seven constants held in (opaque) registers: the compiler otherwise re-materialises them
// from the constant bank through uniform registers on every use (LDCU + IMAD.U32 per constant per block).
struct ExpRegs {
    double inv, nhi, nlo, c5, c4, c3, c2;
};
__device__ __forceinline__ ExpRegs load_exp_regs()
{
    ExpRegs r = {c_exp.inv, -c_exp.hi, -c_exp.lo, c_exp.c5, c_exp.c4, c_exp.c3, c_exp.c2};
    asm volatile("" : "+d"(r.inv), "+d"(r.nhi), "+d"(r.nlo), "+d"(r.c5), "+d"(r.c4), "+d"(r.c3), "+d"(r.c2));
    return r;
}
__device__ __forceinline__ double exp_scaled_r(double u, int Mq, uint32_t tab_addr, const ExpRegs &c)
{
    const double kd0 = fma(u, c.inv, FHMC_EXP_MAGIC);
    const int k = __double2loint(kd0);
    const double kd = kd0 - FHMC_EXP_MAGIC;
    double r = fma(kd, c.nhi, u);
    r = fma(kd, c.nlo, r);
#ifdef FHMC_EXP_ESTRIN
    // Estrin evaluation: 2 more fp64 ops than Horner but a dependency chain that is 2 levels shorter
    const double r2 = r * r;
    const double pa = fma(c.c3, r, c.c2);
    const double pb = fma(c.c5, r, c.c4);
    const double pq = r + 1.0;
    const double r4 = r2 * r2;
    double p = fma(r2, pa, pq);
    p = fma(r4, pb, p);
#else
    double p = fma(r, c.c5, c.c4);
    p = fma(p, r, c.c3);
    p = fma(p, r, c.c2);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#endif
    const double T = lds_f64(tab_addr + ((k & 63) << 3));
    const double v = T * p;
    const int q = max((k >> 6) - Mq, -1022);
    const int hi = __double2hiint(v) + (q << 20);
    return __hiloint2double(hi, __double2loint(v));
}

// exp(t) for t <= ~0 (generic paths): same algorithm, no exponent offset
__device__ __forceinline__ double exp_nonpos(double t, uint32_t tab_addr) { return exp_scaled(t, 0, tab_addr); }

// smallest integer Mq with Mq*ln2 >= m  (so that every term exp(u - Mq*ln2) <= 1 when u <= m)
__device__ __forceinline__ int shift_for_max(double m) { return (int)ceil(m * c_exp.inv_ln2 + 1e-9); }

// Mq*ln2 + x with the product carried in two pieces (Mq*ln2_hi is exact for |Mq| < 2^20)
__device__ __forceinline__ double add_shift(int Mq, double x)
{
    const double q = (double)Mq;
    return fma(q, c_exp.ln2_hi, x) + q * c_exp.ln2_lo;
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
