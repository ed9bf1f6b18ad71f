// fhmc_solver_lean.cuh -- kernel + launcher templates of the warp-per-solve coexistence solver (instantiated in
// fhmc_solver_lean.cu and fhmc_solver_lean2.cu: two translation units so that the instantiations compile in parallel).
#pragma once
#include <stdlib.h>

#include "fhmc_lean.cuh"
#include "fhmc_solver.cuh"

namespace fhmc {

static __device__ unsigned g_solve_counter[64];   // (per translation unit; slots are handed out round-robin per launch)

template <bool TAYLOR>
__device__ __noinline__ unsigned lean_fallback(const SweepArgs &a, const double *sm, const double *s_tab, int lane, double mu1,
                                               double beta, double dmu, long long rec, int &P_now)
{
    PointEval<32, TAYLOR> pe(a, sm, lane, s_tab);
    pe.setup(mu1, beta, dmu);
    const unsigned st = pe.run(rec);
    __syncwarp();
    P_now = pe.P;
    return st;
}

template <int NC, int NSEL, int NT, int CTA>
__global__ void __launch_bounds__(CTA, 1) k_solve_lean(const __grid_constant__ SolveArgs sa, unsigned *counter)
{
    constexpr bool TAYLOR = (NC > 0) || (NT > 1);
    const SweepArgs &a = sa.sw;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_tab;
    const double *sm = stage_histogram(a, smem_raw, s_tab);   // [blob | mbarrier | exp table | per-warp scratch]
    LeanScratch *ws = reinterpret_cast<LeanScratch *>(reinterpret_cast<unsigned char *>(s_tab) + 512) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    LeanEval<NC, NSEL, NT> le(a, sm, ws, lane, s_tab);
    const long long T = a.st.n_states;
    const double n_mid = 0.5 * (sm[a.d.n_pad] + sm[a.d.n_pad + a.d.n - 1]);
    for (;;) {
        unsigned nxt = 0;
        if (lane == 0) nxt = atomicAdd(counter, 1u);
        const long long q = (long long)__shfl_sync(0xffffffffu, nxt, 0);
        if (q >= T) break;
        const fhmc_states &st = a.st;
        long long rec = q;
        long long dep_l = -1, dep_r = -1;
        if (sa.cont_stride > 1) {
            // queue order: the seeds (every cont_stride-th solve and the last one) first, then the rest in list order, so the
            // seeds an item waits for were always handed out before it
            const long long S = sa.cont_stride, n_seed = (T - 1) / S + 1, extra = ((T - 1) % S) ? 1 : 0;
            if (q < n_seed) rec = q * S;
            else if (q < n_seed + extra) rec = T - 1;
            else {
                const long long j = q - n_seed - extra;          // j-th solve that is not a multiple of S ...
                rec = j + j / (S - 1) + 1;
                // (T-1, when it is the extra seed, is the largest such solve and lies beyond the end of this part of the queue)
                dep_l = (rec / S) * S;
                dep_r = dep_l + S < T ? dep_l + S : T - 1;
            }
        }
        double mu = st.mu1[(rec / st.mu1_div) % st.n_mu1];
        const double beta = st.beta ? st.beta[(rec / st.beta_div) % st.n_beta] : a.d.beta_ref;
        const double dmu = st.dmu ? st.dmu[(rec / st.dmu_div) % st.n_dmu] : a.d.dmu_ref;
        if (dep_l >= 0) {
            if (lane == 0) {
                while (*reinterpret_cast<volatile int *>(sa.iters + dep_l) == 0) __nanosleep(100);
                while (*reinterpret_cast<volatile int *>(sa.iters + dep_r) == 0) __nanosleep(100);
                __threadfence();
            }
            __syncwarp();
            const unsigned sl = __ldcg(a.out.status + dep_l), sr = __ldcg(a.out.status + dep_r);
            const bool gl = (sl & (FHMC_ST_CODE_MASK | FHMC_ST_JUMP)) == 0, gr = (sr & (FHMC_ST_CODE_MASK | FHMC_ST_JUMP)) == 0;
            const double ml = __ldcg(sa.mu_coex + dep_l), mr = __ldcg(sa.mu_coex + dep_r);
            const double bl = st.beta ? st.beta[(dep_l / st.beta_div) % st.n_beta] : a.d.beta_ref;
            const double br = st.beta ? st.beta[(dep_r / st.beta_div) % st.n_beta] : a.d.beta_ref;
            if (gl && gr && br != bl) mu = ml + (mr - ml) * ((beta - bl) / (br - bl));
            else if (gl) mu = ml;
            else if (gr) mu = mr;
        }
        bool in_scratch = false;
#ifdef FHMC_LEAN_PROFILE
        const long long t_solve0 = clock64();
#endif
#ifdef FHMC_LEAN_PROFILE
        long long t_last = t_solve0;
#endif
        solve_one(sa, rec, mu, beta, n_mid, lane == 0, [&](double mm, int &P_now, EvalView &v) {
#ifdef FHMC_LEAN_PROFILE
            {   // slot 8: cycles between the end of one evaluation and the start of the next (solve_one's own logic)
                const long long t_now = clock64();
                if (lane == 0) atomicAdd(&g_lean_prof[8], (unsigned long long)(t_now - t_last));
            }
#endif
            le.setup(mm, beta, dmu);
            unsigned st_ = 0;
#ifdef FHMC_LEAN_PROFILE
            const long long t_run0 = clock64();
            const bool lean_ok = le.run(st_);
            t_last = clock64();
            if (lane == 0) atomicAdd(&g_lean_prof[9], (unsigned long long)(t_last - t_run0));   // slot 9: all of run()
            if (lean_ok) {
#else
            if (le.run(st_)) {
#endif
                in_scratch = true;
                P_now = le.P;
                v.fe = ws->fe;
                v.bl = ws->bl;
                v.av = ws->avg;
                return st_;
            }
            in_scratch = false;
            v.fe = a.out.fe + rec * a.d.pmax;
            v.bl = a.out.bounds + rec * a.d.pmax * 2;
            v.av = a.out.avg + rec * a.d.pmax * a.d.n_sel;
            return lean_fallback<TAYLOR>(a, sm, s_tab, lane, mm, beta, dmu, rec, P_now);
        }, [&]() { if (in_scratch) le.commit(rec); });
#ifdef FHMC_LEAN_PROFILE
        if (lane == 0) atomicAdd(&g_lean_stats[7], (unsigned long long)(clock64() - t_solve0));   // (slot 7 doubles as total solve cycles)
#endif
        __syncwarp();
    }
}

template <int NC, int NSEL, int NT, int CTA>
static int try_launch(const SolveArgs &sa, size_t blob_smem, int sm_count, int smem_optin, unsigned *counter, cudaStream_t stream,
                      bool &launched)
{
    launched = false;
    const size_t smem = blob_smem + (CTA / 32) * sizeof(LeanScratch);
    if (smem > (size_t)smem_optin) return 0;
    auto kern = k_solve_lean<NC, NSEL, NT, CTA>;
    // attribute + occupancy are queried once per (kernel, shared-memory size, device): a solver call is latency sensitive
    static thread_local size_t cached_smem[16];
    static thread_local int cached_occ[16];
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    int occ = 0;
    if (dev >= 0 && dev < 16 && cached_smem[dev] == smem && cached_occ[dev] > 0) {
        occ = cached_occ[dev];
    } else {
        if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
        if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA, smem), "occupancy query")) return 1;
        if (dev >= 0 && dev < 16) { cached_smem[dev] = smem; cached_occ[dev] = occ; }
    }
    if (occ < 1) return 0;
    const long long T = sa.sw.st.n_states, per_cta = CTA / 32;
    long long grid = (long long)sm_count * occ;
    const long long need = (T + per_cta - 1) / per_cta;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, CTA, smem, stream>>>(sa, counter);
    launched = true;
    note_kernel("k_solve_lean");
    return check_cuda(cudaGetLastError(), "k_solve_lean launch");
}

template <int NC, int NSEL, int NT>
static int launch_lean(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream)
{
    static unsigned next_slot = 0;
    static thread_local unsigned *bases[16];
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    unsigned *base = (dev >= 0 && dev < 16) ? bases[dev] : nullptr;
    if (!base) {
        if (check_cuda(cudaGetSymbolAddress((void **)&base, g_solve_counter), "cudaGetSymbolAddress")) return 1;
        if (dev >= 0 && dev < 16) bases[dev] = base;
    }
    unsigned *counter = base + (next_slot++ & 63u);
    if (check_cuda(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream), "cudaMemsetAsync")) return 1;
    const size_t blob_smem = (size_t)sa.sw.d.n_rows * sa.sw.d.n_pad * 8 + 16 + 512;
    // (measured on B200, config 4, 10^4 solves: 1.47 / 1.56 / 1.68 ms with 512 / 768 / 1024 threads per CTA, i.e. 128 / 80 / 64
    // registers per thread -- the extra warps do not pay for the spills)
    bool done = false;
    int rc = try_launch<NC, NSEL, NT, 512>(sa, blob_smem, sm_count, smem_optin, counter, stream, done);
    if (rc || done) return rc;
    return -1;
}


// this translation unit's copy of the diagnostic counters: out[0..7] += g_lean_stats (and out[8..23] += cycle counters when
// built with -DFHMC_LEAN_PROFILE); optionally cleared
static int lean_stats_tu(unsigned long long *out, int reset)
{
    unsigned long long buf[24] = {0};
    if (check_cuda(cudaMemcpyFromSymbol(buf, g_lean_stats, 8 * sizeof(unsigned long long)), "cudaMemcpyFromSymbol")) return 1;
#ifdef FHMC_LEAN_PROFILE
    if (check_cuda(cudaMemcpyFromSymbol(buf + 8, g_lean_prof, 16 * sizeof(unsigned long long)), "cudaMemcpyFromSymbol")) return 1;
#endif
    if (out)
        for (int k = 0; k < 24; ++k) out[k] += buf[k];
    if (reset) {
        unsigned long long z[16] = {0};
        if (check_cuda(cudaMemcpyToSymbol(g_lean_stats, z, 8 * sizeof(unsigned long long)), "cudaMemcpyToSymbol")) return 1;
#ifdef FHMC_LEAN_PROFILE
        if (check_cuda(cudaMemcpyToSymbol(g_lean_prof, z, sizeof(z)), "cudaMemcpyToSymbol")) return 1;
#endif
    }
    return 0;
}

}  // namespace fhmc
