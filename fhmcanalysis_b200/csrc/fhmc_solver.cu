// fhmc_solver.cu -- K4: batched find_phase_eq.
//
// This file: the C entry point and the general kernel (PointEval groups of 1/4/32 lanes per solve), which takes whatever
// the default warp-per-solve kernel on the lean evaluator (fhmc_solver_lean.cu) has no instantiation for.
// Reference: histogram.find_phase_eq (gc_hist.pyx:598-668) minimises, per temperature and one at a
// time, phase_eq_error(mu) = min over pairs of phases at least 2*smooth bins wide of
// (F.E._i - F.E._j)^2 (gc_hist.pyx:2570-2630) with scipy's Nelder-Mead.  Here every solve is an
// independent group of G lanes that finds the ROOT of the signed difference d(mu) = F.E._i - F.E._j of
// the pair that objective selects.  d is smooth and monotone while the pair exists:
//     d'(mu) = beta * (<N>_j - <N>_i)          (averages of the N row, free from K2)
// so a bracketed Newton iteration converges in a handful of evaluations; each evaluation is the
// full fused state-point pass of fhmc_point.cuh (reweight + Taylor + phase split + thermo).
#include <stdlib.h>

#include "fhmc_solver.cuh"

namespace fhmc {

template <int G, bool TAYLOR, int CTA>
__global__ void __launch_bounds__(CTA) k_find_phase_eq(const __grid_constant__ SolveArgs sa)
{
    const SweepArgs &a = sa.sw;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_tab;
    const double *sm = stage_histogram(a, smem_raw, s_tab);

    constexpr int GPC = CTA / G;
    const int grp = threadIdx.x / G;
    const long long T = a.st.n_states;
    const long long ntiles = (T + GPC - 1) / GPC;
    PointEval<G, TAYLOR> pe(a, sm, threadIdx.x & 31, s_tab);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long rec = tile * GPC + grp;
        if (rec >= T) continue;
        double mu, beta, dmu;
        const fhmc_states &st = a.st;
        mu = st.mu1[(rec / st.mu1_div) % st.n_mu1];
        beta = st.beta ? st.beta[(rec / st.beta_div) % st.n_beta] : a.d.beta_ref;
        dmu = st.dmu ? st.dmu[(rec / st.dmu_div) % st.n_dmu] : a.d.dmu_ref;

        const double n_mid = 0.5 * (sm[a.d.n_pad] + sm[a.d.n_pad + a.d.n - 1]);
        solve_one(sa, rec, mu, beta, n_mid, pe.g == 0, [&](double m, int &P_now, EvalView &v) {
            pe.setup(m, beta, dmu);
            const unsigned st_ = pe.run(rec);
            if (G > 1) __syncwarp(pe.member);
            P_now = pe.P;
            v.fe = a.out.fe + rec * a.d.pmax;
            v.bl = a.out.bounds + rec * a.d.pmax * 2;
            v.av = a.out.avg + rec * a.d.pmax * a.d.n_sel;
            return st_;
        }, []() {});
    }
}

// warp-per-solve kernels on the lean evaluator (fhmc_solver_lean.cu): the default
int launch_solver_lean(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream);

struct DevCaps {
    int sm_count, smem_optin;
};
static int get_caps(DevCaps &c)
{
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 1;
    if (check_cuda(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev), "device attribute")) return 1;
    if (check_cuda(cudaDeviceGetAttribute(&c.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), "device attribute")) return 1;
    return 0;
}

template <int G, bool TAYLOR, int CTA>
static int launch_solver_cta(const SolveArgs &sa, size_t smem, const DevCaps &caps, cudaStream_t stream, int *occ_out, bool dry)
{
    auto kern = k_find_phase_eq<G, TAYLOR, CTA>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA, smem), "occupancy query")) return 1;
    *occ_out = occ;
    if (dry) return 0;
    if (occ < 1) { set_error("solver kernel does not fit on an SM"); return 1; }
    const long long gpc = CTA / G;
    const long long ntiles = (sa.sw.st.n_states + gpc - 1) / gpc;
    long long grid = (long long)caps.sm_count * occ;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, CTA, smem, stream>>>(sa);
    return check_cuda(cudaGetLastError(), "k_find_phase_eq launch");
}

template <int G, bool TAYLOR>
static int launch_solver(const SolveArgs &sa, size_t smem, const DevCaps &caps, cudaStream_t stream)
{
    // shared memory leaves room for one 256-thread CTA per SM (config 4: 2001 bins x 10 rows): run 512 threads
    int occ = 0;
    if (launch_solver_cta<G, TAYLOR, FHMC_CTA>(sa, smem, caps, stream, &occ, true)) return 1;
    const long long tiles512 = (sa.sw.st.n_states + 512 / G - 1) / (512 / G);
    if (occ == 1 && tiles512 >= caps.sm_count) {
        int occ2 = 0;
        if (launch_solver_cta<G, TAYLOR, 512>(sa, smem, caps, stream, &occ2, true)) return 1;
        if (occ2 >= 1) return launch_solver_cta<G, TAYLOR, 512>(sa, smem, caps, stream, &occ2, false);
    }
    return launch_solver_cta<G, TAYLOR, FHMC_CTA>(sa, smem, caps, stream, &occ, false);
}

}  // namespace fhmc

using namespace fhmc;

static int find_phase_eq_impl(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states, double lnz_tol, double mu_step,
                              int max_iter, int cont_stride, double *mu_coex, double *dfe, int *iters, const fhmc_sweep_out *out,
                              void *stream)
{
    if (!desc || !blob || !states || !out || !mu_coex || !dfe || !iters) { set_error("null pointer"); return 1; }
    if (desc->n_sel < 1) { set_error("solver needs quantity 0 to be N_tot (its phase averages give the Newton slope)"); return 1; }
    if (desc->complete || desc->smooth < 1 || desc->pmax < 2) { set_error("solver needs complete=0, smooth>=1, pmax>=2"); return 1; }
    if (!out->status || !out->nphase || !out->nmin || !out->lnnorm || !out->fe || !out->avg || !out->bounds ||
        !out->max_idx || !out->min_idx) { set_error("missing output buffer"); return 1; }
    if (!(lnz_tol >= 0.0) || !(mu_step > 0.0) || max_iter < 1) { set_error("bad solver controls"); return 1; }
    if (states->n_states == 0) return 0;
    if (((uintptr_t)blob & 15) || (desc->n_pad & 1)) { set_error("blob must be 16-byte aligned with even n_pad"); return 1; }
    DevCaps caps;
    if (get_caps(caps)) return 1;
    size_t smem = (size_t)desc->n_rows * desc->n_pad * 8 + 16 + 512;
    SolveArgs sa;
    sa.sw.blob_global = 0;
    if (smem > (size_t)caps.smem_optin) {
        sa.sw.blob_global = 1;
        smem = 16 + 512;
    }
    sa.sw.d = *desc;
    sa.sw.blob = blob;
    sa.sw.st = *states;
    sa.sw.out = *out;
    sa.lnz_tol = lnz_tol;
    sa.mu_step = mu_step;
    sa.max_iter = max_iter;
    sa.mu_coex = mu_coex;
    sa.dfe = dfe;
    sa.iters = iters;
    sa.cont_stride = cont_stride > 1 ? cont_stride : 0;
    if (sa.cont_stride && check_cuda(cudaMemsetAsync(iters, 0, sizeof(int) * (size_t)states->n_states, (cudaStream_t)stream), "cudaMemsetAsync")) return 1;
    const bool taylor = desc->n_coef > 0 || desc->n_term > 1;
    const long long T = states->n_states;
    cudaStream_t s = (cudaStream_t)stream;
    // a solve is ~10 dependent state-point passes: prefer wide groups unless there are very many solves
    const char *force = getenv("FHMC_SOLVER_LANES");  // tuning/test override: 1/4/32 = the PointEval group kernel with that many lanes per solve
    const int forced = force ? atoi(force) : 0;
    if (forced == 0) {   // default: one warp per solve on the lean evaluator (fhmc_solver_lean.cu)
        const int rc = launch_solver_lean(sa, caps.sm_count, caps.smem_optin, s);
        if (rc >= 0) return rc;
    }
    sa.cont_stride = 0;   // the group kernels solve every record from its own guess
    note_kernel("k_find_phase_eq");
    if (forced == 1) return taylor ? launch_solver<1, true>(sa, smem, caps, s) : launch_solver<1, false>(sa, smem, caps, s);
    if (forced == 4) return taylor ? launch_solver<4, true>(sa, smem, caps, s) : launch_solver<4, false>(sa, smem, caps, s);
    if (forced == 32 || T * 32 <= (long long)caps.sm_count * 2048 * 4)
        return taylor ? launch_solver<32, true>(sa, smem, caps, s) : launch_solver<32, false>(sa, smem, caps, s);
    if (T * 4 <= (long long)caps.sm_count * 2048 * 4)
        return taylor ? launch_solver<4, true>(sa, smem, caps, s) : launch_solver<4, false>(sa, smem, caps, s);
    return taylor ? launch_solver<1, true>(sa, smem, caps, s) : launch_solver<1, false>(sa, smem, caps, s);
}

extern "C" int fhmc_find_phase_eq_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                                     double lnz_tol, double mu_step, int max_iter, double *mu_coex, double *dfe,
                                     int *iters, const fhmc_sweep_out *out, void *stream)
{
    return find_phase_eq_impl(desc, blob, states, lnz_tol, mu_step, max_iter, 0, mu_coex, dfe, iters, out, stream);
}

extern "C" int fhmc_find_phase_eq_curve(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                                        double lnz_tol, double mu_step, int max_iter, int seed_stride, double *mu_coex, double *dfe,
                                        int *iters, const fhmc_sweep_out *out, void *stream)
{
    if (seed_stride < 2) { set_error("seed_stride must be >= 2"); return 1; }
    return find_phase_eq_impl(desc, blob, states, lnz_tol, mu_step, max_iter, seed_stride, mu_coex, dfe, iters, out, stream);
}
