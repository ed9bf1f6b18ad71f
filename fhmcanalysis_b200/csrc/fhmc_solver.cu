// fhmc_solver.cu -- K4: batched find_phase_eq.
//
// Reference: histogram.find_phase_eq (gc_hist.pyx:598-668) minimises, per temperature and one at a
// time, phase_eq_error(mu) = min over pairs of phases at least 2*smooth bins wide of
// (F.E._i - F.E._j)^2 (gc_hist.pyx:2570-2630) with scipy's Nelder-Mead.  Here every solve is an
// independent group of G lanes that finds the ROOT of the signed difference d(mu) = F.E._i - F.E._j of
// the pair that objective selects.  d is smooth and monotone while the pair exists:
//     d'(mu) = beta * (<N>_j - <N>_i)          (averages of the N row, free from K2)
// so a bracketed Newton iteration converges in a handful of evaluations; each evaluation is the
// full fused state-point pass of fhmc_point.cuh (reweight + Taylor + phase split + thermo).
#include <stdlib.h>

#include "fhmc_point.cuh"

namespace fhmc {

struct SolveArgs {
    SweepArgs sw;
    double lnz_tol, mu_step;
    int max_iter;
    double *mu_coex, *dfe;
    int *iters;
};

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
    const int pmax = a.d.pmax, nsel = a.d.n_sel;
    const int min_width = a.d.min_width > 0 ? a.d.min_width : 2 * a.d.smooth;  // ntot/gc_hist.pyx:652, n1/gc_hist.pyx:1479
    PointEval<G, TAYLOR> pe(a, sm, threadIdx.x & 31, s_tab);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long rec = tile * GPC + grp;
        if (rec >= T) continue;
        double mu, beta, dmu;
        const fhmc_states &st = a.st;
        mu = st.mu1[(rec / st.mu1_div) % st.n_mu1];
        beta = st.beta ? st.beta[(rec / st.beta_div) % st.n_beta] : a.d.beta_ref;
        dmu = st.dmu ? st.dmu[(rec / st.dmu_div) % st.n_dmu] : a.d.dmu_ref;

        bool have_lo = false, have_hi = false, converged = false, located = false, have_glo = false, have_ghi = false;
        double lo = 0.0, hi = 0.0, mu_good = mu, d = 0.0, glo = 0.0, ghi = 0.0, step = sa.mu_step;
        const double n_mid = 0.5 * (sm[a.d.n_pad] + sm[a.d.n_pad + a.d.n - 1]);
        unsigned status = 0;
        int code = FHMC_E_NO_COEX, it = 0, nevals = 0;
        for (it = 0; it < sa.max_iter; ++it) {
            pe.setup(mu, beta, dmu);
            status = pe.run(rec);
            ++nevals;
            if (G > 1) __syncwarp(pe.member);
            // pair selection of gc_hist.pyx:2614-2630 (every lane, uniform)
            bool ok = false;
            double slope = 0.0;
            if ((status & FHMC_ST_CODE_MASK) == FHMC_OK) {
                const double *fe = a.out.fe + rec * pmax;
                const int *bl = a.out.bounds + rec * pmax * 2;
                double best = 1.7976931348623157e308;
                int bi = -1, bj = -1;
                for (int i = 0; i < pe.P; ++i) {
                    if (bl[2 * i + 1] - bl[2 * i] < min_width) continue;
                    for (int j = i + 1; j < pe.P; ++j) {
                        if (bl[2 * j + 1] - bl[2 * j] < min_width) continue;
                        const double dd = fe[i] - fe[j];
                        if (dd * dd < best) { best = dd * dd; bi = i; bj = j; d = dd; }
                    }
                }
                if (bi >= 0) {
                    ok = true;
                    const double *av = a.out.avg + rec * pmax * nsel;
                    slope = beta * (av[bj * nsel] - av[bi * nsel]);
                }
            }
            if (!ok) {
                if ((status & FHMC_ST_CODE_MASK) != FHMC_OK) { code = (int)(status & FHMC_ST_CODE_MASK); break; }
                if (located) {
                    mu = 0.5 * (mu + mu_good);  // stepped out of the two-phase window: come back half way
                    continue;
                }
                // ---- locate the two-phase window: <N>_total(mu) is monotone, the window is where it crosses the
                // middle of the N range.  Expand geometrically from the guess until bracketed, then bisect.
                const double *fe = a.out.fe + rec * pmax;
                const double *av = a.out.avg + rec * pmax * nsel;
                double fmin_ = fe[0];
                for (int p = 1; p < pe.P; ++p) fmin_ = fmin(fmin_, fe[p]);
                double wsum = 0.0, nsum = 0.0;
                for (int p = 0; p < pe.P; ++p) {
                    const double wgt = exp(-(fe[p] - fmin_));
                    wsum += wgt;
                    nsum += wgt * av[p * nsel];
                }
                const double g = nsum / wsum - n_mid;
                if (g < 0.0) { glo = mu; have_glo = true; } else { ghi = mu; have_ghi = true; }
                if (have_glo && have_ghi) {
                    if (fabs(ghi - glo) <= 1e-13 * fmax(1.0, fmax(fabs(glo), fabs(ghi)))) { code = FHMC_E_NO_COEX; break; }
                    mu = 0.5 * (glo + ghi);
                } else {
                    mu += (g < 0.0) ? step : -step;
                    step *= 2.0;
                }
                continue;
            }
            located = true;
            mu_good = mu;
            if (fabs(d) <= sa.lnz_tol) { converged = true; code = FHMC_OK; break; }
            const bool below = (slope >= 0.0) ? (d < 0.0) : (d > 0.0);
            if (below) { lo = mu; have_lo = true; } else { hi = mu; have_hi = true; }
            double mu_n = mu;
            if (slope != 0.0) {
                double dm = -d / slope;
                const double cap = 16.0 * sa.mu_step;  // Newton steps are trusted further than the blind search step
                if (dm > cap) dm = cap;
                if (dm < -cap) dm = -cap;
                mu_n = mu + dm;
            } else {
                mu_n = mu + (below ? sa.mu_step : -sa.mu_step);
            }
            if (have_lo && have_hi) {
                const double l = fmin(lo, hi), h = fmax(lo, hi);
                if (!(mu_n > l && mu_n < h)) mu_n = 0.5 * (lo + hi);
                if (mu_n == mu || h - l <= 4.0 * 2.220446049250313e-16 * fmax(fabs(l), fabs(h))) {
                    converged = true;  // bracket exhausted at fp64 resolution
                    code = FHMC_OK;
                    break;
                }
            }
            mu = mu_n;
        }
        if (!converged && code == FHMC_E_NO_COEX && it >= sa.max_iter) code = FHMC_E_NO_COEX + 1;  // iteration cap
        if (pe.g == 0) {
            sa.mu_coex[rec] = mu_good;
            sa.dfe[rec] = d;
            sa.iters[rec] = nevals;
            if (code != FHMC_OK) a.out.status[rec] = (a.out.status[rec] & ~FHMC_ST_CODE_MASK) | (unsigned)code;
        }
    }
}

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

extern "C" int fhmc_find_phase_eq_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                                     double lnz_tol, double mu_step, int max_iter, double *mu_coex, double *dfe,
                                     int *iters, const fhmc_sweep_out *out, void *stream)
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
    const bool taylor = desc->n_coef > 0 || desc->n_term > 1;
    const long long T = states->n_states;
    cudaStream_t s = (cudaStream_t)stream;
    // a solve is ~10 dependent state-point passes: prefer wide groups unless there are very many solves
    const char *force = getenv("FHMC_SOLVER_LANES");  // tuning/debug override
    const int forced = force ? atoi(force) : 0;
    if (forced == 1) return taylor ? launch_solver<1, true>(sa, smem, caps, s) : launch_solver<1, false>(sa, smem, caps, s);
    if (forced == 4) return taylor ? launch_solver<4, true>(sa, smem, caps, s) : launch_solver<4, false>(sa, smem, caps, s);
    if (forced == 32 || T * 32 <= (long long)caps.sm_count * 2048 * 4)
        return taylor ? launch_solver<32, true>(sa, smem, caps, s) : launch_solver<32, false>(sa, smem, caps, s);
    if (T * 4 <= (long long)caps.sm_count * 2048 * 4)
        return taylor ? launch_solver<4, true>(sa, smem, caps, s) : launch_solver<4, false>(sa, smem, caps, s);
    return taylor ? launch_solver<1, true>(sa, smem, caps, s) : launch_solver<1, false>(sa, smem, caps, s);
}
