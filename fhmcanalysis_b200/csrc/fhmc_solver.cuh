// fhmc_solver.cuh -- the coexistence iteration shared by the generic (group-per-solve) and the thread-per-solve kernels.
#pragma once
#include "fhmc_point.cuh"

namespace fhmc {

struct SolveArgs {
    SweepArgs sw;
    double lnz_tol, mu_step;
    int max_iter;
    double *mu_coex, *dfe;
    int *iters;
};


// One coexistence solve for record `rec`, executed by every lane of the group that owns it (uniform control flow).
// eval(mu) runs one full state-point evaluation into record `rec` and returns (status word, number of phases).
template <class Eval>
__device__ __forceinline__ void solve_one(const SolveArgs &sa, long long rec, double mu, double beta, double n_mid, bool leader,
                                          Eval &&eval)
{
    const SweepArgs &a = sa.sw;
    const int pmax = a.d.pmax, nsel = a.d.n_sel;
    const int min_width = a.d.min_width > 0 ? a.d.min_width : 2 * a.d.smooth;  // ntot/gc_hist.pyx:652, n1/gc_hist.pyx:1479
    bool have_lo = false, have_hi = false, converged = false, located = false, have_glo = false, have_ghi = false;
    double lo = 0.0, hi = 0.0, mu_good = mu, d = 0.0, glo = 0.0, ghi = 0.0, step = sa.mu_step, d_lo = 0.0, d_hi = 0.0;
    unsigned status = 0;
    int code = FHMC_E_NO_COEX, it = 0, nevals = 0;
    for (it = 0; it < sa.max_iter; ++it) {
        int P_now = 0;
        status = eval(mu, P_now);
        ++nevals;
        // pair selection of gc_hist.pyx:2614-2630 (every lane, uniform)
        bool ok = false;
        double slope = 0.0;
        if ((status & FHMC_ST_CODE_MASK) == FHMC_OK) {
            const double *fe = a.out.fe + rec * pmax;
            const int *bl = a.out.bounds + rec * pmax * 2;
            double best = 1.7976931348623157e308;
            int bi = -1, bj = -1;
            for (int i = 0; i < P_now; ++i) {
                if (bl[2 * i + 1] - bl[2 * i] < min_width) continue;
                for (int j = i + 1; j < P_now; ++j) {
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
            for (int p = 1; p < P_now; ++p) fmin_ = fmin(fmin_, fe[p]);
            double wsum = 0.0, nsum = 0.0;
            for (int p = 0; p < P_now; ++p) {
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
        if (below) { lo = mu; d_lo = d; have_lo = true; } else { hi = mu; d_hi = d; have_hi = true; }
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
            // d(mu) jumps where the integer phase boundaries move (noisy ln(PI)): when the slope says d can change by
            // less than a thousandth of what is left at BOTH ends of the bracket, no root lies inside and bisecting on
            // to the last bit (~45 evaluations) locates nothing.  Stop here (the reference's simplex stops at 1e-4).
            if (fabs(slope) * (h - l) < 1e-3 * fmin(fabs(d_lo), fabs(d_hi)) && h - l <= 1e-6 * fmax(1.0, fmax(fabs(l), fabs(h)))) {
                converged = true;
                code = FHMC_OK;
                break;
            }
        }
        mu = mu_n;
    }
    if (!converged && code == FHMC_E_NO_COEX && it >= sa.max_iter) code = FHMC_E_NO_COEX + 1;  // iteration cap
    if (leader) {
        sa.mu_coex[rec] = mu_good;
        sa.dfe[rec] = d;
        sa.iters[rec] = nevals;
        if (code != FHMC_OK) a.out.status[rec] = (a.out.status[rec] & ~FHMC_ST_CODE_MASK) | (unsigned)code;
    }
}

}  // namespace fhmc
