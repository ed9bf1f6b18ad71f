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
    // Continuation along a coexistence CURVE (fhmc_find_phase_eq_curve; 0 = off): the solves form a list ordered in beta;
    // every cont_stride-th one (and the last) is a seed solved from its own guess, every other solve starts from the linear
    // interpolation in beta of the roots of the two seeds around it.  iters[] doubles as the "finished" flag (zeroed
    // before the launch, written last).
    int cont_stride;
};


// where the record of the evaluation just made can be read (the caller's arrays, or an evaluator's scratch)
struct EvalView {
    const double *fe;   // [P]
    const int *bl;      // [P][2]
    const double *av;   // [P][n_sel]
};

// One coexistence solve for record `rec`, executed by every lane of the group that owns it (uniform control flow).
// eval(mu, P, view) runs one full state-point evaluation and returns the status word, the number of phases and where its
// F.E. / bounds / averages are; commit() makes sure the last evaluation's record is in record `rec` of the caller's arrays.
template <class Eval, class Commit>
__device__ __forceinline__ void solve_one(const SolveArgs &sa, long long rec, double mu, double beta, double n_mid, bool leader,
                                          Eval &&eval, Commit &&commit)
{
    const SweepArgs &a = sa.sw;
    const int nsel = a.d.n_sel;
    const int min_width = a.d.min_width > 0 ? a.d.min_width : 2 * a.d.smooth;  // ntot/gc_hist.pyx:652, n1/gc_hist.pyx:1479
    bool have_lo = false, have_hi = false, converged = false, located = false, have_glo = false, have_ghi = false;
    double lo = 0.0, hi = 0.0, mu_good = mu, d = 0.0, glo = 0.0, ghi = 0.0, step = sa.mu_step, d_lo = 0.0, d_hi = 0.0;
    unsigned status = 0;
    int code = FHMC_E_NO_COEX, it = 0, nevals = 0;
    for (it = 0; it < sa.max_iter; ++it) {
        int P_now = 0;
        EvalView v;
        status = eval(mu, P_now, v);
        ++nevals;
        // pair selection of gc_hist.pyx:2614-2630 (every lane, uniform)
        bool ok = false;
        double slope = 0.0;
        if ((status & FHMC_ST_CODE_MASK) == FHMC_OK) {
            const double *fe = v.fe;
            const int *bl = v.bl;
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
                slope = beta * (v.av[bj * nsel] - v.av[bi * nsel]);
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
            const double *fe = v.fe;
            const double *av = v.av;
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
                converged = true;  // bracket exhausted at fp64 resolution: d jumps across zero here (status bit JUMP below)
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
            // The same conclusion earlier: while the pair of phases persists d is smooth with d' = slope (exact, from the
            // averages), so over a narrow bracket d can change by ~|slope| (h - l).  Residuals of opposite sign that are more
            // than twice that apart cannot be joined by a smooth branch: the sign change is a jump (a phase boundary moved by
            // a bin between l and h).  Newton lands on both sides of such a jump within two steps; without this test the
            // bracket was then bisected to the last bit (15-20 evaluations that located nothing).
            if (h - l <= 1e-6 * fmax(1.0, fmax(fabs(l), fabs(h))) && 2.0 * fabs(slope) * (h - l) < fabs(d_hi - d_lo) &&
                fmin(fabs(d_lo), fabs(d_hi)) > sa.lnz_tol) {
                converged = true;
                code = FHMC_OK;
                break;
            }
        }
        mu = mu_n;
    }
    if (!converged && code == FHMC_E_NO_COEX && it >= sa.max_iter) code = FHMC_E_NO_COEX + 1;  // iteration cap
    commit();
    if (sa.cont_stride > 0) {   // other warps wait for this record: everything above must be visible before the flag below
        __threadfence();
        __syncwarp();
    }
    if (leader) {
        sa.mu_coex[rec] = mu_good;
        sa.dfe[rec] = d;
        if (code != FHMC_OK) a.out.status[rec] = (a.out.status[rec] & ~FHMC_ST_CODE_MASK) | (unsigned)code;
        // the search ended on a JUMP of the free-energy difference (a phase boundary moved by a bin, a phase appeared or
        // vanished): mu_coex is the edge of the jump, |dfe| > lnz_tol.  Callers tell these from converged roots by this bit.
        else if (!(fabs(d) <= sa.lnz_tol)) a.out.status[rec] |= FHMC_ST_JUMP;
        if (sa.cont_stride > 0) __threadfence();
        *reinterpret_cast<volatile int *>(sa.iters + rec) = nevals;   // last: the record's "finished" flag
    }
}

}  // namespace fhmc
