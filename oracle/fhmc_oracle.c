/*
 * fhmc_oracle.c -- plain-C restatement of the reference's 1-D histogram-reweighting hot path.
 *
 * TEST INFRASTRUCTURE ONLY (parity oracle + "port" CPU baseline).  The product
 * (fhmcanalysis_b200/, libfhmc_b200.so) never links, loads or calls this file.
 *
 * Reference: jeetain/FHMCAnalysis, moments/histogram/one_dim/ntot/gc_hist.pyx ("GH").  Every
 * function cites the GH lines it restates.  The arithmetic ORDER of the reference is kept
 * (sequential spec_exp folds, un-fused multiply/add) so that, linked against the same libm,
 * normalised ln(PI) values, extrema indices and is_safe decisions are bit-identical to the
 * compiled reference (checked in tests/test_oracle_vs_reference.py).
 *
 * Parity status: PINNED against (i) the reference's own unit-test known answers
 * (unittests/moments_histogram_one_dim_gc_ntot.py:155-291), (ii) outputs of the compiled
 * reference itself (oracle/_ref) on seeded inputs, committed under tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o liboracle.so fhmc_oracle.c -lm
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* status codes returned by relextrema/thermo ("the reference would raise here") */
enum {
    FO_OK = 0,
    FO_TOO_SHORT = 1,        /* GH:326-327 */
    FO_BAD_FRONT = 2,        /* GH:341-342 */
    FO_BAD_BACK = 3,         /* GH:350-351 */
    FO_COUNT_MISMATCH = 4,   /* GH:403-404 (also NumPy slice-assign size errors GH:408-412) */
    FO_NOT_SORTED = 5,       /* GH:414-415 */
    FO_INDEX_ERROR = 6,      /* IndexError out of ln(PI)_minima_idx in GH:504/511 */
    FO_RAGGED_GAP = 7,       /* GH:355-363 / 370-378 with tied gap extrema: np.array(ragged) */
    FO_CAPACITY = 8          /* caller's index buffers too small (oracle limit, not reference) */
};

/* flag bit OR-ed into *info when the GH:355-363 / 370-378 branch was taken with exactly one
 * position per gap: defined under the NumPy the reference was written for, raises on NumPy>=1.24 */
#define FO_INFO_GAP_FILL 1

static double dmax2(double a, double b) { return a > b ? a : b; }

/* GH:35-53 */
double fo_spec_exp(double a, double b) { return dmax2(a, b) + log(1.0 + exp(-fabs(a - b))); }

/* GH:57-67: lnNorm = fold(spec_exp, lnPI, -DBL_MAX); lnPI -= lnNorm.  Returns lnNorm. */
double fo_normalize(double *lnpi, int n)
{
    double ln_norm = -DBL_MAX;
    for (int i = 0; i < n; ++i) ln_norm = fo_spec_exp(ln_norm, lnpi[i]);
    for (int i = 0; i < n; ++i) lnpi[i] = lnpi[i] - ln_norm;
    return ln_norm;
}

/* GH:71-78: lnPI += ((mu1_new - curr_mu0) * curr_beta) * ntot ; normalize */
double fo_reweight(double *lnpi, const long long *ntot, int n, double mu1_new, double curr_mu0, double curr_beta)
{
    const double s = (mu1_new - curr_mu0) * curr_beta;
    for (int i = 0; i < n; ++i) {
        const double t = s * (double)ntot[i];
        lnpi[i] = lnpi[i] + t;
    }
    return fo_normalize(lnpi, n);
}

/* scipy.signal.argrelextrema(x, cmp, 0, order, 'clip') as called at GH:329-330:
 * i qualifies iff cmp(x[i], x[clip(i+s)]) and cmp(x[i], x[clip(i-s)]) for s = 1..order */
static int raw_extrema(const double *x, int n, int order, int greater, int *out, int cap)
{
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        int ok = 1;
        for (int s = 1; s <= order && ok; ++s) {
            int ip = i + s; if (ip > n - 1) ip = n - 1;
            int im = i - s; if (im < 0) im = 0;
            if (greater) ok = (x[i] > x[ip]) && (x[i] > x[im]);
            else         ok = (x[i] < x[ip]) && (x[i] < x[im]);
        }
        if (ok) { if (cnt >= cap) return -1; out[cnt++] = i; }
    }
    return cnt;
}

static int contains(const int *a, int n, int v) { for (int i = 0; i < n; ++i) if (a[i] == v) return 1; return 0; }
static void prepend(int *a, int *n, int v) { memmove(a + 1, a, sizeof(int) * (size_t)(*n)); a[0] = v; ++*n; }

/* GH:317-415.  maxima/minima must each hold cap >= n+2 ints.  Returns FO_* status. */
int fo_relextrema(const double *x, int n, int smooth, int *maxima, int *n_max, int *minima, int *n_min, int cap, int *info)
{
    const int last = n - 1;
    *info = 0; *n_max = 0; *n_min = 0;
    if (last <= 1) return FO_TOO_SHORT;
    int nM = raw_extrema(x, n, smooth, 1, maxima, cap - 2);
    int nm = raw_extrema(x, n, smooth, 0, minima, cap - 2);
    if (nM < 0 || nm < 0) return FO_CAPACITY;

    if (nM > 0 && nm > 0) {                                   /* GH:333-351 */
        if (!contains(maxima, nM, 0) && !contains(minima, nm, 0)) {
            if (maxima[0] < minima[0]) prepend(minima, &nm, 0);
            else if (maxima[0] > minima[0]) prepend(maxima, &nM, 0);
            else return FO_BAD_FRONT;
        }
        if (!contains(maxima, nM, last) && !contains(minima, nm, last)) {
            if (maxima[nM - 1] < minima[nm - 1]) maxima[nM++] = last;
            else if (maxima[nM - 1] > minima[nm - 1]) minima[nm++] = last;
            else return FO_BAD_BACK;
        }
    } else if (nM > 0 && nm == 0) {                           /* GH:352-366 */
        if (nM > 1) {
            minima[nm++] = 0;
            for (int i = 0; i < nM - 1; ++i) {
                int l = maxima[i], r = maxima[i + 1], pos = l, ties = 0;
                double v = x[l];
                for (int j = l; j < r; ++j) if (x[j] < v) v = x[j];
                for (int j = l; j < r; ++j) if (x[j] == v) { if (!ties) pos = j; ++ties; }
                if (ties != 1) return FO_RAGGED_GAP;
                minima[nm++] = pos;
            }
            minima[nm++] = last;
            *info |= FO_INFO_GAP_FILL;
        } else { minima[0] = 0; minima[1] = last; nm = 2; }
    } else if (nM == 0 && nm > 0) {                           /* GH:367-381 */
        if (nm > 1) {
            maxima[nM++] = 0;
            for (int i = 0; i < nm - 1; ++i) {
                int l = minima[i], r = minima[i + 1], pos = l, ties = 0;
                double v = x[l];
                for (int j = l; j < r; ++j) if (x[j] > v) v = x[j];
                for (int j = l; j < r; ++j) if (x[j] == v) { if (!ties) pos = j; ++ties; }
                if (ties != 1) return FO_RAGGED_GAP;
                maxima[nM++] = pos;
            }
            maxima[nM++] = last;
            *info |= FO_INFO_GAP_FILL;
        } else { maxima[0] = 0; maxima[1] = last; nM = 2; }
    } else {                                                  /* GH:382-386 */
        double vmax = x[0], vmin = x[0];
        for (int j = 1; j < n; ++j) { if (x[j] > vmax) vmax = x[j]; if (x[j] < vmin) vmin = x[j]; }
        for (int j = 0; j < n; ++j) {
            if (x[j] == vmax) { if (nM >= cap) return FO_CAPACITY; maxima[nM++] = j; }
            if (x[j] == vmin) { if (nm >= cap) return FO_CAPACITY; minima[nm++] = j; }
        }
    }
    *n_max = nM; *n_min = nm;

    /* GH:403-415 */
    if (abs(nM - nm) > 1) return FO_COUNT_MISMATCH;
    const int total = nM + nm;
    const int *ev, *od; int nev, nod;
    if (maxima[0] < minima[0]) { ev = maxima; nev = nM; od = minima; nod = nm; }
    else                       { ev = minima; nev = nm; od = maxima; nod = nM; }
    if (nev != (total + 1) / 2 || nod != total / 2) return FO_COUNT_MISMATCH; /* NumPy slice-assign ValueError */
    int prev = -1;
    for (int k = 0; k < total; ++k) {
        int v = (k & 1) ? od[k / 2] : ev[k / 2];
        if (k > 0 && prev > v) return FO_NOT_SORTED;
        prev = v;
    }
    return FO_OK;
}

/* GH:498-520: phase bounds from extrema; bounds[2*p], bounds[2*p+1]. */
int fo_phase_bounds(int n, const int *maxima, int n_max, const int *minima, int n_min, int *bounds)
{
    int min_ctr = 0;
    for (int p = 0; p < n_max; ++p) {
        int left, right;
        if (maxima[p] > 0) { if (min_ctr >= n_min) return FO_INDEX_ERROR; left = minima[min_ctr]; ++min_ctr; }
        else left = 0;
        if (maxima[p] < n - 1) { if (min_ctr >= n_min) return FO_INDEX_ERROR; right = minima[min_ctr]; }
        else right = n;
        if (right == n - 1) right += 1;
        bounds[2 * p] = left; bounds[2 * p + 1] = right;
    }
    return FO_OK;
}

/* GH:523-526: F.E./kT = -fold(spec_exp, lnPI[j]-lnPI[0], j in [left,right)) */
double fo_free_energy(const double *lnpi, int left, int right)
{
    double ln_x = -DBL_MAX;
    for (int j = left; j < right; ++j) ln_x = fo_spec_exp(ln_x, lnpi[j] - lnpi[0]);
    return -ln_x;
}

/* GH:530-541: avg[a] = sum(prob * mom[a, left:right]) / sum(prob), prob = exp(lnPI[left:right]).
 * mom is [n_arrays][n] row-major.  Returns sum_prob. */
double fo_phase_averages(const double *lnpi, int n, int left, int right, const double *mom, int n_arrays, double *avg)
{
    double sum_prob = 0.0;
    for (int j = left; j < right; ++j) sum_prob += exp(lnpi[j]);
    for (int a = 0; a < n_arrays; ++a) {
        const double *x = mom + (size_t)a * (size_t)n;
        double acc = 0.0;
        for (int j = left; j < right; ++j) acc += exp(lnpi[j]) * x[j];
        avg[a] = acc / sum_prob;
    }
    return sum_prob;
}

/* GH:556-596 (with fresh extrema): returns 1 safe / 0 not safe */
int fo_is_safe(const double *lnpi, int n, const int *maxima, int n_max, double cutoff, int complete)
{
    if (!complete) {
        if (n_max <= 0) return 0;
        return (lnpi[maxima[n_max - 1]] - lnpi[n - 1] < cutoff) ? 0 : 1;
    }
    double vmax = lnpi[0];
    for (int j = 1; j < n; ++j) if (lnpi[j] > vmax) vmax = lnpi[j];
    return (vmax - lnpi[n - 1] < cutoff) ? 0 : 1;
}

/*
 * One full "state point" the way the reference's users drive it (notebook loop
 * example/ntot/square_well/example.ipynb cell 9): fresh copy -> reweight(mu1) -> thermo() ->
 * is_safe().  coef == NULL: plain reweight.  Otherwise lnPI += sum_c xi[c]*coef[c][:] is applied
 * between the reweight and the thermo (the Taylor terms of GH:1023-1025 / 1157-1160 in the
 * closed coefficient form of SURVEY 8(a) row 9) followed by the renormalisation of GH:885/964.
 *
 * sel: n_sel rows of [n] arrays to average per phase (e.g. N, N^2, U).
 * Outputs sized for pmax phases: fe[pmax], avg[pmax][n_sel], bounds[pmax][2], max_idx[pmax], min_idx[pmax+1].
 * Returns FO_* status; *nphase = number of phases (may exceed pmax -> FO_CAPACITY).
 */
int fo_state_point(const double *lnpi_ref, const long long *ntot, int n, double beta_ref, double mu1_ref,
                   double mu1, const double *coef, int n_coef, const double *xi,
                   int smooth, double cutoff, const double *sel, int n_sel, int pmax,
                   double *work /* n */, int *iwork /* 2*(n+2) */,
                   int *nphase, int *n_minima, double *fe, double *avg, int *bounds, int *max_idx, int *min_idx,
                   int *safe, int *info)
{
    memcpy(work, lnpi_ref, sizeof(double) * (size_t)n);
    fo_reweight(work, ntot, n, mu1, mu1_ref, beta_ref);
    if (coef && n_coef > 0) {
        for (int c = 0; c < n_coef; ++c) {
            const double *a = coef + (size_t)c * (size_t)n;
            for (int i = 0; i < n; ++i) { const double t = xi[c] * a[i]; work[i] = work[i] + t; }
        }
        fo_normalize(work, n);
    }
    fo_normalize(work, n);                                    /* GH:475 */
    int *M = iwork, *m = iwork + (n + 2), nM = 0, nm = 0;
    int st = fo_relextrema(work, n, smooth, M, &nM, m, &nm, n + 2, info);
    *nphase = nM; *n_minima = nm; *safe = 0;
    if (st != FO_OK) return st;
    if (nM > pmax || nm > pmax + 1) return FO_CAPACITY;
    st = fo_phase_bounds(n, M, nM, m, nm, bounds);
    if (st != FO_OK) return st;
    for (int p = 0; p < nM; ++p) {
        max_idx[p] = M[p];
        fe[p] = fo_free_energy(work, bounds[2 * p], bounds[2 * p + 1]);
        if (n_sel > 0) fo_phase_averages(work, n, bounds[2 * p], bounds[2 * p + 1], sel, n_sel, avg + (size_t)p * (size_t)n_sel);
    }
    for (int q = 0; q < nm; ++q) min_idx[q] = m[q];
    *safe = fo_is_safe(work, n, M, nM, cutoff, 0);
    return FO_OK;
}

/* GH:2614-2630: min over pairs of wide-enough phases of (dF.E.)^2, default 100.0 */
double fo_phase_eq_err2(const double *fe, const int *bounds, int nphase, int min_width)
{
    double best = 100.0;
    if (nphase <= 1) return best;
    for (int i = 0; i < nphase; ++i) {
        if (bounds[2 * i + 1] - bounds[2 * i] < min_width) continue;
        for (int j = i + 1; j < nphase; ++j) {
            if (bounds[2 * j + 1] - bounds[2 * j] < min_width) continue;
            const double d = fe[i] - fe[j];
            if (d * d < best) best = d * d;
        }
    }
    return best;
}

/* 2-D joint histogram reweight (NEW capability; no reference implementation -- SURVEY 8(a) row 12).
 * Semantics anchored on pore_hist._cy_normalize / thermo (two_dim/h_ntot/pore_hist.pyx:57-80,154-184):
 * logsumexp over the ragged support [lo[i],hi[i]) of each row, -inf entries contribute nothing.
 * out[0]=lnZ (log sum exp of shifted surface), out[1]=<N1>, out[2]=<N2>, out[3..3+n_prop)=<prop>. */
void fo_reweight_2d(const double *lnpi, const int *bounds, int n1, int n2, const double *op1, const double *op2,
                    double b_dmu1, double b_dmu2, const double *props, int n_prop, double *out)
{
    double vmax = -INFINITY;
    for (int i = 0; i < n1; ++i)
        for (int j = bounds[2 * i]; j < bounds[2 * i + 1]; ++j) {
            double v = lnpi[(size_t)i * n2 + j] + b_dmu1 * op1[i] + b_dmu2 * op2[j];
            if (v > vmax) vmax = v;
        }
    double s = 0.0, s1 = 0.0, s2 = 0.0;
    double *sp = (double *)calloc((size_t)(n_prop > 0 ? n_prop : 1), sizeof(double));
    for (int i = 0; i < n1; ++i)
        for (int j = bounds[2 * i]; j < bounds[2 * i + 1]; ++j) {
            double v = lnpi[(size_t)i * n2 + j] + b_dmu1 * op1[i] + b_dmu2 * op2[j];
            if (!(v > -INFINITY)) continue;
            double e = exp(v - vmax);
            s += e; s1 += e * op1[i]; s2 += e * op2[j];
            for (int q = 0; q < n_prop; ++q) sp[q] += e * props[((size_t)q * n1 + i) * n2 + j];
        }
    out[0] = vmax + log(s); out[1] = s1 / s; out[2] = s2 / s;
    for (int q = 0; q < n_prop; ++q) out[3 + q] = sp[q] / s;
    free(sp);
}
