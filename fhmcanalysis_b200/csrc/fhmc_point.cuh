// fhmc_point.cuh -- evaluation of ONE state point by a group of G lanes (G = 1..32):
//   reweight (+Taylor terms) -> max-shifted sums -> windowed extrema -> repair -> phase bounds ->
//   per-phase free energies / averages -> normalisation constant -> is_safe.
// Shared by the sweep kernel (K1+K3+K2) and the batched coexistence solver (K4).
//
// Reference semantics restated here (jeetain/FHMCAnalysis, moments/histogram/one_dim/ntot/gc_hist.pyx = GH):
//   reweight GH:71-78, normalize GH:57-67, relextrema GH:317-415 (scipy argrelextrema 'clip'),
//   thermo GH:451-554, is_safe GH:556-596.   SURVEY.md Appendix A spells out the index rules.
//
// Exactness strategy for the integer outputs.  The reference compares NORMALISED values
// x_i = fl(u_i - c).  Subtraction of a constant is monotone, so strict extrema of x are a subset of
// the strict extrema of u.  Fast path: detect on u, compute c in the same pass that integrates
// the phases, then re-test the detected extrema (and only those) on x.  If one of them collapses to
// a tie, or if the detection found no maxima or no minima (the repair branches need ties/argmin on
// x), the slow path re-runs the detection on x with c known.  Every comparison that decides an
// output index is therefore evaluated on fl(u_i - c), exactly as the reference forms it.
#pragma once
#include "fhmc_common.cuh"

namespace fhmc {

#define FHMC_COMPACT_PMAX 8   // phase slots per state point the compact-record kernels handle
#define FHMC_COMPACT_DST 8    // destination buffers of one launch (this GPU's and, for a fused gather, its NVLink peers')

// Compact-record output of the sweep kernels (fhmc_sweep_1d_compact): records leave the kernel in the phase-major narrow
// form of fhmc_pack_phase_soa16, written straight to every destination buffer.
struct CompactArgs {
    unsigned char *dst[FHMC_COMPACT_DST];
    int n_dst;
    long long n_total;   // records a destination buffer is laid out for
    long long first;     // record index of state point 0 of this launch
    int fill_dead;       // write NaN / -1 into the phase slots >= nphase
    int *max_nphase;     // device int raised (atomicMax) to the largest phase count; nullable
    // k_sweep_cell (fhmc_cell.cu): state points it leaves to the table walk -- list[0 .. *count) (it appends; the indexed
    // instantiation of k_sweep_tab2 reads).  Null everywhere else.
    long long *ix_list;
    int *ix_count;
};

struct SweepArgs {
    fhmc_hist_desc d;
    const double *blob;
    fhmc_states st;
    fhmc_sweep_out out;   // compact-record launches: scratch records for the general evaluator, one per resident warp
    int blob_global;  // histogram too large for shared memory: rows are read from HBM/L2 (all lanes read the same bin)
    CompactArgs c;
};

// Shared-memory prologue of the 1-D kernels: [blob | mbarrier(16 B) | 2^(j/64) table(512 B) | ...].  Returns the
// pointer the evaluator reads rows through (shared copy, or the global blob when it does not fit).
__device__ __forceinline__ const double *stage_histogram(const SweepArgs &a, unsigned char *smem_raw, double *&s_tab)
{
    const uint32_t blob_bytes = a.blob_global ? 0u : (uint32_t)a.d.n_rows * (uint32_t)a.d.n_pad * 8u;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + blob_bytes);
    s_tab = reinterpret_cast<double *>(smem_raw + blob_bytes + 16);
    stage_exp_table(s_tab);
    if (a.blob_global) {
        __syncthreads();
        return a.blob;
    }
    stage_blob(smem_raw, a.blob, blob_bytes, bar);
    return reinterpret_cast<const double *>(smem_raw);
}

#define FHMC_NEED_SLOW 0x7fffffff

template <int G, bool TAYLOR>
struct PointEval {
    const SweepArgs &a;
    const double *sm;   // staged blob (shared memory)
    const int g;        // lane within group
    const unsigned member;  // warp mask of my group
    const unsigned gshift;  // first lane of my group
    const int n, last, w, npad, pmax, R;
    const uint32_t tab;  // shared-memory address of the 2^(j/64) table

    // state point (group-uniform)
    double s, xi[FHMC_MAX_TERMS], ts[FHMC_MAX_TERMS];
    // results (group-uniform)
    int P, nmin;
    double m, c;

    __device__ PointEval(const SweepArgs &args, const double *smem, int lane, const double *exp_table)
        : a(args), sm(smem), g(lane & (G - 1)),
          member(G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)))), gshift(lane & ~(G - 1)),
          n(args.d.n), last(args.d.n - 1), w(args.d.smooth), npad(args.d.n_pad), pmax(args.d.pmax),
          R((args.d.n + G - 1) / G), tab(smem_u32(exp_table))
    {
    }

    __device__ __forceinline__ void setup(double mu1, double beta, double dmu)
    {
        s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);  // GH:77, evaluated left to right
        if (TAYLOR) {
            const double dB = beta - a.d.beta_ref, dD = dmu - a.d.dmu_ref;
#pragma unroll
            for (int t = 0; t < FHMC_MAX_TERMS; ++t) {
                xi[t] = (t < a.d.n_coef) ? monomial(a.d.coef_kind[t], dB, dD, mu1) : 0.0;
                ts[t] = (t < a.d.n_term) ? monomial(a.d.sel_kind[t], dB, dD, mu1) : 0.0;
            }
        }
    }

    // u_i = fl(lnPI_i + fl(s*N_i)) (+ Taylor terms)
    __device__ __forceinline__ double U(int i) const
    {
        double u = __dadd_rn(sm[i], __dmul_rn(s, sm[npad + i]));
        if (TAYLOR) {
            // fully unrolled so that xi[] stays in registers.  Multi-lane groups (few registers to spare, short
            // Taylor series in the solver) leave at the first unused term; the one-lane evaluator that is inlined into
            // the thread-per-state-point kernel keeps the branch-free predicated form (measured: the early exit
            // costs that kernel spills and 30 % at n_coef = 6).
#pragma unroll
            for (int t = 0; t < FHMC_MAX_TERMS; ++t) {
                if (G > 1) {
                    if (t >= a.d.n_coef) break;
                    u = fma(xi[t], sm[a.d.coef_row[t] * npad + i], u);
                } else {
                    if (t < a.d.n_coef) u = fma(xi[t], sm[a.d.coef_row[t] * npad + i], u);
                }
            }
        }
        return u;
    }
    __device__ __forceinline__ double Xsel(int q, int i) const
    {
        const double *row = sm + a.d.sel_row[q] * npad + i;
        double x = row[0];
        if (TAYLOR) {
#pragma unroll
            for (int t = 1; t < FHMC_MAX_TERMS; ++t) {
                if (G > 1) {
                    if (t >= a.d.n_term) break;
                    x = fma(ts[t], row[t * npad], x);
                } else {
                    if (t < a.d.n_term) x = fma(ts[t], row[t * npad], x);
                }
            }
        }
        return x;
    }

    // value the reference compares: fl(u - c) once c is known, u before
    __device__ __forceinline__ double V(int i, bool use_c, double cc) const
    {
        const double u = U(i);
        return use_c ? __dsub_rn(u, cc) : u;
    }

    // remaining shifts 2..w of the argrelextrema test (shift 1 already passed)
    __device__ bool window_ok(int i, double xc, bool is_max, bool use_c, double cc, int d0) const
    {
        for (int d = d0; d <= w; ++d) {
            const int jl = (i - d < 0) ? 0 : i - d;
            const int jr = (i + d > last) ? last : i + d;
            const double xl = V(jl, use_c, cc), xr = V(jr, use_c, cc);
            const bool ok = is_max ? (xc > xl && xc > xr) : (xc < xl && xc < xr);
            if (!ok) return false;
        }
        return true;
    }

    // ---- K3a: raw windowed extrema (GH:329-330).  Raw hit k is stored at list position 1+k
    // (position 0 is reserved for a prepended 0).  Returns counts and the max/min of u. ----------
    __device__ void detect(bool use_c, double cc, int *maxl, int *minl, int &cntM, int &cntm, double &umax, double &umin) const
    {
        cntM = 0;
        cntm = 0;
        double mx = -CUDART_INF, mn = CUDART_INF;
        if (G == 1) {
            double u1 = U(0);
            mx = u1;
            mn = u1;
            double xm = 0.0, xc = use_c ? __dsub_rn(u1, cc) : u1;
            u1 = U(1);
            mx = fmax(mx, u1);
            mn = fmin(mn, u1);
            double xp = use_c ? __dsub_rn(u1, cc) : u1;
            for (int i = 1; i < last; ++i) {
                xm = xc;
                xc = xp;
                const double un = U(i + 1);
                mx = fmax(mx, un);
                mn = fmin(mn, un);
                xp = use_c ? __dsub_rn(un, cc) : un;
                if (xc > xm && xc > xp) {
                    if (window_ok(i, xc, true, use_c, cc, 2)) {
                        if (1 + cntM <= pmax - 1) maxl[1 + cntM] = i;
                        ++cntM;
                    }
                } else if (xc < xm && xc < xp) {
                    if (window_ok(i, xc, false, use_c, cc, 2)) {
                        if (1 + cntm <= pmax) minl[1 + cntm] = i;
                        ++cntm;
                    }
                }
            }
        } else {
            const unsigned gmask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
            const unsigned below = (1u << g) - 1u;
            // every u_i is evaluated once: the neighbours i-1 / i+1 live in the adjacent lanes, the left neighbour of
            // lane 0 is the last lane's value of the previous round, the right neighbour of the last lane is lane 0's
            // value of the next round (computed one round ahead)
            const int lane = gshift + g;
            double u_cur = (g < n) ? U(g) : 0.0, u_prev_last = 0.0;
            for (int r = 0; r < R; ++r) {
                const int i = r * G + g;
                const double u_next = (i + G < n) ? U(i + G) : 0.0;
                double ul = __shfl_sync(member, u_cur, (lane + 31) & 31);
                double ur = __shfl_sync(member, u_cur, (lane + 1) & 31);
                const double u_next0 = __shfl_sync(member, u_next, gshift);
                const double u_last = __shfl_sync(member, u_cur, gshift + G - 1);
                if (g == 0) ul = u_prev_last;
                if (g == G - 1) ur = u_next0;
                bool isM = false, ism = false;
                if (i < n) {
                    const double u = u_cur;
                    mx = fmax(mx, u);
                    mn = fmin(mn, u);
                    if (i > 0 && i < last) {
                        const double xc = use_c ? __dsub_rn(u, cc) : u;
                        const double xl = use_c ? __dsub_rn(ul, cc) : ul, xr = use_c ? __dsub_rn(ur, cc) : ur;
                        if (xc > xl && xc > xr) isM = window_ok(i, xc, true, use_c, cc, 2);
                        else if (xc < xl && xc < xr) ism = window_ok(i, xc, false, use_c, cc, 2);
                    }
                }
                u_prev_last = u_last;
                u_cur = u_next;
                const unsigned bM = (__ballot_sync(member, isM) >> gshift) & gmask;
                const unsigned bm = (__ballot_sync(member, ism) >> gshift) & gmask;
                if (isM) {
                    const int pos = 1 + cntM + __popc(bM & below);
                    if (pos <= pmax - 1) maxl[pos] = i;
                }
                if (ism) {
                    const int pos = 1 + cntm + __popc(bm & below);
                    if (pos <= pmax) minl[pos] = i;
                }
                cntM += __popc(bM);
                cntm += __popc(bm);
            }
            mx = group_max<G>(mx, member);
            mn = group_min<G>(mn, member);
            __syncwarp(member);
        }
        umax = mx;
        umin = mn;
    }

    // ---- K3b: endpoint insertion / repair (GH:333-386), validation (GH:403-415), bounds (GH:498-520).
    // Leader lane only.  Returns the status code, FHMC_NEED_SLOW when the branch needs values on the
    // normalised array and c is not known yet.  Lists end up starting at position 0. -------------
    // All lanes of the group: how many bins tie with the max / min of the normalised array, and the first of each.
    __device__ void tie_scan(double cc, double umax, double umin, int &cM, int &cm, int &pM, int &pm) const
    {
        const double vmax = __dsub_rn(umax, cc), vmin = __dsub_rn(umin, cc);
        cM = 0; cm = 0; pM = 0x7fffffff; pm = 0x7fffffff;
        for (int j = g; j < n; j += G) {
            const double v = V(j, true, cc);
            if (v == vmax) { ++cM; pM = min(pM, j); }
            if (v == vmin) { ++cm; pm = min(pm, j); }
        }
        if (G > 1) {
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                cM += __shfl_xor_sync(member, cM, o);
                cm += __shfl_xor_sync(member, cm, o);
                pM = min(pM, __shfl_xor_sync(member, pM, o));
                pm = min(pm, __shfl_xor_sync(member, pm, o));
            }
        }
    }

    __device__ int repair(bool use_c, double cc, int cntM, int cntm, double umax, double umin, int *maxl, int *minl,
                          int *bl, int &nM_out, int &nm_out, unsigned &flags, bool &partition,
                          int tcM = -1, int tcm = -1, int tpM = 0, int tpm = 0) const
    {
        int nM = 0, nm = 0;
        nM_out = 0;
        nm_out = 0;
        partition = false;
        if (cntM > pmax - 1 || cntm > pmax) return FHMC_E_CAPACITY;
        // raw hits sit at positions 1..cnt; compact() moves them to 0..cnt-1
        auto compact = [](int *lst, int cnt) { for (int k = 0; k < cnt; ++k) lst[k] = lst[k + 1]; };
        if (cntM > 0 && cntm > 0) {  // GH:333-351 (raw lists never contain 0 or last)
            const int fM = maxl[1], fm = minl[1];
            if (fM < fm) { minl[0] = 0; nm = cntm + 1; compact(maxl, cntM); nM = cntM; }
            else if (fM > fm) { maxl[0] = 0; nM = cntM + 1; compact(minl, cntm); nm = cntm; }
            else return FHMC_E_BAD_FRONT;
            const int lM = maxl[nM - 1], lm = minl[nm - 1];
            if (lM < lm) { if (nM >= pmax) return FHMC_E_CAPACITY; maxl[nM++] = last; }
            else if (lM > lm) { if (nm >= pmax + 1) return FHMC_E_CAPACITY; minl[nm++] = last; }
            else return FHMC_E_BAD_BACK;
        } else {
            // exactly one interior extremum of one kind needs no value of the normalised array (GH:364-366, 379-381);
            // gap filling and the "all tied with the max/min" branch do
            if (!use_c && !((cntM == 1 && cntm == 0) || (cntM == 0 && cntm == 1))) return FHMC_NEED_SLOW;
            if (cntM > 0) {  // GH:352-366
                compact(maxl, cntM);
                nM = cntM;
                if (cntM > 1) {
                    if (cntM + 1 > pmax + 1) return FHMC_E_CAPACITY;
                    minl[nm++] = 0;
                    for (int k = 0; k < cntM - 1; ++k) {
                        const int l = maxl[k], r = maxl[k + 1];
                        double v = V(l, true, cc);
                        for (int j = l; j < r; ++j) v = fmin(v, V(j, true, cc));
                        int pos = l, ties = 0;
                        for (int j = l; j < r; ++j) if (V(j, true, cc) == v) { if (!ties) pos = j; ++ties; }
                        if (ties != 1) return FHMC_E_RAGGED_GAP;
                        minl[nm++] = pos;
                    }
                    minl[nm++] = last;
                    flags |= FHMC_ST_GAP_FILL;
                } else { minl[0] = 0; minl[1] = last; nm = 2; }
            } else if (cntm > 0) {  // GH:367-381
                compact(minl, cntm);
                nm = cntm;
                if (cntm > 1) {
                    if (cntm + 1 > pmax) return FHMC_E_CAPACITY;
                    maxl[nM++] = 0;
                    for (int k = 0; k < cntm - 1; ++k) {
                        const int l = minl[k], r = minl[k + 1];
                        double v = V(l, true, cc);
                        for (int j = l; j < r; ++j) v = fmax(v, V(j, true, cc));
                        int pos = l, ties = 0;
                        for (int j = l; j < r; ++j) if (V(j, true, cc) == v) { if (!ties) pos = j; ++ties; }
                        if (ties != 1) return FHMC_E_RAGGED_GAP;
                        maxl[nM++] = pos;
                    }
                    maxl[nM++] = last;
                    flags |= FHMC_ST_GAP_FILL;
                } else {
                    if (pmax < 2) return FHMC_E_CAPACITY;
                    maxl[0] = 0; maxl[1] = last; nM = 2;
                }
            } else {  // GH:382-386: all positions tied with the max / min of the normalised array
                const double vmax = __dsub_rn(umax, cc), vmin = __dsub_rn(umin, cc);
                bool over = false;
                // a unique argmax and a unique argmin (the usual case) come from the group-wide tie_scan() of run();
                // ties (or no scan) take the ordered serial scan
                if (tcM == 1 && tcm == 1) {
                    if (nM < pmax) maxl[nM] = tpM; else over = true;
                    ++nM;
                    if (nm < pmax + 1) minl[nm] = tpm; else over = true;
                    ++nm;
                } else {
                    for (int j = 0; j < n; ++j) {
                        const double v = V(j, true, cc);
                        if (v == vmax) { if (nM < pmax) maxl[nM] = j; else over = true; ++nM; }
                        if (v == vmin) { if (nm < pmax + 1) minl[nm] = j; else over = true; ++nm; }
                    }
                }
                if (over) return FHMC_E_CAPACITY;
            }
        }
        nM_out = nM;
        nm_out = nm;
        // GH:403-415
        const int diff = nM - nm;
        if (diff > 1 || diff < -1) return FHMC_E_COUNT_MISMATCH;
        const int total = nM + nm;
        const bool max_first = maxl[0] < minl[0];
        const int nev = max_first ? nM : nm, nod = max_first ? nm : nM;
        if (nev != (total + 1) / 2 || nod != total / 2) return FHMC_E_COUNT_MISMATCH;
        int prev = -1;
        for (int k = 0; k < total; ++k) {
            const int *lst = ((k & 1) == 0) == max_first ? maxl : minl;
            const int v = lst[k >> 1];
            if (k > 0 && prev > v) return FHMC_E_NOT_SORTED;
            prev = v;
        }
        // GH:498-520
        int ctr = 0;
        bool part = true;
        int prev_right = 0;
        for (int p = 0; p < nM; ++p) {
            int left, right;
            if (maxl[p] > 0) { if (ctr >= nm) return FHMC_E_INDEX; left = minl[ctr]; ++ctr; }
            else left = 0;
            if (maxl[p] < last) { if (ctr >= nm) return FHMC_E_INDEX; right = minl[ctr]; }
            else right = n;
            if (right == last) right += 1;
            bl[2 * p] = left;
            bl[2 * p + 1] = right;
            part = part && (left == prev_right) && (right >= left);
            prev_right = right;
        }
        partition = part && (prev_right == n) && nM > 0;
        return FHMC_OK;
    }

    // ---- K2: sums over an arbitrary bin range about an explicit shift ---------------------------
    __device__ void range_sums(int left, int right, double shift, double &Sg, double (&Ag)[FHMC_MAX_SEL]) const
    {
        double S = 0.0, A[FHMC_MAX_SEL];
#pragma unroll
        for (int q = 0; q < FHMC_MAX_SEL; ++q) A[q] = 0.0;
        for (int i = left + g; i < right; i += G) {
            const double e = exp_nonpos(U(i) - shift, tab);
            S += e;
#pragma unroll
            for (int q = 0; q < FHMC_MAX_SEL; ++q)
                if (q < a.d.n_sel) A[q] = fma(e, Xsel(q, i), A[q]);
        }
        Sg = group_sum<G>(S, member);
#pragma unroll
        for (int q = 0; q < FHMC_MAX_SEL; ++q) Ag[q] = (q < a.d.n_sel) ? group_sum<G>(A[q], member) : 0.0;
    }
    __device__ double range_max(int left, int right) const
    {
        double mx = -CUDART_INF;
        for (int i = left + g; i < right; i += G) mx = fmax(mx, U(i));
        return group_max<G>(mx, member);
    }

    __device__ __forceinline__ void write_phase(long long rec, int p, double lnS, double u0, double Sg,
                                                const double (&Ag)[FHMC_MAX_SEL]) const
    {
        if (g == 0) {
            a.out.fe[rec * pmax + p] = -((lnS)-u0);  // F.E./kT = -ln sum exp(x_j - x_0), GH:523-526
#pragma unroll
            for (int q = 0; q < FHMC_MAX_SEL; ++q)
                if (q < a.d.n_sel) a.out.avg[(rec * pmax + p) * a.d.n_sel + q] = Ag[q] / Sg;
        }
    }

    // phase p integrated about its own maximum (arbitrary bounds, arbitrarily unlikely phases)
    __device__ double phase_general(long long rec, int p, int left, int right, double u0) const
    {
        double Sg, Ag[FHMC_MAX_SEL];
        if (right <= left) {
            if (g == 0) {
                a.out.fe[rec * pmax + p] = 1.7976931348623157e308;  // fold over an empty range, GH:523
                for (int q = 0; q < a.d.n_sel; ++q) a.out.avg[(rec * pmax + p) * a.d.n_sel + q] = CUDART_NAN;
            }
            return 0.0;
        }
        const double ml = range_max(left, right);
        range_sums(left, right, ml, Sg, Ag);
        write_phase(rec, p, ml + log(Sg), u0, Sg, Ag);
        return Sg * exp(ml - m);
    }

    // ---- K2: all phases in ONE pass when they tile [0,n): accumulate, flush at each boundary ----
    __device__ double partition_sums(long long rec, const int *bl, double u0, unsigned &rescue) const
    {
        double S = 0.0, A[FHMC_MAX_SEL], Stot = 0.0;
#pragma unroll
        for (int q = 0; q < FHMC_MAX_SEL; ++q) A[q] = 0.0;
        int cur = 0;
        int nb = bl[1];
        rescue = 0;
        auto flush = [&]() {
            const double Sg = group_sum<G>(S, member);
            double Ag[FHMC_MAX_SEL];
#pragma unroll
            for (int q = 0; q < FHMC_MAX_SEL; ++q) Ag[q] = (q < a.d.n_sel) ? group_sum<G>(A[q], member) : 0.0;
            if (Sg < 1e-280) rescue |= 1u << (cur < 31 ? cur : 31);
            else write_phase(rec, cur, m + log(Sg), u0, Sg, Ag);
            Stot += Sg;
            S = 0.0;
#pragma unroll
            for (int q = 0; q < FHMC_MAX_SEL; ++q) A[q] = 0.0;
            ++cur;
            nb = (cur < P) ? bl[2 * cur + 1] : 0x7fffffff;
        };
        for (int r = 0; r < R; ++r) {
            const int i = r * G + g;
            double e = 0.0, X[FHMC_MAX_SEL];
#pragma unroll
            for (int q = 0; q < FHMC_MAX_SEL; ++q) X[q] = 0.0;
            if (G == 1 || i < n) {
                e = exp_nonpos(U(i) - m, tab);
#pragma unroll
                for (int q = 0; q < FHMC_MAX_SEL; ++q)
                    if (q < a.d.n_sel) X[q] = Xsel(q, i);
            }
            const int rowend = r * G + G;
            while (nb < rowend) {
                if (i < nb) {
                    S += e;
#pragma unroll
                    for (int q = 0; q < FHMC_MAX_SEL; ++q) A[q] = fma(e, X[q], A[q]);
                    e = 0.0;
                }
                flush();
            }
            S += e;
#pragma unroll
            for (int q = 0; q < FHMC_MAX_SEL; ++q) A[q] = fma(e, X[q], A[q]);
        }
        while (cur < P) flush();
        return Stot;
    }

    // re-test the detected interior extrema on the normalised values (see file header)
    __device__ bool verify(const int *maxl, const int *minl, double cc) const
    {
        bool bad = false;
        for (int k = g; k < P + nmin; k += G) {
            const bool is_max = k < P;
            const int idx = is_max ? maxl[k] : minl[k - P];
            if (idx > 0 && idx < last) {
                const double xc = V(idx, true, cc);
                if (!window_ok(idx, xc, is_max, true, cc, 1)) bad = true;
            }
        }
        return __any_sync(member, bad);
    }

    // ---- one state point, written to record `rec`.  Returns the status word. --------------------
    __device__ unsigned run(long long rec)
    {
        int *maxl = a.out.max_idx + rec * pmax;
        int *minl = a.out.min_idx + rec * (pmax + 1);
        int *bl = a.out.bounds + rec * pmax * 2;
        unsigned flags = 0;
        int code = FHMC_OK;
        P = 0;
        nmin = 0;
        c = 0.0;
        const double u0 = U(0);
        const double ulast = U(last);

        if (n < 3 && !a.d.complete) {
            code = FHMC_E_TOO_SHORT;
        } else if (a.d.complete) {  // thermo(complete=True), GH:489-494, 518-520; is_safe GH:593-596
            m = range_max(0, n);
            double Sg, Ag[FHMC_MAX_SEL];
            range_sums(0, n, m, Sg, Ag);
            c = m + log(Sg);
            P = 1;
            if (g == 0) { bl[0] = 0; bl[1] = n; }
            write_phase(rec, 0, c, u0, Sg, Ag);
            if (!(__dsub_rn(__dsub_rn(m, c), __dsub_rn(ulast, c)) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
        } else {
            int cntM, cntm;
            double umin;
            bool partition = false, done = false;
            // ---------------- fast path: detect on u ------------------------------------------
            detect(false, 0.0, maxl, minl, cntM, cntm, m, umin);
            int packed = 0;
            if (g == 0) {
                unsigned f = 0;
                int nM, nm;
                bool part;
                const int rc = repair(false, 0.0, cntM, cntm, m, umin, maxl, minl, bl, nM, nm, f, part);
                packed = (rc == FHMC_NEED_SLOW) ? -1 : (rc | (part ? 0x100 : 0) | (nM << 9));
                nmin = nm;
            }
            if (G > 1) {
                __syncwarp(member);
                packed = __shfl_sync(member, packed, gshift);
                nmin = __shfl_sync(member, nmin, gshift);
            }
            if (packed >= 0) {
                code = packed & 0xff;
                partition = (packed & 0x100) != 0;
                P = packed >> 9;
                if (code == FHMC_OK) {
                    unsigned rescue = 0;
                    double Stot = 0.0;
                    if (partition) Stot = partition_sums(rec, bl, u0, rescue);
                    else for (int p = 0; p < P; ++p) Stot += phase_general(rec, p, bl[2 * p], bl[2 * p + 1], u0);
                    if (rescue) {
                        flags |= FHMC_ST_RESCUED;
                        for (int p = 0; p < P; ++p)
                            if ((rescue >> (p < 31 ? p : 31)) & 1u) {
                                const double Sp = partition_sum_probe(rec, p, bl, u0);
                                (void)Sp;
                            }
                    }
                    if (!partition) {  // phases do not tile [0,n): c needs its own full pass
                        double Sg, Ag[FHMC_MAX_SEL];
                        range_sums(0, n, m, Sg, Ag);
                        Stot = Sg;
                    }
                    c = m + log(Stot);
                    done = a.d.compare_raw ? true : !verify(maxl, minl, c);
                } else {
                    done = true;  // the reference raises; nothing to integrate
                    double Sg, Ag[FHMC_MAX_SEL];
                    range_sums(0, n, m, Sg, Ag);
                    c = m + log(Sg);
                }
            }
            // ---------------- slow path: everything on the normalised array --------------------
            if (!done) {
                flags |= FHMC_ST_SLOW_PATH;
                double Sfull, Afull[FHMC_MAX_SEL];
                range_sums(0, n, m, Sfull, Afull);
                c = m + log(Sfull);
                const double cc = a.d.compare_raw ? 0.0 : c;  // fl(u - 0) == u: relextrema() on the raw array
                detect(true, cc, maxl, minl, cntM, cntm, m, umin);
                packed = 0;
                unsigned f = 0;
                int tcM = -1, tcm = -1, tpM = 0, tpm = 0;
                if (cntM == 0 && cntm == 0) tie_scan(cc, m, umin, tcM, tcm, tpM, tpm);  // GH:382-386, group-wide
                if (g == 0) {
                    int nM, nm;
                    bool part;
                    const int rc = repair(true, cc, cntM, cntm, m, umin, maxl, minl, bl, nM, nm, f, part, tcM, tcm, tpM, tpm);
                    packed = rc | (part ? 0x100 : 0) | (nM << 9);
                    nmin = nm;
                }
                if (G > 1) {
                    __syncwarp(member);
                    packed = __shfl_sync(member, packed, gshift);
                    nmin = __shfl_sync(member, nmin, gshift);
                    f = __shfl_sync(member, f, gshift);
                }
                flags |= f;
                code = packed & 0xff;
                partition = (packed & 0x100) != 0;
                P = packed >> 9;
                if (code == FHMC_OK) {
                    if (P == 1 && bl[0] == 0 && bl[1] == n && Sfull >= 1e-280) {
                        write_phase(rec, 0, c, u0, Sfull, Afull);
                    } else if (partition) {
                        unsigned rescue = 0;
                        partition_sums(rec, bl, u0, rescue);
                        if (rescue) {
                            flags |= FHMC_ST_RESCUED;
                            for (int p = 0; p < P; ++p)
                                if ((rescue >> (p < 31 ? p : 31)) & 1u) partition_sum_probe(rec, p, bl, u0);
                        }
                    } else {
                        for (int p = 0; p < P; ++p) phase_general(rec, p, bl[2 * p], bl[2 * p + 1], u0);
                    }
                }
            }
            if (code == FHMC_OK && P > 0) {  // is_safe, GH:586-591
                const double xM = __dsub_rn(U(maxl[P - 1]), c), xl = __dsub_rn(ulast, c);
                if (!(__dsub_rn(xM, xl) < a.d.cutoff)) flags |= FHMC_ST_SAFE;
            }
        }
        const unsigned status = (unsigned)code | flags;
        if (g == 0) {
            a.out.status[rec] = status;
            a.out.nphase[rec] = P;
            a.out.nmin[rec] = nmin;
            a.out.lnnorm[rec] = c;
        }
        return status;
    }

    // a phase whose max-shifted weight underflowed: integrate it again about its own maximum.
    // (phases 31.. share one rescue bit, so re-test the weight before doing the work)
    __device__ double partition_sum_probe(long long rec, int p, const int *bl, double u0) const
    {
        return phase_general(rec, p, bl[2 * p], bl[2 * p + 1], u0);
    }
};

}  // namespace fhmc
