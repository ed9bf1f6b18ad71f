/*
 * fhmc_b200.h -- C ABI of libfhmc_b200.so: B200 (sm_100a) kernels for the histogram-reweighting
 * hot path of FHMCAnalysis.
 *
 * The reference (jeetain/FHMCAnalysis) has no FFI of its own: its native layer is a handful of
 * Cython cdef functions bound onto the Python class `histogram`
 * (moments/histogram/one_dim/ntot/gc_hist.pyx, "GH").  Each entry point below names the
 * reference routine(s) it replaces; INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to call it.
 *
 * Conventions
 *   - every pointer inside the fhmc_* structs is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all buffers; no call allocates device memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, non-zero = error (text via fhmc_last_error());
 *   - all floating point is IEEE fp64; indices are 32-bit signed.
 */
#ifndef FHMC_B200_H
#define FHMC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHMC_ABI_VERSION 1
#define FHMC_MAX_TERMS 8   /* Taylor terms per array                      */
#define FHMC_MAX_SEL 4     /* quantities averaged inside the fused sweep   */

/* Monomials in (dB = beta - beta_ref, dD = dmu2 - dmu2_ref, mu1) multiplying a coefficient row.
 * They are the terms of the second/third-order Taylor expansion the reference builds in
 * GH:968-1239 (_temp_dmu_extrap_{1,2}[_multi]) and GH:2208 (_dB3).                            */
enum fhmc_monomial {
    FHMC_M_DB = 0,        /* dB            */
    FHMC_M_DD = 1,        /* dD            */
    FHMC_M_DB2 = 2,       /* dB^2 / 2      */
    FHMC_M_DBDD = 3,      /* dB * dD       */
    FHMC_M_DD2 = 4,       /* dD^2 / 2      */
    FHMC_M_DB3 = 5,       /* dB^3 / 6      */
    FHMC_M_DB_MU1 = 6,    /* dB * mu1      (the mu1*N part of d lnPI / d beta, GH:1660-1722)    */
    FHMC_M_ONE = 7        /* 1             */
};

/* per-state-point status word (fhmc_sweep_out.status) */
#define FHMC_ST_CODE_MASK   0xFFu   /* 0 ok, else "the reference would raise": codes below        */
#define FHMC_ST_SAFE        0x100u  /* is_safe(cutoff) == True (GH:556-596, fresh extrema)        */
#define FHMC_ST_GAP_FILL    0x200u  /* GH:355-363 / 370-378 branch taken (unique gap extremum)    */
#define FHMC_ST_SLOW_PATH   0x400u  /* extrema had to be re-evaluated on the normalised array     */
#define FHMC_ST_RESCUED     0x800u  /* a phase with negligible weight was re-summed about its own max */
#define FHMC_ST_FAST        0x1000u /* produced by the one-thread-per-point, one-exp-pass kernel (diagnostic)     */
#define FHMC_ST_JUMP        0x2000u /* fhmc_find_phase_eq_1d: code 0 but |dfe| > lnz_tol -- the search ended on a jump of dF.E. */
#define FHMC_ST_LEAN        0x4000u /* produced by the lean warp-per-point evaluator (diagnostic)                  */
enum fhmc_status_code {
    FHMC_OK = 0,
    FHMC_E_TOO_SHORT = 1,       /* GH:326-327                                                    */
    FHMC_E_BAD_FRONT = 2,       /* GH:341-342                                                    */
    FHMC_E_BAD_BACK = 3,        /* GH:350-351                                                    */
    FHMC_E_COUNT_MISMATCH = 4,  /* GH:403-404 (and the NumPy slice-assignment errors of 408-412) */
    FHMC_E_NOT_SORTED = 5,      /* GH:414-415                                                    */
    FHMC_E_INDEX = 6,           /* IndexError in GH:504 / 511                                    */
    FHMC_E_RAGGED_GAP = 7,      /* GH:355-363 / 370-378 with tied gap extrema                    */
    FHMC_E_CAPACITY = 8         /* more extrema than pmax: call again with a larger pmax         */
};

/*
 * One histogram, packed by the host as a dense "blob" of fp64 rows of stride n_pad (n_pad even,
 * n_pad >= n; the blob base must be 16-byte aligned so one TMA bulk copy can stage it in shared
 * memory):
 *   row 0                    ln(PI)(N)                       (GH:148)
 *   row 1                    N_tot as fp64                   (GH:154)
 *   rows 2 .. n_rows-1       Taylor coefficient rows / rows of quantities to average.
 * lnPI'(N) = row0 + fl(s*row1) + sum_c mono(coef_kind[c]) * row[coef_row[c]],   s = fl(fl(mu1-mu1_ref)*beta_ref)
 * (the un-fused s*N product and add reproduce GH:77 bit-for-bit).
 * Quantity q (q < n_sel):   X_q(N) = sum_t mono(sel_kind[t]) * row[sel_row[q] + t],  t < n_term
 * (sel_kind[0] must be FHMC_M_ONE).
 */
typedef struct fhmc_hist_desc {
    int n;                          /* bins                                                      */
    int n_pad;                      /* row stride (doubles), even                                */
    int n_rows;                     /* rows in the blob                                          */
    int n_coef;                     /* Taylor terms added to lnPI, <= FHMC_MAX_TERMS             */
    int coef_row[FHMC_MAX_TERMS];
    int coef_kind[FHMC_MAX_TERMS];
    int n_sel;                      /* <= FHMC_MAX_SEL                                           */
    int n_term;                     /* rows per quantity, <= FHMC_MAX_TERMS                      */
    int sel_row[FHMC_MAX_SEL];
    int sel_kind[FHMC_MAX_TERMS];
    int smooth;                     /* extrema window, metadata['smooth'] (GH:329-330), >= 1     */
    int pmax;                       /* capacity of the per-state-point phase arrays              */
    int complete;                   /* thermo(complete=True): one phase [0,n), no extrema        */
    int compare_raw;                /* relextrema() on the array as given (GH:329): compare u, not fl(u-c) */
    double cutoff;                  /* is_safe cutoff (GH:556)                                   */
    double beta_ref;                /* data['curr_beta'] of the stored histogram                 */
    double mu1_ref;                 /* data['curr_mu'][0]                                        */
    double dmu_ref;                 /* data['curr_mu'][1]-data['curr_mu'][0] (0 for 1 species)   */
    /* Optional accelerator for pure mu sweeps (n_coef == 0): upper concave envelope of the points
     * (N_i, lnPI_i).  Row hull_row holds the hull_len-1 edge slopes (strictly decreasing), row
     * hull_row+1 the hull_len vertex bin indices (as fp64).  hull_len == 0: not provided.          */
    int hull_row;
    int hull_len;
    /* Caller's promise that row 1 (N) is uniformly spaced and |lnPI_{i+4} - lnPI_i| < 300 for all i: pure mu sweeps may
     * then advance exp(lnPI_i + s N_i - shift) along four interleaved bin chains by two multiplications per bin
     * (e_i = e_{i-4} * exp(4 s dN) * exp(lnPI_i - lnPI_{i-4})), re-anchored with a true exp every 64 bins (value 1).
     * Value 2: product form -- exp(lnPI_i - A_seg) tabulated per CTA, the sums of a 4-bin block are Horner polynomials in
     * exp(s dN), re-anchored every 128 bins (fhmc_fast_prod.cu).  Value 3: as 2 with two state points per thread when the
     * sweep has more than 2*256 state points per SM (fhmc_prod.cuh).                                                   */
    int mu_recurrence;
    /* fhmc_find_phase_eq_1d only: minimum width (bins) of a phase that counts in the coexistence objective.
     * 0 = 2*smooth (ntot/gc_hist.pyx:652); n1/gc_hist.pyx:1479 passes smooth itself.                                  */
    int min_width;
    /* Optional accelerator for pure mu sweeps with compact records (fhmc_sweep_1d_compact, fhmc_sweep_host_compact*): per-histogram
     * tables built once by fhmc_mu_tables_build() -- the product tables of the sweep kernel and, per elementary interval of
     * the tilt -s dN between two consecutive chord slopes of ln(PI), the outcome of relextrema() + the phase bounds of
     * thermo() (GH:317-415, 498-520).  NULL: not provided (the sweep kernel builds its tables per launch and tests for
     * extrema as it walks).  The buffer must stay valid and unchanged while sweeps use it, and belongs to exactly this
     * blob / n / smooth / sel_row combination.                                                                        */
    const void *mu_tables;
    /* Optional second level on top of mu_tables, for DENSE mu sweeps (fhmc_sweep_1d_compact): moment expansions of the per-phase
     * sums about the centres of small tilt cells, built by fhmc_mu_cells_build() for one range of mu.  Inside an elementary
     * interval the phase bounds are fixed, so every per-phase sum  sum_i exp(lnPI_i + s N_i) X_q(i)  is an entire function of s:
     * a state point then costs one degree-7 polynomial per phase and quantity instead of a walk over the bins (GH:71-78,
     * 498-554).  State points outside the range, or that fail a rounding-margin test, take the table walk.  NULL: not provided.
     * The buffer also holds the counter of the state points a sweep leaves to the table walk: sweeps that share one cells buffer
     * must be ordered on one stream (or serialised by events); the tables above carry no such restriction.                    */
    const void *mu_cells;
} fhmc_hist_desc;

/*
 * State points.  State point s (0 <= s < n_states) uses
 *   mu1  = mu1 [(s / mu1_div ) % n_mu1 ],  beta = beta[(s / beta_div) % n_beta] (NULL: beta_ref),
 *   dmu2 = dmu [(s / dmu_div ) % n_dmu ]   (NULL: dmu_ref).
 * Flat list: all div = 1 and n_* = n_states (or 1 to broadcast).  (beta x dmu) grid like
 * temp_dmu_extrap_multi (GH:813-887): beta_div = n_dmu, dmu_div = 1, n_states = n_beta*n_dmu.
 */
typedef struct fhmc_states {
    long long n_states;
    const double *mu1;  long long n_mu1;  long long mu1_div;
    const double *beta; long long n_beta; long long beta_div;
    const double *dmu;  long long n_dmu;  long long dmu_div;
} fhmc_states;

/* Results, one record per state point.  status, nphase, lnnorm, fe, bounds, max_idx, min_idx, nmin
 * are required (max_idx/min_idx/bounds double as kernel work space); avg may be NULL iff n_sel==0. */
typedef struct fhmc_sweep_out {
    unsigned *status;   /* [S]               see FHMC_ST_*                                       */
    int *nphase;        /* [S]               len(data['ln(PI)_maxima_idx'])                      */
    int *nmin;          /* [S]               len(data['ln(PI)_minima_idx'])                      */
    double *lnnorm;     /* [S]               c:  normalised lnPI = lnPI' - c          (GH:57-67) */
    double *fe;         /* [S][pmax]         thermo[p]['F.E./kT']                    (GH:523-526) */
    double *avg;        /* [S][pmax][n_sel]  phase averages                          (GH:530-541) */
    int *bounds;        /* [S][pmax][2]      thermo[p]['bound_idx']                  (GH:498-520) */
    int *max_idx;       /* [S][pmax]                                                             */
    int *min_idx;       /* [S][pmax+1]                                                           */
} fhmc_sweep_out;

int fhmc_version(void);
const char *fhmc_last_error(void);
/* Diagnostic: name of the sweep / solver kernel the calling thread launched last ("k_sweep_prod2", "k_sweep_fast<prod>",
 * "k_sweep_1d<32>", "k_solve_tab", ...): lets tests and bench.py state which kernel produced the records they check. */
const char *fhmc_last_kernel(void);

/* Device/SM query used by the host layer to size grids. */
int fhmc_device_info(int *sm_count, int *max_smem_optin);

/*
 * K1+K3+K2 fused: for every state point  reweight (GH:268-289, 71-78) [+ Taylor extrapolation,
 * GH:813-1239] + normalise (GH:57-67) + relextrema (GH:317-415) + phase bounds / free energies /
 * selected averages (thermo, GH:451-554) + is_safe (GH:556-596).
 * lanes_per_point in {1,4,32}: threads cooperating on one state point (0 = choose).
 */
int fhmc_sweep_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                  const fhmc_sweep_out *out, int lanes_per_point, void *stream);

/* Normalised, reweighted ln(PI) rows: lnpi_out[s][i] = fl(lnPI'_s(i) - lnnorm[s]); what
 * histogram.reweight()/normalize() leave in data['ln(PI)'] (GH:67, 77). */
int fhmc_lnpi_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                 const double *lnnorm, double *lnpi_out, void *stream);

/*
 * K2 for the drop-in thermo(props=True): phase averages of EVERY moment array (GH:530-541).
 *   lnpi  [n]            normalised ln(PI)
 *   mom   [n_arrays][n]  moment tensor flattened over (i,j,k,m,p)
 *   bounds[n_phase][2]
 *   avg   [n_phase][n_arrays]   out (may be NULL when n_arrays == 0)
 *   lnsum [n_phase]             out: ln sum_{j in phase} exp(lnpi_j)  (F.E./kT = -(lnsum - lnpi[0]), GH:523-526)
 */
int fhmc_phase_moments(const double *lnpi, int n, const double *mom, int n_arrays, const int *bounds,
                       int n_phase, double *avg, double *lnsum, void *stream);
/* The same with the phase count still on the device: `bounds`, `status`, `nphase` point into the record of a sweep
 * (fhmc_sweep_out of ONE state point, capacity pmax); phases >= *nphase (or all, if the status code is not FHMC_OK) are skipped. */
int fhmc_phase_moments_dev(const double *lnpi, int n, const double *mom, int n_arrays, const int *bounds, const unsigned *status,
                           const int *nphase, int pmax, double *avg, double *lnsum, void *stream);

/*
 * One scalar drop-in call (histogram.reweight / normalize / relextrema / thermo for ONE state point, GH:260-289, 317-415,
 * 451-554) as ONE host call: uploads what changed on the host (lnpi_host / ntot_host non-NULL: pinned rows of desc->n
 * doubles copied into blob rows 0 / 1), sets the target mu_1, runs fhmc_sweep_1d (one warp, general evaluator),
 * fhmc_lnpi_1d (want_row) and fhmc_phase_moments_dev (want_moments: averages of the n_arrays rows of `mom` over the phases
 * of the record), copies the contiguous device range [out_dev, out_dev + out_bytes) -- which the caller lays out to hold
 * rec, row, avg and lnsum -- to out_host (pinned) and synchronises `stream`.
 */
typedef struct fhmc_scalar_io {
    double *blob;              /* device [>= 2][n_pad]: row 0 ln(PI), row 1 N (+ whatever desc describes)   */
    double *mu1_dev;           /* device [1]                                                               */
    double *mu1_pinned;        /* pinned host [1]: staging of mu1                                          */
    double mu1;                /* target mu_1 of the state point                                           */
    const double *lnpi_host;   /* pinned host [n] or NULL (blob row 0 is current)                          */
    const double *ntot_host;   /* pinned host [n] or NULL (blob row 1 is current)                          */
    fhmc_sweep_out rec;        /* device record of one state point, capacity desc->pmax                    */
    double *row;               /* device [n]: normalised reweighted ln(PI)            (want_row)            */
    const double *mom;         /* device [n_arrays][n] or NULL                        (want_moments)        */
    int n_arrays;
    double *avg;               /* device [pmax][n_arrays]                                                   */
    double *lnsum;             /* device [pmax]                                                             */
    void *out_dev;             /* device range that holds rec / row / avg / lnsum ...                       */
    void *out_host;            /* ... and its pinned host mirror                                            */
    size_t out_bytes;
} fhmc_scalar_io;
int fhmc_scalar_point(const fhmc_hist_desc *desc, const fhmc_scalar_io *io, int want_row, int want_moments, void *stream);


/*
 * Phase-major repack of sweep records for the trip to the host (new; no reference counterpart).  fhmc_sweep_out keeps
 * pmax slots per state point; most state points have one or two phases, so most of fe/avg/bounds is padding that need
 * not cross PCIe.  After this call the fields of phase p of ALL state points are contiguous:
 *   packed (bytes, S = n_states, R = 16 + 8*n_sel):
 *     { u32 status; i32 nphase; }[S]  |  for p in 0..pmax-1: { f64 fe; f64 avg[n_sel]; i32 bounds[2]; }[S]
 *   slots p >= nphase[s] (or every slot when the status code is not FHMC_OK) hold NaN / -1.
 *   max_nphase: device int the caller zeroes; raised to max_s nphase[s] -> the first 8*S + max_nphase*R*S bytes are
 *   all that needs to be copied.
 * packed must be 16-byte aligned and hold fhmc_pack_bytes(n_states, pmax, n_sel) = 8*S + pmax*R*S bytes.
 */
long long fhmc_pack_bytes(long long n_states, int pmax, int n_sel);
int fhmc_pack_phase_major(const fhmc_sweep_out *out, long long n_states, int pmax, int n_sel, void *packed,
                          int *max_nphase, void *stream);

/*
 * Narrow variant of the repack: 4 + P (8 + 8 n_sel + 4) bytes per state point with P live phases.
 *   packed: { u16 status (FHMC_ST_* fit 15 bits); u8 nphase; u8 path (diagnostic: 1 = written by the tilt cells, else 0) }[S], padded to 16 bytes  |
 *           for p in 0..pmax-1: { f64 fe; f64 avg[n_sel]; }[S]  |  for p in 0..pmax-1: { i16 bounds[2]; }[S]
 * (bin indices must fit int16: n <= 32767).  Slots p >= nphase[s] hold NaN / -1 as in fhmc_pack_phase_major.
 */
long long fhmc_pack_soa16_bytes(long long n_states, int pmax, int n_sel);
int fhmc_pack_phase_soa16(const fhmc_sweep_out *out, long long n_states, int pmax, int n_sel, void *packed,
                          int *max_nphase, void *stream);

/*
 * K1+K3+K2 with COMPACT records (new; the result set {status, nphase, fe, avg, bounds} of the reference's reweight()/thermo()/
 * is_safe() loop, GH:268-289, 451-596, for many state points): every state point of `states` is written as a phase-major
 * narrow record (layout of fhmc_pack_phase_soa16 for n_total records) at index first + s, by the sweep kernel itself, into
 * EVERY destination buffer.  With n_dst == 1 this is the sweep + repack of the host pipeline in one kernel; with the peers'
 * buffers in dst[] (pointers a process obtained by mapping its NVLink peers' memory) it is the sweep fused with the final
 * gather of a sharded sweep: no collective afterwards, only a barrier.
 *   fill_dead: also write NaN / -1 into the phase slots >= nphase (otherwise the caller has filled the buffers with 0xFF
 *   bytes, which read as NaN / -1);  max_nphase: nullable device int, raised to the largest phase count (caller zeroes it).
 * workspace: device, 256-byte aligned, fhmc_sweep_compact_workspace() bytes.  desc->pmax <= 8 and desc->n <= 32767.
 */
typedef struct fhmc_compact_out {
    void *dst[8];
    int n_dst;
    long long n_total;
    long long first;
    int fill_dead;
    int *max_nphase;
} fhmc_compact_out;
size_t fhmc_sweep_compact_workspace(const fhmc_hist_desc *desc, long long n_states);
/*
 * Per-histogram tables for fhmc_hist_desc.mu_tables (new; the reference recomputes relextrema()/thermo() bounds from
 * scratch at every mu, GH:317-415, 498-520 -- on uniformly spaced N they are piecewise constant in mu and change only at
 * chord slopes of ln(PI), so they are evaluated once per interval by the general evaluator and looked up afterwards).
 * desc must describe a pure mu sweep in product form (mu_recurrence >= 2, hull rows present, n_coef == 0, n_term <= 1,
 * n_sel <= 2, n <= 32767); desc->mu_tables is ignored.  tables: device, 256-byte aligned, fhmc_mu_tables_bytes() bytes
 * (0 = not applicable).  Asynchronous on `stream`; returns 0 ok, 1 error, 2 not applicable for this descriptor.
 */
size_t fhmc_mu_tables_bytes(const fhmc_hist_desc *desc);
int fhmc_mu_tables_build(const fhmc_hist_desc *desc, const double *blob, void *tables, size_t tables_bytes, void *stream);
int fhmc_sweep_1d_compact(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states, const fhmc_compact_out *out,
                          void *workspace, size_t workspace_bytes, void *stream);
/*
 * Tilt cells for fhmc_hist_desc.mu_cells (new; see the field): moment expansions of the per-phase sums of reweight() + thermo()
 * (GH:71-78, 498-554) about the centres of small cells of the tilt, for the state points with mu1 in [mu_lo, mu_hi].  desc must
 * carry the mu_tables the cells refine (same blob); desc->mu_cells is ignored.  cells: device, 256-byte aligned,
 * fhmc_mu_cells_bytes(desc, extra_pieces) bytes -- room for one cell per elementary interval plus extra_pieces (a one-phase
 * interval needs a cell per 0.2 / n of tilt; intervals that do not fit any more are left to the table walk).  Asynchronous on
 * `stream`; returns 0 ok, 1 error, 2 not applicable.  The buffer belongs to exactly this blob / tables combination.
 */
size_t fhmc_mu_cells_bytes(const fhmc_hist_desc *desc, int extra_pieces);
int fhmc_mu_cells_build(const fhmc_hist_desc *desc, const double *blob, void *cells, size_t cells_bytes, int extra_pieces,
                        double mu_lo, double mu_hi, void *stream);
/* the same for the range of the finite values of a device array of mu1 (read on the device: nothing synchronises) */
int fhmc_mu_cells_build_for(const fhmc_hist_desc *desc, const double *blob, void *cells, size_t cells_bytes, int extra_pieces,
                            const double *mu_dev, long long n_mu, void *stream);

/*
 * Host-buffer mu sweep (new; replaces the user's Python loop of reweight()/thermo()/is_safe() calls on host arrays,
 * GH:268-289, 451-596, README.md:60-85).  mu_host[n_states] and out_host are PINNED host memory; everything in between is
 * pipelined on three private streams (upload, compute, download) in chunks of `chunk` state points: H2D(mu) -> fhmc_sweep_1d -> fhmc_pack_phase_major
 * -> D2H of the head and of the phase blocks that exist.  Returns when the results are in out_host, laid out as
 * fhmc_pack_phase_major describes for S = n_states (phase blocks >= *max_nphase_out are not written).
 *   desc->pmax, desc->n_sel size the records; workspace: device, 256-byte aligned, fhmc_sweep_host_workspace() bytes
 *   flags_host: pinned, one int per chunk (receives max nphase of the chunk)
 *   guess_nphase: phase blocks to copy speculatively per chunk (a chunk that needs more is topped up later)
 *   stream: work already queued there (e.g. the blob upload) is waited for before the first chunk starts
 */
size_t fhmc_sweep_host_workspace(long long chunk, int pmax, int n_sel);
int fhmc_sweep_host_compact(const fhmc_hist_desc *desc, const double *blob, const double *mu_host, long long n_states,
                            int lanes_per_point, long long chunk, void *workspace, size_t workspace_bytes,
                            void *out_host, int *flags_host, int guess_nphase, int *max_nphase_out,
                            long long *d2h_bytes_out, void *stream);
/* the same with the records of fhmc_pack_phase_soa16 in out_host (fhmc_pack_soa16_bytes(n_states, ...) bytes) */
int fhmc_sweep_host_compact16(const fhmc_hist_desc *desc, const double *blob, const double *mu_host, long long n_states,
                              int lanes_per_point, long long chunk, void *workspace, size_t workspace_bytes,
                              void *out_host, int *flags_host, int guess_nphase, int *max_nphase_out,
                              long long *d2h_bytes_out, void *stream);

/*
 * Pointwise Taylor update of a stack of arrays (moment extrapolation, GH:1027-1034 / 1162-1171,
 * and histogram.mix, GH:244-252):  out[a][i] = sum_t w[t] * src[t][a][i],  t < n_src.
 */
int fhmc_axpy_rows(const double *const *src, const double *w_host, int n_src, long long count,
                   double *out, void *stream);

/*
 * K4: batched find_phase_eq (GH:598-668 with the objective of GH:2570-2630).  One solve per
 * entry of `states` (mu1 there is the initial guess).  Root of the signed dF.E./kT between the
 * two phases the reference objective would select, by bracketing + bisection/Newton, to
 * |dF.E.| <= lnz_tol.  Outputs: mu_coex[T], dfe[T] (signed residual), iters[T], status as above
 * (code FHMC_E_* or 100 = no two wide phases / no bracket).  A solve whose bracket closed on a JUMP of
 * dF.E. (integer phase bounds moving with mu; no root exists at fp64 resolution) reports code 0 with
 * FHMC_ST_JUMP set and the residual in dfe.  The thermo at mu_coex is written through `out` (same
 * layout as fhmc_sweep_1d).
 */
#define FHMC_E_NO_COEX 100
int fhmc_find_phase_eq_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                          double lnz_tol, double mu_step, int max_iter,
                          double *mu_coex, double *dfe, int *iters,
                          const fhmc_sweep_out *out, void *stream);
/* The same for a coexistence CURVE: the flat list of solves is ordered along the curve (beta monotone).  Every
 * seed_stride-th solve and the last one are "seeds", solved from their own mu_guess; every other solve starts from the
 * linear interpolation in beta of the roots of the two seeds around it (a seed that did not converge is left out; with no
 * usable seed the solve's own guess is used) -- what a notebook does when it feeds the previous temperature's mu into the
 * next find_phase_eq call (GH:598-668 needs a good guess), done inside ONE launch: seeds are handed out first, the other
 * solves wait for their two seeds only.  Same outputs as fhmc_find_phase_eq_1d (roots agree to lnz_tol; deterministic). */
int fhmc_find_phase_eq_curve(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states,
                             double lnz_tol, double mu_step, int max_iter, int seed_stride, double *mu_coex, double *dfe,
                             int *iters, const fhmc_sweep_out *out, void *stream);

/*
 * K5: 2-D joint histogram lnPI(op1,op2) (container: two_dim/joint_hist.pyx:201-247) reweighted to
 * S state points:  v = lnPI[i][j] + a1[s]*op1[i] + a2[s]*op2[j]  over the support
 * [bounds[i][0], bounds[i][1]) of each row (-inf padding ignored);
 * out[s][0] = ln sum exp v, out[s][1] = <op1>, out[s][2] = <op2>, out[s][3+q] = <prop_q>  (n_prop <= 2).
 * `workspace` (device, fhmc_reweight_2d_workspace() bytes) holds the per-row-chunk partial sums.
 */
size_t fhmc_reweight_2d_workspace(int n1, int n2, int n_prop, long long n_states);
int fhmc_reweight_2d(const double *lnpi, const int *bounds, int n1, int n2, const double *op1,
                     const double *op2, const double *props, int n_prop, const double *a1,
                     const double *a2, long long n_states, double *out, double *workspace,
                     size_t workspace_bytes, void *stream);

/*
 * K5 in product form: same outputs as fhmc_reweight_2d.  The caller promises that op2 is uniformly spaced and that
 * max_s |a2[s]| * (op2[n2-1] - op2[0]) < 300; then exp(lnPI_ij + a2 op2_j) factorises into a tabulated exp(lnPI_ij - A) and
 * a geometric sequence along the row, the sums of a 4-bin block become Horner polynomials in exp(a2 d2) and a true exp
 * is only evaluated once per 128-bin row segment.  workspace additionally holds the tables (built by the call).
 */
size_t fhmc_reweight_2d_prod_workspace(int n1, int n2, int n_prop, long long n_states);
int fhmc_reweight_2d_prod(const double *lnpi, const int *bounds, int n1, int n2, const double *op1,
                          const double *op2, const double *props, int n_prop, const double *a1,
                          const double *a2, long long n_states, double *out, double *workspace,
                          size_t workspace_bytes, void *stream);

/*
 * Ragged / masked 2-D log-sum-exp with averages, one surface per call (HBM-bound streaming reduction):
 *   pore_hist.normalize (two_dim/h_ntot/pore_hist.pyx:57-80, 147-152): edge[i] = last valid column of row i, mask NULL,
 *       lnpi_shifted (nullable) receives lnPI - ln sum exp lnPI;
 *   pore_hist.thermo(mask) (pore_hist.pyx:154-184): mask[n1][n2] (1 = bin belongs to the phase), props [n_prop][n1][n2].
 * out[0] = ln sum_sel exp lnPI, out[1] = max_sel lnPI, out[2+q] = <prop_q>;  peak[0] = number of selected bins equal to
 * the maximum, peak[1..peak_cap] their flat indices (unordered; np.where(lp == np.max(lp)), PH:182).  n_prop <= 8.
 */
size_t fhmc_masked_lse_2d_workspace(int n1, int n2, int n_prop);
int fhmc_masked_lse_2d(const double *lnpi, const unsigned char *mask, const int *edge, int n1, int n2,
                       const double *props, int n_prop, double *out, long long *peak, int peak_cap,
                       double *lnpi_shifted, double *workspace, size_t workspace_bytes, void *stream);

/*
 * Window patching shift solve, batched (moments/win_patch/fhmc_patch.pyx:640-709): pair w owns a[offsets[w]:offsets[w+1]]
 * (the upper window's overlap slice) and b[...] (the lower window's); shift[w] = argmin_x sum ((a_i + x) - b_i)^2 =
 * mean(b - a) (the reference finds it with scipy.optimize.fmin), err2[w] = sum ((a_i + shift) - b_i)^2 / len.
 */
int fhmc_patch_shifts(const double *a, const double *b, const long long *offsets, int n_pairs, double *shift,
                      double *err2, void *stream);

/* Diagnostic counters of the lean warp-per-point evaluator (K4), out24 = 24 values: [0] evaluations it finished, [1..7]
 * evaluations handed to the general evaluator (too many candidates / repair needs the normalised array or the reference
 * raises / phase count / underflowing phase / failed re-test on the normalised values / monotone ln(PI) with ties / monotone
 * ln(PI) that is not one phase); [8..23] cycle counters, zero unless built with -DFHMC_LEAN_PROFILE.  Synchronises the device. */
int fhmc_lean_stats(unsigned long long *out24, int reset);

/* Roofline micro-benchmarks (register-resident): return ops executed; time with CUDA events. */
long long fhmc_bench_dfma(int iters, double *sink, void *stream);
long long fhmc_bench_exp(int iters, double *sink, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FHMC_B200_H */
