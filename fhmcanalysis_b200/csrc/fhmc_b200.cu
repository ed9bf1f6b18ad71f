// fhmc_b200.cu -- kernels + C ABI of libfhmc_b200.so (sm_100a only).
//
//   K1+K3+K2  k_sweep_1d        fused reweight / normalise / phase split / thermo / is_safe sweep
//   K1        k_lnpi_1d         normalised reweighted ln(PI) rows (what reweight() leaves behind)
//   K2        k_phase_moments   per-phase averages of every moment array (drop-in thermo(props=True))
//             k_axpy_rows       pointwise Taylor / mixing update of array stacks
//   roofline  k_bench_dfma / k_bench_exp
// K4 (solver) and K5 (2-D) live in fhmc_solver.cu / fhmc_2d.cu.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "fhmc_point.cuh"
#include "fhmc_fast.cuh"

namespace fhmc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local const char *g_last_kernel = "";
void note_kernel(const char *name) { g_last_kernel = name; }

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
}

// ---------------------------------------------------------------------------------------------
// state point s -> (mu1, beta, dmu)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_state(const SweepArgs &a, long long sp, double &mu1, double &beta, double &dmu)
{
    const fhmc_states &st = a.st;
    mu1 = st.mu1[(sp / st.mu1_div) % st.n_mu1];
    beta = st.beta ? st.beta[(sp / st.beta_div) % st.n_beta] : a.d.beta_ref;
    dmu = st.dmu ? st.dmu[(sp / st.dmu_div) % st.n_dmu] : a.d.dmu_ref;
}

// ---------------------------------------------------------------------------------------------
// K1+K3+K2: persistent CTAs; the histogram blob is staged ONCE per CTA by a TMA bulk copy, then the
// CTA walks tiles of FHMC_CTA/G state points.
// ---------------------------------------------------------------------------------------------
template <int G, bool TAYLOR, int CTA>
__global__ void __launch_bounds__(CTA) k_sweep_1d(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_tab;
    const double *sm = stage_histogram(a, smem_raw, s_tab);

    constexpr int GPC = CTA / G;  // state points per CTA tile
    const int grp = threadIdx.x / G;
    const long long S = a.st.n_states;
    const long long ntiles = (S + GPC - 1) / GPC;
    PointEval<G, TAYLOR> pe(a, sm, threadIdx.x & 31, s_tab);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long sp = tile * GPC + grp;
        if (sp >= S) continue;  // whole groups drop out together; collectives use the group mask
        double mu1, beta, dmu;
        load_state(a, sp, mu1, beta, dmu);
        pe.setup(mu1, beta, dmu);
        pe.run(sp);
    }
}

// ---------------------------------------------------------------------------------------------
// K1 output: lnpi_out[s][i] = fl(u_s(i) - lnnorm[s]).  One CTA per (state point, chunk of bins):
// coalesced 8-byte stores, HBM-write bound.
// ---------------------------------------------------------------------------------------------
template <bool TAYLOR>
__global__ void __launch_bounds__(FHMC_CTA) k_lnpi_1d(const __grid_constant__ SweepArgs a, const double *__restrict__ lnnorm,
                                                      double *__restrict__ lnpi_out)
{
    const int n = a.d.n, npad = a.d.n_pad;
    const long long S = a.st.n_states;
    for (long long sp = blockIdx.y; sp < S; sp += gridDim.y) {
        double mu1, beta, dmu;
        load_state(a, sp, mu1, beta, dmu);
        const double s = __dmul_rn(__dsub_rn(mu1, a.d.mu1_ref), a.d.beta_ref);
        double xi[FHMC_MAX_TERMS];
        if (TAYLOR) {
            const double dB = beta - a.d.beta_ref, dD = dmu - a.d.dmu_ref;
#pragma unroll
            for (int t = 0; t < FHMC_MAX_TERMS; ++t) xi[t] = (t < a.d.n_coef) ? monomial(a.d.coef_kind[t], dB, dD, mu1) : 0.0;
        }
        const double c = lnnorm[sp];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            double u = __dadd_rn(a.blob[i], __dmul_rn(s, a.blob[npad + i]));
            if (TAYLOR) {
#pragma unroll
                for (int t = 0; t < FHMC_MAX_TERMS; ++t)
                    if (t < a.d.n_coef) u = fma(xi[t], a.blob[(size_t)a.d.coef_row[t] * npad + i], u);
            }
            lnpi_out[sp * n + i] = __dsub_rn(u, c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2 (drop-in): avg[p][arr] = sum_{j in phase p} exp(lnpi_j) mom[arr][j] / sum_{j in phase p} exp(lnpi_j)
// One warp per (phase, array); rows are read once from HBM with coalesced 8-byte loads.
// exp(lnpi_j) of a normalised distribution is <= 1; the sums are shifted by the phase maximum so
// that phases of negligible total weight keep full relative precision.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_phase_moments(const double *__restrict__ lnpi, int n, const double *__restrict__ mom,
                                                       int n_arrays, const int *__restrict__ bounds, int n_phase,
                                                       double *__restrict__ avg, double *__restrict__ lnsum,
                                                       const unsigned *__restrict__ status_dev, const int *__restrict__ nphase_dev)
{
    __shared__ double s_tab[64];
    stage_exp_table(s_tab);
    __syncthreads();
    if (nphase_dev) {   // bounds are the record of a sweep that is still on the device: its phase count, none if it raised
        const int P = ((*status_dev & FHMC_ST_CODE_MASK) == FHMC_OK) ? *nphase_dev : 0;
        n_phase = min(n_phase, max(P, 0));
    }
    const uint32_t tab = smem_u32(s_tab);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int per_phase = n_arrays + 1;  // job `n_arrays` of each phase is the ln-sum itself
    for (int job = warp; job < n_phase * per_phase; job += nwarps) {
        const int p = job / per_phase, arr = job % per_phase;
        const int left = bounds[2 * p], right = bounds[2 * p + 1];
        double mx = -CUDART_INF;
        for (int j = left + lane; j < right; j += 32) mx = fmax(mx, lnpi[j]);
        mx = group_max<32>(mx, 0xffffffffu);
        double S = 0.0, A = 0.0;
        if (arr == n_arrays) {
            for (int j = left + lane; j < right; j += 32) S += exp_nonpos(lnpi[j] - mx, tab);
            S = group_sum<32>(S, 0xffffffffu);
            if (lane == 0 && lnsum) lnsum[p] = (right > left) ? mx + log(S) : -1.7976931348623157e308;
            continue;
        }
        const double *row = mom + (size_t)arr * n;
        for (int j = left + lane; j < right; j += 32) {
            const double e = exp_nonpos(lnpi[j] - mx, tab);
            S += e;
            A = fma(e, row[j], A);
        }
        S = group_sum<32>(S, 0xffffffffu);
        A = group_sum<32>(A, 0xffffffffu);
        if (lane == 0) avg[(size_t)p * n_arrays + arr] = A / S;
    }
}

// out[i] = sum_t w[t] * src[t][i]
struct AxpyArgs {
    const double *src[FHMC_MAX_TERMS];
    double w[FHMC_MAX_TERMS];
    int n_src;
};
__global__ void __launch_bounds__(256) k_axpy_rows(const __grid_constant__ AxpyArgs a, long long count, double *__restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        double v = a.w[0] * a.src[0][i];
#pragma unroll
        for (int t = 1; t < FHMC_MAX_TERMS; ++t)
            if (t < a.n_src) v = fma(a.w[t], a.src[t][i], v);
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// roofline micro-benchmarks: register-resident fp64 FMA chains / exp_nonpos chains
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bench_dfma(int iters, double *sink)
{
    double x0 = 1.0 + threadIdx.x * 1e-9, x1 = 0.5, x2 = 0.25, x3 = 0.125, x4 = 1.5, x5 = 1.25, x6 = 1.125, x7 = 0.75;
    const double a = 0.999999, b = 1e-7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == 123.456) sink[0] = r;
}

__global__ void __launch_bounds__(256) k_bench_exp(int iters, double *sink)
{
    __shared__ double s_tab[64];
    stage_exp_table(s_tab);
    __syncthreads();
    const uint32_t tab = smem_u32(s_tab);
    double t0 = -1e-3 * (threadIdx.x + 1), t1 = t0 - 0.3, t2 = t0 - 1.7, t3 = t0 - 11.0;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (int i = 0; i < iters; ++i) {
        acc0 += exp_nonpos(t0, tab); acc1 += exp_nonpos(t1, tab); acc2 += exp_nonpos(t2, tab); acc3 += exp_nonpos(t3, tab);
        t0 -= 1e-6; t1 -= 1e-6; t2 -= 1e-6; t3 -= 1e-6;
    }
    const double r = acc0 + acc1 + acc2 + acc3;
    if (r == 123.456) sink[0] = r;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DevInfo {
    int sm_count = 0, smem_optin = 0, ok = 0;
};
static DevInfo g_dev[64];

static const DevInfo *dev_info()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    DevInfo &d = g_dev[dev];
    if (!d.ok) {
        if (cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
        if (cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return nullptr;
        d.ok = 1;
    }
    return &d;
}

static int validate_desc(const fhmc_hist_desc *d)
{
    if (!d) { set_error("null descriptor"); return 1; }
    if (d->n < 1 || d->n_pad < d->n || (d->n_pad & 1)) { set_error("bad n/n_pad (%d/%d): n_pad must be even and >= n", d->n, d->n_pad); return 1; }
    if (d->n_rows < 2) { set_error("blob needs at least the ln(PI) and N rows"); return 1; }
    if (d->n_coef < 0 || d->n_coef > FHMC_MAX_TERMS || d->n_term < 0 || d->n_term > FHMC_MAX_TERMS) { set_error("too many Taylor terms"); return 1; }
    if (d->n_sel < 0 || d->n_sel > FHMC_MAX_SEL) { set_error("n_sel must be in [0,%d]", FHMC_MAX_SEL); return 1; }
    if (d->n_sel > 0 && d->n_term < 1) { set_error("n_term must be >= 1 when n_sel > 0"); return 1; }
    for (int c = 0; c < d->n_coef; ++c)
        if (d->coef_row[c] < 0 || d->coef_row[c] >= d->n_rows) { set_error("coef_row[%d] out of range", c); return 1; }
    for (int q = 0; q < d->n_sel; ++q)
        if (d->sel_row[q] < 0 || d->sel_row[q] + d->n_term > d->n_rows) { set_error("sel_row[%d] out of range", q); return 1; }
    if (!d->complete && d->smooth < 1) { set_error("smooth must be >= 1 (scipy argrelextrema order)"); return 1; }
    if (d->pmax < 1) { set_error("pmax must be >= 1"); return 1; }
    return 0;
}

static int validate_states(const fhmc_states *st)
{
    if (!st || st->n_states < 0 || !st->mu1 || st->n_mu1 < 1 || st->mu1_div < 1) { set_error("bad state-point description"); return 1; }
    if (st->beta && (st->n_beta < 1 || st->beta_div < 1)) { set_error("bad beta array description"); return 1; }
    if (st->dmu && (st->n_dmu < 1 || st->dmu_div < 1)) { set_error("bad dmu array description"); return 1; }
    return 0;
}

template <int G, bool TAYLOR, int CTA>
static int launch_sweep_cta(const SweepArgs &args, size_t smem, const DevInfo *di, cudaStream_t stream, int *occ_out, bool dry)
{
    auto kern = k_sweep_1d<G, TAYLOR, CTA>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA, smem), "occupancy query")) return 1;
    *occ_out = occ;
    if (dry) return 0;
    if (occ < 1) { set_error("sweep kernel does not fit on an SM (smem %zu bytes)", smem); return 1; }
    const long long gpc = CTA / G;
    const long long ntiles = (args.st.n_states + gpc - 1) / gpc;
    long long grid = (long long)di->sm_count * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, CTA, smem, stream>>>(args);
    note_kernel(G == 32 ? "k_sweep_1d<32>" : (G == 4 ? "k_sweep_1d<4>" : "k_sweep_1d<1>"));
    return check_cuda(cudaGetLastError(), "k_sweep_1d launch");
}

template <int G, bool TAYLOR>
static int launch_sweep(const SweepArgs &args, size_t smem, const DevInfo *di, cudaStream_t stream)
{
    // a histogram blob that leaves room for one 256-thread CTA per SM only (Taylor sweeps at 2001 bins) runs as one
    // 512-thread CTA instead: twice the warps to hide the fp64 latencies behind
    int occ = 0;
    if (launch_sweep_cta<G, TAYLOR, FHMC_CTA>(args, smem, di, stream, &occ, true)) return 1;
    const long long tiles512 = (args.st.n_states + 512 / G - 1) / (512 / G);
    if (occ == 1 && tiles512 >= di->sm_count) {
        int occ2 = 0;
        if (launch_sweep_cta<G, TAYLOR, 512>(args, smem, di, stream, &occ2, true)) return 1;
        if (occ2 >= 1) return launch_sweep_cta<G, TAYLOR, 512>(args, smem, di, stream, &occ2, false);
    }
    return launch_sweep_cta<G, TAYLOR, FHMC_CTA>(args, smem, di, stream, &occ, false);
}

template <int G>
static int launch_sweep_t(bool taylor, const SweepArgs &args, size_t smem, const DevInfo *di, cudaStream_t stream)
{
    return taylor ? launch_sweep<G, true>(args, smem, di, stream) : launch_sweep<G, false>(args, smem, di, stream);
}

// one-pass mu-sweep kernel (fhmc_fast.cuh, NC = 0): one thread per state point
static int launch_fast_mu(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    if (args.d.mu_recurrence >= 2) {
        const int rc = launch_fast_mu_prod(args, sm_count, smem_optin, stream);   // fhmc_fast_prod.cu
        if (rc >= 0) return rc;
    }
    if (args.d.mu_recurrence) {
        const int rc = launch_fast_mu_rec(args, sm_count, smem_optin, stream);   // fhmc_fast_rec.cu
        if (rc >= 0) return rc;
    }
    const bool s0n = args.d.n_sel > 0 && args.d.sel_row[0] == 1;
    switch (args.d.n_sel) {
    case 0: return launch_fast<0, false, 0, 1>(args, sm_count, smem_optin, stream);
    case 1: return s0n ? launch_fast<1, true, 0, 1>(args, sm_count, smem_optin, stream) : launch_fast<1, false, 0, 1>(args, sm_count, smem_optin, stream);
    case 2: return s0n ? launch_fast<2, true, 0, 1>(args, sm_count, smem_optin, stream) : launch_fast<2, false, 0, 1>(args, sm_count, smem_optin, stream);
    case 3: return s0n ? launch_fast<3, true, 0, 1>(args, sm_count, smem_optin, stream) : launch_fast<3, false, 0, 1>(args, sm_count, smem_optin, stream);
    default: return s0n ? launch_fast<4, true, 0, 1>(args, sm_count, smem_optin, stream) : launch_fast<4, false, 0, 1>(args, sm_count, smem_optin, stream);
    }
}


// ---------------------------------------------------------------------------------------------
// Phase-major repack of the sweep records (see fhmc_pack_phase_major in the header).  Pointwise, HBM bound:
// reads one record (<= 4+4+pmax*(8+8*nsel+8) B) and writes as much per state point, all stores coalesced.
// ---------------------------------------------------------------------------------------------
struct PackArgs {
    fhmc_sweep_out out;
    long long S;
    int pmax, nsel;
    unsigned char *packed;
    int *max_nphase;
    CompactArgs c;   // k_pack_phase_soa16: destinations (c.n_dst >= 1), layout size c.n_total, first record c.first
};

__global__ void __launch_bounds__(256) k_pack_phase_major(const __grid_constant__ PackArgs a)
{
    int2 *head = reinterpret_cast<int2 *>(a.packed);
    unsigned char *base = a.packed + 8 * a.S;
    const int rec = 16 + 8 * a.nsel;
    int pm = 0;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < a.S; s += (long long)gridDim.x * blockDim.x) {
        const unsigned st = a.out.status[s];
        const int P = a.out.nphase[s];
        head[s] = make_int2((int)st, P);
        const int Pe = ((st & FHMC_ST_CODE_MASK) == FHMC_OK) ? min(max(P, 0), a.pmax) : 0;
        pm = max(pm, Pe);
        for (int p = 0; p < a.pmax; ++p) {
            double *r = reinterpret_cast<double *>(base + ((long long)p * a.S + s) * rec);
            const bool live = p < Pe;
            r[0] = live ? a.out.fe[s * a.pmax + p] : CUDART_NAN;
            for (int q = 0; q < a.nsel; ++q) r[1 + q] = live ? a.out.avg[(s * a.pmax + p) * a.nsel + q] : CUDART_NAN;
            const int2 b = live ? *reinterpret_cast<const int2 *>(a.out.bounds + (s * a.pmax + p) * 2) : make_int2(-1, -1);
            *reinterpret_cast<int2 *>(r + 1 + a.nsel) = b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pm = max(pm, __shfl_xor_sync(0xffffffffu, pm, o));
    if ((threadIdx.x & 31) == 0 && pm > 0) atomicMax(a.max_nphase, pm);
}

// Narrow variant of the repack (fhmc_pack_phase_soa16): 4-byte head {u16 status, u8 nphase, u8 0}, then per phase the fp64
// fields {fe, avg[nsel]}[S] and, in a separate array, the bounds as int16 pairs -- 4 + P (8 + 8 nsel + 4) bytes per state
// point instead of 8 + P (16 + 8 nsel).
__global__ void __launch_bounds__(256) k_pack_phase_soa16(const __grid_constant__ PackArgs a)
{
    const int nf = 1 + a.nsel;
    const long long NT = a.c.n_total;
    int pm = 0;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < a.S; s += (long long)gridDim.x * blockDim.x) {
        const unsigned st = a.out.status[s];
        const int P = a.out.nphase[s];
        const uchar4 hd = make_uchar4((unsigned char)(st & 0xFFu), (unsigned char)((st >> 8) & 0xFFu), (unsigned char)min(max(P, 0), 255), 0);
        const int Pe = ((st & FHMC_ST_CODE_MASK) == FHMC_OK) ? min(max(P, 0), a.pmax) : 0;
        pm = max(pm, Pe);
        const long long g = a.c.first + s;
        for (int d = 0; d < a.c.n_dst; ++d) {
            unsigned char *base = a.c.dst[d];
            reinterpret_cast<uchar4 *>(base)[g] = hd;
            double *F = reinterpret_cast<double *>(base + ((4 * NT + 15) & ~15ll));
            short2 *B = reinterpret_cast<short2 *>(reinterpret_cast<unsigned char *>(F) + (long long)a.pmax * NT * nf * 8);
            const int pend = a.c.fill_dead ? a.pmax : Pe;
            for (int p = 0; p < pend; ++p) {
                double *r = F + ((long long)p * NT + g) * nf;
                const bool live = p < Pe;
                r[0] = live ? a.out.fe[s * a.pmax + p] : CUDART_NAN;
                for (int q = 0; q < a.nsel; ++q) r[1 + q] = live ? a.out.avg[(s * a.pmax + p) * a.nsel + q] : CUDART_NAN;
                const int2 b = live ? *reinterpret_cast<const int2 *>(a.out.bounds + (s * a.pmax + p) * 2) : make_int2(-1, -1);
                B[(long long)p * NT + g] = make_short2((short)b.x, (short)b.y);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pm = max(pm, __shfl_xor_sync(0xffffffffu, pm, o));
    if ((threadIdx.x & 31) == 0 && pm > 0 && a.max_nphase) atomicMax(a.max_nphase, pm);
}

// compact-record instantiations of the headline kernel (fhmc_fast_prod_compact.cu)
int launch_prod2_compact(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream, int *grid_out, bool dry);
// the same on per-histogram tables (fhmc_tab.cu); needs args.d.mu_tables
int launch_tab2_compact(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream, int *grid_out, bool dry);
// dense sweeps on tilt cells + the table walk over whatever the cells leave (fhmc_cell.cu); needs args.d.mu_cells and args.c.ix_*
int launch_cell_compact(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream);

// Taylor-extrapolated grids, rows combined per (mu_1, beta) (fhmc_rowc.cu)
int launch_rowc(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream);

#define FHMC_FAST_MIN_STATES 4096

int choose_lanes(long long n_states, int bins, const DevInfo *di)
{
    if (bins >= 512) {
        // measured on B200 at 1001 bins (scripts/probe_threshold.py): a warp per state point wins up to ~2.4e4 state
        // points, one lane per state point beyond; four lanes never win
        return n_states >= (long long)di->sm_count * 160 ? 1 : 32;
    }
    // short histograms: enough groups to give every SM >= 1024 busy threads, otherwise widen the groups
    const long long target = (long long)di->sm_count * 1024;
    if (n_states >= target) return 1;
    if (n_states * 4 >= target) return 4;
    return 32;
}

}  // namespace fhmc

using namespace fhmc;

extern "C" {

int fhmc_version(void) { return FHMC_ABI_VERSION; }

const char *fhmc_last_error(void) { return g_err; }

const char *fhmc_last_kernel(void) { return g_last_kernel; }

int fhmc_device_info(int *sm_count, int *max_smem_optin)
{
    const DevInfo *di = dev_info();
    if (!di) { set_error("no CUDA device"); return 1; }
    if (sm_count) *sm_count = di->sm_count;
    if (max_smem_optin) *max_smem_optin = di->smem_optin;
    return 0;
}

int fhmc_sweep_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states, const fhmc_sweep_out *out,
                  int lanes_per_point, void *stream)
{
    if (validate_desc(desc) || validate_states(states)) return 1;
    if (!blob || ((uintptr_t)blob & 15)) { set_error("blob must be a 16-byte aligned device pointer"); return 1; }
    if (!out || !out->status || !out->nphase || !out->nmin || !out->lnnorm || !out->fe || !out->bounds || !out->max_idx ||
        !out->min_idx || (desc->n_sel > 0 && !out->avg)) { set_error("missing output buffer"); return 1; }
    if (states->n_states == 0) return 0;
    const DevInfo *di = dev_info();
    if (!di) { set_error("no CUDA device"); return 1; }
    size_t smem = (size_t)desc->n_rows * desc->n_pad * 8 + 16 + 512;
    SweepArgs args;
    args.d = *desc;
    args.blob = blob;
    args.st = *states;
    args.out = *out;
    args.blob_global = 0;
    memset(&args.c, 0, sizeof(args.c));
    if (smem > (size_t)di->smem_optin) {  // histogram larger than shared memory: read the rows through L1/L2
        args.blob_global = 1;
        smem = 16 + 512;
    }
    const bool taylor = desc->n_coef > 0 || desc->n_term > 1;
    int G = lanes_per_point > 0 ? lanes_per_point : choose_lanes(states->n_states, desc->n, di);
    // the one-thread-per-point kernel has a latency floor of one walk over the bins (~0.15 ms at 1001 bins) and beats
    // the generic kernels from a few thousand state points on (scripts/probe_threshold.py)
    // (pure mu sweeps; the Taylor variant walks the bins twice with a true exp per bin: 4 ms for 10^4 state points at
    // 2001 bins against 0.6 ms for the warp-per-point kernel, so it waits for two full CTAs per SM)
    const long long fast_min = taylor ? (long long)di->sm_count * 512 : FHMC_FAST_MIN_STATES;
    const bool try_fast = lanes_per_point == 1 || (lanes_per_point == 0 && states->n_states >= fast_min);
    cudaStream_t s = (cudaStream_t)stream;
    // One thread per state point + one exp pass (fhmc_fast.cuh): large pure-mu sweeps with a precomputed hull, and
    // Taylor-extrapolated sweeps whose term pattern has an instantiation.  lanes_per_point = -1 forces the generic
    // one-lane kernel (tests compare the two).  -1 from the launchers means "not applicable": fall through.
    if (lanes_per_point == -1) G = 1;
    else if (try_fast && !desc->complete && desc->n >= 3) {
        int rc = -1;
        if (!taylor) {
            if (desc->hull_len >= 2 && desc->hull_row > 1 && desc->hull_row + 2 <= desc->n_rows)
                rc = launch_fast_mu(args, di->sm_count, di->smem_optin, s);
        } else {
            // (beta x dmu_2) grids with dmu_2 fastest: coefficient rows combined once per grid row (fhmc_rowc.cu)
            static const bool rowc_on = !(getenv("FHMC_ROWC") && getenv("FHMC_ROWC")[0] == '0');
            if (rowc_on) rc = launch_rowc(args, di->sm_count, di->smem_optin, s);
            if (rc < 0) rc = launch_fast_taylor(args, di->sm_count, di->smem_optin, s);
        }
        if (rc >= 0) return rc;
    }
    switch (G) {
    case 1: return launch_sweep_t<1>(taylor, args, smem, di, s);
    case 4: return launch_sweep_t<4>(taylor, args, smem, di, s);
    case 32: return launch_sweep_t<32>(taylor, args, smem, di, s);
    default: set_error("lanes_per_point must be 0, 1, 4 or 32"); return 1;
    }
}

// scratch records of a compact-record launch: one per warp that can be resident (fused path), or one per state point (general
// path: plain sweep into the scratch, then the narrow repack)
#define FHMC_COMPACT_SCRATCH_RECORDS 8192
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // extern "C"
namespace fhmc {
size_t carve_records(unsigned char *base, long long c, int pmax, int nsel, fhmc_sweep_out *o)
{
    size_t off = 0;
    auto take = [&](size_t nbytes) { unsigned char *p = base ? base + off : nullptr; off += al256(nbytes); return p; };
    fhmc_sweep_out t;
    t.status = reinterpret_cast<unsigned *>(take(4 * c));
    t.nphase = reinterpret_cast<int *>(take(4 * c));
    t.nmin = reinterpret_cast<int *>(take(4 * c));
    t.lnnorm = reinterpret_cast<double *>(take(8 * c));
    t.fe = reinterpret_cast<double *>(take(8 * c * pmax));
    t.avg = reinterpret_cast<double *>(take(8 * c * pmax * (nsel > 0 ? nsel : 1)));
    t.bounds = reinterpret_cast<int *>(take(8 * c * pmax));
    t.max_idx = reinterpret_cast<int *>(take(4 * c * pmax));
    t.min_idx = reinterpret_cast<int *>(take(4 * c * (pmax + 1)));
    if (o) *o = t;
    return off;
}
}  // namespace fhmc
static size_t cell_index_bytes(long long n_states) { return (((size_t)8 * (size_t)n_states + 255) & ~(size_t)255) + 256; }

extern "C" {

size_t fhmc_sweep_compact_workspace(const fhmc_hist_desc *desc, long long n_states)
{
    if (!desc || n_states < 0 || desc->pmax < 1) return 0;
    // the fused kernel needs FHMC_COMPACT_SCRATCH_RECORDS records; anything it does not cover needs a record per state point
    // (the same size threshold as fhmc_sweep_1d_compact: smaller sweeps take the general path, one scratch record per state point)
    const DevInfo *di = dev_info();
    const bool fused = desc->mu_recurrence >= 2 && desc->n_coef == 0 && desc->n_term <= 1 && !desc->complete && desc->n_sel <= 2 &&
                       desc->pmax <= FHMC_COMPACT_PMAX && desc->n <= 32767 && di && n_states > (long long)di->sm_count * 2 * FHMC_CTA;
    const long long c = (fused && n_states > FHMC_COMPACT_SCRATCH_RECORDS) ? FHMC_COMPACT_SCRATCH_RECORDS
                        : (n_states > FHMC_COMPACT_SCRATCH_RECORDS ? n_states : FHMC_COMPACT_SCRATCH_RECORDS);
    // + the index list of the state points the tilt cells leave to the table walk
    const size_t ix = (fused && desc->mu_tables && desc->mu_cells) ? cell_index_bytes(n_states) : 0;
    return carve_records(nullptr, c, desc->pmax, desc->n_sel, nullptr) + ix;
}

int fhmc_sweep_1d_compact(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states, const fhmc_compact_out *cout,
                          void *workspace, size_t workspace_bytes, void *stream)
{
    if (validate_desc(desc) || validate_states(states)) return 1;
    if (!blob || ((uintptr_t)blob & 15)) { set_error("blob must be a 16-byte aligned device pointer"); return 1; }
    if (!cout || cout->n_dst < 1 || cout->n_dst > FHMC_COMPACT_DST || cout->n_total < 1 || cout->first < 0 ||
        cout->first + states->n_states > cout->n_total) { set_error("bad compact-output description"); return 1; }
    for (int d = 0; d < cout->n_dst; ++d)
        if (!cout->dst[d] || ((uintptr_t)cout->dst[d] & 15)) { set_error("destination %d must be a 16-byte aligned device pointer", d); return 1; }
    if (desc->n > 32767) { set_error("histogram too long for int16 bounds"); return 1; }
    if (!workspace || ((uintptr_t)workspace & 255)) { set_error("workspace must be a 256-byte aligned device pointer"); return 1; }
    if (states->n_states == 0) return 0;
    const DevInfo *di = dev_info();
    if (!di) { set_error("no CUDA device"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    SweepArgs args;
    args.d = *desc;
    args.blob = blob;
    args.st = *states;
    args.blob_global = 0;
    memset(&args.c, 0, sizeof(args.c));
    for (int d = 0; d < cout->n_dst; ++d) args.c.dst[d] = static_cast<unsigned char *>(cout->dst[d]);
    args.c.n_dst = cout->n_dst;
    args.c.n_total = cout->n_total;
    args.c.first = cout->first;
    args.c.fill_dead = cout->fill_dead;
    args.c.max_nphase = cout->max_nphase;
    // fused path: the product-form kernel writes the records itself
    if (states->n_states > (long long)di->sm_count * 2 * FHMC_CTA && !states->beta && !states->dmu) {
        int grid = 0;
        const bool tab = desc->mu_tables != nullptr;
        int rc = tab ? launch_tab2_compact(args, di->sm_count, di->smem_optin, s, &grid, true) : -1;
        const bool use_tab = tab && rc == 0;
        if (!use_tab) rc = launch_prod2_compact(args, di->sm_count, di->smem_optin, s, &grid, true);
        if (rc == 0) {
            const long long need = (long long)grid * (FHMC_CTA / 32);
            if (need <= FHMC_COMPACT_SCRATCH_RECORDS &&
                carve_records(nullptr, FHMC_COMPACT_SCRATCH_RECORDS, desc->pmax, desc->n_sel, nullptr) <= workspace_bytes) {
                const size_t rec_bytes = carve_records(static_cast<unsigned char *>(workspace), FHMC_COMPACT_SCRATCH_RECORDS, desc->pmax,
                                                       desc->n_sel, &args.out);
                if (use_tab && desc->mu_cells && rec_bytes + cell_index_bytes(states->n_states) <= workspace_bytes) {
                    // dense sweep on tilt cells; what they leave goes through the table walk in the same call
                    unsigned char *ix = static_cast<unsigned char *>(workspace) + rec_bytes;
                    args.c.ix_list = reinterpret_cast<long long *>(ix);   // (its counter lives in the cells buffer)
                    rc = launch_cell_compact(args, di->sm_count, di->smem_optin, s);
                    if (rc >= 0) return rc;
                    args.c.ix_list = nullptr;
                    args.c.ix_count = nullptr;
                }
                rc = use_tab ? launch_tab2_compact(args, di->sm_count, di->smem_optin, s, &grid, false)
                             : launch_prod2_compact(args, di->sm_count, di->smem_optin, s, &grid, false);
                if (rc >= 0) return rc;
            }
        } else if (rc == 1) {
            return 1;
        }
    }
    // general path: any kernel of fhmc_sweep_1d into scratch records, then the narrow repack to every destination
    fhmc_sweep_out rec;
    if (carve_records(nullptr, states->n_states, desc->pmax, desc->n_sel, nullptr) > workspace_bytes) {
        set_error("workspace too small: need fhmc_sweep_compact_workspace() bytes");
        return 1;
    }
    carve_records(static_cast<unsigned char *>(workspace), states->n_states, desc->pmax, desc->n_sel, &rec);
    if (fhmc_sweep_1d(desc, blob, states, &rec, 0, stream)) return 1;
    PackArgs a;
    a.out = rec;
    a.S = states->n_states;
    a.pmax = desc->pmax;
    a.nsel = desc->n_sel;
    a.packed = nullptr;
    a.max_nphase = cout->max_nphase;
    a.c = args.c;
    long long blocks = (a.S + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pack_phase_soa16<<<(unsigned)blocks, 256, 0, s>>>(a);
    return check_cuda(cudaGetLastError(), "k_pack_phase_soa16 launch");
}

int fhmc_lnpi_1d(const fhmc_hist_desc *desc, const double *blob, const fhmc_states *states, const double *lnnorm,
                 double *lnpi_out, void *stream)
{
    if (validate_desc(desc) || validate_states(states)) return 1;
    if (!blob || !lnnorm || !lnpi_out) { set_error("null pointer"); return 1; }
    if (states->n_states == 0) return 0;
    SweepArgs args;
    memset(&args, 0, sizeof(args));
    args.d = *desc;
    args.blob = blob;
    args.st = *states;
    dim3 grid((desc->n + FHMC_CTA - 1) / FHMC_CTA, (unsigned)(states->n_states < 65535 ? states->n_states : 65535));
    if (desc->n_coef > 0) k_lnpi_1d<true><<<grid, FHMC_CTA, 0, (cudaStream_t)stream>>>(args, lnnorm, lnpi_out);
    else k_lnpi_1d<false><<<grid, FHMC_CTA, 0, (cudaStream_t)stream>>>(args, lnnorm, lnpi_out);
    return check_cuda(cudaGetLastError(), "k_lnpi_1d launch");
}

int fhmc_phase_moments(const double *lnpi, int n, const double *mom, int n_arrays, const int *bounds, int n_phase,
                       double *avg, double *lnsum, void *stream)
{
    if (!lnpi || !bounds || n < 1 || n_arrays < 0 || n_phase < 0 || (n_arrays > 0 && (!mom || !avg))) { set_error("bad arguments"); return 1; }
    if (n_phase == 0) return 0;
    const long long jobs = (long long)n_phase * (n_arrays + 1);
    long long blocks = (jobs + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_phase_moments<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(lnpi, n, mom, n_arrays, bounds, n_phase, avg, lnsum, nullptr, nullptr);
    return check_cuda(cudaGetLastError(), "k_phase_moments launch");
}

int fhmc_phase_moments_dev(const double *lnpi, int n, const double *mom, int n_arrays, const int *bounds, const unsigned *status,
                           const int *nphase, int pmax, double *avg, double *lnsum, void *stream)
{
    if (!lnpi || n < 1 || !bounds || !status || !nphase || pmax < 1 || n_arrays < 0 || (n_arrays > 0 && (!mom || !avg))) { set_error("bad arguments"); return 1; }
    const long long jobs = (long long)pmax * (n_arrays + 1);
    long long blocks = (jobs + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_phase_moments<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(lnpi, n, mom, n_arrays, bounds, pmax, avg, lnsum, status, nphase);
    return check_cuda(cudaGetLastError(), "k_phase_moments launch");
}

int fhmc_axpy_rows(const double *const *src, const double *w_host, int n_src, long long count, double *out, void *stream)
{
    if (!src || !w_host || !out || n_src < 1 || n_src > FHMC_MAX_TERMS || count < 0) { set_error("bad arguments"); return 1; }
    if (count == 0) return 0;
    AxpyArgs a;
    memset(&a, 0, sizeof(a));
    a.n_src = n_src;
    for (int t = 0; t < n_src; ++t) { a.src[t] = src[t]; a.w[t] = w_host[t]; }
    long long blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_axpy_rows<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, count, out);
    return check_cuda(cudaGetLastError(), "k_axpy_rows launch");
}

long long fhmc_pack_bytes(long long n_states, int pmax, int n_sel)
{
    if (n_states < 0 || pmax < 1 || n_sel < 0) return -1;
    return 8 * n_states + (long long)pmax * n_states * (16 + 8 * (long long)n_sel);
}

int fhmc_pack_phase_major(const fhmc_sweep_out *out, long long n_states, int pmax, int n_sel, void *packed, int *max_nphase,
                          void *stream)
{
    if (!out || !out->status || !out->nphase || !out->fe || !out->bounds || (n_sel > 0 && !out->avg) || !packed || !max_nphase ||
        n_states < 0 || pmax < 1 || n_sel < 0 || n_sel > FHMC_MAX_SEL) { set_error("bad arguments"); return 1; }
    if ((uintptr_t)packed & 15) { set_error("packed buffer must be 16-byte aligned"); return 1; }
    if (n_states == 0) return 0;
    PackArgs a;
    a.out = *out;
    a.S = n_states;
    a.pmax = pmax;
    a.nsel = n_sel;
    a.packed = static_cast<unsigned char *>(packed);
    a.max_nphase = max_nphase;
    memset(&a.c, 0, sizeof(a.c));
    long long blocks = (n_states + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pack_phase_major<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return check_cuda(cudaGetLastError(), "k_pack_phase_major launch");
}

long long fhmc_pack_soa16_bytes(long long n_states, int pmax, int n_sel)
{
    if (n_states < 0 || pmax < 1 || n_sel < 0) return -1;
    return ((4 * n_states + 15) & ~15ll) + (long long)pmax * n_states * (8 * (1 + (long long)n_sel) + 4);
}

int fhmc_pack_phase_soa16(const fhmc_sweep_out *out, long long n_states, int pmax, int n_sel, void *packed, int *max_nphase,
                          void *stream)
{
    if (!out || !out->status || !out->nphase || !out->fe || !out->bounds || (n_sel > 0 && !out->avg) || !packed || !max_nphase ||
        n_states < 0 || pmax < 1 || n_sel < 0 || n_sel > FHMC_MAX_SEL) { set_error("bad arguments"); return 1; }
    if ((uintptr_t)packed & 15) { set_error("packed buffer must be 16-byte aligned"); return 1; }
    if (n_states == 0) return 0;
    PackArgs a;
    a.out = *out;
    a.S = n_states;
    a.pmax = pmax;
    a.nsel = n_sel;
    a.packed = static_cast<unsigned char *>(packed);
    a.max_nphase = max_nphase;
    memset(&a.c, 0, sizeof(a.c));
    a.c.dst[0] = a.packed;
    a.c.n_dst = 1;
    a.c.n_total = n_states;
    a.c.first = 0;
    a.c.fill_dead = 1;
    long long blocks = (n_states + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pack_phase_soa16<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return check_cuda(cudaGetLastError(), "k_pack_phase_soa16 launch");
}

long long fhmc_bench_dfma(int iters, double *sink, void *stream)
{
    const DevInfo *di = dev_info();
    if (!di) return -1;
    const int blocks = di->sm_count * 8;
    k_bench_dfma<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (long long)blocks * 256 * 8 * iters;
}

long long fhmc_bench_exp(int iters, double *sink, void *stream)
{
    const DevInfo *di = dev_info();
    if (!di) return -1;
    const int blocks = di->sm_count * 8;
    k_bench_exp<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (long long)blocks * 256 * 4 * iters;
}

}  // extern "C"
