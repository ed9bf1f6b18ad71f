// fhmc_solver_fast.cu -- K4 for many solves: one coexistence solve per THREAD, every evaluation is the one-pass walk of
// fhmc_fast.cuh (packed shared-memory rows, four-bin blocks, on-the-fly phase sums) instead of the two-pass
// group evaluator.  ~9x fewer instructions per evaluation; irregular evaluations (monotone ties, failed verification,
// capacity) fall back to the generic one-lane evaluator on the spot.  Same iteration as fhmc_solver.cu (solve_one).
#include "fhmc_fast.cuh"
#include "fhmc_solver.cuh"

namespace fhmc {

#define FHMC_SOLVE_CTA 256  // upper bound; the launcher sizes the CTAs so that one round of CTAs covers all solves

template <int NSEL, bool SEL0N, int NC, int NT>
__global__ void __launch_bounds__(FHMC_SOLVE_CTA) k_find_phase_eq_fast(const __grid_constant__ SolveArgs sa)
{
    const SweepArgs &a = sa.sw;
    constexpr bool TAYLOR = (NC > 0) || (NT > 1);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const FastCtx cx = fast_prepare<NSEL, SEL0N, NC, NT, false>(a, smem_raw);
    PointEval<1, TAYLOR> pe(a, a.blob, threadIdx.x & 31, cx.s_tab);
    const ExpRegs ec = load_exp_regs();
    const long long T = a.st.n_states;
    const double n_mid = 0.5 * (a.blob[a.d.n_pad] + a.blob[a.d.n_pad + a.d.n - 1]);
    for (long long rec = (long long)blockIdx.x * blockDim.x + threadIdx.x; rec < T; rec += (long long)gridDim.x * blockDim.x) {
        const fhmc_states &st = a.st;
        const double mu = st.mu1[(rec / st.mu1_div) % st.n_mu1];
        const double beta = st.beta ? st.beta[(rec / st.beta_div) % st.n_beta] : a.d.beta_ref;
        const double dmu = st.dmu ? st.dmu[(rec / st.dmu_div) % st.n_dmu] : a.d.dmu_ref;
        solve_one(sa, rec, mu, beta, n_mid, true, [&](double m, int &P_now) {
            if (!fast_point<NSEL, SEL0N, NC, NT, false>(a, cx, pe, ec, rec, m, beta, dmu))
                run_generic_point<TAYLOR>(a, cx.s_tab, threadIdx.x & 31, m, beta, dmu, rec);
            P_now = a.out.nphase[rec];
            return a.out.status[rec];
        });
    }
}

template <int NSEL, bool SEL0N, int NC, int NT>
static int launch_solve_fast(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream)
{
    // same shared-memory footprint as the sweep kernel minus its fallback queue
    const size_t smem = fast_smem_bytes<NSEL, SEL0N, NC, NT, false>(sa.sw.d.n_pad);
    if (smem > (size_t)smem_optin) return -1;
    auto kern = k_find_phase_eq_fast<NSEL, SEL0N, NC, NT>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    int occ = 0;
    if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FHMC_SOLVE_CTA, smem), "occupancy query")) return 1;
    if (occ < 1) return -1;
    // a solve is a long dependent chain: spread the solves over every resident CTA slot in ONE round when possible
    const long long T = sa.sw.st.n_states, slots = (long long)sm_count * occ;
    long long cta = ((T + slots - 1) / slots + 31) / 32 * 32;
    if (cta < 32) cta = 32;
    if (cta > FHMC_SOLVE_CTA) cta = FHMC_SOLVE_CTA;
    long long grid = (T + cta - 1) / cta;
    if (grid > slots) grid = slots;
    kern<<<(unsigned)grid, (unsigned)cta, smem, stream>>>(sa);
    return check_cuda(cudaGetLastError(), "k_find_phase_eq_fast launch");
}

// returns 0 ok, 1 error, -1 "no instantiation for this term pattern" (caller uses the group-per-solve kernel)
int launch_solver_fast(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = sa.sw.d;
    if (sa.sw.blob_global || d.complete || d.n < 3 || d.n_sel < 1) return -1;
    if (d.n_coef == 0 && d.n_term == 1) {   // pure mu solves: the shift comes from the hull
        if (!(d.hull_len >= 2 && d.hull_row > 1 && d.hull_row + 2 <= d.n_rows) || d.sel_row[0] != 1) return -1;
        for (int q = 1; q < d.n_sel; ++q)
            if (d.sel_row[q] < 2) return -1;
        if (d.n_sel == 1) return launch_solve_fast<1, true, 0, 1>(sa, sm_count, smem_optin, stream);
        if (d.n_sel == 3) return launch_solve_fast<3, true, 0, 1>(sa, sm_count, smem_optin, stream);
        return -1;
    }
    // Taylor pattern of launch_fast_taylor(): term 0 multiplies the N row, every other term owns a distinct row >= 2
    if (d.n_coef < 2 || d.coef_row[0] != 1) return -1;
    for (int c = 1; c < d.n_coef; ++c) {
        if (d.coef_row[c] < 2) return -1;
        for (int e = 1; e < c; ++e)
            if (d.coef_row[e] == d.coef_row[c]) return -1;
    }
    for (int q = 0; q < d.n_sel; ++q)
        if (d.sel_row[q] < 2) return -1;
#define FHMC_TRY(NSEL, NC, NT) \
    if (d.n_sel == NSEL && d.n_coef == NC && d.n_term == NT) return launch_solve_fast<NSEL, false, NC, NT>(sa, sm_count, smem_optin, stream)
    FHMC_TRY(1, 2, 2);
    FHMC_TRY(1, 3, 2);
    FHMC_TRY(1, 3, 3);
    FHMC_TRY(1, 6, 3);
    FHMC_TRY(3, 2, 2);
    FHMC_TRY(3, 3, 2);
    FHMC_TRY(3, 4, 2);
    FHMC_TRY(3, 3, 3);
    FHMC_TRY(3, 6, 3);
#undef FHMC_TRY
    return -1;
}

}  // namespace fhmc
