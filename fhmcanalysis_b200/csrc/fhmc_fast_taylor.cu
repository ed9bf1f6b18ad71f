// fhmc_fast_taylor.cu -- instantiations of the one-thread-per-state-point kernel for Taylor-extrapolated sweeps
// (temp_dmu_extrap_multi-style (beta x dmu_2) grids and flat (mu_1, beta, dmu_2) lists; reference gc_hist.pyx:813-1239).
// Coefficient-term counts produced by histogram.taylor_rows():  1 species: order 1/2/3 -> NC = 2/3/4;
// 2 species: order 1/2 -> NC = 3/6.  Averaged quantities: none (skip_mom) or three with first-order terms (NT = 2 or 3).
#include "fhmc_fast.cuh"

namespace fhmc {

int launch_fast_taylor(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = args.d;
    // pattern the packed layout assumes: term 0 multiplies the N row, every other term owns a distinct row >= 2
    if (d.n_coef < 2 || d.coef_row[0] != 1) return -1;
    for (int c = 1; c < d.n_coef; ++c) {
        if (d.coef_row[c] < 2) return -1;
        for (int e = 1; e < c; ++e)
            if (d.coef_row[e] == d.coef_row[c]) return -1;
    }
    for (int q = 0; q < d.n_sel; ++q)
        if (d.sel_row[q] < 2) return -1;
#define FHMC_TRY(NSEL, NC, NT) \
    if (d.n_sel == NSEL && d.n_coef == NC && d.n_term == NT) return launch_fast<NSEL, false, NC, NT>(args, sm_count, smem_optin, stream)
    if (d.n_sel == 0) {
        FHMC_TRY(0, 2, 1);
        FHMC_TRY(0, 3, 1);
        FHMC_TRY(0, 4, 1);
        FHMC_TRY(0, 6, 1);
        if (d.n_term != 1) {  // n_term is irrelevant without quantities
            SweepArgs b = args;
            b.d.n_term = 1;
            return launch_fast_taylor(b, sm_count, smem_optin, stream);
        }
    }
    FHMC_TRY(3, 2, 2);
    FHMC_TRY(3, 3, 2);
    FHMC_TRY(3, 4, 2);
    FHMC_TRY(3, 3, 3);
    FHMC_TRY(3, 6, 3);
#undef FHMC_TRY
    return -1;
}

}  // namespace fhmc
