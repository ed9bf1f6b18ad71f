// fhmc_solver_lean2.cu -- second half of the instantiations of the warp-per-solve coexistence solver (fhmc_solver_lean.cuh).
#include "fhmc_solver_lean.cuh"

namespace fhmc {

int launch_solver_lean2(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = sa.sw.d;
    const int nt = d.n_sel > 0 ? d.n_term : 1;
#define FHMC_TRY(NC, NSEL, NT) \
    if (d.n_coef == NC && d.n_sel == NSEL && nt == NT) return launch_lean<NC, NSEL, NT>(sa, sm_count, smem_optin, stream)
    FHMC_TRY(2, 1, 2);
    FHMC_TRY(3, 1, 2);
    FHMC_TRY(4, 1, 2);
    FHMC_TRY(2, 3, 2);
    FHMC_TRY(4, 3, 2);
    FHMC_TRY(3, 1, 3);
    FHMC_TRY(6, 1, 3);
    FHMC_TRY(3, 3, 3);
    FHMC_TRY(6, 3, 3);
#undef FHMC_TRY
    return -1;
}

int lean_stats_tu2(unsigned long long *out, int reset) { return lean_stats_tu(out, reset); }

}  // namespace fhmc
