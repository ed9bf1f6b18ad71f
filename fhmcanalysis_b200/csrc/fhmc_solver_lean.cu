// fhmc_solver_lean.cu -- K4, the batched coexistence solver, one WARP per solve on the lean evaluator (fhmc_lean.cuh).
//
// Reference: histogram.find_phase_eq (gc_hist.pyx:598-668), objective phase_eq_error (gc_hist.pyx:2570-2630); iteration =
// solve_one() (fhmc_solver.cuh): bracketed Newton on the signed F.E. difference of the pair the objective selects.
// Each evaluation is LeanEval::run (bit-identical u to PointEval, ~1/6 of its instructions); whatever it declines
// (monotone ln(PI) with ties, underflowing phase, failed re-test on the normalised values ...) is re-run on the spot by
// PointEval<32> on the same staged blob.  Solves are handed out through an atomic counter, so a warp that finishes early
// takes the next solve instead of waiting for the slowest warp of its tile.
#include "fhmc_solver_lean.cuh"

namespace fhmc {

int launch_solver_lean2(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream);   // fhmc_solver_lean2.cu

// returns 0 ok, 1 error, -1 "no lean instantiation for this descriptor" (caller uses the PointEval group kernel)
int launch_solver_lean(const SolveArgs &sa, int sm_count, int smem_optin, cudaStream_t stream)
{
    const fhmc_hist_desc &d = sa.sw.d;
    if (sa.sw.blob_global || d.complete || d.n < 3 || d.n_sel < 1 || d.pmax > FHMC_LEAN_PMAX) return -1;
    if ((((d.n + 31) / 32) | 1) > 64) return -1;   // a lane's chunk must fit its 64-bit candidate mask
    const int nt = d.n_sel > 0 ? d.n_term : 1;
#define FHMC_TRY(NC, NSEL, NT) \
    if (d.n_coef == NC && d.n_sel == NSEL && nt == NT) return launch_lean<NC, NSEL, NT>(sa, sm_count, smem_optin, stream)
    // pure mu solves
    FHMC_TRY(0, 1, 1);
    FHMC_TRY(0, 2, 1);
    FHMC_TRY(0, 3, 1);
    // Taylor-extrapolated solves: 1 species order 1/2/3 -> NC = 2/3/4; 2 species order 1/2 -> NC = 3/6
    FHMC_TRY(2, 1, 1);
    FHMC_TRY(3, 1, 1);
    FHMC_TRY(4, 1, 1);
    FHMC_TRY(6, 1, 1);
    FHMC_TRY(3, 3, 2);
#undef FHMC_TRY
    return launch_solver_lean2(sa, sm_count, smem_optin, stream);
}

}  // namespace fhmc

// diagnostic: sum (and optionally clear) the lean-path outcome counters of both instantiation units, see fhmc_lean.cuh
// (synchronises the device).  out24: 8 counters + 16 cycle counters (zero unless built with -DFHMC_LEAN_PROFILE)
namespace fhmc { int lean_stats_tu2(unsigned long long *out, int reset); }
extern "C" int fhmc_lean_stats(unsigned long long *out24, int reset)
{
    if (out24)
        for (int k = 0; k < 24; ++k) out24[k] = 0ull;
    if (fhmc::lean_stats_tu(out24, reset)) return 1;
    return fhmc::lean_stats_tu2(out24, reset);
}
