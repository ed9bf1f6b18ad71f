// fhmc_fast_prod.cu -- instantiations of the one-thread-per-state-point kernel for pure mu sweeps in product form
// (template value REC = 2 of k_sweep_fast, fhmc_fast.cuh).  exp(lnPI_i + s N_i - shift) = P_i * t_i with P_i =
// exp(lnPI_i - A_g) tabulated per CTA (A_g = max lnPI over a 128-bin segment) and t_i = exp(A_g + s N_i - shift) a geometric
// sequence in the bin index, so the per-phase sums of a 4-bin block are Horner polynomials in exp(s dN): 4 fused
// multiply-adds per block and summed quantity instead of an exp per bin, re-anchored with a true exp every 128 bins and
// after every block that runs the exact extremum tests ("strength reduction ... re-anchored every k bins
// is allowed").  Used when the caller set fhmc_hist_desc.mu_recurrence = 2.
#include "fhmc_prod.cuh"

namespace fhmc {

// two state points per thread (fhmc_prod.cuh): halves the shared-memory traffic per state point
static int launch_fast_mu_prod2(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const bool s0n = args.d.n_sel > 0 && args.d.sel_row[0] == 1;
    switch (args.d.n_sel) {
    case 0: return launch_prod2<0, false>(args, sm_count, smem_optin, stream);
    case 1: return s0n ? launch_prod2<1, true>(args, sm_count, smem_optin, stream) : launch_prod2<1, false>(args, sm_count, smem_optin, stream);
    case 2: return s0n ? launch_prod2<2, true>(args, sm_count, smem_optin, stream) : launch_prod2<2, false>(args, sm_count, smem_optin, stream);
    default: return -1;   // three or four summed quantities: the table registers of two points do not fit; one point per thread
    }
}

int launch_fast_mu_prod(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    // two points per thread as soon as the one-point form would need a second round of tiles on its two CTAs per SM
    // (below that it fills the GPU better); measured crossover on B200: 121 -> 85 us at 2^17 state points
    if (args.d.mu_recurrence >= 3 && args.st.n_states > (long long)sm_count * 2 * FHMC_CTA) {
        const int rc = launch_fast_mu_prod2(args, sm_count, smem_optin, stream);
        if (rc >= 0) return rc;
    }
    const bool s0n = args.d.n_sel > 0 && args.d.sel_row[0] == 1;
    switch (args.d.n_sel) {
    case 0: return launch_fast<0, false, 0, 1, 2>(args, sm_count, smem_optin, stream);
    case 1: return s0n ? launch_fast<1, true, 0, 1, 2>(args, sm_count, smem_optin, stream) : launch_fast<1, false, 0, 1, 2>(args, sm_count, smem_optin, stream);
    case 2: return s0n ? launch_fast<2, true, 0, 1, 2>(args, sm_count, smem_optin, stream) : launch_fast<2, false, 0, 1, 2>(args, sm_count, smem_optin, stream);
    case 3: return s0n ? launch_fast<3, true, 0, 1, 2>(args, sm_count, smem_optin, stream) : launch_fast<3, false, 0, 1, 2>(args, sm_count, smem_optin, stream);
    default: return s0n ? launch_fast<4, true, 0, 1, 2>(args, sm_count, smem_optin, stream) : launch_fast<4, false, 0, 1, 2>(args, sm_count, smem_optin, stream);
    }
}

}  // namespace fhmc
