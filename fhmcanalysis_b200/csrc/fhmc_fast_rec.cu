// fhmc_fast_rec.cu -- instantiations of the one-thread-per-state-point kernel for pure mu sweeps with the exp recurrence
// (template flag REC of k_sweep_fast, fhmc_fast.cuh): exp(lnPI_i + s N_i - shift) is advanced along four interleaved bin
// chains by two multiplications per bin and re-anchored with a true exp every 64 bins ("strength
// reduction ... re-anchored every k bins is allowed").  Used when the caller set fhmc_hist_desc.mu_recurrence.
#include "fhmc_fast.cuh"

namespace fhmc {

int launch_fast_mu_rec(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream)
{
    const bool s0n = args.d.n_sel > 0 && args.d.sel_row[0] == 1;
    switch (args.d.n_sel) {
    case 0: return launch_fast<0, false, 0, 1, true>(args, sm_count, smem_optin, stream);
    case 1: return s0n ? launch_fast<1, true, 0, 1, true>(args, sm_count, smem_optin, stream) : launch_fast<1, false, 0, 1, true>(args, sm_count, smem_optin, stream);
    case 2: return s0n ? launch_fast<2, true, 0, 1, true>(args, sm_count, smem_optin, stream) : launch_fast<2, false, 0, 1, true>(args, sm_count, smem_optin, stream);
    case 3: return s0n ? launch_fast<3, true, 0, 1, true>(args, sm_count, smem_optin, stream) : launch_fast<3, false, 0, 1, true>(args, sm_count, smem_optin, stream);
    default: return s0n ? launch_fast<4, true, 0, 1, true>(args, sm_count, smem_optin, stream) : launch_fast<4, false, 0, 1, true>(args, sm_count, smem_optin, stream);
    }
}

}  // namespace fhmc
