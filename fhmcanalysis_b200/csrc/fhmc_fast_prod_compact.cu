// fhmc_fast_prod_compact.cu -- instantiations of the headline product-form mu-sweep kernel (k_sweep_prod2, fhmc_prod.cuh) with
// COMPACT record output: every state point leaves the kernel as a phase-major narrow record (layout of
// fhmc_pack_phase_soa16), written by the walk itself to one or several destination buffers -- this GPU's, and for a gather
// fused into the sweep the buffers of its NVLink peers.  Replaces "nine strided per-record arrays + a repack kernel"
// (r01b: 262 MB of DRAM traffic per 10^6 points against 192 MB algorithmic, 13.8x read amplification from partial sectors).
#include "fhmc_prod.cuh"

namespace fhmc {

// returns 0 ok, 1 error, -1 not applicable.  dry: only report the resident-CTA bound (*grid_out) for sizing the scratch.
int launch_prod2_compact(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream, int *grid_out, bool dry)
{
    const fhmc_hist_desc &d = args.d;
    if (d.mu_recurrence < 2 || d.n_coef != 0 || d.n_term > 1 || d.complete || d.n < 3 || d.n > 32767 || d.pmax > FHMC_COMPACT_PMAX) return -1;
    if (!(d.hull_len >= 2 && d.hull_row > 1 && d.hull_row + 2 <= d.n_rows)) return -1;
    const bool s0n = d.n_sel > 0 && d.sel_row[0] == 1;
    switch (d.n_sel) {
    case 0: return launch_prod2<0, false, true>(args, sm_count, smem_optin, stream, grid_out, dry);
    case 1: return s0n ? launch_prod2<1, true, true>(args, sm_count, smem_optin, stream, grid_out, dry)
                       : launch_prod2<1, false, true>(args, sm_count, smem_optin, stream, grid_out, dry);
    case 2: return s0n ? launch_prod2<2, true, true>(args, sm_count, smem_optin, stream, grid_out, dry)
                       : launch_prod2<2, false, true>(args, sm_count, smem_optin, stream, grid_out, dry);
    default: return -1;
    }
}

}  // namespace fhmc
