// fhmc_tab.cu -- per-histogram tables for pure mu sweeps (fhmc_tab.cuh): build entry points of the C ABI and the instantiations of
// the table-driven sweep kernel k_sweep_tab2.
#include <string.h>

#include "fhmc_tab.cuh"

namespace fhmc {

size_t carve_records(unsigned char *base, long long c, int pmax, int nsel, fhmc_sweep_out *o);   // fhmc_b200.cu

static bool tab_applicable(const fhmc_hist_desc &d)
{
    if (d.mu_recurrence < 2 || d.n_coef != 0 || d.n_term > 1 || d.complete || d.n < 3 || d.n > 32767 || d.n_sel > 2 || d.n_sel < 0) return false;
    if (!(d.hull_len >= 2 && d.hull_row > 1 && d.hull_row + 2 <= d.n_rows)) return false;
    return true;
}

// returns 0 ok, 1 error, -1 not applicable.  dry: only report the resident-CTA bound (*grid_out) for sizing the scratch.
int launch_tab2_compact(const SweepArgs &args, int sm_count, int smem_optin, cudaStream_t stream, int *grid_out, bool dry)
{
    const fhmc_hist_desc &d = args.d;
    if (!d.mu_tables || !tab_applicable(d) || d.pmax > FHMC_COMPACT_PMAX) return -1;
    const bool s0n = d.n_sel > 0 && d.sel_row[0] == 1;
    switch (d.n_sel) {
    case 0: return launch_tab2<0, false>(args, sm_count, smem_optin, stream, grid_out, dry);
    case 1: return s0n ? launch_tab2<1, true>(args, sm_count, smem_optin, stream, grid_out, dry)
                       : launch_tab2<1, false>(args, sm_count, smem_optin, stream, grid_out, dry);
    case 2: return s0n ? launch_tab2<2, true>(args, sm_count, smem_optin, stream, grid_out, dry)
                       : launch_tab2<2, false>(args, sm_count, smem_optin, stream, grid_out, dry);
    default: return -1;
    }
}

struct TabLayout {
    MuTabHeader h;
    size_t total;
};

static size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static TabLayout tab_layout(const fhmc_hist_desc &d)
{
    TabLayout L;
    memset(&L, 0, sizeof(L));
    const bool s0n = d.n_sel > 0 && d.sel_row[0] == 1;
    const int NX = d.n_sel - (s0n ? 1 : 0), RAW = 2 + NX, PK = RAW + (RAW & 1);
    const int BW = 2 + 4 * (1 + d.n_sel), SEGB = 32, GRPB = 8, QN = FHMC_FAST_QUEUE;
    const int n = d.n, npad = d.n_pad, nb = (n - 2) / 4, nseg = (nb + SEGB - 1) / SEGB, ngrp = (nb + GRPB - 1) / GRPB;
    MuTabHeader &h = L.h;
    h.magic = FHMC_TAB_MAGIC;
    h.n = n;
    h.n_pad = npad;
    h.smooth = d.smooth;
    h.n_sel = d.n_sel;
    h.sel0n = s0n ? 1 : 0;
    h.sel_row[0] = d.n_sel > 0 ? d.sel_row[0] : 0;
    h.sel_row[1] = d.n_sel > 1 ? d.sel_row[1] : 0;
    h.ep_cap = (int)up((size_t)4 * n + d.hull_len, 256);
    h.regA_bytes = npad * 8 * (PK + 1);
    h.regB_off = (int)((((size_t)npad * 8 * (PK + 1) + 16 + 512 + (size_t)QN * 8 + 64) + 15) & ~(size_t)15);   // fast_base_bytes
    h.regB_bytes = (int)up((size_t)nb * BW * 8 + (size_t)(nseg + 2) * 8 + (size_t)ngrp * 8, 16);
    size_t off = 256;
    auto take = [&](size_t bytes) { const size_t o = off; off = up(off + bytes, 256); return (long long)o; };
    h.off_imgA = take(h.regA_bytes);
    h.off_imgB = take(h.regB_bytes);
    h.off_iv = take((size_t)4 * n * 8);
    h.off_raw = take((size_t)h.ep_cap * 8);
    h.off_ep = take((size_t)(h.ep_cap + 1) * 8);
    h.off_rec = take((size_t)(h.ep_cap + 1) * FHMC_TAB_REC_I16 * 2);
    h.off_mu = take((size_t)(h.ep_cap + 1) * 8);
    h.off_scratch = take(carve_records(nullptr, h.ep_cap + 1, FHMC_COMPACT_PMAX, 0, nullptr));
    L.total = off;
    return L;
}

template <int NSEL, bool SEL0N>
static int launch_image(const SweepArgs &args, unsigned char *tables, const MuTabHeader &h, int smem_optin, cudaStream_t stream)
{
    const size_t smem = fast_smem_bytes<NSEL, SEL0N, 0, 1, 2>(args.d.n_pad);
    if (smem > (size_t)smem_optin) return -1;
    if ((size_t)h.regB_off + h.regB_bytes > smem) { set_error("table image does not fit the kernel's shared memory"); return 1; }
    auto kern = k_tab_image<NSEL, SEL0N>;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return 1;
    kern<<<1, FHMC_CTA, smem, stream>>>(args, tables, h);
    return check_cuda(cudaGetLastError(), "k_tab_image launch");
}

}  // namespace fhmc

using namespace fhmc;

extern "C" {

size_t fhmc_mu_tables_bytes(const fhmc_hist_desc *desc)
{
    if (!desc || !tab_applicable(*desc)) return 0;
    return tab_layout(*desc).total;
}

int fhmc_mu_tables_build(const fhmc_hist_desc *desc, const double *blob, void *tables, size_t tables_bytes, void *stream)
{
    if (!desc || !blob || !tables) { set_error("null argument"); return 1; }
    if (!tab_applicable(*desc)) { set_error("mu tables: not a pure mu sweep in product form with hull rows"); return 2; }
    if ((uintptr_t)tables & 255) { set_error("tables must be a 256-byte aligned device pointer"); return 1; }
    const TabLayout L = tab_layout(*desc);
    if (tables_bytes < L.total) { set_error("tables buffer too small: need fhmc_mu_tables_bytes() bytes"); return 1; }
    int sm_count = 0, smem_optin = 0;
    if (fhmc_device_info(&sm_count, &smem_optin)) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *tb = static_cast<unsigned char *>(tables);
    SweepArgs args;
    memset(&args, 0, sizeof(args));
    args.d = *desc;
    args.d.mu_tables = nullptr;
    args.blob = blob;
    const bool s0n = desc->n_sel > 0 && desc->sel_row[0] == 1;
    int rc;
    switch (desc->n_sel) {
    case 0: rc = launch_image<0, false>(args, tb, L.h, smem_optin, s); break;
    case 1: rc = s0n ? launch_image<1, true>(args, tb, L.h, smem_optin, s) : launch_image<1, false>(args, tb, L.h, smem_optin, s); break;
    default: rc = s0n ? launch_image<2, true>(args, tb, L.h, smem_optin, s) : launch_image<2, false>(args, tb, L.h, smem_optin, s); break;
    }
    if (rc < 0) { set_error("mu tables: histogram too large for shared memory"); return 2; }
    if (rc) return rc;
    const int E = L.h.ep_cap;
    k_tab_rank<<<(E + 255) / 256, 256, 0, s>>>(tb);
    k_tab_mu<<<(E + 1 + 255) / 256, 256, 0, s>>>(tb, desc->mu1_ref, desc->beta_ref);
    if (check_cuda(cudaGetLastError(), "k_tab_rank / k_tab_mu launch")) return 1;
    // the representatives through the general evaluator (a warp per state point), records into the scratch
    fhmc_hist_desc d2 = *desc;
    d2.n_sel = 0;
    d2.pmax = FHMC_COMPACT_PMAX;
    d2.mu_tables = nullptr;
    d2.mu_recurrence = 0;
    fhmc_states st;
    memset(&st, 0, sizeof(st));
    st.n_states = E + 1;
    st.mu1 = reinterpret_cast<const double *>(tb + L.h.off_mu);
    st.n_mu1 = E + 1;
    st.mu1_div = st.beta_div = st.dmu_div = 1;
    st.n_beta = st.n_dmu = 1;
    fhmc_sweep_out scratch;
    carve_records(tb + L.h.off_scratch, E + 1, FHMC_COMPACT_PMAX, 0, &scratch);
    if (fhmc_sweep_1d(&d2, blob, &st, &scratch, 32, stream)) return 1;
    const double *hull = blob + (size_t)desc->hull_row * desc->n_pad;
    k_tab_records<<<(E + 1 + 127) / 128, 128, 0, s>>>(tb, scratch, FHMC_COMPACT_PMAX, desc->mu1_ref, desc->beta_ref, hull, hull + desc->n_pad,
                                                      desc->hull_len);
    return check_cuda(cudaGetLastError(), "k_tab_records launch");
}

#ifdef FHMC_TAB_PROFILE
// probe builds only: per-phase cycle sums of k_sweep_tab2 (init, walk, finish, -, warp tiles)
int fhmc_tab_profile(unsigned long long *out8, int reset)
{
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    if (out8 && cudaMemcpyFromSymbol(out8, g_tab_prof, sizeof(g_tab_prof)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_tab_prof, z, sizeof(z)); }
    return 0;
}
#endif

}  // extern "C"
