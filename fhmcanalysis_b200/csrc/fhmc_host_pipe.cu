// fhmc_host_pipe.cu -- the host-buffer entry point of the mu sweep (new; the reference's seam for this is a Python loop of
// reweight()/thermo() calls on host arrays, GH:268-289, 451-554, README.md:60-85).
//
// fhmc_sweep_host_compact / fhmc_sweep_host_compact16: pinned host mu[S] in, pinned host results out, everything in between
// pipelined in chunks on three private streams (upload, compute, download): H2D(mu chunk) -> fhmc_sweep_1d -> repack
// (fhmc_pack_phase_major, or the narrow records of fhmc_pack_phase_soa16) -> D2H of the chunk's head and of the phase blocks
// that exist.  The number of live phase blocks of a chunk is only known after its kernels ran; waiting for it before
// queueing copies would idle the copy engine, so copies are queued at once for `guess` blocks (what the previous chunk
// needed) and a chunk that needed more is topped up when its flag is read, two chunks later, just before its device
// buffers are reused.
// Measured on a B200 (10^6 state points, 2^17-point chunks; FHMC_PIPE_TRACE=1 prints the device timeline): 1.44 ms with the
// 60-byte narrow records, 1.61-1.66 ms with the 72-byte ones.  The download stream is busy from the end of the first
// chunk's kernels on at ~50 GB/s (57 GB/s for one large copy); what is left over the D2H floor is the pipeline fill
// (~0.17 ms).  Tried without gain: a short first chunk, two download streams, driving the loop from Python (host time is
// not the limit).  Chunks below 2^16 points are slower: they hold fewer tiles than the GPU has resident CTAs.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fhmc_common.cuh"

namespace fhmc {

// One stream per engine: all kernels of all chunks on `comp` (a chunk's repack must not queue behind the NEXT chunk's
// persistent sweep CTAs, which is what happened when whole chunks alternated between two streams), mu uploads on `up`,
// result copies on `down`.
struct HostPipe {
    cudaStream_t comp, up, down;
    cudaEvent_t ready, h2d[2], done[2], flag[2], freed[2];
    bool ok;
};

static HostPipe *host_pipe()
{
    static HostPipe pipes[16];
    static bool made[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    HostPipe &p = pipes[dev];
    if (!made[dev]) {
        bool ok = cudaStreamCreateWithFlags(&p.comp, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&p.up, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&p.down, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming) == cudaSuccess;
        for (int b = 0; b < 2 && ok; ++b)
            ok = cudaEventCreateWithFlags(&p.h2d[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.done[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.flag[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.freed[b], cudaEventDisableTiming) == cudaSuccess;
        p.ok = ok;
        made[dev] = true;
    }
    return p.ok ? &p : nullptr;
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// device scratch of ONE of the two buffer sets
struct PipeBuf {
    double *mu;
    fhmc_sweep_out out;
    unsigned char *packed;
    int *flag;
    size_t bytes;
    unsigned char *rec_base;   // the record arrays as one block (scratch of fhmc_sweep_1d_compact)
    size_t rec_bytes;
};

static PipeBuf carve(unsigned char *base, long long c, int pmax, int nsel)
{
    PipeBuf b;
    size_t off = 0;
    auto take = [&](size_t nbytes) { unsigned char *p = base ? base + off : nullptr; off += al256(nbytes); return p; };
    b.mu = reinterpret_cast<double *>(take(8 * c));
    b.out.status = reinterpret_cast<unsigned *>(take(4 * c));
    b.out.nphase = reinterpret_cast<int *>(take(4 * c));
    b.out.nmin = reinterpret_cast<int *>(take(4 * c));
    b.out.lnnorm = reinterpret_cast<double *>(take(8 * c));
    b.out.fe = reinterpret_cast<double *>(take(8 * c * pmax));
    b.out.avg = reinterpret_cast<double *>(take(8 * c * pmax * (nsel > 0 ? nsel : 1)));
    b.out.bounds = reinterpret_cast<int *>(take(8 * c * pmax));
    b.out.max_idx = reinterpret_cast<int *>(take(4 * c * pmax));
    b.out.min_idx = reinterpret_cast<int *>(take(4 * c * (pmax + 1)));
    b.rec_base = reinterpret_cast<unsigned char *>(b.out.status);
    b.rec_bytes = off - al256(8 * c);
    b.packed = take((size_t)fhmc_pack_bytes(c, pmax, nsel));   // (>= fhmc_pack_soa16_bytes)
    b.flag = reinterpret_cast<int *>(take(16));
    b.bytes = off;
    return b;
}

}  // namespace fhmc

using namespace fhmc;

extern "C" size_t fhmc_sweep_host_workspace(long long chunk, int pmax, int n_sel)
{
    if (chunk < 1 || pmax < 1 || n_sel < 0 || n_sel > FHMC_MAX_SEL) return 0;
    return 2 * carve(nullptr, chunk, pmax, n_sel).bytes;
}

static int sweep_host_impl(const fhmc_hist_desc *desc, const double *blob, const double *mu_host, long long n_states,
                           int lanes_per_point, long long chunk, void *workspace, size_t workspace_bytes,
                           void *out_host, int *flags_host, int guess_nphase, int *max_nphase_out,
                           long long *d2h_bytes_out, void *stream, bool narrow)
{
    if (!desc || !blob || !mu_host || !workspace || !out_host || !flags_host || n_states < 0 || chunk < 1) { set_error("bad arguments"); return 1; }
    const int pmax = desc->pmax, nsel = desc->n_sel;
    const size_t need = fhmc_sweep_host_workspace(chunk, pmax, nsel);
    if (need == 0 || workspace_bytes < need) { set_error("workspace too small: need %zu bytes", need); return 1; }
    if ((uintptr_t)workspace & 255) { set_error("workspace must be 256-byte aligned"); return 1; }
    HostPipe *hp = host_pipe();
    if (!hp) { set_error("could not create the pipeline streams"); return 1; }
    if (max_nphase_out) *max_nphase_out = 0;
    if (d2h_bytes_out) *d2h_bytes_out = 0;
    if (n_states == 0) return 0;
    PipeBuf buf[2];
    buf[0] = carve(static_cast<unsigned char *>(workspace), chunk, pmax, nsel);
    buf[1] = carve(static_cast<unsigned char *>(workspace) + buf[0].bytes, chunk, pmax, nsel);
    const long long S = n_states, rec = 16 + 8 * (long long)nsel;
    const long long n_chunks = (S + chunk - 1) / chunk;
    unsigned char *oh = static_cast<unsigned char *>(out_host);
    int guess = guess_nphase < 1 ? 1 : (guess_nphase > pmax ? pmax : guess_nphase);
    int top = 0;
    long long moved = 0;
    // whatever the caller queued on `stream` (the blob upload) comes first
    if (check_cuda(cudaEventRecord(hp->ready, (cudaStream_t)stream), "cudaEventRecord")) return 1;
    if (check_cuda(cudaStreamWaitEvent(hp->comp, hp->ready, 0), "cudaStreamWaitEvent")) return 1;

    // blocks [from, upto] of chunk k: block 0 = head, block 1+p = phase p.  The live phase blocks leave in ONE strided copy
    // per array: rows of m records on the device, S records apart on the host.
    const long long nf8 = 8 * (1 + (long long)nsel);
    auto copies = [&](long long k, int from, int upto) -> int {
        const long long lo = k * chunk, m = (lo + chunk <= S ? chunk : S - lo);
        const PipeBuf &B = buf[k & 1];
        const long long hb = narrow ? 4 : 8;   // head bytes per state point
        if (from == 0) {
            if (check_cuda(cudaMemcpyAsync(oh + hb * lo, B.packed, (size_t)(hb * m), cudaMemcpyDeviceToHost, hp->down), "cudaMemcpyAsync D2H")) return 1;
            moved += hb * m;
            from = 1;
        }
        if (upto < from) return 0;
        const int np = upto - from + 1;
        if (!narrow) {
            const unsigned char *src = B.packed + 8 * m + (long long)(from - 1) * m * rec;
            unsigned char *dst = oh + 8 * S + ((long long)(from - 1) * S + lo) * rec;
            if (check_cuda(cudaMemcpy2DAsync(dst, (size_t)(S * rec), src, (size_t)(m * rec), (size_t)(m * rec), (size_t)np,
                                             cudaMemcpyDeviceToHost, hp->down), "cudaMemcpy2DAsync D2H")) return 1;
            moved += (long long)np * m * rec;
            return 0;
        }
        // narrow layout: fp64 fields and int16 bounds are separate arrays (fhmc_pack_phase_soa16)
        const unsigned char *Fd = B.packed + ((4 * m + 15) & ~15ll), *Bd = Fd + (long long)pmax * m * nf8;
        unsigned char *Fh = oh + ((4 * S + 15) & ~15ll), *Bh = Fh + (long long)pmax * S * nf8;
        if (check_cuda(cudaMemcpy2DAsync(Fh + ((long long)(from - 1) * S + lo) * nf8, (size_t)(S * nf8), Fd + (long long)(from - 1) * m * nf8,
                                         (size_t)(m * nf8), (size_t)(m * nf8), (size_t)np, cudaMemcpyDeviceToHost, hp->down), "cudaMemcpy2DAsync D2H")) return 1;
        if (check_cuda(cudaMemcpy2DAsync(Bh + ((long long)(from - 1) * S + lo) * 4, (size_t)(S * 4), Bd + (long long)(from - 1) * m * 4,
                                         (size_t)(m * 4), (size_t)(m * 4), (size_t)np, cudaMemcpyDeviceToHost, hp->down), "cudaMemcpy2DAsync D2H")) return 1;
        moved += (long long)np * m * (nf8 + 4);
        return 0;
    };
    int sent[2] = {0, 0};   // phase blocks already queued for the chunk that owns buffer set b
    std::vector<int> sent_final((size_t)n_chunks, 0);   // phase blocks chunk k has in out_host when the call returns
    // chunk k's live phase count is known once its flag arrived: top its copies up if the guess was short
    auto settle = [&](long long k) -> int {
        const int b = (int)(k & 1);
        if (check_cuda(cudaEventSynchronize(hp->flag[b]), "cudaEventSynchronize")) return 1;
        int live = flags_host[k];
        live = live < 1 ? 1 : (live > pmax ? pmax : live);
        if (live > top) top = live;
        if (live > guess) guess = live;
        if (live > sent[b]) {   // rare (first call, or a chunk with more phases than the last): re-mark the buffer set
            if (copies(k, 1 + sent[b], live)) return 1;
            sent[b] = live;
            sent_final[(size_t)k] = live;
            return check_cuda(cudaEventRecord(hp->freed[b], hp->down), "cudaEventRecord");
        }
        sent_final[(size_t)k] = sent[b];
        return 0;
    };
    fhmc_hist_desc d = *desc;
    for (long long k = 0; k < n_chunks; ++k) {
        const int b = (int)(k & 1);
        const long long lo = k * chunk, m = (lo + chunk <= S ? chunk : S - lo);
        if (k >= 2) {
            if (settle(k - 2)) return 1;
            // the kernels of chunk k overwrite what chunk k - 2 left in buffer set b: wait for its copies
            if (check_cuda(cudaStreamWaitEvent(hp->comp, hp->freed[b], 0), "cudaStreamWaitEvent")) return 1;
            // its mu buffer is free once the sweep of chunk k - 2 has run
            if (check_cuda(cudaStreamWaitEvent(hp->up, hp->done[b], 0), "cudaStreamWaitEvent")) return 1;
        }
        if (check_cuda(cudaMemcpyAsync(buf[b].mu, mu_host + lo, (size_t)(8 * m), cudaMemcpyHostToDevice, hp->up), "cudaMemcpyAsync H2D")) return 1;
        if (check_cuda(cudaEventRecord(hp->h2d[b], hp->up), "cudaEventRecord")) return 1;
        if (check_cuda(cudaStreamWaitEvent(hp->comp, hp->h2d[b], 0), "cudaStreamWaitEvent")) return 1;
        fhmc_states st;
        st.n_states = m;
        st.mu1 = buf[b].mu; st.n_mu1 = m; st.mu1_div = 1;
        st.beta = nullptr; st.n_beta = 1; st.beta_div = 1;
        st.dmu = nullptr; st.n_dmu = 1; st.dmu_div = 1;
        if (check_cuda(cudaMemsetAsync(buf[b].flag, 0, 4, hp->comp), "cudaMemsetAsync")) return 1;
        if (narrow && pmax <= 8 && lanes_per_point == 0) {
            // the sweep kernel writes the narrow phase-major records itself (k_sweep_prod2<compact>; any other kernel + repack
            // inside fhmc_sweep_1d_compact when the product form does not apply)
            fhmc_compact_out co;
            memset(&co, 0, sizeof(co));
            co.dst[0] = buf[b].packed;
            co.n_dst = 1;
            co.n_total = m;
            co.first = 0;
            co.fill_dead = 1;
            co.max_nphase = buf[b].flag;
            if (fhmc_sweep_1d_compact(&d, blob, &st, &co, buf[b].rec_base, buf[b].rec_bytes, hp->comp)) return 1;
        } else {
            if (fhmc_sweep_1d(&d, blob, &st, &buf[b].out, lanes_per_point, hp->comp)) return 1;
            if (narrow ? fhmc_pack_phase_soa16(&buf[b].out, m, pmax, nsel, buf[b].packed, buf[b].flag, hp->comp)
                       : fhmc_pack_phase_major(&buf[b].out, m, pmax, nsel, buf[b].packed, buf[b].flag, hp->comp)) return 1;
        }
        if (check_cuda(cudaEventRecord(hp->done[b], hp->comp), "cudaEventRecord")) return 1;
        if (check_cuda(cudaStreamWaitEvent(hp->down, hp->done[b], 0), "cudaStreamWaitEvent")) return 1;
        if (check_cuda(cudaMemcpyAsync(&flags_host[k], buf[b].flag, 4, cudaMemcpyDeviceToHost, hp->down), "cudaMemcpyAsync flag")) return 1;
        if (check_cuda(cudaEventRecord(hp->flag[b], hp->down), "cudaEventRecord")) return 1;
        if (copies(k, 0, guess)) return 1;
        sent[b] = guess;
        if (check_cuda(cudaEventRecord(hp->freed[b], hp->down), "cudaEventRecord")) return 1;   // buffer set b free again
    }
    for (long long k = (n_chunks >= 2 ? n_chunks - 2 : 0); k < n_chunks; ++k)
        if (settle(k)) return 1;
    if (check_cuda(cudaStreamSynchronize(hp->down), "cudaStreamSynchronize")) return 1;
    if (check_cuda(cudaStreamSynchronize(hp->comp), "cudaStreamSynchronize")) return 1;
    // A chunk copied max(guess at queue time, its own live count) phase blocks; `top` is the maximum over ALL chunks.  Blocks
    // [sent_final[k], top) of chunk k never crossed PCIe (no state point of the chunk has such a phase): give them the NaN / -1
    // the repack kernel writes into empty slots, so that every block below *max_nphase_out is defined for every state point.
    for (long long k = 0; k < n_chunks; ++k) {
        const long long lo = k * chunk, m = (lo + chunk <= S ? chunk : S - lo);
        for (int p = sent_final[(size_t)k]; p < top; ++p) {
            if (!narrow) {
                unsigned char *r = oh + 8 * S + ((long long)p * S + lo) * rec;
                for (long long j = 0; j < m; ++j, r += rec) {
                    double *f = reinterpret_cast<double *>(r);
                    for (int q = 0; q <= nsel; ++q) f[q] = NAN;
                    int *bi = reinterpret_cast<int *>(f + 1 + nsel);
                    bi[0] = bi[1] = -1;
                }
            } else {
                unsigned char *Fh = oh + ((4 * S + 15) & ~15ll), *Bh = Fh + (long long)pmax * S * nf8;
                double *f = reinterpret_cast<double *>(Fh + ((long long)p * S + lo) * nf8);
                for (long long j = 0; j < m * (1 + nsel); ++j) f[j] = NAN;
                short *bi = reinterpret_cast<short *>(Bh + ((long long)p * S + lo) * 4);
                for (long long j = 0; j < 2 * m; ++j) bi[j] = -1;
            }
        }
    }
    if (max_nphase_out) *max_nphase_out = top;
    if (d2h_bytes_out) *d2h_bytes_out = moved;
    return 0;
}

extern "C" int fhmc_sweep_host_compact(const fhmc_hist_desc *desc, const double *blob, const double *mu_host, long long n_states,
                                       int lanes_per_point, long long chunk, void *workspace, size_t workspace_bytes,
                                       void *out_host, int *flags_host, int guess_nphase, int *max_nphase_out,
                                       long long *d2h_bytes_out, void *stream)
{
    return sweep_host_impl(desc, blob, mu_host, n_states, lanes_per_point, chunk, workspace, workspace_bytes, out_host, flags_host,
                           guess_nphase, max_nphase_out, d2h_bytes_out, stream, false);
}

// Same pipeline with the narrow records of fhmc_pack_phase_soa16 (desc->n must be <= 32767: bounds travel as int16).
extern "C" int fhmc_sweep_host_compact16(const fhmc_hist_desc *desc, const double *blob, const double *mu_host, long long n_states,
                                         int lanes_per_point, long long chunk, void *workspace, size_t workspace_bytes,
                                         void *out_host, int *flags_host, int guess_nphase, int *max_nphase_out,
                                         long long *d2h_bytes_out, void *stream)
{
    if (desc && desc->n > 32767) { set_error("histogram too long for int16 bounds: use fhmc_sweep_host_compact"); return 1; }
    return sweep_host_impl(desc, blob, mu_host, n_states, lanes_per_point, chunk, workspace, workspace_bytes, out_host, flags_host,
                           guess_nphase, max_nphase_out, d2h_bytes_out, stream, true);
}

// ---------------------------------------------------------------------------------------------------------------------
// fhmc_scalar_point: the scalar drop-in calls (histogram.reweight / normalize / relextrema / thermo on ONE state point,
// GH:260-289, 317-415, 451-554) as one host call: H2D of what changed on the host side (ln(PI) row, N row, target mu_1),
// sweep record (warp-per-point general evaluator), normalised row, per-phase averages of every moment array with the phase
// bounds read from the record on the device, ONE D2H of the contiguous result range, one stream synchronisation.
// The notebook loop of the reference costs a dozen synchronising copies per state point when each method uploads and
// downloads on its own (0.86 ms measured); here a state point is two of these calls.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int fhmc_scalar_point(const fhmc_hist_desc *desc, const fhmc_scalar_io *io, int want_row, int want_moments, void *stream)
{
    if (!desc || !io || !io->blob || !io->mu1_dev || !io->mu1_pinned || !io->out_dev || !io->out_host) { fhmc::set_error("fhmc_scalar_point: missing buffer"); return 1; }
    if (want_moments && (!io->row || !io->lnsum || (io->n_arrays > 0 && (!io->mom || !io->avg)))) { fhmc::set_error("fhmc_scalar_point: missing moment buffers"); return 1; }
    if (want_moments && !want_row) { fhmc::set_error("fhmc_scalar_point: the moments are averaged over the normalised row"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t row_bytes = (size_t)desc->n * sizeof(double);
    if (io->lnpi_host && fhmc::check_cuda(cudaMemcpyAsync(io->blob, io->lnpi_host, row_bytes, cudaMemcpyHostToDevice, s), "H2D ln(PI)")) return 1;
    if (io->ntot_host && fhmc::check_cuda(cudaMemcpyAsync(io->blob + desc->n_pad, io->ntot_host, row_bytes, cudaMemcpyHostToDevice, s), "H2D N")) return 1;
    *io->mu1_pinned = io->mu1;
    if (fhmc::check_cuda(cudaMemcpyAsync(io->mu1_dev, io->mu1_pinned, sizeof(double), cudaMemcpyHostToDevice, s), "H2D mu_1")) return 1;
    fhmc_states st;
    memset(&st, 0, sizeof(st));
    st.n_states = 1;
    st.mu1 = io->mu1_dev;
    st.n_mu1 = st.mu1_div = 1;
    st.n_beta = st.beta_div = st.n_dmu = st.dmu_div = 1;
    if (fhmc_sweep_1d(desc, io->blob, &st, &io->rec, 32, stream)) return 1;
    if (want_row && fhmc_lnpi_1d(desc, io->blob, &st, io->rec.lnnorm, io->row, stream)) return 1;
    if (want_moments &&
        fhmc_phase_moments_dev(io->row, desc->n, io->mom, io->n_arrays, io->rec.bounds, io->rec.status, io->rec.nphase, desc->pmax,
                               io->avg, io->lnsum, stream)) return 1;
    if (fhmc::check_cuda(cudaMemcpyAsync(io->out_host, io->out_dev, io->out_bytes, cudaMemcpyDeviceToHost, s), "D2H results")) return 1;
    return fhmc::check_cuda(cudaStreamSynchronize(s), "fhmc_scalar_point");
}
