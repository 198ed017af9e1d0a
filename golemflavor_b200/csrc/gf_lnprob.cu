/*
 * gf_lnprob.cu -- the batched log-posterior kernel (llh.ln_prob, llh.py:121-130, with the
 * Gaussian flavor-ratio likelihood of the reference notebooks) and its siblings
 * (llh.lnprior, fr.flux_averaged_BSMu), plus the host-buffer pipeline gf_lnprob_host.
 *
 * One parameter point per thread; the flattened model travels as a __grid_constant__ kernel
 * parameter (constant bank), theta is read through a strided view so that both the emcee
 * row-major layout and an SoA layout are served, everything else stays in registers.
 * The kernel is fp64-pipe bound (~2.5e3 DFMA-pipe instructions per point for 20 energy bins
 * against 64 B of traffic), see DESIGN.md.  theta is loaded directly (no shared-memory staging):
 * an A/B test of a persistent-grid variant that double-buffered theta tiles through shared memory
 * with cp.async was 1-12 % SLOWER the coarser its work granularity (0.787 ms at one 32-point tile
 * per warp ... 0.874 ms fully persistent, vs 0.780 ms here) -- the exposed first-load latency is
 * covered by other warps' arithmetic, while the hardware block scheduler's fine-grained dynamic
 * balancing over SMs is worth more than the prefetch.
 */
#include <atomic>
#include <mutex>
#include <string.h>

#include "gf_common.cuh"

extern std::atomic<unsigned long long> g_gf_launches;

#ifndef GF_LP_THREADS
#define GF_LP_THREADS 64
#endif
#ifndef GF_LP_MIN_BLOCKS
#define GF_LP_MIN_BLOCKS 8 /* resident blocks per SM the register allocation is tuned for */
#endif

enum { GF_K_LNPROB = 0, GF_K_FR = 1, GF_K_LNPRIOR = 2 };

template <int KIND, int SPEC>
__global__ void __launch_bounds__(GF_LP_THREADS, GF_SPEC_IS_SM(SPEC) ? 16 : GF_LP_MIN_BLOCKS)
    k_lnprob(const __grid_constant__ gf_dev_model m, const gf_theta_view th, const int64_t n, double* __restrict__ lnp,
             double* __restrict__ fr_out, uint8_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* __restrict__ row = th.p + i * th.ld_point;
    if (SPEC == GF_SPEC_SM && KIND != GF_K_LNPRIOR) {
        /* SM-only models are bound by instruction issue, and a third of their instructions were 64-bit
         * address arithmetic of theta reads through the runtime column map (every value is read twice: by
         * the prior loop and by the physics).  Stage the point's row ONCE in shared memory, column-major
         * per block (each thread reads back only what it wrote: no barrier, no bank conflicts); a read is
         * then one LDS at a warp-uniform offset. */
        extern __shared__ double sh_theta[];
        const double* src = row;
        /* left to the compiler's default unrolling: forcing it fully (or not at all) costs the SoA layout 40 % */
        for (int k = 0; k < m.ndim; ++k, src += th.ld_dim) sh_theta[k * GF_LP_THREADS + threadIdx.x] = __ldg(src);
        auto get = [&](int k) { return sh_theta[k * GF_LP_THREADS + threadIdx.x]; };
        double fr[3];
        unsigned st = 0u;
        if (KIND == GF_K_FR) {
            gf_point q;
            gf_resolve_point<SPEC>(m, get, q);
            st = gf_point_fr<SPEC, 1>(m, q, fr);
        } else {
            lnp[i] = gf_point_lnprob<SPEC, 1>(m, get, fr, st);
        }
        if (fr_out) {
            fr_out[3 * i] = fr[0];
            fr_out[3 * i + 1] = fr[1];
            fr_out[3 * i + 2] = fr[2];
        }
        if (status) status[i] = (uint8_t)st;
        return;
    }
    /* one 64-bit multiply for the row, then a warp-uniform offset per column (k and ld_dim are uniform; for
     * GF_SPEC_SM6 every k is a compile-time constant and repeated reads of a column are one load) */
    auto get = [&](int k) { return __ldg(row + (int64_t)k * th.ld_dim); };
    if (KIND == GF_K_LNPRIOR) {
        lnp[i] = gf_point_lnprior(m, get);
        return;
    }
    double fr[3];
    unsigned st = 0u;
    if (KIND == GF_K_FR) {
        gf_point q;
        gf_resolve_point<SPEC>(m, get, q);
        st = gf_point_fr<SPEC, 2>(m, q, fr);
    } else {
        lnp[i] = gf_point_lnprob<SPEC, 2>(m, get, fr, st);
    }
    if (fr_out) {
        fr_out[3 * i] = fr[0];
        fr_out[3 * i + 1] = fr[1];
        fr_out[3 * i + 2] = fr[2];
    }
    if (status) status[i] = (uint8_t)st;
}

static int check_view(const char* fn, const gf_model* model, const double* d_theta, int64_t n, int64_t ld_point, int64_t ld_dim) {
    GF_REQUIRE(n >= 0, "%s: n = %lld", fn, (long long)n);
    GF_REQUIRE(n == 0 || d_theta != nullptr, "%s: theta is NULL", fn);
    GF_REQUIRE(ld_point >= 1 && ld_dim >= 1, "%s: leading dimensions (%lld, %lld) must be positive", fn, (long long)ld_point, (long long)ld_dim);
    (void)model;
    return GF_OK;
}

template <int KIND>
static int launch(const char* fn, const gf_model* model, const double* d_theta, int64_t n, int64_t ld_point, int64_t ld_dim,
                  double* d_lnp, double* d_fr, uint8_t* d_status, cudaStream_t stream) {
    gf_dev_model d;
    if (int rc = gf_build_dev_model(model, &d)) return rc;
    if (int rc = check_view(fn, model, d_theta, n, ld_point, ld_dim)) return rc;
    if (n == 0) return GF_OK;
    const gf_theta_view th{d_theta, ld_point, ld_dim};
    const int spec = KIND == GF_K_LNPRIOR ? GF_SPEC_GENERIC : gf_model_spec(d); /* NPFREE falls through to GENERIC */
    const unsigned blocks = gf_blocks_for(n, GF_LP_THREADS);
    if (spec == GF_SPEC_FIXED)
        k_lnprob<KIND, GF_SPEC_FIXED><<<blocks, GF_LP_THREADS, 0, stream>>>(d, th, n, d_lnp, d_fr, d_status);
    else if (spec == GF_SPEC_FIXED7)
        k_lnprob<KIND, GF_SPEC_FIXED7><<<blocks, GF_LP_THREADS, 0, stream>>>(d, th, n, d_lnp, d_fr, d_status);
    else if (spec == GF_SPEC_FIXED12)
        k_lnprob<KIND, GF_SPEC_FIXED12><<<blocks, GF_LP_THREADS, 0, stream>>>(d, th, n, d_lnp, d_fr, d_status);
    else if (spec == GF_SPEC_SM)
        k_lnprob<KIND, GF_SPEC_SM><<<blocks, GF_LP_THREADS, (size_t)d.ndim * GF_LP_THREADS * sizeof(double), stream>>>(d, th, n, d_lnp, d_fr, d_status);
    else if (spec == GF_SPEC_SM6)
        k_lnprob<KIND, GF_SPEC_SM6><<<blocks, GF_LP_THREADS, 0, stream>>>(d, th, n, d_lnp, d_fr, d_status);
    else
        k_lnprob<KIND, GF_SPEC_GENERIC><<<blocks, GF_LP_THREADS, 0, stream>>>(d, th, n, d_lnp, d_fr, d_status);
    ++g_gf_launches;
    GF_LAUNCH_CHECK(fn);
    return GF_OK;
}

extern "C" int gf_lnprob(const gf_model* model, const double* d_theta, int64_t n, int64_t ld_point, int64_t ld_dim, double* d_lnprob,
                         double* d_fr, uint8_t* d_status, void* stream) {
    GF_REQUIRE(n == 0 || d_lnprob != nullptr, "gf_lnprob: output is NULL");
    return launch<GF_K_LNPROB>("gf_lnprob", model, d_theta, n, ld_point, ld_dim, d_lnprob, d_fr, d_status, (cudaStream_t)stream);
}

extern "C" int gf_flux_averaged_fr(const gf_model* model, const double* d_theta, int64_t n, int64_t ld_point, int64_t ld_dim,
                                   double* d_fr, uint8_t* d_status, void* stream) {
    GF_REQUIRE(n == 0 || d_fr != nullptr, "gf_flux_averaged_fr: output is NULL");
    return launch<GF_K_FR>("gf_flux_averaged_fr", model, d_theta, n, ld_point, ld_dim, nullptr, d_fr, d_status, (cudaStream_t)stream);
}

extern "C" int gf_lnprior(const gf_model* model, const double* d_theta, int64_t n, int64_t ld_point, int64_t ld_dim, double* d_lnprior,
                          void* stream) {
    GF_REQUIRE(n == 0 || d_lnprior != nullptr, "gf_lnprior: output is NULL");
    return launch<GF_K_LNPRIOR>("gf_lnprior", model, d_theta, n, ld_point, ld_dim, d_lnprior, nullptr, nullptr, (cudaStream_t)stream);
}

/* ------------------------------------------------------------------ host-buffer pipeline */

namespace {

constexpr int kSlots = 3;               /* H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c */
constexpr int64_t kChunkPoints = 1 << 18;

struct HostPipe {
    int device = -1;
    cudaStream_t stream[kSlots] = {};
    cudaEvent_t done[kSlots] = {};
    double* d_theta[kSlots] = {};
    double* d_lnp[kSlots] = {};
    double* d_fr[kSlots] = {};
    uint8_t* d_st[kSlots] = {};
    /* pinned staging for pageable caller buffers */
    double* s_theta[kSlots] = {};
    double* s_lnp[kSlots] = {};
    double* s_fr[kSlots] = {};
    uint8_t* s_st[kSlots] = {};
    bool ready = false;
};

std::mutex g_pipe_mutex;
HostPipe g_pipe;

int pipe_init(HostPipe& p) {
    int dev = 0;
    GF_CUDA(cudaGetDevice(&dev));
    if (p.ready && p.device == dev) return GF_OK;
    GF_REQUIRE(!p.ready, "gf_lnprob_host: the host pipeline is bound to device %d, current device is %d", p.device, dev);
    for (int s = 0; s < kSlots; ++s) {
        GF_CUDA(cudaStreamCreateWithFlags(&p.stream[s], cudaStreamNonBlocking));
        GF_CUDA(cudaEventCreateWithFlags(&p.done[s], cudaEventDisableTiming));
        GF_CUDA(cudaMalloc(&p.d_theta[s], kChunkPoints * GF_MAX_DIM * sizeof(double)));
        GF_CUDA(cudaMalloc(&p.d_lnp[s], kChunkPoints * sizeof(double)));
        GF_CUDA(cudaMalloc(&p.d_fr[s], kChunkPoints * 3 * sizeof(double)));
        GF_CUDA(cudaMalloc(&p.d_st[s], kChunkPoints));
    }
    p.device = dev;
    p.ready = true;
    return GF_OK;
}

bool is_pinned(const void* ptr) {
    if (!ptr) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int staging_init(HostPipe& p) {
    if (p.s_theta[0]) return GF_OK;
    for (int s = 0; s < kSlots; ++s) {
        GF_CUDA(cudaHostAlloc(&p.s_theta[s], kChunkPoints * GF_MAX_DIM * sizeof(double), cudaHostAllocDefault));
        GF_CUDA(cudaHostAlloc(&p.s_lnp[s], kChunkPoints * sizeof(double), cudaHostAllocDefault));
        GF_CUDA(cudaHostAlloc(&p.s_fr[s], kChunkPoints * 3 * sizeof(double), cudaHostAllocDefault));
        GF_CUDA(cudaHostAlloc(&p.s_st[s], kChunkPoints, cudaHostAllocDefault));
    }
    return GF_OK;
}

}  // namespace

extern "C" int gf_lnprob_host(const gf_model* model, const double* h_theta, int64_t n, double* h_lnprob, double* h_fr, uint8_t* h_status) {
    gf_dev_model d;
    if (int rc = gf_build_dev_model(model, &d)) return rc;
    GF_REQUIRE(n >= 0, "gf_lnprob_host: n = %lld", (long long)n);
    if (n == 0) return GF_OK;
    GF_REQUIRE(h_theta && h_lnprob, "gf_lnprob_host: null pointer");
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    HostPipe& p = g_pipe;
    if (int rc = pipe_init(p)) return rc;
    const bool direct = is_pinned(h_theta) && is_pinned(h_lnprob) && is_pinned(h_fr) && is_pinned(h_status);
    if (!direct)
        if (int rc = staging_init(p)) return rc;
    const int ndim = d.ndim;
    const int64_t nchunks = (n + kChunkPoints - 1) / kChunkPoints;
    /* results of chunk c staged in slot c % kSlots are copied out before the slot is reused */
    auto drain = [&](int64_t c) -> int {
        const int s = (int)(c % kSlots);
        GF_CUDA(cudaEventSynchronize(p.done[s]));
        if (!direct) {
            const int64_t off = c * kChunkPoints, cnt = (n - off < kChunkPoints) ? n - off : kChunkPoints;
            memcpy(h_lnprob + off, p.s_lnp[s], cnt * sizeof(double));
            if (h_fr) memcpy(h_fr + 3 * off, p.s_fr[s], cnt * 3 * sizeof(double));
            if (h_status) memcpy(h_status + off, p.s_st[s], cnt);
        }
        return GF_OK;
    };
    for (int64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % kSlots);
        if (c >= kSlots)
            if (int rc = drain(c - kSlots)) return rc;
        const int64_t off = c * kChunkPoints, cnt = (n - off < kChunkPoints) ? n - off : kChunkPoints;
        const double* src = h_theta + off * ndim;
        if (!direct) {
            memcpy(p.s_theta[s], src, cnt * ndim * sizeof(double));
            src = p.s_theta[s];
        }
        GF_CUDA(cudaMemcpyAsync(p.d_theta[s], src, cnt * ndim * sizeof(double), cudaMemcpyHostToDevice, p.stream[s]));
        const gf_theta_view th{p.d_theta[s], ndim, 1};
        const int spec = gf_model_spec(d);
        if (spec == GF_SPEC_FIXED)
            k_lnprob<GF_K_LNPROB, GF_SPEC_FIXED><<<gf_blocks_for(cnt, GF_LP_THREADS), GF_LP_THREADS, 0, p.stream[s]>>>(
                d, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr);
        else if (spec == GF_SPEC_FIXED7)
            k_lnprob<GF_K_LNPROB, GF_SPEC_FIXED7><<<gf_blocks_for(cnt, GF_LP_THREADS), GF_LP_THREADS, 0, p.stream[s]>>>(
                d, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr);
        else if (spec == GF_SPEC_FIXED12)
            k_lnprob<GF_K_LNPROB, GF_SPEC_FIXED12><<<gf_blocks_for(cnt, GF_LP_THREADS), GF_LP_THREADS, 0, p.stream[s]>>>(
                d, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr);
        else if (spec == GF_SPEC_SM)
            k_lnprob<GF_K_LNPROB, GF_SPEC_SM><<<gf_blocks_for(cnt, GF_LP_THREADS), GF_LP_THREADS, (size_t)ndim * GF_LP_THREADS * sizeof(double), p.stream[s]>>>(
                d, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr);
        else if (spec == GF_SPEC_SM6)
            k_lnprob<GF_K_LNPROB, GF_SPEC_SM6><<<gf_blocks_for(cnt, GF_LP_THREADS), GF_LP_THREADS, 0, p.stream[s]>>>(
                d, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr);
        else
            k_lnprob<GF_K_LNPROB, GF_SPEC_GENERIC><<<gf_blocks_for(cnt, GF_LP_THREADS), GF_LP_THREADS, 0, p.stream[s]>>>(
                d, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr);
        ++g_gf_launches;
        GF_LAUNCH_CHECK("gf_lnprob_host");
        GF_CUDA(cudaMemcpyAsync(direct ? h_lnprob + off : p.s_lnp[s], p.d_lnp[s], cnt * sizeof(double), cudaMemcpyDeviceToHost, p.stream[s]));
        if (h_fr)
            GF_CUDA(cudaMemcpyAsync(direct ? h_fr + 3 * off : p.s_fr[s], p.d_fr[s], cnt * 3 * sizeof(double), cudaMemcpyDeviceToHost, p.stream[s]));
        if (h_status)
            GF_CUDA(cudaMemcpyAsync(direct ? h_status + off : p.s_st[s], p.d_st[s], cnt, cudaMemcpyDeviceToHost, p.stream[s]));
        GF_CUDA(cudaEventRecord(p.done[s], p.stream[s]));
    }
    for (int64_t c = (nchunks > kSlots ? nchunks - kSlots : 0); c < nchunks; ++c)
        if (int rc = drain(c)) return rc;
    return GF_OK;
}
