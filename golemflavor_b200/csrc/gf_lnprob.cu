/*
 * gf_lnprob.cu -- the batched log-posterior kernel (llh.ln_prob, llh.py:121-130, with the
 * Gaussian flavor-ratio likelihood of the reference notebooks) and its siblings
 * (llh.lnprior, fr.flux_averaged_BSMu), plus the host-buffer pipeline gf_lnprob_host.
 *
 * One parameter point per thread; the flattened model travels as a __grid_constant__ kernel
 * parameter (constant bank), theta is read through a strided view so that both the emcee
 * row-major layout and an SoA layout are served, everything else stays in registers.
 * The BSM kernel is fp64-pipe bound (~1.9e3 fp64-pipe instructions per point for 20 energy bins
 * against 64 B of traffic), see DESIGN.md.  The kernels are specialised at compile time on the model
 * (gf_model_spec: which quantities are sampled, and -- for the reference's own column layouts -- where)
 * and on the theta view (LAYOUT).  theta is loaded directly by the BSM kernels (no shared-memory
 * staging): an A/B test of a persistent-grid variant that double-buffered theta tiles through shared
 * memory with cp.async was 1-12 % SLOWER the coarser its work granularity (0.787 ms at one 32-point
 * tile per warp ... 0.874 ms fully persistent, vs 0.780 ms at the time) -- the exposed first-load
 * latency is covered by other warps' arithmetic, while the hardware block scheduler's fine-grained
 * dynamic balancing over SMs is worth more than the prefetch.
 */
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "gf_common.cuh"

extern std::atomic<unsigned long long> g_gf_launches;

#ifndef GF_LP_THREADS
#define GF_LP_THREADS 64
#endif
#ifndef GF_LP_MIN_BLOCKS
#define GF_LP_MIN_BLOCKS 8 /* resident blocks per SM the register allocation is tuned for */
#endif

enum { GF_K_LNPROB = 0, GF_K_FR = 1, GF_K_LNPRIOR = 2 };

/* LAYOUT (compile-time layouts only): 0 = strided theta (runtime ld_point / ld_dim); 1 = contiguous rows
 * (ld_dim == 1): every read is one load at a constant offset from the row pointer; 2 = packed rows of an even
 * number of doubles on a 16-byte boundary: the row arrives in ndim/2 128-bit loads. */
/* Points per thread of the BSM kernels: a thread evaluates point i and then point i + blockDim.x, and asks for the
 * second point's row (prefetch.global.L1) before it starts on the first.  Every thread otherwise begins with a
 * ~1 us wait for its theta row (the first use of theta -- the prior's box test -- held 12 % of the kernel's stall
 * samples, profiles/r02a_lnprob.md): with two points per thread that wait is paid once per two points, while the work
 * granularity stays fine (128 points per block). */
#ifndef GF_LP_PTS
#define GF_LP_PTS 2
#endif
/* The SM-only kernel in its compile-time layout (GF_SPEC_SM6; every theta view) runs ~260 instructions per point against 48 + 8
 * bytes: a thread that loads its row, waits, computes and stores has nothing in flight for most of its life, and 32
 * resident warps x 1.5 KB per SM are short of the ~47 KB per SM that ~1 us of latency at the HBM rate needs (long_scoreboard
 * is its first stall reason, profiles/r02c_k1_sm_lnprob.md).  Each thread therefore loads the rows of GF_K1_PTS points into
 * registers back to back before it evaluates the first one. */
#ifndef GF_K1_PTS
#define GF_K1_PTS 2
#endif
#ifndef GF_K1_MIN_BLOCKS
#define GF_K1_MIN_BLOCKS 18 /* resident 64-thread blocks per SM the SM-only kernels are compiled for (56 registers: no spill; 16: -2 %, 20 spills) */
#endif
#define GF_LP_PTS_FOR(SPEC, KIND)                                                                                   \
    (((KIND) == GF_K_LNPRIOR) ? 1 : GF_SPEC_IS_FIXED(SPEC) ? GF_LP_PTS : (SPEC) == GF_SPEC_SM6 ? GF_K1_PTS : 1) /* the generic specialisation sits at the register limit already */

template <int KIND, int SPEC, int LAYOUT = 0>
__global__ void __launch_bounds__(GF_LP_THREADS, GF_SPEC_IS_SM(SPEC) ? (KIND == GF_K_LNPROB ? GF_K1_MIN_BLOCKS : 16) : GF_LP_MIN_BLOCKS)
    k_lnprob(const __grid_constant__ gf_dev_model m, const gf_theta_view th, const int64_t n, double* __restrict__ lnp,
             double* __restrict__ fr_out, uint8_t* __restrict__ status) {
    constexpr int PTS = GF_LP_PTS_FOR(SPEC, KIND); /* as in launch_lnprob */
    int64_t i = (int64_t)blockIdx.x * (blockDim.x * PTS) + threadIdx.x;
    if (i >= n) return;
    if constexpr (SPEC == GF_SPEC_SM6 && PTS > 1) {
        /* all rows of this thread first (packed rows: 3 x 128-bit loads per point, 3 PTS loads in flight per thread), then
         * the points one after the other from registers; rows past the end re-read the last row and are not evaluated */
        constexpr int ND = GF_SPEC_STATIC_NDIM(SPEC);
        double v[PTS][ND];
#pragma unroll
        for (int pt = 0; pt < PTS; ++pt) {
            const int64_t j = i + (int64_t)pt * blockDim.x;
            const double* __restrict__ r = th.p + (j < n ? j : n - 1) * th.ld_point;
            if constexpr (LAYOUT == 2) {
                const double2* __restrict__ r2 = reinterpret_cast<const double2*>(r);
#pragma unroll
                for (int k = 0; k < ND / 2; ++k) {
                    const double2 t = __ldg(r2 + k);
                    v[pt][2 * k] = t.x;
                    v[pt][2 * k + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < ND; ++k) v[pt][k] = __ldg(r + (LAYOUT == 1 ? (int64_t)k : (int64_t)k * th.ld_dim));
            }
        }
#pragma unroll
        for (int pt = 0; pt < PTS; ++pt) {
            const int64_t j = i + (int64_t)pt * blockDim.x;
            if (j >= n) return;
            double fr[3];
            unsigned st = 0u;
            const auto get = [&](int k) { return v[pt][k]; };
            if (KIND == GF_K_FR) {
                gf_point q;
                gf_resolve_point<SPEC>(m, get, q);
                st = gf_point_fr<SPEC, 1, 1>(m, q, fr);
            } else {
                lnp[j] = gf_point_lnprob<SPEC, 1, 1>(m, get, fr, st);
            }
            if (fr_out) {
                fr_out[3 * j] = fr[0];
                fr_out[3 * j + 1] = fr[1];
                fr_out[3 * j + 2] = fr[2];
            }
            if (status) status[j] = (uint8_t)st;
        }
        return;
    }
    if constexpr (PTS > 1 && GF_SPEC_IS_FIXED(SPEC)) {
        if (LAYOUT != 0 || th.ld_dim == 1) { /* contiguous rows: one or two 32-byte sectors ahead of time */
            const int64_t nxt = i + blockDim.x;
            if (nxt < n) {
                const char* r = reinterpret_cast<const char*>(th.p + nxt * th.ld_point);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(r));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(r + (m.ndim - 1) * 8));
            }
        }
    }
#pragma unroll 1
    for (int pt = 0; pt < (GF_SPEC_IS_FIXED(SPEC) ? PTS : 1); ++pt, i += blockDim.x) {
    if (i >= n) return;
    const double* __restrict__ row = th.p + i * th.ld_point;
    /* the point: prior + physics + likelihood on whatever `get` reads theta from */
    auto evaluate = [&](auto get, auto ilp) {
        constexpr int ILP = decltype(ilp)::value;
        if (KIND == GF_K_LNPRIOR) {
            lnp[i] = gf_point_lnprior(m, get);
            return;
        }
        double fr[3];
        unsigned st = 0u;
        /* the rare refinement path of the bin loop re-reads the row instead of H0 / T waiting in local memory */
        const gf_src_row again{row, LAYOUT == 0 ? th.ld_dim : (int64_t)1};
        if (KIND == GF_K_FR) {
            gf_point q;
            gf_resolve_point<SPEC>(m, get, q);
            st = gf_point_fr<SPEC, ILP, 1>(m, q, fr, 0, again);
        } else {
            lnp[i] = gf_point_lnprob<SPEC, ILP, 1>(m, get, fr, st, 0, again);
        }
        if (fr_out) {
            fr_out[3 * i] = fr[0];
            fr_out[3 * i + 1] = fr[1];
            fr_out[3 * i + 2] = fr[2];
        }
        if (status) status[i] = (uint8_t)st;
    };
    if constexpr (SPEC == GF_SPEC_SM && KIND != GF_K_LNPRIOR) {
        /* SM-only models are bound by instruction issue, and a third of their instructions were 64-bit
         * address arithmetic of theta reads through the runtime column map (every value is read twice: by
         * the prior loop and by the physics).  Stage the point's row ONCE in shared memory, column-major
         * per block (each thread reads back only what it wrote: no barrier, no bank conflicts); a read is
         * then one LDS at a warp-uniform offset. */
        extern __shared__ double sh_theta[];
        const double* src = row;
        /* left to the compiler's default unrolling: forcing it fully (or not at all) costs the SoA layout 40 % */
        for (int k = 0; k < m.ndim; ++k, src += th.ld_dim) sh_theta[k * GF_LP_THREADS + threadIdx.x] = __ldg(src);
        evaluate([&](int k) { return sh_theta[k * GF_LP_THREADS + threadIdx.x]; }, std::integral_constant<int, 1>{});
    } else if constexpr (LAYOUT == 2) {
        constexpr int ND = GF_SPEC_STATIC_NDIM(SPEC);
        static_assert(ND > 0 && ND % 2 == 0, "LAYOUT 2 needs a compile-time layout with an even number of columns");
        double v[ND];
        const double2* __restrict__ r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
        for (int k = 0; k < ND / 2; ++k) {
            const double2 t = __ldg(r2 + k);
            v[2 * k] = t.x;
            v[2 * k + 1] = t.y;
        }
        evaluate([&](int k) { return v[k]; }, std::integral_constant<int, 2>{});
    } else if constexpr (LAYOUT == 1) {
        evaluate([&](int k) { return __ldg(row + k); }, std::integral_constant<int, 2>{});
    } else {
        /* one 64-bit multiply for the row, then a warp-uniform offset per column (k and ld_dim are uniform;
         * with a compile-time layout every k is a constant and repeated reads of a column are one load) */
        evaluate([&](int k) { return __ldg(row + (int64_t)k * th.ld_dim); }, std::integral_constant<int, 2>{});
    }
    } /* points of this thread */
}

/* one launch of k_lnprob: specialisation from the model, theta layout from the view */
template <int KIND>
static void launch_lnprob(const gf_dev_model& d, int spec, const gf_theta_view& th, int64_t n, double* d_lnp, double* d_fr, uint8_t* d_status,
                          cudaStream_t stream) {
    const bool rows = th.ld_dim == 1;
    const bool packed16 = rows && th.ld_point == d.ndim && (reinterpret_cast<uintptr_t>(th.p) & 15u) == 0;
    const unsigned blocks = gf_blocks_for(n, GF_LP_THREADS * GF_LP_PTS_FOR(spec, KIND)); /* points per thread: as in k_lnprob */
#define GF_LP_LAUNCH(SPEC, LAYOUT, SMEM) k_lnprob<KIND, SPEC, LAYOUT><<<blocks, GF_LP_THREADS, SMEM, stream>>>(d, th, n, d_lnp, d_fr, d_status)
    switch (spec) {
        case GF_SPEC_FIXED: GF_LP_LAUNCH(GF_SPEC_FIXED, 0, 0); break;
        case GF_SPEC_FIXED7:
            if (rows) GF_LP_LAUNCH(GF_SPEC_FIXED7, 1, 0); else GF_LP_LAUNCH(GF_SPEC_FIXED7, 0, 0);
            break;
        case GF_SPEC_FIXED12:
            if (packed16) GF_LP_LAUNCH(GF_SPEC_FIXED12, 2, 0); else if (rows) GF_LP_LAUNCH(GF_SPEC_FIXED12, 1, 0); else GF_LP_LAUNCH(GF_SPEC_FIXED12, 0, 0);
            break;
        case GF_SPEC_SM: GF_LP_LAUNCH(GF_SPEC_SM, 0, (size_t)d.ndim * GF_LP_THREADS * sizeof(double)); break;
        case GF_SPEC_SM6:
            if (packed16) GF_LP_LAUNCH(GF_SPEC_SM6, 2, 0); else if (rows) GF_LP_LAUNCH(GF_SPEC_SM6, 1, 0); else GF_LP_LAUNCH(GF_SPEC_SM6, 0, 0);
            break;
        default: GF_LP_LAUNCH(GF_SPEC_GENERIC, 0, 0); break; /* GENERIC, and NPFREE (a scan-only specialisation) */
    }
#undef GF_LP_LAUNCH
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
    const int spec = KIND == GF_K_LNPRIOR ? GF_SPEC_GENERIC : gf_model_spec(d);
    launch_lnprob<KIND>(d, spec, th, n, d_lnp, d_fr, d_status, stream);
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

/* Ring of (H2D, kernel, D2H) slots: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c.  Defaults: 3
 * slots of 2^18 points; GF_HOST_SLOTS (2..8) and GF_HOST_CHUNK_LOG2 (12..24) in the environment override them when a
 * process first uses the pipeline (deeper rings ride out host-thread wake-up jitter when many ranks share one host). */
constexpr int kMaxSlots = 8;
#ifndef GF_HOST_CHUNK_LOG2
#define GF_HOST_CHUNK_LOG2 18
#endif
constexpr int kMaxPipeDevices = 64;

int env_int(const char* name, int fallback, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return fallback;
    const long x = strtol(v, nullptr, 10);
    return x < lo ? lo : x > hi ? hi : (int)x;
}

/* One pipeline per device ordinal (a process may drive several GPUs); created on first use, kept for the
 * life of the process. */
struct HostPipe {
    int slots = 3;
    int64_t chunk = 1ll << GF_HOST_CHUNK_LOG2;
    cudaStream_t stream[kMaxSlots] = {};
    cudaEvent_t done[kMaxSlots] = {};
    double* d_theta[kMaxSlots] = {};
    double* d_lnp[kMaxSlots] = {};
    double* d_fr[kMaxSlots] = {};
    uint8_t* d_st[kMaxSlots] = {};
    /* pinned staging for pageable caller buffers */
    double* s_theta[kMaxSlots] = {};
    double* s_lnp[kMaxSlots] = {};
    double* s_fr[kMaxSlots] = {};
    uint8_t* s_st[kMaxSlots] = {};
    bool ready = false, staged = false;
};

std::mutex g_pipe_mutex;
HostPipe g_pipes[kMaxPipeDevices];

void pipe_release(HostPipe& p) {
    for (int s = 0; s < kMaxSlots; ++s) {
        if (p.stream[s]) cudaStreamDestroy(p.stream[s]);
        if (p.done[s]) cudaEventDestroy(p.done[s]);
        cudaFree(p.d_theta[s]);
        cudaFree(p.d_lnp[s]);
        cudaFree(p.d_fr[s]);
        cudaFree(p.d_st[s]);
    }
    cudaGetLastError();
    const HostPipe fresh;
    /* staging buffers (if any) survive: they are host memory, independent of what failed */
    double* st[kMaxSlots]; double* sl[kMaxSlots]; double* sf[kMaxSlots]; uint8_t* ss[kMaxSlots];
    for (int s = 0; s < kMaxSlots; ++s) { st[s] = p.s_theta[s]; sl[s] = p.s_lnp[s]; sf[s] = p.s_fr[s]; ss[s] = p.s_st[s]; }
    const bool staged = p.staged;
    const int slots = p.slots;
    const int64_t chunk = p.chunk;
    p = fresh;
    for (int s = 0; s < kMaxSlots; ++s) { p.s_theta[s] = st[s]; p.s_lnp[s] = sl[s]; p.s_fr[s] = sf[s]; p.s_st[s] = ss[s]; }
    p.staged = staged;
    p.slots = slots;
    p.chunk = chunk;
}

int pipe_init_unchecked(HostPipe& p) {
    if (!p.staged) { /* the geometry is fixed once any buffer of this pipeline exists */
        p.slots = env_int("GF_HOST_SLOTS", 3, 2, kMaxSlots);
        p.chunk = 1ll << env_int("GF_HOST_CHUNK_LOG2", GF_HOST_CHUNK_LOG2, 12, 24);
    }
    for (int s = 0; s < p.slots; ++s) {
        GF_CUDA(cudaStreamCreateWithFlags(&p.stream[s], cudaStreamNonBlocking));
        GF_CUDA(cudaEventCreateWithFlags(&p.done[s], cudaEventDisableTiming));
        GF_CUDA(cudaMalloc(&p.d_theta[s], p.chunk * GF_MAX_DIM * sizeof(double)));
        GF_CUDA(cudaMalloc(&p.d_lnp[s], p.chunk * sizeof(double)));
        GF_CUDA(cudaMalloc(&p.d_fr[s], p.chunk * 3 * sizeof(double)));
        GF_CUDA(cudaMalloc(&p.d_st[s], p.chunk));
    }
    return GF_OK;
}

int pipe_init(HostPipe& p) {
    if (p.ready) return GF_OK;
    const int rc = pipe_init_unchecked(p);
    if (rc != GF_OK) {
        pipe_release(p); /* nothing half-built is kept: the next call starts from scratch */
        return rc;
    }
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
    if (p.staged) return GF_OK;
    for (int s = 0; s < p.slots; ++s) {
        if (!p.s_theta[s]) GF_CUDA(cudaHostAlloc(&p.s_theta[s], p.chunk * GF_MAX_DIM * sizeof(double), cudaHostAllocDefault));
        if (!p.s_lnp[s]) GF_CUDA(cudaHostAlloc(&p.s_lnp[s], p.chunk * sizeof(double), cudaHostAllocDefault));
        if (!p.s_fr[s]) GF_CUDA(cudaHostAlloc(&p.s_fr[s], p.chunk * 3 * sizeof(double), cudaHostAllocDefault));
        if (!p.s_st[s]) GF_CUDA(cudaHostAlloc(&p.s_st[s], p.chunk, cudaHostAllocDefault));
    }
    p.staged = true;
    return GF_OK;
}

/* the chunk loop proper; on an error return copies may still be in flight -- the caller quiesces the streams */
int pipe_run(HostPipe& p, const gf_dev_model& d, bool direct, const double* h_theta, int64_t n, double* h_lnprob, double* h_fr,
             uint8_t* h_status) {
    const int ndim = d.ndim;
    const int spec = gf_model_spec(d);
    const int kSlots = p.slots;
    const int64_t kChunkPoints = p.chunk;
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
        launch_lnprob<GF_K_LNPROB>(d, spec, th, cnt, p.d_lnp[s], h_fr ? p.d_fr[s] : nullptr, h_status ? p.d_st[s] : nullptr, p.stream[s]);
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

}  // namespace

extern "C" int gf_lnprob_host(const gf_model* model, const double* h_theta, int64_t n, double* h_lnprob, double* h_fr, uint8_t* h_status) {
    gf_dev_model d;
    if (int rc = gf_build_dev_model(model, &d)) return rc;
    GF_REQUIRE(n >= 0, "gf_lnprob_host: n = %lld", (long long)n);
    if (n == 0) return GF_OK;
    GF_REQUIRE(h_theta && h_lnprob, "gf_lnprob_host: null pointer");
    int dev = 0;
    GF_CUDA(cudaGetDevice(&dev));
    GF_REQUIRE(dev >= 0 && dev < kMaxPipeDevices, "gf_lnprob_host: device ordinal %d outside [0, %d)", dev, kMaxPipeDevices);
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    HostPipe& p = g_pipes[dev];
    if (int rc = pipe_init(p)) return rc;
    const bool direct = is_pinned(h_theta) && is_pinned(h_lnprob) && is_pinned(h_fr) && is_pinned(h_status);
    if (!direct)
        if (int rc = staging_init(p)) return rc;
    const int rc = pipe_run(p, d, direct, h_theta, n, h_lnprob, h_fr, h_status);
    if (rc != GF_OK) {
        /* the caller is about to be told that the call failed: no copy may still be writing into its buffers
         * (or reading the staging ring) after we return */
        for (int s = 0; s < p.slots; ++s) cudaStreamSynchronize(p.stream[s]);
        cudaGetLastError();
    }
    return rc;
}
