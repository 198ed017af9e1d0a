/*
 * gf_scan.cu -- Monte-Carlo prior scans (scripts/mc_unitary.py, mc_x.py, mc_texture.py) fused
 * with the ternary flavor histogram of plot.flavor_contour (plot.py:364-370).
 *
 * The reference draws prior samples by running emcee on a flat likelihood, maps the flavor
 * functions over the chain in Python and histograms later.  Here every sample is drawn directly
 * from the priors with a counter-based Philox4x32-10 stream (counter = global sample index, so
 * the result does not depend on launch geometry or on how the index range is sharded over GPUs),
 * pushed through the same per-point physics as the log-posterior kernel and counted into a
 * block-private shared-memory histogram; nothing but the final counts touches HBM.
 */
#include <atomic>
#include <string.h>

#include "gf_common.cuh"
#include "gf_scan_dev.cuh"

extern std::atomic<unsigned long long> g_gf_launches;

#define GF_SCAN_THREADS 256

/* ------------------------------------------------------------------ kernels */

/* Source of compositions: drawn samples (SCAN) or a given array (GIVEN). */
template <bool SCAN, bool SMEM_HIST, int SPEC>
__global__ void __launch_bounds__(GF_SCAN_THREADS, 2) /* <= 128 registers: two 256-thread blocks (2 x 70 KB histograms) per SM */
    k_hist(const __grid_constant__ gf_dev_model m, const uint64_t seed, const uint64_t first_index, const uint64_t count,
           const double* __restrict__ fr_in, const int nb1, const double step, unsigned long long* __restrict__ hist,
           unsigned long long* __restrict__ accepted) {
    extern __shared__ unsigned int sh_hist[];
    const int cells = nb1 * nb1 * nb1;
    if (SMEM_HIST) {
        for (int c = threadIdx.x; c < cells; c += blockDim.x) sh_hist[c] = 0u;
        __syncthreads();
    }
    unsigned int kept = 0u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        double fr[3];
        if (SCAN) {
            double theta[GF_MAX_DIM];
            gf_draw_theta(m, seed, first_index + j, theta);
            gf_point q;
            gf_resolve_point<SPEC>(m, [&](int k) { return theta[k]; }, q);
            gf_point_fr<SPEC>(m, q, fr);
        } else {
            fr[0] = fr_in[3 * j];
            fr[1] = fr_in[3 * j + 1];
            fr[2] = fr_in[3 * j + 2];
        }
        const int cell = gf_cell_index(fr, nb1, step);
        if (cell >= 0) {
            ++kept;
            if (SMEM_HIST)
                atomicAdd(&sh_hist[cell], 1u);
            else
                atomicAdd(&hist[cell], 1ull);
        }
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (int c = threadIdx.x; c < cells; c += blockDim.x) {
            const unsigned int v = sh_hist[c];
            if (v) atomicAdd(&hist[c], (unsigned long long)v);
        }
    }
    if (accepted) {
        /* warp-shuffle reduction, one atomic per warp */
        for (int o = 16; o > 0; o >>= 1) kept += __shfl_down_sync(0xffffffffu, kept, o);
        if ((threadIdx.x & 31) == 0 && kept) atomicAdd(accepted, (unsigned long long)kept);
    }
}

__global__ void __launch_bounds__(GF_SCAN_THREADS)
    k_scan_samples(const __grid_constant__ gf_dev_model m, const uint64_t seed, const uint64_t first_index, const uint64_t count,
                   double* __restrict__ theta_out, double* __restrict__ fr_out, uint8_t* __restrict__ status) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    double theta[GF_MAX_DIM];
    gf_draw_theta(m, seed, first_index + j, theta);
    gf_point q;
    gf_resolve_point(m, [&](int k) { return theta[k]; }, q);
    double fr[3];
    const unsigned st = gf_point_fr(m, q, fr);
    if (theta_out)
        for (int k = 0; k < m.ndim; ++k) theta_out[j * m.ndim + k] = theta[k];
    if (fr_out) {
        fr_out[3 * j] = fr[0];
        fr_out[3 * j + 1] = fr[1];
        fr_out[3 * j + 2] = fr[2];
    }
    if (status) status[j] = (uint8_t)st;
}

/* ------------------------------------------------------------------ C ABI */

namespace {

/* samples per launch: keeps the 32-bit block-private counters far from overflow */
constexpr uint64_t kMaxPerLaunch = 1ull << 36;
constexpr size_t kSmemHistLimit = 200 * 1024;

template <bool SCAN>
int launch_hist(const char* fn, const gf_dev_model& d, uint64_t seed, uint64_t first, uint64_t count, const double* d_fr, int nb,
                unsigned long long* d_hist, unsigned long long* d_accepted, cudaStream_t stream) {
    GF_REQUIRE(nb >= 0 && nb <= 1023, "%s: nb = %d outside [0, 1023]", fn, nb);
    GF_REQUIRE(d_hist != nullptr, "%s: histogram is NULL", fn);
    if (count == 0) return GF_OK;
    const int nb1 = nb + 1;
    const double step = 1.0 / (double)nb1;
    const size_t smem = (size_t)nb1 * nb1 * nb1 * sizeof(unsigned int);
    const bool use_smem = smem <= kSmemHistLimit;
    int sms = 0;
    if (int rc = gf_sm_count(&sms)) return rc;
    const bool fixed = SCAN && gf_model_is_fixed_spec(d);
    auto kern_s = fixed ? k_hist<SCAN, true, GF_SPEC_FIXED> : k_hist<SCAN, true, GF_SPEC_GENERIC>;
    auto kern_g = fixed ? k_hist<SCAN, false, GF_SPEC_FIXED> : k_hist<SCAN, false, GF_SPEC_GENERIC>;
    int per_sm = 1;
    if (use_smem) {
        GF_CUDA(cudaFuncSetAttribute(kern_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_s, GF_SCAN_THREADS, smem));
    } else {
        GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_g, GF_SCAN_THREADS, 0));
    }
    GF_REQUIRE(per_sm >= 1, "%s: kernel does not fit on an SM (smem %zu B)", fn, smem);
    for (uint64_t done = 0; done < count; done += kMaxPerLaunch) {
        const uint64_t cnt = (count - done < kMaxPerLaunch) ? count - done : kMaxPerLaunch;
        uint64_t want = (cnt + GF_SCAN_THREADS - 1) / GF_SCAN_THREADS;
        const uint64_t persistent = (uint64_t)sms * per_sm;
        const unsigned blocks = (unsigned)(want < persistent ? want : persistent);
        const double* frp = d_fr ? d_fr + 3 * done : nullptr;
        if (use_smem)
            kern_s<<<blocks, GF_SCAN_THREADS, smem, stream>>>(d, seed, first + done, cnt, frp, nb1, step, d_hist, d_accepted);
        else
            kern_g<<<blocks, GF_SCAN_THREADS, 0, stream>>>(d, seed, first + done, cnt, frp, nb1, step, d_hist, d_accepted);
        ++g_gf_launches;
        GF_LAUNCH_CHECK(fn);
    }
    return GF_OK;
}

}  // namespace

extern "C" int gf_scan_hist(const gf_model* model, const gf_scan_config* cfg, unsigned long long* d_hist,
                            unsigned long long* d_accepted, void* stream) {
    GF_REQUIRE(cfg != nullptr, "gf_scan_hist: cfg is NULL");
    gf_dev_model d;
    if (int rc = gf_build_dev_model(model, &d)) return rc;
    return launch_hist<true>("gf_scan_hist", d, cfg->seed, cfg->first_index, cfg->count, nullptr, cfg->nb, d_hist, d_accepted,
                             (cudaStream_t)stream);
}

extern "C" int gf_ternary_hist(const double* d_fr, int64_t n, int32_t nb, unsigned long long* d_hist, void* stream) {
    GF_REQUIRE(n >= 0, "gf_ternary_hist: n = %lld", (long long)n);
    GF_REQUIRE(n == 0 || d_fr != nullptr, "gf_ternary_hist: fr is NULL");
    gf_dev_model d;
    memset(&d, 0, sizeof(d));
    return launch_hist<false>("gf_ternary_hist", d, 0, 0, (uint64_t)n, d_fr, nb, d_hist, nullptr, (cudaStream_t)stream);
}

extern "C" int gf_scan_samples(const gf_model* model, const gf_scan_config* cfg, double* d_theta, double* d_fr, uint8_t* d_status,
                               void* stream) {
    GF_REQUIRE(cfg != nullptr, "gf_scan_samples: cfg is NULL");
    gf_dev_model d;
    if (int rc = gf_build_dev_model(model, &d)) return rc;
    if (cfg->count == 0) return GF_OK;
    GF_REQUIRE(cfg->count < (1ull << 40), "gf_scan_samples: count too large");
    k_scan_samples<<<gf_blocks_for((int64_t)cfg->count, GF_SCAN_THREADS), GF_SCAN_THREADS, 0, (cudaStream_t)stream>>>(
        d, cfg->seed, cfg->first_index, cfg->count, d_theta, d_fr, d_status);
    ++g_gf_launches;
    GF_LAUNCH_CHECK("gf_scan_samples");
    return GF_OK;
}
