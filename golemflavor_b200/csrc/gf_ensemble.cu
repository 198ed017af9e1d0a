/*
 * gf_ensemble.cu -- device-resident affine-invariant ensemble sampler (the stretch move that
 * golemflavor/mcmc.py:27-53 runs through emcee's host loop), batched over independent chains.
 *
 * One thread owns walker w of the first half and walker half + w of the second half of one chain.
 * A step is two half-steps: every walker of half h proposes q = c_j - z (c_j - p) with a random
 * partner c_j from the OTHER half (which is not modified during that half-step), scores q with the
 * same per-point log-posterior as k_lnprob, and accepts or rejects in place.  Half-steps are
 * separated by a grid-wide barrier: one cooperative launch runs the whole chain when the batch is
 * co-resident (launch-latency free: the emcee shapes of 512..2048 points per half-step are far too
 * small to amortise a launch per half-step), otherwise the host issues one launch per half-step.
 */
#include <atomic>
#include <cooperative_groups.h>

#include "gf_common.cuh"
#include "gf_ensemble_dev.cuh"

namespace cg = cooperative_groups;
extern std::atomic<unsigned long long> g_gf_launches;

#define GF_ENS_THREADS 128

/* COOP: the whole run in one cooperative launch; otherwise one (step, half) per launch. */
template <bool COOP>
__global__ void __launch_bounds__(GF_ENS_THREADS)
    k_ensemble(const __grid_constant__ gf_dev_model m, const gf_ens_args A, const int64_t one_step, const int one_half) {
    const int half = A.nwalkers / 2;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = A.nchains * half;
    const bool active = tid < total;
    const int64_t c = active ? tid / half : 0;
    const int w = active ? (int)(tid % half) : 0;
    const int64_t nstore = A.nsteps / A.thin;
    unsigned acc0 = 0u, acc1 = 0u;
    if (COOP) {
        cg::grid_group grid = cg::this_grid();
        for (int64_t s = 0; s < A.nsteps; ++s) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                if (active) {
                    const unsigned acc = gf_ens_update(m, A, c, h * half + w, h, A.step0 + s);
                    acc0 += h ? 0u : acc;
                    acc1 += h ? acc : 0u;
                }
                __threadfence();
                grid.sync();
            }
            if (active && (s + 1) % A.thin == 0 && (s + 1) / A.thin <= nstore) {
                gf_ens_store(m, A, c, w, (s + 1) / A.thin - 1, nstore);
                gf_ens_store(m, A, c, half + w, (s + 1) / A.thin - 1, nstore);
            }
        }
    } else if (active) {
        const int64_t s = one_step;
        if (one_half < 2) {
            const unsigned acc = gf_ens_update(m, A, c, one_half * half + w, one_half, A.step0 + s);
            acc0 = one_half ? 0u : acc;
            acc1 = one_half ? acc : 0u;
        }
        if (one_half == 2 && (s + 1) % A.thin == 0 && (s + 1) / A.thin <= nstore) { /* store pass */
            gf_ens_store(m, A, c, w, (s + 1) / A.thin - 1, nstore);
            gf_ens_store(m, A, c, half + w, (s + 1) / A.thin - 1, nstore);
        }
    }
    if (active && A.naccept) {
        if (acc0) atomicAdd(A.naccept + c * A.nwalkers + w, (unsigned long long)acc0);
        if (acc1) atomicAdd(A.naccept + c * A.nwalkers + half + w, (unsigned long long)acc1);
    }
}

/*
 * Block-per-chain variant: chains are independent, so only the walkers of ONE chain have to agree on
 * half-step boundaries.  When a half-ensemble fits a thread block (each thread may own a few walker
 * pairs) the chain lives in one block and half-steps are separated by __syncthreads() instead of a
 * grid-wide barrier; different chains drift apart freely.  This is the latency-optimal shape for the
 * emcee configurations (60 ... 1024 walkers): no 2-3 us grid barrier twice per step.
 */
#define GF_ENS_BLOCK_MAX 256
__global__ void __launch_bounds__(GF_ENS_BLOCK_MAX)
    k_ensemble_block(const __grid_constant__ gf_dev_model m, const gf_ens_args A) {
    const int half = A.nwalkers / 2;
    const int64_t c = blockIdx.x;
    const int64_t nstore = A.nsteps / A.thin;
    for (int64_t s = 0; s < A.nsteps; ++s) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            for (int w = threadIdx.x; w < half; w += blockDim.x) {
                const unsigned acc = gf_ens_update(m, A, c, h * half + w, h, A.step0 + s);
                if (acc && A.naccept) A.naccept[c * A.nwalkers + h * half + w] += 1ull; /* owned by this thread */
            }
            __threadfence_block();
            __syncthreads();
        }
        if ((s + 1) % A.thin == 0 && (s + 1) / A.thin <= nstore)
            for (int k = threadIdx.x; k < A.nwalkers; k += blockDim.x) gf_ens_store(m, A, c, k, (s + 1) / A.thin - 1, nstore);
    }
}

extern "C" int gf_ensemble_run(const gf_model* model, const gf_ensemble_config* cfg, double* d_pos, double* d_lnp, double* d_chain,
                               double* d_lnp_chain, unsigned long long* d_naccept, void* stream) {
    GF_REQUIRE(cfg != nullptr, "gf_ensemble_run: cfg is NULL");
    gf_dev_model d;
    if (int rc = gf_build_dev_model(model, &d)) return rc;
    GF_REQUIRE(cfg->nchains >= 0 && cfg->nsteps >= 0, "gf_ensemble_run: negative nchains / nsteps");
    GF_REQUIRE(cfg->nwalkers >= 2 && cfg->nwalkers % 2 == 0, "gf_ensemble_run: the number of walkers must be even, got %d", cfg->nwalkers);
    GF_REQUIRE(cfg->nfree >= 1 && cfg->nfree <= d.ndim, "gf_ensemble_run: nfree = %d outside [1, ndim]", cfg->nfree);
    GF_REQUIRE(cfg->thin >= 1, "gf_ensemble_run: thin must be >= 1");
    GF_REQUIRE(cfg->chain0 >= 0 && (cfg->chain0 + cfg->nchains) * (int64_t)cfg->nwalkers < (1ll << 32),
               "gf_ensemble_run: (chain0 + nchains) * nwalkers must stay below 2^32");
    GF_REQUIRE(cfg->a > 1.0, "gf_ensemble_run: stretch scale a = %g must exceed 1", cfg->a);
    if (cfg->nchains == 0 || cfg->nsteps == 0) return GF_OK;
    GF_REQUIRE(d_pos && d_lnp, "gf_ensemble_run: null pointer");
    gf_ens_args A;
    A.nchains = cfg->nchains; A.nsteps = cfg->nsteps; A.step0 = cfg->step0; A.thin = cfg->thin;
    A.nwalkers = cfg->nwalkers; A.nfree = cfg->nfree; A.a = cfg->a; A.seed = cfg->seed; A.chain0 = cfg->chain0;
    A.pos = d_pos; A.lnp = d_lnp; A.chain = d_chain; A.lnp_chain = d_lnp_chain; A.naccept = d_naccept;
    const int64_t total = cfg->nchains * (cfg->nwalkers / 2);
    const unsigned blocks = gf_blocks_for(total, GF_ENS_THREADS);
    cudaStream_t st = (cudaStream_t)stream;

    const int half = cfg->nwalkers / 2;
    GF_REQUIRE(cfg->mode >= 0 && cfg->mode <= 2, "gf_ensemble_run: mode = %d outside [0, 2]", cfg->mode);
    /* auto: block mode only when every thread owns ONE walker pair -- a second sequential pass per
     * half-step costs more than the grid barrier it saves (measured: 1024 walkers, 24 vs 12 us / step) */
    if (cfg->mode == 2 || (cfg->mode == 0 && half <= GF_ENS_BLOCK_MAX)) {
        const int threads = half >= GF_ENS_BLOCK_MAX ? GF_ENS_BLOCK_MAX : ((half + 31) / 32) * 32;
        k_ensemble_block<<<(unsigned)cfg->nchains, threads, 0, st>>>(d, A);
        ++g_gf_launches;
        GF_LAUNCH_CHECK("k_ensemble_block");
        return GF_OK;
    }
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    GF_CUDA(cudaGetDevice(&dev));
    GF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (int rc = gf_sm_count(&sms)) return rc;
    GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ensemble<true>, GF_ENS_THREADS, 0));
    if (coop && (int64_t)blocks <= (int64_t)sms * per_sm) {
        int64_t one_step = 0;
        int one_half = 0;
        void* params[] = {(void*)&d, (void*)&A, (void*)&one_step, (void*)&one_half};
        GF_CUDA(cudaLaunchCooperativeKernel((const void*)k_ensemble<true>, dim3(blocks), dim3(GF_ENS_THREADS), params, 0, st));
        ++g_gf_launches;
        return GF_OK;
    }
    for (int64_t s = 0; s < cfg->nsteps; ++s) {
        for (int h = 0; h < 2; ++h) {
            k_ensemble<false><<<blocks, GF_ENS_THREADS, 0, st>>>(d, A, s, h);
            ++g_gf_launches;
        }
        if ((d_chain || d_lnp_chain) && (s + 1) % cfg->thin == 0) {
            k_ensemble<false><<<blocks, GF_ENS_THREADS, 0, st>>>(d, A, s, 2);
            ++g_gf_launches;
        }
        GF_LAUNCH_CHECK("k_ensemble");
    }
    return GF_OK;
}
