/*
 * gf_ensemble.cu -- device-resident affine-invariant ensemble sampler (the stretch move that
 * golemflavor/mcmc.py:27-53 runs through emcee's host loop), batched over independent chains.
 *
 * One thread owns walker w of the first half and walker half + w of the second half of one chain.
 * A step is two half-steps: every walker of half h proposes q = c_j - z (c_j - p) with a random
 * partner c_j from the OTHER half (which is not modified during that half-step), scores q with the
 * same per-point log-posterior as k_lnprob, and accepts or rejects in place.  Half-steps are
 * separated by a barrier.  Launch shapes, all giving identical chains:
 *   cluster : one thread-block CLUSTER per chain (up to 16 CTAs on 16 SMs of one GPC); the ensemble lives
 *             in the distributed shared memory of the cluster for the whole run, partners are read with
 *             ld.shared::cluster, half-steps are separated by the hardware cluster barrier.  The
 *             latency-optimal shape for the emcee configurations (60 ... 8192 walkers).
 *   block   : one block per chain with several walker pairs per thread (positions in global memory).
 *   grid    : grid-wide barrier -- one cooperative launch when the batch is co-resident, otherwise one
 *             launch per half-step.
 * The per-point log-posterior is the same specialisation (gf_model_spec) that gf_lnprob launches.
 */
#include <atomic>
#include <cooperative_groups.h>

#include "gf_common.cuh"
#include "gf_ensemble_dev.cuh"

namespace cg = cooperative_groups;
extern std::atomic<unsigned long long> g_gf_launches;

#define GF_ENS_THREADS 128
/* energy bins interleaved per thread (gf_bin_loop): the sampler runs about one warp per SM sub-partition,
 * so the latency of the eigen-stage chain is the step time -- interleave as many bins as registers allow */
#ifndef GF_ENS_ILP_FIXED
#define GF_ENS_ILP_FIXED 4
#endif
#ifndef GF_ENS_ILP_GENERIC
#define GF_ENS_ILP_GENERIC 2
#endif

/* COOP: the whole run in one cooperative launch; otherwise one (step, half) per launch. */
template <bool COOP, int SPEC, int ILP>
__global__ void __launch_bounds__(GF_ENS_THREADS)
    k_ensemble(const __grid_constant__ gf_dev_model m, const gf_ens_args A, const int64_t one_step, const int one_half) {
    constexpr int LANES = GF_ENS_LANES(SPEC);
    const int half = A.nwalkers / 2;
    const int64_t tid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; /* walker pair of this thread */
    const int lane = (int)(threadIdx.x % LANES);
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
                    const unsigned acc = gf_ens_update<SPEC, ILP, LANES>(m, A, c, h * half + w, h, A.step0 + s, lane);
                    acc0 += h ? 0u : acc;
                    acc1 += h ? acc : 0u;
                }
                __threadfence();
                grid.sync();
            }
            if (active && lane == 0 && (s + 1) % A.thin == 0 && (s + 1) / A.thin <= nstore) {
                gf_ens_store(m, A, c, w, (s + 1) / A.thin - 1, nstore);
                gf_ens_store(m, A, c, half + w, (s + 1) / A.thin - 1, nstore);
            }
        }
    } else if (active) {
        const int64_t s = one_step;
        if (one_half < 2) {
            const unsigned acc = gf_ens_update<SPEC, ILP, LANES>(m, A, c, one_half * half + w, one_half, A.step0 + s, lane);
            acc0 = one_half ? 0u : acc;
            acc1 = one_half ? acc : 0u;
        }
        if (one_half == 2 && lane == 0 && (s + 1) % A.thin == 0 && (s + 1) / A.thin <= nstore) { /* store pass */
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
template <int SPEC, int ILP>
__global__ void __launch_bounds__(GF_ENS_BLOCK_MAX)
    k_ensemble_block(const __grid_constant__ gf_dev_model m, const gf_ens_args A) {
    constexpr int LANES = GF_ENS_LANES(SPEC);
    const int half = A.nwalkers / 2;
    const int lane = (int)(threadIdx.x % LANES);
    const int64_t c = blockIdx.x;
    const int64_t nstore = A.nsteps / A.thin;
    for (int64_t s = 0; s < A.nsteps; ++s) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            for (int w = threadIdx.x / LANES; w < half; w += blockDim.x / LANES) {
                const unsigned acc = gf_ens_update<SPEC, ILP, LANES>(m, A, c, h * half + w, h, A.step0 + s, lane);
                if (acc && A.naccept) A.naccept[c * A.nwalkers + h * half + w] += 1ull; /* owned by this thread */
            }
            __threadfence_block();
            __syncthreads();
        }
        if ((s + 1) % A.thin == 0 && (s + 1) / A.thin <= nstore)
            for (int k = threadIdx.x; k < A.nwalkers; k += blockDim.x) gf_ens_store(m, A, c, k, (s + 1) / A.thin - 1, nstore);
    }
}


/*
 * Cluster-per-chain variant.  CTA `rank` of the cluster owns walker pairs [rank*T, (rank+1)*T): walker w of BOTH
 * halves; positions and log-posteriors stay in its shared memory (sh = pos[2][T][ndim], lnp[2][T]) from the first
 * step to the last.  During half-step h every thread reads its partner from the OTHER half through distributed
 * shared memory -- that half is not written during the half-step -- and updates its own walker of half h in place;
 * the cluster barrier (barrier.cluster arrive.release / wait.acquire) orders the two.  No global-memory round trip
 * and no grid barrier on the critical path: a half-step costs one log-posterior latency plus the barrier.
 * On the BSM path two adjacent lanes share one walker pair (LANES = 2, see gf_bin_loop): both run the same update and
 * split the energy bins of its log-posterior; lane 0 writes.
 *
 * Measured and dropped (round 2): a dedicated producer warp per CTA that computes the next half-step's draws into a
 * double-buffered shared-memory table while the consumers evaluate.  It took 4-6 % off a single chain (C2 2.99 ->
 * 2.82, C3 13.2 -> 12.6 us per step) but the extra warp and the 224-register budget it forces cut the resident CTAs
 * per SM from 5 to 3, so the 600 single-warp chains of the sensitivity sweep no longer ran in one wave (+15 %).
 */
/* consumer threads per CTA: 256 walker pairs for the register-light SM-only models, 128 walker pairs x 2 lanes on the
 * BSM path (<= 255 registers per thread: no spills on the latency-critical path) */
#define GF_ENS_CL_MAX_CONSUMERS 256
template <int SPEC, int ILP>
__global__ void __launch_bounds__(GF_ENS_CL_MAX_CONSUMERS, 1)
    k_ensemble_cluster(const __grid_constant__ gf_dev_model m, const gf_ens_args A) {
    extern __shared__ double sh_ens[];
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int ND = GF_SPEC_STATIC_NDIM(SPEC); /* compile-time layouts: constant dimension count */
    constexpr int LANES = GF_ENS_LANES(SPEC);
    /* T = walker pairs of this CTA: with two lanes per walker (BSM) threads 2 wl and 2 wl + 1 share walker pair wl */
    const int T = (int)blockDim.x / LANES, ndim = ND > 0 ? ND : m.ndim, half = A.nwalkers / 2;
    const int nc = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int64_t c = blockIdx.x / nc;
    const int wl = (int)threadIdx.x / LANES, lane = (int)threadIdx.x % LANES, w = rank * T + wl;
    const bool active = w < half;
    const bool writer = active && lane == 0;
    double* pos_s = sh_ens;
    double* lnp_s = sh_ens + 2 * T * ndim;
    const int64_t nstore = A.nsteps / A.thin;
    if (writer) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int64_t k = c * A.nwalkers + h * half + w;
            for (int d = 0; d < ndim; ++d) pos_s[(h * T + wl) * ndim + d] = A.pos[k * ndim + d];
            lnp_s[h * T + wl] = A.lnp[k];
        }
    }
    cluster.sync();
    unsigned acc0 = 0u, acc1 = 0u;
    const uint64_t gid0 = (uint64_t)(A.chain0 + c) * (uint64_t)A.nwalkers + (uint64_t)w;
#ifdef GF_ENS_PROFILE
    long long t_draw = 0, t_eval = 0, t_sync = 0;
#define GF_TICK(var) { const long long now_ = clock64(); var += now_ - tick_; tick_ = now_; }
    long long tick_ = clock64();
#else
#define GF_TICK(var)
#endif
    /* software pipeline: the draws of the NEXT half-step (Philox, stretch factor, both logarithms -- nothing
     * that depends on a walker position) are computed between the arrive and the wait of the cluster barrier */
    /* chain output of this thread's two walkers (emcee layout [walker][slot][dim]) */
    int64_t until_store = A.thin, stored = 0;
    double* out0 = A.chain ? A.chain + (c * A.nwalkers + w) * nstore * ndim : nullptr;
    double* out1 = A.chain ? A.chain + (c * A.nwalkers + half + w) * nstore * ndim : nullptr;
    double* lout0 = A.lnp_chain ? A.lnp_chain + (c * A.nwalkers + w) * nstore : nullptr;
    double* lout1 = A.lnp_chain ? A.lnp_chain + (c * A.nwalkers + half + w) * nstore : nullptr;
    gf_ens_draw dr;
    if (active) {
        dr = gf_ens_draws(A, gid0, A.step0, half);
        gf_ens_finish_draw(A, dr);
    }
    GF_TICK(t_draw)
    for (int64_t s = 0; s < A.nsteps; ++s) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            if (active) {
                /* j / T for j < 4096, T <= 256: (j + 1/2) / T stays >= 1/512 away from every integer, so the
                 * approximate fp32 quotient truncates to the exact floor (an integer division costs ~150 cycles) */
                const int rj = __float2int_rz(__fdividef((float)dr.j + 0.5f, (float)T)), jl = dr.j - rj * T;
                const double* cj = cluster.map_shared_rank(pos_s + ((1 - h) * T + jl) * ndim, rj);
                double* p = pos_s + (h * T + wl) * ndim;
                double q[GF_MAX_DIM];
                double lnew;
                const bool accept = gf_ens_move<SPEC, ILP, true, LANES>(m, A, dr, [&](int d) { return cj[d]; }, [&](int d) { return p[d]; },
                                                                         lnp_s[h * T + wl], q, lnew, lane);
                if (LANES > 1) __syncwarp(3u << (((unsigned)threadIdx.x & 31u) & ~1u)); /* the partner lane has read the old position */
                if (accept && lane == 0) {
                    _Pragma("unroll") for (int d = 0; d < (ND > 0 ? ND : GF_MAX_DIM); ++d)
                        if (d < ndim) p[d] = q[d]; /* static indices keep q in registers */
                    lnp_s[h * T + wl] = lnew;
                    acc0 += h ? 0u : 1u;
                    acc1 += h ? 1u : 0u;
                }
                GF_STAGE(11);
                GF_TICK(t_eval)
            }
            cluster.barrier_arrive();
            if (active) { /* next half-step: (s, 1) after (s, 0), (s + 1, 0) after (s, 1) */
                dr = gf_ens_draws(A, gid0 + (uint64_t)((1 - h) * half), A.step0 + s + h, half);
                gf_ens_finish_draw(A, dr);
                GF_TICK(t_draw)
                /* the step is complete for this thread's two walkers once its own second-half update is
                 * done (only their owner writes them): store them in the shadow of the barrier as well */
                if (lane == 0 && h == 1 && --until_store == 0) {
                    until_store = A.thin;
                    if (stored < nstore) { /* running output pointers: no 64-bit index arithmetic per store */
                        if (A.chain) {
#pragma unroll
                            for (int d = 0; d < (ND > 0 ? ND : GF_MAX_DIM); ++d) {
                                if (d < ndim) {
                                    out0[d] = pos_s[wl * ndim + d];
                                    out1[d] = pos_s[(T + wl) * ndim + d];
                                }
                            }
                            out0 += ndim;
                            out1 += ndim;
                        }
                        if (A.lnp_chain) {
                            *lout0++ = lnp_s[wl];
                            *lout1++ = lnp_s[T + wl];
                        }
                        ++stored;
                    }
                }
            }
            cluster.barrier_wait();
            GF_STAGE(12);
            GF_TICK(t_sync)
        }
    }
#ifdef GF_ENS_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        printf("cluster sampler profile (cycles per half-step, thread 0): draws %.0f  move %.0f  barrier %.0f\n",
               (double)t_draw / (2.0 * A.nsteps), (double)t_eval / (2.0 * A.nsteps), (double)t_sync / (2.0 * A.nsteps));
        const char* names[13] = {"(from barrier)", "partner+own loads", "stretch", "lnprior", "trig (7 sqrt + sincos)", "cols + H0", "exp10", "pencil_P",
                                 "bin loop", "mix tail + llh", "accept test", "write-back", "arrive + draws + wait"};
        for (int i = 0; i < 13; ++i) printf("   stage %2d %-24s %8.0f\n", i, names[i], (double)gf_stage_acc[i] / (2.0 * A.nsteps));
        for (int i = 0; i < 16; ++i) gf_stage_acc[i] = 0;
    }
#endif
    if (writer) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int64_t k = c * A.nwalkers + h * half + w;
            for (int d = 0; d < ndim; ++d) A.pos[k * ndim + d] = pos_s[(h * T + wl) * ndim + d];
            A.lnp[k] = lnp_s[h * T + wl];
        }
        if (A.naccept) { /* owned by this thread */
            A.naccept[c * A.nwalkers + w] += (unsigned long long)acc0;
            A.naccept[c * A.nwalkers + half + w] += (unsigned long long)acc1;
        }
    }
    /* the last cluster.sync() of the step loop (or the one after the load) already guarantees that no
     * CTA reads a peer's shared memory after this point */
}

/* cluster geometry for (nchains, half): spread a chain as widely as possible -- down to one warp per CTA,
 * i.e. per SM (measured, 1024 walkers: 6.2 us / step at 4 CTAs x 128 threads, 5.4 at 16 x 32) -- as long
 * as the clusters of all chains fit the SMs at once; never more than 16 CTAs (the non-portable maximum)
 * or max_threads per CTA.  Returns false if the ensemble does not fit a cluster. */
static bool cluster_geometry(int64_t nchains, int half, int sms, int want_nc, int max_threads, int lanes, int* nc_out, int* threads_out) {
    /* `lanes` threads work on one walker pair: a CTA of `threads` threads holds threads / lanes pairs; warps stay whole,
     * so the pairs per CTA come in multiples of 32 / lanes */
    const int max_pairs = max_threads / lanes, quantum = 32 / lanes;
    auto pairs_for = [&](int nc_) { return (((half + nc_ - 1) / nc_) + quantum - 1) / quantum * quantum; };
    int nc = 1;
    if (want_nc > 0) {
        while (nc < want_nc && nc < 16) nc *= 2;
    } else {
        while (nc < 16 && (half + nc - 1) / nc > quantum && nchains * (nc * 2) <= (int64_t)sms) nc *= 2;
    }
    while (nc < 16 && (half + nc - 1) / nc > max_pairs) nc *= 2;
    const int pairs = pairs_for(nc);
    if (pairs > max_pairs) return false;
    while (nc > 1 && (nc / 2) * pairs >= half) nc /= 2; /* rounding to warps may have emptied the last CTAs */
    *nc_out = nc;
    *threads_out = pairs * lanes;
    return true;
}

template <int SPEC, int ILP>
static int launch_cluster(const gf_dev_model& d, const gf_ens_args& A, int nc, int threads, cudaStream_t st, bool* launched) {
    auto kern = k_ensemble_cluster<SPEC, ILP>;
    const int pairs = threads / GF_ENS_LANES(SPEC);
    const size_t smem = (size_t)(2 * pairs * d.ndim + 2 * pairs) * sizeof(double);
    GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (nc > 8) GF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(A.nchains * nc));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError(); /* this cluster size cannot be scheduled here: the caller falls back */
        *launched = false;
        return GF_OK;
    }
    GF_CUDA(cudaLaunchKernelEx(&cfg, kern, d, A));
    ++g_gf_launches;
    *launched = true;
    return GF_OK;
}

template <int SPEC, int ILP>
static int run_spec(const gf_dev_model& d, const gf_ens_args& A, const gf_ensemble_config* cfg, cudaStream_t st) {
    constexpr int LANES = GF_ENS_LANES(SPEC);
    const int half = cfg->nwalkers / 2;
    const int64_t total = cfg->nchains * half * LANES; /* threads of the grid-barrier shape */
    const unsigned blocks = gf_blocks_for(total, GF_ENS_THREADS);
    int sms = 0;
    if (int rc = gf_sm_count(&sms)) return rc;
    /* auto: the cluster shape whenever the ensemble fits one (<= 16 x 256 walker pairs, 16 x 128 pairs x 2 lanes on the BSM path) */
    if (cfg->mode == 3 || cfg->mode == 0) {
        int nc = 0, threads = 0;
        if (cluster_geometry(cfg->nchains, half, sms, cfg->cluster_blocks, GF_ENS_CL_MAX_CONSUMERS, LANES, &nc, &threads)) {
            for (;;) {
                bool launched = false;
                if (int rc = launch_cluster<SPEC, ILP>(d, A, nc, threads, st, &launched)) return rc;
                if (launched) {
                    GF_LAUNCH_CHECK("k_ensemble_cluster");
                    return GF_OK;
                }
                /* this cluster size cannot be scheduled on the device: halve it while the CTA still holds its share */
                nc /= 2;
                threads = nc >= 1 ? ((((half + nc - 1) / nc) * LANES) + 31) / 32 * 32 : 0;
                if (nc < 1 || threads > GF_ENS_CL_MAX_CONSUMERS) break;
            }
        }
        GF_REQUIRE(cfg->mode == 0, "gf_ensemble_run: mode 3 (cluster per chain) cannot hold %d walkers per chain", cfg->nwalkers);
    }
    /* block mode only when every thread owns ONE walker pair -- a second sequential pass per half-step
     * costs more than the grid barrier it saves (measured: 1024 walkers, 24 vs 12 us / step) */
    if (cfg->mode == 2 || (cfg->mode == 0 && half * LANES <= GF_ENS_BLOCK_MAX)) {
        const int threads = half * LANES >= GF_ENS_BLOCK_MAX ? GF_ENS_BLOCK_MAX : ((half * LANES + 31) / 32) * 32;
        k_ensemble_block<SPEC, ILP><<<(unsigned)cfg->nchains, threads, 0, st>>>(d, A);
        ++g_gf_launches;
        GF_LAUNCH_CHECK("k_ensemble_block");
        return GF_OK;
    }
    int dev = 0, coop = 0, per_sm = 0;
    GF_CUDA(cudaGetDevice(&dev));
    GF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    GF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ensemble<true, SPEC, ILP>, GF_ENS_THREADS, 0));
    if (coop && (int64_t)blocks <= (int64_t)sms * per_sm) {
        int64_t one_step = 0;
        int one_half = 0;
        void* params[] = {(void*)&d, (void*)&A, (void*)&one_step, (void*)&one_half};
        GF_CUDA(cudaLaunchCooperativeKernel((const void*)k_ensemble<true, SPEC, ILP>, dim3(blocks), dim3(GF_ENS_THREADS), params, 0, st));
        ++g_gf_launches;
        return GF_OK;
    }
    for (int64_t s = 0; s < cfg->nsteps; ++s) {
        for (int h = 0; h < 2; ++h) {
            k_ensemble<false, SPEC, ILP><<<blocks, GF_ENS_THREADS, 0, st>>>(d, A, s, h);
            ++g_gf_launches;
        }
        if ((A.chain || A.lnp_chain) && (s + 1) % cfg->thin == 0) {
            k_ensemble<false, SPEC, ILP><<<blocks, GF_ENS_THREADS, 0, st>>>(d, A, s, 2);
            ++g_gf_launches;
        }
        GF_LAUNCH_CHECK("k_ensemble");
    }
    return GF_OK;
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
    cudaStream_t st = (cudaStream_t)stream;
    GF_REQUIRE(cfg->mode >= 0 && cfg->mode <= 3, "gf_ensemble_run: mode = %d outside [0, 3]", cfg->mode);
    GF_REQUIRE(cfg->cluster_blocks >= 0 && cfg->cluster_blocks <= 16, "gf_ensemble_run: cluster_blocks = %d outside [0, 16]", cfg->cluster_blocks);
    /* the per-point log-posterior specialisation gf_lnprob launches for this model (NPFREE -> GENERIC) */
    const int spec = gf_model_spec(d);
    if (spec == GF_SPEC_SM) return run_spec<GF_SPEC_SM, 1>(d, A, cfg, st);
    if (spec == GF_SPEC_SM6) return run_spec<GF_SPEC_SM6, 1>(d, A, cfg, st);
    if (spec == GF_SPEC_FIXED) return run_spec<GF_SPEC_FIXED, GF_ENS_ILP_FIXED>(d, A, cfg, st);
    if (spec == GF_SPEC_FIXED7) return run_spec<GF_SPEC_FIXED7, GF_ENS_ILP_FIXED>(d, A, cfg, st);
    if (spec == GF_SPEC_FIXED12) return run_spec<GF_SPEC_FIXED12, GF_ENS_ILP_FIXED>(d, A, cfg, st);
    return run_spec<GF_SPEC_GENERIC, GF_ENS_ILP_GENERIC>(d, A, cfg, st);
}
