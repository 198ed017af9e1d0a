/*
 * gf_ensemble_dev.cuh -- one stretch-move update of the device-resident ensemble sampler
 * (see gf_ensemble.cu and gf_ensemble_config in the C header for the RNG / proposal convention).
 */
#ifndef GF_ENSEMBLE_DEV_CUH
#define GF_ENSEMBLE_DEV_CUH

#include "gf_scan_dev.cuh"

/* round-to-nearest operations that the compiler must not contract into FMAs, and L2 loads */
#ifdef __CUDA_ARCH__
#define GF_ADD_RN(a, b) __dadd_rn(a, b)
#define GF_SUB_RN(a, b) __dsub_rn(a, b)
#define GF_MUL_RN(a, b) __dmul_rn(a, b)
#define GF_DIV_RN(a, b) __ddiv_rn(a, b)
#define GF_LDCG(p) __ldcg(p)
#else /* host instantiation (tests/host_harness, built with -ffp-contract=off) */
#define GF_ADD_RN(a, b) ((a) + (b))
#define GF_SUB_RN(a, b) ((a) - (b))
#define GF_MUL_RN(a, b) ((a) * (b))
#define GF_DIV_RN(a, b) ((a) / (b))
#define GF_LDCG(p) (*(p))
#endif

struct gf_ens_args {
    int64_t nchains, nsteps, step0, thin, chain0;
    int32_t nwalkers, nfree;
    double a;
    uint64_t seed;
    double* pos;
    double* lnp;
    double* chain;
    double* lnp_chain;
    unsigned long long* naccept;
};

/* The three Philox words of one update (see gf_ensemble_config): stretch factor z, partner index j in
 * the other half, and the uniform of the acceptance test.
 *
 * Latency matters more than throughput here (about one warp per SM sub-partition, in-order issue): the
 * update is written so that the partner's coordinates are REQUESTED first -- all dimensions back to back,
 * not one round trip per dimension -- and the division and the two logarithms execute while they are in
 * flight. */
struct gf_ens_draw {
    double u_z, u_accept;
    int j;
    double z, lz, lu; /* filled by gf_ens_finish_draw: stretch factor, (nfree-1) ln z, ln u_accept */
};

GF_HD gf_ens_draw gf_ens_draws(const gf_ens_args& A, uint64_t gid, int64_t step, int half) {
    const uint64_t s = (uint64_t)step;
    const gf_u4 r = gf_philox4x32_10((uint32_t)gid, (uint32_t)s, (uint32_t)(s >> 32), 0u, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
    gf_ens_draw d;
    d.u_z = gf_u01(r.x);
    int j = (int)(gf_u01(r.y) * (double)half);
    d.j = j < half ? j : half - 1;
    d.u_accept = gf_u01(r.z);
    return d;
}

/* z = ((a-1) u + 1)^2 / a without FMA contraction: bit-reproducible with NumPy.  Dividing by a power of
 * two (emcee's default a = 2) is an exact multiplication. */
GF_HD double gf_ens_z(const gf_ens_args& A, double u) {
    const double t = GF_ADD_RN(GF_MUL_RN(A.a - 1.0, u), 1.0);
    const double tt = GF_MUL_RN(t, t);
    return A.a == 2.0 ? GF_MUL_RN(tt, 0.5) : GF_DIV_RN(tt, A.a);
}

/* the arithmetic on the draws that does not depend on any walker position */
GF_HD void gf_ens_finish_draw(const gf_ens_args& A, gf_ens_draw& dr) {
    dr.z = gf_ens_z(A, dr.u_z);
    dr.lz = (double)(A.nfree - 1) * gfp_log_pos(dr.z); /* z in [1/a, a], u in (0, 1): positive and normal */
    dr.lu = gfp_log_pos(dr.u_accept);
}

/* q = c_j - z (c_j - p), contraction-free */
GF_HD double gf_ens_stretch(double cd, double pd, double z) { return GF_SUB_RN(cd, GF_MUL_RN(z, GF_SUB_RN(cd, pd))); }

/*
 * One stretch-move update given accessors for the partner's and the walker's own coordinates
 * (global memory through L2, or distributed shared memory).  Returns true and leaves the proposal in q /
 * its log-posterior in lnew when the move is accepted: accept iff (nfree-1) ln z + lnp(q) - lnp(p) > ln u
 * (false for NaN).  FINISHED: gf_ens_finish_draw has already run (the cluster kernel does it in the shadow
 * of the barrier); otherwise it runs here, in the shadow of the loads.
 */
/* LANES = 2: lanes 2w and 2w + 1 of a warp run the same update (same draws, same proposal) and share the energy bins of
 * its log-posterior (gf_bin_loop); both obtain the same decision.  `lane` = threadIdx.x & 1. */
#ifndef GF_ENS_BSM_LANES
#define GF_ENS_BSM_LANES 2 /* developer builds: 1 = one thread per walker pair on the BSM path as well */
#endif
#define GF_ENS_LANES(SPEC) (GF_SPEC_IS_SM(SPEC) ? 1 : GF_ENS_BSM_LANES)

template <int SPEC, int ILP, bool FINISHED, int LANES = 1, class LoadPartner, class LoadOwn>
GF_HD bool gf_ens_move(const gf_dev_model& m, const gf_ens_args& A, gf_ens_draw& dr, LoadPartner partner, LoadOwn own, double lold,
                       double* q, double& lnew, int lane = 0) {
    /* with a compile-time layout the dimension count is a constant: no guarded work on the 16 - ndim unused slots */
    constexpr int ND = GF_SPEC_STATIC_NDIM(SPEC);
    const int ndim = ND > 0 ? ND : m.ndim;
    double cv[GF_MAX_DIM], pv[GF_MAX_DIM];
    GF_STAGE(0);
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : GF_MAX_DIM); ++d) {
        if (d < ndim) {
            cv[d] = partner(d);
            pv[d] = own(d);
        }
    }
    if (!FINISHED) gf_ens_finish_draw(A, dr);
    GF_STAGE(1);
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : GF_MAX_DIM); ++d)
        if (d < ndim) q[d] = gf_ens_stretch(cv[d], pv[d], dr.z);
    GF_STAGE(2);
    double fr[3];
    unsigned st = 0u;
    lnew = gf_point_lnprob<SPEC, ILP, LANES>(m, [&](int d) { return q[d]; }, fr, st, lane);
    const double diff = dr.lz + lnew - lold;
    GF_STAGE(10);
    return diff > dr.lu;
}

/* one stretch-move update of walker k (in half h) of chain c, positions in global memory */
template <int SPEC = GF_SPEC_GENERIC, int ILP = 1, int LANES = 1>
GF_HD unsigned gf_ens_update(const gf_dev_model& m, const gf_ens_args& A, int64_t c, int k, int h, int64_t step, int lane = 0) {
    const int ndim = m.ndim, half = A.nwalkers / 2;
    const uint64_t gid = (uint64_t)(A.chain0 + c) * (uint64_t)A.nwalkers + (uint64_t)k;
    gf_ens_draw dr = gf_ens_draws(A, gid, step, half);
    double* p = A.pos + (c * A.nwalkers + k) * ndim;
    const double* cj = A.pos + (c * A.nwalkers + (1 - h) * half + dr.j) * ndim;
    double q[GF_MAX_DIM];
    double lnew;
    /* positions of other walkers were written by other SMs before the last grid barrier: read them
     * through L2 (ld.global.cg), not through this SM's non-coherent L1 */
    const bool accept = gf_ens_move<SPEC, ILP, false, LANES>(
        m, A, dr, [&](int d) { return GF_LDCG(cj + d); }, [&](int d) { return GF_LDCG(p + d); }, GF_LDCG(A.lnp + c * A.nwalkers + k), q, lnew, lane);
    if (accept && lane == 0) { /* with two lanes per walker both hold the same proposal and decision: one writes */
        _Pragma("unroll") for (int d = 0; d < (GF_SPEC_STATIC_NDIM(SPEC) > 0 ? GF_SPEC_STATIC_NDIM(SPEC) : GF_MAX_DIM); ++d)
                        if (d < ndim) p[d] = q[d]; /* static indices keep q in registers */
        A.lnp[c * A.nwalkers + k] = lnew;
    }
    return (accept && lane == 0) ? 1u : 0u;
}

GF_HD void gf_ens_store(const gf_dev_model& m, const gf_ens_args& A, int64_t c, int k, int64_t slot, int64_t nstore) {
    const int ndim = m.ndim;
    const double* p = A.pos + (c * A.nwalkers + k) * ndim;
    if (A.chain) {
        double* o = A.chain + ((c * A.nwalkers + k) * nstore + slot) * ndim;
        for (int d = 0; d < ndim; ++d) o[d] = p[d];
    }
    if (A.lnp_chain) A.lnp_chain[(c * A.nwalkers + k) * nstore + slot] = A.lnp[c * A.nwalkers + k];
}

#endif /* GF_ENSEMBLE_DEV_CUH */
