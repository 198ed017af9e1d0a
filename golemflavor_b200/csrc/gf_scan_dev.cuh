/*
 * gf_scan_dev.cuh -- per-sample pieces of the Monte-Carlo scans: the Philox4x32-10 counter-based
 * generator, the prior draws and the bit-exact np.histogramdd bin index.
 */
#ifndef GF_SCAN_DEV_CUH
#define GF_SCAN_DEV_CUH

#include "gf_model.cuh"

/* ------------------------------------------------------------------ Philox4x32-10 */

struct gf_u4 {
    uint32_t x, y, z, w;
};

GF_HD gf_u4 gf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return gf_u4{c0, c1, c2, c3};
}

GF_HD double gf_u01(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

/* The inverse normal CDF is a large routine (rational approximations with dozens of fp64 immediates):
 * keep ONE out-of-line copy per kernel instead of one per unrolled parameter slot -- the scan kernels
 * were 88 KB of code with 83 % instruction-cache hit rate before. */
#ifdef __CUDA_ARCH__
__device__ __noinline__ double gf_normcdfinv(double p) { return normcdfinv(p); }
#else
inline double gf_normcdfinv(double p) { return NAN * p; } /* evaluated on the device only */
#endif

/* Draw parameter k of sample `index` from its prior (see gf_scan_config in the C header). */
GF_HD double gf_draw_dim(const gf_dev_model& m, int k, double u) {
    if (m.kind[k] == GF_PRIOR_UNIFORM) return fma(u, m.hi[k] - m.lo[k], m.lo[k]);
    const double p = fma(u, m.cdf_span[k], m.cdf_lo[k]);
    const double x = fma(m.sigma[k], gf_normcdfinv(p), m.mu[k]);
    return fmin(fmax(x, m.lo[k]), m.hi[k]);
}

/* UNROLL: fully unrolled with an early exit on the (uniform) dimension count -- with a compile-time k the
 * prior tables are constant-bank operands and the word of the Philox block is a register, not a select
 * chain: +12 % on the unitary scan, +8 % on the x scan.  The BSM scan kernels keep the rolled loop: they sit
 * at the 128-register limit and the unrolled draws made them spill. */
template <bool UNROLL = false, int NDIM = 0>
GF_HD void gf_draw_theta(const gf_dev_model& m, uint64_t seed, uint64_t index, double* theta) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32);
    gf_u4 r = {0u, 0u, 0u, 0u};
    if constexpr (UNROLL) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < (NDIM > 0 ? NDIM : GF_MAX_DIM); ++k) {
            if (NDIM == 0 && k >= m.ndim) break;
            const int j = k & 3;
            if (j == 0) r = gf_philox4x32_10(c0, c1, (uint32_t)(k >> 2), 0u, k0, k1);
            const uint32_t w = j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w;
            theta[k] = gf_draw_dim(m, k, gf_u01(w));
        }
    } else {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (int k = 0; k < m.ndim; ++k) {
            const int j = k & 3;
            if (j == 0) r = gf_philox4x32_10(c0, c1, (uint32_t)(k >> 2), 0u, k0, k1);
            const uint32_t w = j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w;
            theta[k] = gf_draw_dim(m, k, gf_u01(w));
        }
    }
}

/* ------------------------------------------------------------------ bit-exact np.histogramdd bin */

/* Index of x in np.linspace(0, 1, nb1 + 1) bins with the semantics of np.histogramdd
 * (searchsorted side='right', x == 1 joins the last bin, everything else outside is dropped):
 * edges e_k = k * step for k < nb1 and e_nb1 = 1.0 exactly (numpy linspace: arange * step, end
 * point overwritten).  Returns -1 for dropped values. */
GF_HD int gf_bin_index(double x, int nb1, double step) {
    if (!(x >= 0.0 && x <= 1.0)) return -1;
    int k = (int)(x * (double)nb1);
    k = k > nb1 - 1 ? nb1 - 1 : k;
    while (k > 0 && x < (double)k * step) --k;
    while (k < nb1 - 1 && x >= (double)(k + 1) * step) ++k;
    return k;
}

GF_HD int gf_cell_index(const double* fr, int nb1, double step) {
    const int b0 = gf_bin_index(fr[0], nb1, step);
    const int b1 = gf_bin_index(fr[1], nb1, step);
    const int b2 = gf_bin_index(fr[2], nb1, step);
    if ((b0 | b1 | b2) < 0) return -1;
    return (b0 * nb1 + b1) * nb1 + b2;
}

#endif /* GF_SCAN_DEV_CUH */
