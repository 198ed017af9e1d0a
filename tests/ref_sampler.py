"""NumPy restatement of the device sampler's convention (include/golemflavor_b200.h,
gf_ensemble_config): emcee's red/blue stretch move (emcee EnsembleSampler._propose_stretch /
StretchMove, as driven by golemflavor/mcmc.py:27-53) with the counter-based Philox draws of the
device kernel.  TEST INFRASTRUCTURE: lets the tests replay a device chain step by step with any
batched log-posterior (the oracle's, the host harness', or the CUDA kernel's)."""

import numpy as np

from oracle import golem_oracle as go


def _uniforms(seed, gids, step):
    ctr = np.zeros((len(gids), 4), dtype=np.uint32)
    ctr[:, 0] = np.asarray(gids, dtype=np.uint64).astype(np.uint32)
    ctr[:, 1] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32((step >> 32) & 0xFFFFFFFF)
    x = go.philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return (x.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def run(lnprob_fn, pos, lnp, nsteps, nfree=None, step0=0, a=2.0, seed=0, chain_index=0):
    """pos[nwalkers, ndim], lnp[nwalkers] -> (pos, lnp, chain[nwalkers, nsteps, ndim], naccept[nwalkers])."""
    pos = np.array(pos, dtype=np.float64)
    lnp = np.array(lnp, dtype=np.float64)
    k, ndim = pos.shape
    half = k // 2
    nfree = ndim if nfree is None else nfree
    chain = np.zeros((k, nsteps, ndim))
    nacc = np.zeros(k, dtype=np.int64)
    for s in range(nsteps):
        for h in (0, 1):
            idx = np.arange(h * half, (h + 1) * half)
            u = _uniforms(seed, chain_index * k + idx, step0 + s)
            t = (a - 1.0) * u[:, 0] + 1.0
            z = (t * t) / a
            j = np.minimum((u[:, 1] * half).astype(np.int64), half - 1) + (1 - h) * half
            c = pos[j]
            q = c - z[:, None] * (c - pos[idx])
            new = np.asarray(lnprob_fn(q), dtype=np.float64)
            with np.errstate(invalid='ignore', divide='ignore'):
                accept = (nfree - 1.0) * np.log(z) + new - lnp[idx] > np.log(u[:, 2])
            pos[idx[accept]] = q[accept]
            lnp[idx[accept]] = new[accept]
            nacc[idx[accept]] += 1
        chain[:, s] = pos
    return pos, lnp, chain, nacc
