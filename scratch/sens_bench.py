"""Developer timing of the config-5 sensitivity sweep (600 grid points x 60 walkers x 1200 steps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from golemflavor_b200 import sens
sens.sweep(segments=100, nwalkers=60, burnin=5, nsteps=5, distributed=False)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sw = sens.sweep(segments=100, nwalkers=60, burnin=200, nsteps=1000, distributed=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('C5 sweep: %.4f s, %d grid points, acceptance %.4f, checksum %.10e' % (dt, len(sw['scale']), sw['acceptance'].mean(), float(np.nansum(sw['mean_lnprob']))))
