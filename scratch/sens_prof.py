"""cProfile of the host side of sens.sweep (config 5) + CUDA-event time of its kernels."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from golemflavor_b200 import sens
sens.sweep(segments=100, nwalkers=60, burnin=5, nsteps=5, distributed=False)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
sw = sens.sweep(segments=100, nwalkers=60, burnin=200, nsteps=1000, distributed=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print('sweep wall %.4f s' % dt)
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
