"""Developer timing of sens.evidence_grid (600 grid points x 10^6 prior samples)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from golemflavor_b200 import sens
sens.evidence_grid(dimensions=(6,), segments=5, samples=10000)
torch.cuda.synchronize(); t0 = time.perf_counter()
ev = sens.evidence_grid(segments=100, samples=1_000_000)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('evidence grid: %.3f s for %d points x 1e6 samples (%.3g samples/s); lnZ[6][:3] = %s' % (dt, sum(len(v) for v in ev.values()), 6e8 / dt, ev[6][:3, 1]))
