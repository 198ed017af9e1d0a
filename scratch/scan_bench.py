"""Developer timing of the fused scan kernel for the four modes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from golemflavor_b200 import scan
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['unitary', 'x', 'texture', 'anarchic']
for mode in modes:
    for nb in (25, 200):
        fm = scan.scan_model(mode)
        scan.scan_histogram(fm, 10 ** 6, nb=nb, distributed=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h, kept = scan.scan_histogram(fm, n, nb=nb, distributed=False, return_tensor=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('%-9s nb=%3d  %.3g samples/s  (%.1f ms, kept %d)' % (mode, nb, n / ms * 1e3, ms, int(kept)))
