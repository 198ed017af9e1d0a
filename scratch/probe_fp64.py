"""Developer probe: fp64 pipe throughput by operand pattern (see gf_fp64_peak_probe)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from golemflavor_b200 import _lib
lib = _lib.load()
sink = torch.zeros(8, dtype=torch.float64, device='cuda')
flops = C.c_double()
for mode in (0, 1, 2, 3):
    best = 0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.gf_fp64_peak_probe(mode, 100000, _lib.ptr(sink), C.byref(flops), None))
        e1.record(); torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print('mode', mode, 'TFLOP/s %.2f' % best, ' warp-inst/clk/SMSP @1.965GHz: %.3f' % (best * 1e12 / (1 if mode == 2 else 2) / 32 / 592 / 1.965e9))
