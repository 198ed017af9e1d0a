"""Multi-GPU check run under torchrun (not collected by pytest): every rank computes the sharded scan
with the NCCL all-reduce and compares it bit-for-bit with the single-rank scan it computes itself.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_scan_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from golemflavor_b200 import scan

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok = True
for mode, count, nb in (('unitary', 50_000_001, 25), ('texture', 20_000_003, 25), ('anarchic', 5_000_001, 200)):
    fm = scan.scan_model(mode)
    sharded, kept = scan.scan_histogram(fm, count, nb=nb, seed=26, return_tensor=True)
    single, kept1 = scan.scan_histogram(fm, count, nb=nb, seed=26, distributed=False, return_tensor=True)
    same = bool(torch.equal(sharded, single)) and int(kept) == int(kept1) == count
    ok = ok and same
    if dist.get_rank() == 0:
        print('%-9s count=%d nb=%d world=%d bit-identical=%s' % (mode, count, nb, dist.get_world_size(), same))
# the evidence grid: samples sharded over the ranks, (max, sum-exp) merged with two all-reduces -- against one rank alone
import numpy as np
from argparse import Namespace
from golemflavor_b200 import model, sens
from golemflavor_b200.enums import Texture
from golemflavor_b200.param import ParamSet
args = Namespace(source_ratio=np.array([1, 2, 0.]) / 3, dimension=6, texture=Texture.OET, binning=scan.DEFAULT_BINNING, no_bsm=False,
                 injected_ratio=[1 / 3, 1 / 3, 1 / 3], smearing=0.02, fixed_scale=-100.0)
fm = model.flatten(args, None, ParamSet(scan.sm_paramset(with_mass=True)))
scales = sens.scale_grid(6, 30)
sharded = scan.scan_evidence_grid(fm, scales, 400_003, seed=26)
single = scan.scan_evidence_grid(fm, scales, 400_003, seed=26, distributed=False)
rel = float(np.max(np.abs(sharded - single) / np.abs(single)))
same = rel < 1e-12
ok = ok and same
if dist.get_rank() == 0:
    print('evidence grid: 30 scales x 400003 samples world=%d max rel |lnZ sharded - single| = %.2e ok=%s' % (dist.get_world_size(), rel, same))
# the sensitivity sweep: grid points split over the ranks, one all-reduce of the summaries -- rank-count invariant
sw = sens.sweep(dimensions=(3, 6), segments=8, nwalkers=60, burnin=20, nsteps=40)
sw1 = sens.sweep(dimensions=(3, 6), segments=8, nwalkers=60, burnin=20, nsteps=40, distributed=False)
same = all(np.array_equal(sw[k], sw1[k]) for k in sw)
ok = ok and same
if dist.get_rank() == 0:
    print('sens sweep: 16 chains world=%d identical to the single-rank sweep=%s' % (dist.get_world_size(), same))
flag = torch.tensor([int(ok)], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
