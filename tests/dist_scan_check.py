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
flag = torch.tensor([int(ok)], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
