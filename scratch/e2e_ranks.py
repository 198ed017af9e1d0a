"""End-to-end host pipeline under torchrun: per-rank input rate of gf_lnprob_host against the CONCURRENT plain pinned-copy rate
(all ranks copy at the same time after a barrier).  GF_HOST_SLOTS / GF_HOST_CHUNK_LOG2 select the ring geometry."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, torch.distributed as dist, models
from golemflavor_b200 import llh
from golemflavor_b200.enums import Texture
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
args, asimov, pset = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
fn = llh.LnProb(args, asimov, pset)
n = 1 << 22
from golemflavor_b200 import _lib
if os.environ.get('E2E_WC', '0') == '1':   # theta in write-combined pinned memory from the library
    hb = _lib.HostBuffer((n, 7), write_combined=True)
    hb.array[...] = models.draw_in_ranges(pset, n, np.random.default_rng(25 + rank))
    th = torch.from_numpy(hb.array)
else:
    th = torch.as_tensor(models.draw_in_ranges(pset, n, np.random.default_rng(25 + rank))).pin_memory()
out = torch.empty(n, dtype=torch.float64).pin_memory()
dev = torch.empty_like(th, device='cuda')
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device='cuda')
    if world == 1: return [x]
    outl = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outl, t)
    return [float(o) for o in outl]
for _ in range(2): fn.evaluate_host(th.numpy(), out=out.numpy())
res = []
for rep in range(3):
    barrier(); t0 = time.perf_counter()
    for _ in range(10): fn.evaluate_host(th.numpy(), out=out.numpy())
    dt = time.perf_counter() - t0
    res.append(10 * n * 56 / dt / 1e9)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dev.copy_(th, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    link = 10 * n * 56 / (e0.elapsed_time(e1) * 1e-3) / 1e9
pipe, links = gather(max(res)), gather(link)
if rank == 0:
    print('wc %s slots %s chunk 2^%s world %d: pipeline h2d GB/s per rank min %.1f mean %.1f | concurrent plain copy min %.1f mean %.1f | ratio of means %.3f' % (
        os.environ.get('E2E_WC', '0'), os.environ.get('GF_HOST_SLOTS', '3'), os.environ.get('GF_HOST_CHUNK_LOG2', '18'), world, min(pipe), np.mean(pipe), min(links), np.mean(links), np.mean(pipe) / np.mean(links)), flush=True)
if world > 1: dist.destroy_process_group()
