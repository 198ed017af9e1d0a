"""world_size-2 tests of the N>1 scan path on CPU (gloo): index-range sharding + the single
all-reduce of the histograms.  The per-sample work is done by the host harness (test
infrastructure) with the kernels' own draw / physics / bin-index functions, so the test checks the
property the GPU path relies on: the summed histogram is bit-identical to the single-rank one."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _local_hist(fm, start, n, nb, seed):
    import host_harness as hh
    theta = hh.draw(fm, seed, start, n)
    fr, _ = hh.fr(fm, theta)
    return hh.hist(fr, nb)


def _worker(rank, world, port, count, nb, seed, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from golemflavor_b200 import scan
    fm = scan.scan_model('unitary')
    start, n = scan.shard_range(count, rank, world, first_index=123)
    hist = torch.as_tensor(_local_hist(fm, start, n, nb, seed).reshape(-1))
    kept = torch.tensor([int(hist.sum())])
    scan.allreduce_counts(hist, kept)
    np.save(os.path.join(out_dir, 'hist_%d.npy' % rank), hist.numpy())
    np.save(os.path.join(out_dir, 'kept_%d.npy' % rank), kept.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_scan_matches_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from golemflavor_b200 import scan
    count, nb, seed = 20001, 25, 26
    mp.spawn(_worker, args=(2, _free_port(), count, nb, seed, str(tmp_path)), nprocs=2, join=True)
    full = _local_hist(scan.scan_model('unitary'), 123, count, nb, seed).reshape(-1)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ('hist_%d.npy' % r)), full)
        assert int(np.load(tmp_path / ('kept_%d.npy' % r))[0]) == count == full.sum()


def test_allreduce_is_noop_without_process_group():
    from golemflavor_b200 import scan
    t = torch.arange(5)
    assert scan.allreduce_counts(t)[0] is t and t.tolist() == [0, 1, 2, 3, 4]
