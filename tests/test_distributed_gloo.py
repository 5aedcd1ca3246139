"""World-size-2 (and 3, uneven) gloo runs of the multi-GPU host logic on CPU: image sharding covers the batch exactly
once and the detection all-gather reproduces the unsharded result byte for byte on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_eager_object_detection_b200 import distributed as bxd


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full(num_images, kmax):
    rng = np.random.default_rng(123)
    counts = rng.integers(0, kmax + 1, num_images).astype(np.int32)
    rec = rng.standard_normal((num_images, kmax, 6)).astype(np.float32)
    for i, c in enumerate(counts):
        rec[i, c:] = 0
    return torch.from_numpy(rec), torch.from_numpy(counts)


def _worker(rank, world, port, num_images, kmax, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rec, counts = _full(num_images, kmax)
        loc_rec, loc_cnt = bxd.shard_images([rec, counts], rank, world)
        all_rec, all_cnt = bxd.allgather_detections(loc_rec.contiguous(), loc_cnt.contiguous())
        ok = torch.equal(all_rec, rec) and torch.equal(all_cnt, counts)
        sizes = [hi - lo for lo, hi in (bxd.shard_bounds(num_images, r, world) for r in range(world))]
        rec2, cnt2 = bxd.allgather_detections(loc_rec.contiguous(), loc_cnt.contiguous(), sizes=sizes)   # static sizes
        ok = ok and torch.equal(rec2, rec) and torch.equal(cnt2, counts)
        try:
            bxd.allgather_detections(loc_rec.contiguous(), loc_cnt.contiguous(), sizes=[1] * world if num_images != world else [2] * world)
            ok = False                                           # wrong sizes must be rejected
        except ValueError:
            pass
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,num_images', [(2, 16), (2, 7), (3, 8)])
def test_sharded_allgather_equals_unsharded(world, num_images):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_images, 30, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


def test_shard_bounds_partition_the_batch():
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 4, 8):
            spans = [bxd.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_detections_layout():
    b = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    s = torch.ones(2, 3) * 0.5
    lab = torch.full((2, 3), 7, dtype=torch.int32)
    r = bxd.pack_detections(b, s, lab)
    assert r.shape == (2, 3, 6) and r[1, 2].tolist() == [20.0, 21.0, 22.0, 23.0, 0.5, 7.0]
