"""2+ ranks: bx_allgather_detections (C ABI, framework's ncclComm_t) against torch.distributed.all_gather_into_tensor.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/_allgather_ranks.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_eager_object_detection_b200 import distributed as bxd
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
b = 3 + (rank % 2)                                   # uneven shards
rec = torch.randn((b, 150, 6), device=dev, generator=g)
cnt = torch.randint(0, 151, (b,), device=dev, dtype=torch.int32, generator=g)
assert bxd.nccl_comm_ptr() is not None, 'no ncclComm_t from the process group'
ra, ca = bxd.allgather_detections(rec, cnt)                                   # packed torch collective, sizes exchanged
sizes = [3 + (r % 2) for r in range(world)]
rs, cs = bxd.allgather_detections(rec, cnt, sizes=sizes)                      # static sizes: no host sync
grp = bxd.new_detection_group()                                                # dedicated communicator for the C-ABI path
rn, cn = bxd.allgather_detections(rec, cnt, sizes=sizes, native_group=grp)    # bx_allgather_detections
# the native call on a side stream, interleaved with torch collectives on the default group (different communicators)
side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
pending = dist.all_reduce(torch.ones(1 << 20, device=dev), async_op=True)
with torch.cuda.stream(side):
    rn2, cn2 = bxd.allgather_detections(rec, cnt, sizes=sizes, native_group=grp)
pending.wait()
torch.cuda.current_stream(dev).wait_stream(side)
# reference: plain torch collectives on padded blocks
bmax = 4
rp = torch.cat([rec, rec.new_zeros((bmax - b, 150, 6))]); cp = torch.cat([cnt, cnt.new_zeros((bmax - b,))])
r2 = torch.empty((world * bmax, 150, 6), device=dev); c2 = torch.empty((world * bmax,), device=dev, dtype=torch.int32)
dist.all_gather_into_tensor(r2, rp); dist.all_gather_into_tensor(c2, cp)
keep = torch.cat([torch.arange(r * bmax, r * bmax + 3 + (r % 2), device=dev) for r in range(world)])
for name, (x, y) in dict(packed=(ra, ca), static=(rs, cs), native=(rn, cn), native_side=(rn2, cn2)).items():
    assert torch.equal(x, r2[keep]) and torch.equal(y, c2[keep]), name
torch.cuda.synchronize()
if rank == 0:
    print('allgather ok: world %d, %d images, records %s' % (world, ra.shape[0], tuple(ra.shape)), flush=True)
dist.destroy_process_group()
