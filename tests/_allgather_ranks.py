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


# ---- evaluation.get_prediction_files over two ranks: uneven shards of the three golden images, files written by rank 0
import tempfile
import numpy as np
from oracle.voc_fixture import eval_loop_inputs
from tf_eager_object_detection_b200 import evaluation as ev
imgs = eval_loop_inputs()
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)


class _Model:
    def im_detect_batched(self, i):
        im = imgs[i]
        return (cu(im['scores'])[None], cu(im['deltas'].reshape(300, -1))[None], cu(im['rois'])[None],
                torch.tensor([300], dtype=torch.int32, device=dev))


lo, hi = bxd.shard_bounds(len(imgs), rank, world)
names = ['%06d' % (i + 1) for i in range(len(imgs))]
with tempfile.TemporaryDirectory() as d:
    rec, cnt = ev.get_prediction_files(_Model(), [(i, imgs[i]['scale'], imgs[i]['raw_h'], imgs[i]['raw_w']) for i in range(lo, hi)],
                                       names, os.path.join(d, '{:s}.txt'), score_threshold=0.05, iou_threshold=0.3,
                                       max_objects_per_class=50, max_objects_per_image=50, min_size=10)
    assert cnt.tolist() == [50, 50, 76]
    if rank == 0:
        golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_on_shim.npz'))
        got = ''.join(open(os.path.join(d, '%s.txt' % c)).read() for c in ev.PASCAL_CLASSES[1:]).splitlines()
        ref = bytes(golden['eval_voc_files']).decode().splitlines()
        assert len(got) == len(ref) and all(x.split()[:2] == y.split()[:2] for x, y in zip(got, ref))
        print('prediction files ok: %d lines from %d ranks' % (len(got), world), flush=True)
torch.cuda.synchronize()
dist.destroy_process_group()
