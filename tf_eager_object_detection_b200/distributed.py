"""Image-level data parallelism for the box path: every function on the path is per image, so a batch is
split into contiguous blocks of images, one block per GPU, with NO collective on the data path.  The only exchange is
the evaluation-time all-gather of fixed-size, padded per-image detection records, which replaces the reference's
single-process accumulation loops (evaluation/pascal_eval_files_utils.py:73-107, scripts/eval_coco.py:116-164).

torch.distributed is plumbing only (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist

__all__ = ['shard_bounds', 'shard_images', 'pack_detections', 'allgather_detections', 'allgather_native',
           'new_detection_group', 'nccl_comm_ptr']


def shard_bounds(num_images, rank, world_size):
    """Contiguous block [lo, hi) of images owned by `rank`; blocks differ in size by at most one image."""
    base, rem = divmod(int(num_images), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_images(tensors, rank, world_size):
    """Slice the leading (image) axis of each tensor to this rank's block."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world_size)
    return [t[lo:hi] for t in tensors]


def pack_detections(boxes, scores, labels):
    """[b,k,4], [b,k], [b,k] -> fp32 records [b,k,6] = (x1,y1,x2,y2,score,label): the layout that is all-gathered."""
    return torch.cat([boxes.to(torch.float32), scores.to(torch.float32).unsqueeze(-1),
                      labels.to(torch.float32).unsqueeze(-1)], dim=-1).contiguous()


def nccl_comm_ptr(group=None):
    """Raw ncclComm_t of the group's NCCL backend on the current CUDA device, or None (gloo / no NCCL / not initialised /
    a torch build without ProcessGroupNCCL._comm_ptr)."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    try:
        backend = pg._get_backend(torch.device('cuda', torch.cuda.current_device()))
    except (RuntimeError, ValueError):           # no CUDA backend registered for this group (gloo)
        return None
    get = getattr(backend, '_comm_ptr', None)
    if get is None:
        return None
    ptr = get()
    return int(ptr) if ptr else None


def new_detection_group():
    """A process group (and hence an ncclComm_t) used by NOTHING but bx_allgather_detections.  NCCL requires the
    operations of one communicator to be issued in one order on every rank; the native all-gather enqueues on the raw
    communicator behind ProcessGroupNCCL's back, so it must never share a communicator with torch collectives that may be
    in flight (async_op=True, or issued from other streams).  A dedicated group makes that structural."""
    return dist.new_group(ranks=list(range(dist.get_world_size())), backend='nccl')


def allgather_native(records, counts, rec_all, cnt_all, group, stream=None):
    """bx_allgather_detections (C ABI) on `group`'s ncclComm_t: records + counts in ONE fused NCCL group, enqueued on
    `stream` (default: torch's current stream).  `group` should come from new_detection_group().  Raises when the group is
    not NCCL-backed; never falls back."""
    import ctypes
    from . import _lib
    from ._tensor import stream_ptr
    if not (records.is_cuda and records.dtype == torch.float32 and counts.dtype == torch.int32 and records.dim() == 3
            and records.is_contiguous() and counts.is_contiguous()):
        raise TypeError('allgather_native: records must be contiguous CUDA fp32 [b,k,f], counts contiguous int32 [b]')
    comm = nccl_comm_ptr(group)
    if comm is None:
        raise RuntimeError('allgather_native: the process group has no NCCL communicator on this device')
    world = dist.get_world_size(group)
    dev = records.device.index
    st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else stream_ptr(dev)
    h = _lib.handle(dev, st.value)
    b, k, f = records.shape
    _lib.check(_lib.load().bx_allgather_detections(h, ctypes.c_void_p(comm), records.data_ptr(), counts.data_ptr(), b, k, f,
                                                   world, rec_all.data_ptr(), cnt_all.data_ptr(), st))
    return rec_all, cnt_all


def allgather_detections(records, counts, group=None, sizes=None, native_group=None):
    """records [b_local, kmax, f] fp32 (zero padded), counts [b_local] int32 ->
    (records_all [b_total, kmax, f], counts_all [b_total]) in rank order, identical on every rank.

    `sizes`: the number of images every rank owns (a list of world_size ints, e.g. from shard_bounds) — static shard
    sizes make the call free of host synchronisation.  Without it the sizes are exchanged first (one small all-gather and
    one device->host read per call).  Uneven shards are padded to the largest for a single fixed-size all-gather and the
    padding is dropped afterwards.

    Transport: ONE torch.distributed all_gather_into_tensor of a packed [b, kmax*f + 1] buffer (the int32 counts travel
    bit-cast in the last column) on `group`, ordered by ProcessGroupNCCL like every other torch collective.  Pass
    `native_group=new_detection_group()` to use the C-ABI entry point bx_allgather_detections on that dedicated
    communicator instead (no packing copy; see new_detection_group for why it must be dedicated)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records, counts
    world = dist.get_world_size(group)
    b_local = records.shape[0]
    if sizes is None:
        nloc = torch.tensor([b_local], dtype=torch.int32, device=records.device)
        szt = torch.empty((world,), dtype=torch.int32, device=records.device)
        dist.all_gather_into_tensor(szt, nloc, group=group)
        sizes = szt.tolist()                                   # the one host sync of the dynamic-size form
    sizes = [int(v) for v in sizes]
    if len(sizes) != world or sizes[dist.get_rank(group)] != b_local:
        raise ValueError('allgather_detections: sizes %s do not match world size %d / local block %d' % (sizes, world, b_local))
    bmax = max(sizes)
    pad = bmax - b_local
    if pad:
        records = torch.cat([records, records.new_zeros((pad,) + tuple(records.shape[1:]))])
        counts = torch.cat([counts, counts.new_zeros((pad,))])
    k, f = records.shape[1], records.shape[2]
    if native_group is not None:
        rec_all = torch.empty((world * bmax, k, f), dtype=records.dtype, device=records.device)
        cnt_all = torch.empty((world * bmax,), dtype=counts.dtype, device=counts.device)
        allgather_native(records.contiguous(), counts.contiguous(), rec_all, cnt_all, native_group)
    else:
        packed = torch.cat([records.reshape(bmax, k * f), counts.to(torch.int32).view(torch.float32).reshape(bmax, 1)], dim=1)
        out = torch.empty((world * bmax, k * f + 1), dtype=torch.float32, device=records.device)
        dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
        rec_all = out[:, :k * f].reshape(world * bmax, k, f)
        cnt_all = out[:, k * f].contiguous().view(torch.int32)
    if all(s_ == bmax for s_ in sizes):
        return rec_all, cnt_all
    keep = torch.cat([torch.arange(r * bmax, r * bmax + s_, device=records.device) for r, s_ in enumerate(sizes)])
    return rec_all.index_select(0, keep), cnt_all.index_select(0, keep)
