"""Image-level data parallelism for the box path: every function on the path is per image, so a batch is
split into contiguous blocks of images, one block per GPU, with NO collective on the data path.  The only exchange is
the evaluation-time all-gather of fixed-size, padded per-image detection records, which replaces the reference's
single-process accumulation loops (evaluation/pascal_eval_files_utils.py:73-107, scripts/eval_coco.py:116-164).

torch.distributed is plumbing only (NCCL over NVLink on the GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist

__all__ = ['shard_bounds', 'shard_images', 'pack_detections', 'allgather_detections']


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
    """Raw ncclComm_t of the group's NCCL backend on the current CUDA device, or None (gloo / no NCCL / not initialised)."""
    try:
        pg = group if group is not None else dist.distributed_c10d._get_default_group()
        backend = pg._get_backend(torch.device('cuda', torch.cuda.current_device()))
        ptr = backend._comm_ptr()
        return int(ptr) if ptr else None
    except Exception:
        return None


def _allgather_native(records, counts, rec_all, cnt_all, world, group):
    """bx_allgather_detections on the framework's own communicator: both tensors in one fused NCCL group on the current
    stream.  Returns False when the group is not NCCL-backed (the gloo CPU tests), so the caller uses torch.distributed."""
    if not (records.is_cuda and records.dtype == torch.float32 and counts.dtype == torch.int32 and records.dim() == 3):
        return False
    comm = nccl_comm_ptr(group)
    if comm is None:
        return False
    import ctypes
    from . import _lib
    from ._tensor import stream_ptr
    dev = records.device.index
    st = stream_ptr(dev)
    h = _lib.handle(dev, st.value)
    b, k, f = records.shape
    _lib.check(_lib.load().bx_allgather_detections(h, ctypes.c_void_p(comm), records.data_ptr(), counts.data_ptr(), b, k, f,
                                                   world, rec_all.data_ptr(), cnt_all.data_ptr(), st))
    return True


def allgather_detections(records, counts, group=None, max_images_per_rank=None):
    """records [b_local, kmax, f] fp32 (zero padded), counts [b_local] int32 ->
    (records_all [b_total, kmax, f], counts_all [b_total]) in rank order, identical on every rank.

    Ranks may own different numbers of images (uneven shards): blocks are padded to `max_images_per_rank`
    (default: all-reduced max) for a single fixed-size all-gather, then the padding is dropped."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records, counts
    world = dist.get_world_size(group)
    b_local = records.shape[0]
    nloc = torch.tensor([b_local], dtype=torch.int32, device=records.device)
    sizes = torch.empty((world,), dtype=torch.int32, device=records.device)
    dist.all_gather_into_tensor(sizes, nloc, group=group)
    if max_images_per_rank is None:
        max_images_per_rank = int(sizes.max().item())
    pad = max_images_per_rank - b_local
    if pad:
        records = torch.cat([records, records.new_zeros((pad,) + tuple(records.shape[1:]))])
        counts = torch.cat([counts, counts.new_zeros((pad,))])
    rec_all = torch.empty((world * max_images_per_rank,) + tuple(records.shape[1:]), dtype=records.dtype,
                          device=records.device)
    cnt_all = torch.empty((world * max_images_per_rank,), dtype=counts.dtype, device=counts.device)
    if not _allgather_native(records.contiguous(), counts.contiguous(), rec_all, cnt_all, world, group):
        dist.all_gather_into_tensor(rec_all, records.contiguous(), group=group)
        dist.all_gather_into_tensor(cnt_all, counts.contiguous(), group=group)
    sizes = sizes.tolist()
    if all(s == max_images_per_rank for s in sizes):
        return rec_all, cnt_all
    keep = torch.cat([torch.arange(r * max_images_per_rank, r * max_images_per_rank + s, device=records.device)
                      for r, s in enumerate(sizes)])
    return rec_all.index_select(0, keep), cnt_all.index_select(0, keep)
