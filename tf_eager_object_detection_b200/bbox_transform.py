"""Mirror of the reference's `object_detection/utils/bbox_transform.py` (same names, argument meaning)."""
from . import ops

__all__ = ['encode_bbox_with_mean_and_std', 'decode_bbox_with_mean_and_std']


def encode_bbox_with_mean_and_std(src_bbox, dst_bbox, target_means, target_stds):
    """utils/bbox_transform.py:4-29 -> [n,4] deltas."""
    return ops.encode(src_bbox, dst_bbox, target_means, target_stds)


def decode_bbox_with_mean_and_std(anchors, bboxes_txtytwth, target_means, target_stds):
    """utils/bbox_transform.py:32-55 -> [n,4] boxes (x2 = x1 + w: no -1, as the reference)."""
    return ops.decode_clip(anchors, bboxes_txtytwth, target_means, target_stds, image_shape=None)
