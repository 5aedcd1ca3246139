"""Mirror of the reference's `object_detection/utils/bbox_tf.py` (same names, argument meaning, return types)."""
import torch

from . import ops

__all__ = ['pairwise_iou', 'bboxes_clip_filter', 'bboxes_range_filter']


def pairwise_iou(boxlist1, boxlist2):
    """utils/bbox_tf.py:37-56 -> [N,M] fp32 ("+1" areas; 0 where the intersection is 0)."""
    return ops.pairwise_iou(boxlist1, boxlist2)


def bboxes_clip_filter(rpn_proposals, min_value, max_height, max_width, min_edge=None):
    """utils/bbox_tf.py:59-84.  min_edge None: (clipped boxes, arange int32) with no host sync.  Otherwise the ragged
    (boxes [n',4], idx [n'] int64) pair of the reference — one explicit sync to read n'."""
    if min_edge is None:
        b = ops.to_device(rpn_proposals, torch.float32)
        lo = float(min_value)
        # clip only (:71-74); elementwise on the caller's tensor layout, done by the same decode/clip kernel family
        ob, _, _ = ops.clip_filter(b, lo, max_height, max_width, float('-inf'))
        return ob, torch.arange(b.shape[0], dtype=torch.int32, device=b.device)
    ob, oi, oc = ops.clip_filter(rpn_proposals, min_value, max_height, max_width, min_edge)
    n = int(oc.item())
    return ob[:n], oi[:n].to(torch.int64)


def bboxes_range_filter(anchors, max_height, max_width):
    """utils/bbox_tf.py:87-101 -> int64 indices of anchors inside the image (one sync for the ragged length)."""
    oi, oc = ops.range_filter(anchors, max_height, max_width)
    return oi[:int(oc.item())].to(torch.int64)
