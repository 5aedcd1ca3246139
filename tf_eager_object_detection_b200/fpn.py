"""Mirror of `BaseFPN._assign_levels` / `BaseFPN._get_roi_features`
(reference: object_detection/model/fpn/base_fpn_model.py:303-324, :152-161)."""
import torch

from . import ops
from .roi_pooling import RoiPoolingCropAndResize2

__all__ = ['assign_levels', 'get_roi_features', 'fpn_roi_features']


def assign_levels(all_rois, min_level=2, max_level=5):
    """-> (rois_list [P2..P5] ragged, assign_level_idx [R] int64), as `_assign_levels` returns.  One sync (the ragged
    per-level counts are read back)."""
    rois = ops.to_device(all_rois, torch.float32)
    _, order, counts = ops.fpn_assign_levels(rois, min_level, max_level)
    order = order.to(torch.int64)
    gathered = rois.index_select(0, order) if rois.shape[0] else rois
    return list(torch.split(gathered, counts.tolist())), order


def get_roi_features(rois_list, p_list, image_shape, pool_size=7):
    """`_get_roi_features`: per non-empty level RoiPoolingCropAndResize2, concatenated on axis 0 (level-major)."""
    pool = RoiPoolingCropAndResize2(pool_size)
    outs = [pool((p, r, image_shape)) for r, p in zip(rois_list, p_list) if r.shape[0] > 0]
    return torch.cat(outs, dim=0)


def fpn_roi_features(all_rois, p_list, image_shape, pool_size=7, min_level=2, box_ind=None):
    """Fused `_assign_levels` + `_get_roi_features` in two launches with no host sync:
    -> (features [R,P,P,C] level-major, assign_level_idx [R] int32, level [R] int32, counts [L] int32)."""
    feats, lv, order, counts = ops.fpn_roi_features(p_list, all_rois, image_shape, pool_size, min_level, box_ind)
    return feats, order, lv, counts
