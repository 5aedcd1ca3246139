"""Mirror of the reference's `object_detection/model/anchor_target.py` — same constructor and call signature."""
import torch

from . import ops

__all__ = ['AnchorTarget']


class AnchorTarget:
    """model/anchor_target.py:7-107.  The reference subsamples with an unseeded `tf.random_shuffle`; here the shuffle
    order is a priority array: pass `perm` ([N] ints, lower = earlier) for reproducible targets, else a fresh random
    permutation is drawn per call (torch.randperm on the device, seeded by `seed` when given)."""

    def __init__(self, pos_iou_threshold=0.7, neg_iou_threshold=0.3, total_num_samples=256, max_pos_samples=128,
                 target_means=None, target_stds=None):
        self._pos_iou_threshold = pos_iou_threshold
        self._neg_iou_threshold = neg_iou_threshold
        self._total_num_samples = total_num_samples
        self._max_pos_samples = max_pos_samples
        self._target_means = [0, 0, 0, 0] if target_means is None else target_means
        self._target_stds = [1, 1, 1, 1] if target_stds is None else target_stds

    def _perm(self, b, n, device, perm, seed):
        if perm is not None:
            return ops.to_device(perm, torch.int32, device).reshape(b, n)
        g = None
        if seed is not None:
            g = torch.Generator(device=device)
            g.manual_seed(int(seed))
        return torch.stack([torch.randperm(n, device=device, generator=g) for _ in range(b)]).to(torch.int32)

    def call_batched(self, inputs, perm=None, seed=None, gt_counts=None):
        """inputs = (gt [b,m,4], image_shape, anchors [n,4]) -> labels [b,n] fp32, targets/in_w/out_w [b,n,4], counts [b,2]."""
        gt, image_shape, anchors = inputs
        anchors = ops.to_device(anchors, torch.float32)
        gt = ops.to_device(gt, torch.float32, anchors.device)
        p = self._perm(gt.shape[0], anchors.shape[0], anchors.device, perm, seed)
        return ops.anchor_target(anchors, gt, p, image_shape, self._pos_iou_threshold, self._neg_iou_threshold,
                                 self._total_num_samples, self._max_pos_samples, self._target_means, self._target_stds,
                                 gt_counts)

    def call(self, inputs, training=None, mask=None, perm=None, seed=None):
        """inputs = (gt_bboxes [m,4], image_shape [H,W], all_anchors [n,4]) -> (labels [n] fp32 in {-1,0,1},
        bbox_targets [n,4], bbox_inside_weights [n,4], bbox_outside_weights [n,4]).  No host sync."""
        gt, image_shape, anchors = inputs
        gt = ops.to_device(gt, torch.float32)
        lab, tg, iw, ow, _ = self.call_batched((gt.unsqueeze(0), image_shape, anchors), perm, seed)
        return lab[0], tg[0], iw[0], ow[0]

    __call__ = call
