"""Mirror of the reference's `object_detection/model/proposal_target.py` — same constructor and call signature."""
import torch

from . import ops

__all__ = ['ProposalTarget']


class ProposalTarget:
    """model/proposal_target.py:8-124.  Sampling priority `perm` replaces the unseeded tf.random_shuffle /
    np.random.choice (see AnchorTarget).  Raises ValueError, like np.random.choice at :77, when background padding is
    needed but no background roi exists (`check=True`, one host sync; `check=False` skips the sync and the check)."""

    def __init__(self, num_classes=21, pos_iou_threshold=0.5, neg_iou_threshold=0.5, total_num_samples=128,
                 max_pos_samples=32, target_means=None, target_stds=None):
        self._num_classes = num_classes
        self._pos_iou_threshold = pos_iou_threshold
        self._neg_iou_threshold = neg_iou_threshold
        self._total_num_samples = total_num_samples
        self._max_pos_samples = max_pos_samples
        self._target_means = [0, 0, 0, 0] if target_means is None else target_means
        self._target_stds = [1, 1, 1, 1] if target_stds is None else target_stds

    def call_batched(self, inputs, perm=None, seed=None, roi_counts=None, gt_counts=None):
        """inputs = (rois [b,k,4], gt [b,m,4], gt_labels [b,m]) ->
        (rois [b,S,4], labels [b,S], targets, in_w, out_w [b,S,4C], keep [b,S], counts [b,2] = (#fg, status))."""
        rois, gt, gt_labels = inputs
        rois = ops.to_device(rois, torch.float32)
        b, k = rois.shape[0], rois.shape[1]
        if perm is None:
            g = None
            if seed is not None:
                g = torch.Generator(device=rois.device)
                g.manual_seed(int(seed))
            perm = torch.stack([torch.randperm(k, device=rois.device, generator=g) for _ in range(b)]).to(torch.int32)
        else:
            perm = ops.to_device(perm, torch.int32, rois.device).reshape(b, k)
        return ops.proposal_target(rois, gt, gt_labels, perm, self._num_classes, self._pos_iou_threshold,
                                   self._neg_iou_threshold, self._total_num_samples, self._max_pos_samples,
                                   self._target_means, self._target_stds, roi_counts, gt_counts)

    def call(self, inputs, training=None, mask=None, perm=None, seed=None, check=True):
        """inputs = (rois [k,4], gt_bboxes [m,4], gt_labels [m]) -> (final_rois [S,4], final_labels [S],
        final_bbox_targets [S,4C], bbox_inside_weights [S,4C], bbox_outside_weights [S,4C])."""
        rois, gt, gt_labels = inputs
        rois = ops.to_device(rois, torch.float32)
        gt = ops.to_device(gt, torch.float32, rois.device)
        gt_labels = ops.to_device(gt_labels, torch.int32, rois.device)
        perm_b = None if perm is None else ops.to_device(perm, torch.int32, rois.device).unsqueeze(0)
        r, lab, tg, iw, ow, _, cnt = self.call_batched((rois.unsqueeze(0), gt.unsqueeze(0), gt_labels.unsqueeze(0)),
                                                       perm_b, seed)
        if check and int(cnt[0, 1].item()) != 0:
            raise ValueError("'a' cannot be empty unless no samples are taken")  # np.random.choice, proposal_target.py:77
        return r[0], lab[0], tg[0], iw[0], ow[0]

    __call__ = call
