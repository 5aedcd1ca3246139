"""Mirror of the reference's `object_detection/model/region_proposal.py` — same constructor and call signature."""
from . import ops

__all__ = ['RegionProposal']


class RegionProposal:
    """model/region_proposal.py:11-81.  As in the reference, `num_pre_nms_*` are stored and NOT applied (the top-k block
    at :65-69 is commented out) unless `apply_pre_nms_top_k=True`; `min_size` <= 0 keeps the reference's min_edge=None."""

    def __init__(self, num_anchors=9, num_pre_nms_train=12000, num_post_nms_train=2000, num_pre_nms_test=6000,
                 num_post_nms_test=300, nms_iou_threshold=0.7, target_means=None, target_stds=None,
                 apply_pre_nms_top_k=False, min_size=0.0):
        self._num_anchors = num_anchors
        self._num_pre_nms_train = num_pre_nms_train
        self._num_post_nms_train = num_post_nms_train
        self._num_pre_nms_test = num_pre_nms_test
        self._num_post_nms_test = num_post_nms_test
        self._nms_iou_threshold = nms_iou_threshold
        self._target_means = [0, 0, 0, 0] if target_means is None else target_means
        self._target_stds = [1, 1, 1, 1] if target_stds is None else target_stds
        self._apply_pre_nms_top_k = apply_pre_nms_top_k
        self._min_size = min_size

    def _knobs(self, training):
        post = self._num_post_nms_train if training else self._num_post_nms_test
        pre = (self._num_pre_nms_train if training else self._num_pre_nms_test) if self._apply_pre_nms_top_k else 0
        return post, pre

    def call_batched(self, inputs, training=None):
        """inputs = (deltas [b,n,4], anchors [n,4], scores [b,n], image_shape) ->
        (rois [b,post,4] zero padded, idx [b,post] int32 -1 padded, count [b]) with no host synchronisation."""
        deltas, anchors, scores, image_shape = inputs
        post, pre = self._knobs(training)
        return ops.proposals(anchors, deltas, scores, image_shape, post, self._nms_iou_threshold,
                             self._target_means, self._target_stds, pre, self._min_size)

    def call(self, inputs, training=None, mask=None):
        """inputs = (bboxes_txtytwth [n,4], anchors [n,4], scores [n], image_shape [H,W]) -> rois [K<=post_nms, 4] in
        selection (descending score) order.  One explicit sync reads K for the ragged view."""
        deltas, anchors, scores, image_shape = inputs
        deltas = ops.to_device(deltas, ops.f32)
        scores = ops.to_device(scores, ops.f32, deltas.device)
        rois, _, count = self.call_batched((deltas.unsqueeze(0), anchors, scores.unsqueeze(0), image_shape), training)
        return rois[0, :int(count[0].item())]

    __call__ = call
