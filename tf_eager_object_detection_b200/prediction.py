"""Mirror of the reference's `object_detection/model/prediction.py` (post-head detection filtering)."""
import torch

from . import ops

__all__ = ['post_ops_prediction', 'post_ops_prediction_batched']


def post_ops_prediction_batched(roi_scores_softmax, roi_txtytwth, rois, image_shape, target_means, target_stds,
                                max_num_per_class=50, max_num_per_image=150, nms_iou_threshold=0.3,
                                score_threshold=0.05, extractor_stride=16, roi_counts=None):
    """Batched form: scores [b,r,C], deltas [b,r,C,4], rois [b,r,4] -> (records [b,max_num_per_image,6] =
    (x1,y1,x2,y2,score,class), count [b]); padded, no host sync — the layout distributed.allgather_detections ships."""
    means = [0, 0, 0, 0] if target_means is None else target_means
    stds = [1, 1, 1, 1] if target_stds is None else target_stds
    return ops.post_ops_prediction(roi_scores_softmax, roi_txtytwth, rois, image_shape, means, stds, max_num_per_class,
                                   max_num_per_image, nms_iou_threshold, score_threshold, extractor_stride, roi_counts)


def post_ops_prediction(roi_scores_softmax, roi_txtytwth, rois, image_shape, target_means, target_stds,
                        max_num_per_class=50, max_num_per_image=150, nms_iou_threshold=0.3, score_threshold=0.05,
                        extractor_stride=16, num_classes=21):
    """model/prediction.py:103-163: -> (bboxes [n,4], classes [n] int32, scores [n]) with n <= max_num_per_image, or
    (None, None, None) when nothing survives.  Order: descending score (the reference's `top_k(sorted=False)` leaves
    it unspecified).  One host sync reads n."""
    scores = ops.to_device(roi_scores_softmax, torch.float32)
    if num_classes is not None and int(num_classes) != scores.shape[1]:
        raise ValueError('post_ops_prediction: num_classes=%d but the score tensor has %d columns' % (num_classes, scores.shape[1]))
    deltas = ops.to_device(roi_txtytwth, torch.float32, scores.device)
    rois = ops.to_device(rois, torch.float32, scores.device)
    det, cnt = post_ops_prediction_batched(scores.unsqueeze(0), deltas.reshape(1, scores.shape[0], -1, 4), rois.unsqueeze(0),
                                           image_shape, target_means, target_stds, max_num_per_class, max_num_per_image,
                                           nms_iou_threshold, score_threshold, extractor_stride)
    n = int(cnt[0].item())
    if n == 0:
        return None, None, None
    d = det[0, :n]
    return d[:, :4], d[:, 5].to(torch.int32), d[:, 4]
