"""Mirror of the reference's `object_detection/model/prediction.py` (post-head detection filtering)."""
import torch

from . import ops

__all__ = ['post_ops_prediction', 'post_ops_prediction_batched', 'eval_loop_detections']


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


def eval_loop_detections(roi_scores_softmax, roi_txtytwth, rois, img_scale, raw_hw, target_means=None, target_stds=None,
                         score_threshold=0.05, iou_threshold=0.3, max_objects_per_class=50, max_objects_per_image=50,
                         min_size=10, loop='voc', out_rows=None, roi_counts=None):
    """The per-image body of the reference's two evaluation loops, batched over images with no host synchronisation:
    `evaluation/pascal_eval_files_utils.py:76-106` (loop='voc': per-image cut `score >= k-th largest`, ties kept) and
    `scripts/eval_coco.py:116-153` (loop='coco': `tf.nn.top_k`).  Inputs are what `im_detect` returns BEFORE its final
    `rois / img_scale` (`faster_rcnn/base_faster_rcnn_model.py:304`), which runs inside the kernel: scores [b,r,C] softmax,
    roi_txtytwth [b,r,4C] or [b,r,C,4], rois [b,r,4] in network-input pixels, img_scale [b], raw_hw [b,2] = (raw_h, raw_w).
    -> (records [b,rows,6] = (x1,y1,x2,y2,score,class) in raw-image pixels, count [b]) — the layout
    `distributed.allgather_detections` ships and `evaluation.write_voc_results` / `coco_results` consume.
    The loops default to roi-head stds (0.1, 0.1, 0.2, 0.2) (pascal_eval_files_utils.py:68-71)."""
    from . import _lib
    means = [0, 0, 0, 0] if target_means is None else target_means
    stds = [0.1, 0.1, 0.2, 0.2] if target_stds is None else target_stds
    if loop not in ('voc', 'coco'):
        raise ValueError("loop must be 'voc' or 'coco'")
    cut = _lib.CUT_SCORE_GE if loop == 'voc' else _lib.CUT_TOP_K
    if max_objects_per_image <= 0:                       # pascal_eval_files_utils.py:98: no per-image cut at all
        b, r, c = roi_scores_softmax.shape
        max_objects_per_image = (c - 1) * max_objects_per_class
    return ops.eval_detections(roi_scores_softmax, roi_txtytwth, rois, raw_hw, img_scale, (0, 0), means, stds,
                               max_objects_per_class, max_objects_per_image, iou_threshold, score_threshold, min_size, cut,
                               out_rows, roi_counts)
