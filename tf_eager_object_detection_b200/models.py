"""Mirrors of the two CALLERS of the box path: `BaseFasterRcnn` (object_detection/model/faster_rcnn/
base_faster_rcnn_model.py:16-306) and `BaseFPN` (object_detection/model/fpn/base_fpn_model.py:17-390) — same constructor
arguments, same methods (`call`, `im_detect`, `predict_roi(s)`, `_get_rpn_loss`, `_get_roi_loss`, `_get_anchors`,
`_assign_levels`, `_get_roi_features`), same order of operations, with every box-processing step on libboxpath.

What stays outside (dense convolutions, cuDNN territory) is supplied by the subclass exactly as in the
reference: `_get_extractor()`, `_get_roi_head()`, `_get_rpn_head()` (and `_get_neck()` for FPN) return callables on torch
CUDA tensors in the reference's layouts — images and feature maps NHWC `[1,h,w,C]`; the RPN head returns
(`rpn_score`, `rpn_bbox_txtytwth`) = (`[h*w, 2A]` caffe layout, `[h*w*A, 4]`) for Faster R-CNN
(base_faster_rcnn_model.py:342-349) and (`[h*w*A, 2]`, `[h*w*A, 4]`) per level for FPN (base_fpn_model.py:428-433); the RoI
head maps `[R,7,7,C]` to (`[R,num_classes]`, `[R,4*num_classes]`).

Differences from the reference, all on the host side of the path:
  * the RPN softmax + layout dance (base_faster_rcnn_model.py:149-152, base_fpn_model.py:223) runs inside the proposal kernel
    (`ops.proposals_rpn`), anchors are generated on the device once per image shape;
  * evaluation keeps padded tensors + device counts between the stages (proposals -> RoI features -> head -> post-processing)
    and synchronises once, for the ragged result the reference's signature returns;
  * `BaseFPN._get_roi_loss` returns its two losses (the reference's has no `return`, base_fpn_model.py:291-301, so its
    training `call` raises on unpacking `None`);
  * sampling is reproducible when `perm_anchor` / `perm_roi` (or `seed`) are given (DESIGN.md "Sampling")."""
import math

import torch

from . import _lib, ops
from .anchor_generator import generate_anchor_base, generate_by_anchor_base_tf, make_fpn_anchors
from .anchor_target import AnchorTarget
from .fpn import assign_levels, get_roi_features
from .losses import cls_loss, smooth_l1_loss
from .prediction import post_ops_prediction_batched
from .proposal_target import ProposalTarget
from .region_proposal import RegionProposal
from .roi_pooling import RoiPoolingCropAndResize, RoiPoolingCropAndResize2

__all__ = ['BaseFasterRcnn', 'BaseFPN']


def _ragged_detections(det, cnt):
    """[1,K,6] records + count -> (boxes [n,4], labels [n] int32, scores [n]) or (None, None, None): the triple
    post_ops_prediction returns (prediction.py:161-163).  The one host synchronisation of an evaluation forward."""
    n = int(cnt[0].item())
    if n == 0:
        return None, None, None
    d = det[0, :n]
    return d[:, :4], d[:, 5].to(torch.int32), d[:, 4]


class _BoxPathModel:
    """What the two skeletons share: constructor bookkeeping, the two loss helpers, the prediction tail."""

    def _init_common(self, num_classes, weight_decay, rpn_sigma, roi_sigma, roi_proposal_means, roi_proposal_stds,
                     prediction_max_objects_per_image, prediction_max_objects_per_class, prediction_nms_iou_threshold,
                     prediction_score_threshold):
        self.num_classes = num_classes
        self.weight_decay = weight_decay
        self._rpn_sigma = rpn_sigma
        self._roi_sigma = roi_sigma
        self._roi_proposal_means = roi_proposal_means
        self._roi_proposal_stds = roi_proposal_stds
        self._prediction_max_objects_per_image = prediction_max_objects_per_image
        self._prediction_max_objects_per_class = prediction_max_objects_per_class
        self._prediction_nms_iou_threshold = prediction_nms_iou_threshold
        self._prediction_score_threshold = prediction_score_threshold

    def _get_roi_head(self):
        raise NotImplementedError

    def _get_extractor(self):
        raise NotImplementedError

    def _get_rpn_head(self):
        raise NotImplementedError

    def _get_roi_loss(self, roi_score, roi_bbox_txtytwth, proposal_target_labels, proposal_target_bboxes_txtytwth,
                      proposal_target_in_weights, proposal_target_out_weights):
        """base_faster_rcnn_model.py:213-224."""
        roi_cls_loss = cls_loss(logits=roi_score, labels=proposal_target_labels)
        roi_reg_loss = smooth_l1_loss(roi_bbox_txtytwth, proposal_target_bboxes_txtytwth, proposal_target_in_weights,
                                      proposal_target_out_weights, sigma=self._roi_sigma)
        return roi_cls_loss, roi_reg_loss

    def _predict(self, roi_score, roi_bboxes_txtytwth, rois, roi_counts, image_shape, extractor_stride):
        """softmax + post_ops_prediction (base_faster_rcnn_model.py:186-198): padded records + device count."""
        roi_score_softmax = torch.softmax(roi_score, dim=-1)
        r = rois.shape[1]
        return post_ops_prediction_batched(roi_score_softmax.reshape(1, r, self.num_classes),
                                           roi_bboxes_txtytwth.reshape(1, r, self.num_classes, 4), rois, image_shape,
                                           self._roi_proposal_means, self._roi_proposal_stds,
                                           max_num_per_class=self._prediction_max_objects_per_class,
                                           max_num_per_image=self._prediction_max_objects_per_image,
                                           nms_iou_threshold=self._prediction_nms_iou_threshold,
                                           score_threshold=self._prediction_score_threshold,
                                           extractor_stride=extractor_stride, roi_counts=roi_counts)

    def __call__(self, inputs, training=None, **kw):
        return self.call(inputs, training=training, **kw)


class BaseFasterRcnn(_BoxPathModel):
    """faster_rcnn/base_faster_rcnn_model.py:16-306 (C4 models: VGG16 / ResNet)."""

    def __init__(self, num_classes, weight_decay, ratios, scales, extractor_stride, rpn_proposal_means, rpn_proposal_stds,
                 rpn_proposal_num_pre_nms_train, rpn_proposal_num_post_nms_train, rpn_proposal_num_pre_nms_test,
                 rpn_proposal_num_post_nms_test, rpn_proposal_nms_iou_threshold, rpn_sigma, rpn_training_pos_iou_threshold,
                 rpn_training_neg_iou_threshold, rpn_training_total_num_samples, rpn_training_max_pos_samples,
                 roi_proposal_means, roi_proposal_stds, roi_pool_size, roi_pooling_max_pooling_flag, roi_sigma,
                 roi_training_pos_iou_threshold, roi_training_neg_iou_threshold, roi_training_total_num_samples,
                 roi_training_max_pos_samples, prediction_max_objects_per_image, prediction_max_objects_per_class,
                 prediction_nms_iou_threshold, prediction_score_threshold):
        self._init_common(num_classes, weight_decay, rpn_sigma, roi_sigma, roi_proposal_means, roi_proposal_stds,
                          prediction_max_objects_per_image, prediction_max_objects_per_class,
                          prediction_nms_iou_threshold, prediction_score_threshold)
        self._ratios = ratios
        self._scales = scales
        self._num_anchors = len(ratios) * len(scales)
        self._extractor_stride = extractor_stride
        self._anchor_generator = generate_by_anchor_base_tf
        self._anchor_base = generate_anchor_base(extractor_stride, ratios, scales)                       # :83
        self._rpn_proposal = RegionProposal(num_anchors=self._num_anchors, num_pre_nms_train=rpn_proposal_num_pre_nms_train,
                                            num_post_nms_train=rpn_proposal_num_post_nms_train,
                                            num_pre_nms_test=rpn_proposal_num_pre_nms_test,
                                            num_post_nms_test=rpn_proposal_num_post_nms_test,
                                            nms_iou_threshold=rpn_proposal_nms_iou_threshold,
                                            target_means=rpn_proposal_means, target_stds=rpn_proposal_stds)    # :88-97
        self._anchor_target = AnchorTarget(pos_iou_threshold=rpn_training_pos_iou_threshold,
                                           neg_iou_threshold=rpn_training_neg_iou_threshold,
                                           total_num_samples=rpn_training_total_num_samples,
                                           max_pos_samples=rpn_training_max_pos_samples, target_means=rpn_proposal_means,
                                           target_stds=rpn_proposal_stds)                                      # :98-105
        self._roi_pooling = RoiPoolingCropAndResize(pool_size=roi_pool_size, max_pooling_flag=roi_pooling_max_pooling_flag)
        self._proposal_target = ProposalTarget(num_classes=num_classes, pos_iou_threshold=roi_training_pos_iou_threshold,
                                               neg_iou_threshold=roi_training_neg_iou_threshold,
                                               total_num_samples=roi_training_total_num_samples,
                                               max_pos_samples=roi_training_max_pos_samples,
                                               target_means=roi_proposal_means, target_stds=roi_proposal_stds)  # :108-115
        self._extractor = self._get_extractor()
        self._roi_head = self._get_roi_head()
        self._rpn_head = self._get_rpn_head()

    # ---- the shared front of call / im_detect / predict_roi (:132-153, :280-300, :252-263)
    def _front(self, image, training):
        image_shape = [int(image.shape[1]), int(image.shape[2])]
        shared_features = self._extractor(image, training=training)
        anchors = self._anchor_generator(self._anchor_base, self._extractor_stride,
                                         math.ceil(image_shape[0] / self._extractor_stride),
                                         math.ceil(image_shape[1] / self._extractor_stride), device=image.device)
        rpn_score, rpn_bbox_txtytwth = self._rpn_head(shared_features, training=training)
        post, pre = self._rpn_proposal._knobs(training)
        # :149-153 — the four-statement softmax / layout dance and RegionProposal.call, one launch on the raw logits
        rois, _, counts = ops.proposals_rpn(anchors, rpn_bbox_txtytwth.detach().reshape(1, -1, 4),
                                            rpn_score.detach().reshape(1, -1, 2 * self._num_anchors), _lib.RPN_CAFFE,
                                            self._num_anchors, image_shape, post, self._rpn_proposal._nms_iou_threshold,
                                            self._rpn_proposal._target_means, self._rpn_proposal._target_stds, pre,
                                            self._rpn_proposal._min_size)
        return image_shape, shared_features, anchors, rpn_score, rpn_bbox_txtytwth, rois, counts

    def call(self, inputs, training=None, mask=None, perm_anchor=None, perm_roi=None, seed=None):
        if training:
            image, gt_bboxes, gt_labels = inputs
        else:
            image = inputs
        image_shape, shared_features, anchors, rpn_score, rpn_bbox_txtytwth, rois, counts = self._front(image, training)
        if training:
            rpn_labels, rpn_bbox_targets, rpn_in_weights, rpn_out_weights = self._anchor_target(
                (gt_bboxes, image_shape, anchors), training, perm=perm_anchor, seed=seed)                 # :157-161
            rpn_cls_loss, rpn_reg_loss = self._get_rpn_loss(rpn_score, rpn_bbox_txtytwth, rpn_labels, rpn_bbox_targets,
                                                            rpn_in_weights, rpn_out_weights)
            k = int(counts[0].item())                                 # the ragged proposal list ProposalTarget samples from
            final_rois, roi_labels, roi_bbox_target, roi_in_weights, roi_out_weights = self._proposal_target(
                (rois[0, :k], gt_bboxes, gt_labels), training, perm=None if perm_roi is None else perm_roi[:k], seed=seed)
            roi_features = self._roi_pooling((shared_features, final_rois, self._extractor_stride), training=training)
            roi_score, roi_bboxes_txtytwth = self._roi_head(roi_features, training=training)
            roi_cls_loss, roi_reg_loss = self._get_roi_loss(roi_score, roi_bboxes_txtytwth, roi_labels, roi_bbox_target,
                                                            roi_in_weights, roi_out_weights)
            return rpn_cls_loss, rpn_reg_loss, roi_cls_loss, roi_reg_loss
        roi_features = ops.roi_pool(_lib.ROI_STRIDE_NORM,
                                    _lib.POOL_MAX2 if self._roi_pooling._max_pooling_flag else _lib.POOL_NONE,
                                    self._roi_pooling._pool_size, shared_features.detach(), rois,
                                    stride=float(self._extractor_stride), roi_counts=counts)              # :182
        roi_score, roi_bboxes_txtytwth = self._roi_head(roi_features, training=training)
        det, cnt = self._predict(roi_score, roi_bboxes_txtytwth, rois, counts, image_shape, self._extractor_stride)
        return _ragged_detections(det, cnt)

    def _get_rpn_loss(self, rpn_score, rpn_bbox_txtytwth, anchor_target_labels, anchor_target_bboxes_txtytwth,
                      anchor_target_in_weights, anchor_target_out_weights):
        """:200-211 (the `labels >= 0` gather is folded into cls_loss)."""
        a = self._num_anchors
        rpn_score = rpn_score.reshape(-1, 2, a).permute(0, 2, 1).reshape(-1, 2)
        rpn_cls_loss = cls_loss(logits=rpn_score, labels=anchor_target_labels)
        rpn_reg_loss = smooth_l1_loss(rpn_bbox_txtytwth, anchor_target_bboxes_txtytwth, anchor_target_in_weights,
                                      anchor_target_out_weights, self._rpn_sigma, dim=[0, 1])
        return rpn_cls_loss, rpn_reg_loss

    def predict_roi(self, image, gt_bboxes, gt_labels, perm=None, seed=None):
        """:243-265: the ProposalTarget outputs for one training image."""
        _, _, _, _, _, rois, counts = self._front(image, True)
        k = int(counts[0].item())
        return self._proposal_target((rois[0, :k], gt_bboxes, gt_labels), training=True, perm=perm, seed=seed)

    def im_detect_batched(self, preprocessed_image):
        """im_detect (:279-306) without its last statement and without a host sync: (roi_score_softmax [1,R,C],
        roi_bboxes_txtytwth [1,R,4C], rois [1,R,4] in network-input pixels, counts [1]) — what
        `prediction.eval_loop_detections` consumes together with img_scale and the raw image size."""
        image_shape, shared_features, _, _, _, rois, counts = self._front(preprocessed_image, False)
        roi_features = ops.roi_pool(_lib.ROI_STRIDE_NORM,
                                    _lib.POOL_MAX2 if self._roi_pooling._max_pooling_flag else _lib.POOL_NONE,
                                    self._roi_pooling._pool_size, shared_features.detach(), rois,
                                    stride=float(self._extractor_stride), roi_counts=counts)
        roi_score, roi_bboxes_txtytwth = self._roi_head(roi_features, training=False)
        r = rois.shape[1]
        return torch.softmax(roi_score, dim=-1).reshape(1, r, -1), roi_bboxes_txtytwth.reshape(1, r, -1), rois, counts

    def im_detect(self, preprocessed_image, img_scale):
        """:279-306 -> (roi_score_softmax [K,C], roi_bboxes_txtytwth [K,4C], rois / img_scale [K,4])."""
        sm, tx, rois, counts = self.im_detect_batched(preprocessed_image)
        k = int(counts[0].item())
        return sm[0, :k], tx[0, :k], rois[0, :k] / float(img_scale)


class BaseFPN(_BoxPathModel):
    """fpn/base_fpn_model.py:17-390 (ResNet FPN)."""

    def __init__(self, num_classes, weight_decay, ratios, scales, rpn_proposal_means, rpn_proposal_stds,
                 rpn_proposal_num_pre_nms_train, rpn_proposal_num_post_nms_train, rpn_proposal_num_pre_nms_test,
                 rpn_proposal_num_post_nms_test, rpn_proposal_nms_iou_threshold, rpn_sigma, rpn_training_pos_iou_threshold,
                 rpn_training_neg_iou_threshold, rpn_training_total_num_samples, rpn_training_max_pos_samples,
                 roi_proposal_means, roi_proposal_stds, roi_pool_size, roi_sigma, roi_training_pos_iou_threshold,
                 roi_training_neg_iou_threshold, roi_training_total_num_samples, roi_training_max_pos_samples,
                 prediction_max_objects_per_image, prediction_max_objects_per_class, prediction_nms_iou_threshold,
                 prediction_score_threshold, level_name_list=('p2', 'p3', 'p4', 'p5', 'p6'), min_level=2, max_level=5,
                 top_down_dims=256, anchor_stride_list=(4, 8, 16, 32, 64), base_anchor_size_list=(32, 64, 128, 256, 512),
                 roi_pooling_max_pooling_flag=True):
        self._init_common(num_classes, weight_decay, rpn_sigma, roi_sigma, roi_proposal_means, roi_proposal_stds,
                          prediction_max_objects_per_image, prediction_max_objects_per_class,
                          prediction_nms_iou_threshold, prediction_score_threshold)
        self._ratios = ratios
        self._scales = scales
        self._num_anchors = len(ratios) * len(scales)
        self._level_name_list = list(level_name_list)
        self._min_level, self._max_level = min_level, max_level
        self._top_down_dims = top_down_dims
        self._anchor_stride_list = list(anchor_stride_list)
        self._base_anchor_size_list = list(base_anchor_size_list)
        self._rpn_proposal = RegionProposal(num_anchors=self._num_anchors, num_pre_nms_train=rpn_proposal_num_pre_nms_train,
                                            num_post_nms_train=rpn_proposal_num_post_nms_train,
                                            num_pre_nms_test=rpn_proposal_num_pre_nms_test,
                                            num_post_nms_test=rpn_proposal_num_post_nms_test,
                                            nms_iou_threshold=rpn_proposal_nms_iou_threshold,
                                            target_means=rpn_proposal_means, target_stds=rpn_proposal_stds)    # :112-121
        self._anchor_target = AnchorTarget(pos_iou_threshold=rpn_training_pos_iou_threshold,
                                           neg_iou_threshold=rpn_training_neg_iou_threshold,
                                           total_num_samples=rpn_training_total_num_samples,
                                           max_pos_samples=rpn_training_max_pos_samples, target_means=rpn_proposal_means,
                                           target_stds=rpn_proposal_stds)
        self._roi_pooling = RoiPoolingCropAndResize2(pool_size=roi_pool_size)      # the max-pooling flag is ignored, :122
        self._roi_pool_size = roi_pool_size
        self._proposal_target = ProposalTarget(num_classes=num_classes, pos_iou_threshold=roi_training_pos_iou_threshold,
                                               neg_iou_threshold=roi_training_neg_iou_threshold,
                                               total_num_samples=roi_training_total_num_samples,
                                               max_pos_samples=roi_training_max_pos_samples,
                                               target_means=roi_proposal_means, target_stds=roi_proposal_stds)
        self._extractor = self._get_extractor()
        self._neck = self._get_neck()
        self._roi_head = self._get_roi_head()
        self._rpn_head = self._get_rpn_head()

    def _get_neck(self):
        raise NotImplementedError

    def _get_anchors(self, image_shape, device=None):
        """:163-186: P2..P6 concatenated (one launch, cached per image shape)."""
        return make_fpn_anchors(image_shape, self._base_anchor_size_list, self._anchor_stride_list, self._scales, self._ratios,
                                device=device)

    def _get_fpn_head_results(self, p_list):
        """:188-200."""
        scores, preds = zip(*[self._rpn_head(p) for p in p_list])
        return torch.cat(scores, dim=0), torch.cat(preds, dim=0)

    def _assign_levels(self, all_rois):
        """:303-324."""
        return assign_levels(all_rois, self._min_level, self._max_level)

    def _get_roi_features(self, rois_list, p_list, image_shape):
        """:152-161 (zip stops at the shorter list: P6 carries anchors only)."""
        return get_roi_features(rois_list, p_list, image_shape, self._roi_pool_size)

    def _front(self, image, training):
        image_shape = [int(image.shape[1]), int(image.shape[2])]
        c_list = self._extractor(image, training=training)
        p_list = self._neck(c_list, training=training)
        all_fpn_scores, all_fpn_bbox_pred = self._get_fpn_head_results(p_list)
        all_anchors = self._get_anchors(image_shape, device=image.device)
        post, pre = self._rpn_proposal._knobs(training)
        # :223-225 — softmax(...)[:, 1] inside the proposal stage
        rois, _, counts = ops.proposals_rpn(all_anchors, all_fpn_bbox_pred.detach().reshape(1, -1, 4),
                                            all_fpn_scores.detach().reshape(1, -1, 2), _lib.RPN_PAIRS, 1, image_shape, post,
                                            self._rpn_proposal._nms_iou_threshold, self._rpn_proposal._target_means,
                                            self._rpn_proposal._target_stds, pre, self._rpn_proposal._min_size)
        return image_shape, p_list, all_anchors, all_fpn_scores, all_fpn_bbox_pred, rois, counts

    def _eval_tail(self, p_list, rois, counts, image_shape):
        """Level assignment + RoI features + head on the padded proposal list; rows behind `counts` are dropped (their
        level / features are whatever the zero box gets, never read).  -> (roi_score, txtytwth, rois level-major [1,K,4], K)"""
        k = int(counts[0].item())
        feats, _, order, _ = ops.fpn_roi_features([p.detach() for p in p_list[:self._max_level - self._min_level + 1]],
                                                  rois[0, :k], image_shape, self._roi_pool_size, self._min_level)
        roi_score, roi_bboxes_txtytwth = self._roi_head(feats, training=False)
        final_rois = rois[0, :k].index_select(0, order.long())                     # tf.concat(rois_list, axis=0), :262
        return roi_score, roi_bboxes_txtytwth, final_rois.unsqueeze(0), k

    def call(self, inputs, training=None, mask=None, perm_anchor=None, perm_roi=None, seed=None):
        if training:
            image, gt_bboxes, gt_labels = inputs
        else:
            image = inputs
        image_shape, p_list, all_anchors, all_fpn_scores, all_fpn_bbox_pred, rois, counts = self._front(image, training)
        if training:
            rpn_labels, rpn_bbox_targets, rpn_in_weights, rpn_out_weights = self._anchor_target(
                (gt_bboxes, image_shape, all_anchors), training, perm=perm_anchor, seed=seed)             # :229-232
            rpn_cls_loss, rpn_reg_loss = self._get_rpn_loss(all_fpn_scores, all_fpn_bbox_pred, rpn_labels, rpn_bbox_targets,
                                                            rpn_in_weights, rpn_out_weights)
            k = int(counts[0].item())
            final_rois, roi_labels, roi_bbox_target, roi_in_weights, roi_out_weights = self._proposal_target(
                (rois[0, :k], gt_bboxes, gt_labels), training, perm=None if perm_roi is None else perm_roi[:k], seed=seed)
            rois_list, selected_idx = self._assign_levels(final_rois)                                    # :244
            roi_features = self._get_roi_features(rois_list, p_list, image_shape)
            roi_score, roi_bboxes_txtytwth = self._roi_head(roi_features, training=training)
            roi_labels = roi_labels.index_select(0, selected_idx)                                        # :249-252
            roi_bbox_target = roi_bbox_target.index_select(0, selected_idx)
            roi_in_weights = roi_in_weights.index_select(0, selected_idx)
            roi_out_weights = roi_out_weights.index_select(0, selected_idx)
            roi_cls_loss, roi_reg_loss = self._get_roi_loss(roi_score, roi_bboxes_txtytwth, roi_labels, roi_bbox_target,
                                                            roi_in_weights, roi_out_weights)
            return rpn_cls_loss, rpn_reg_loss, roi_cls_loss, roi_reg_loss
        roi_score, roi_bboxes_txtytwth, final_rois, k = self._eval_tail(p_list, rois, counts, image_shape)
        if k == 0:
            return None, None, None
        det, cnt = self._predict(roi_score, roi_bboxes_txtytwth, final_rois, None, image_shape, 16)      # extractor_stride=16, :272
        return _ragged_detections(det, cnt)

    def _get_rpn_loss(self, rpn_score, rpn_bbox_txtytwth, anchor_target_labels, anchor_target_bboxes_txtytwth,
                      anchor_target_in_weights, anchor_target_out_weights):
        """:278-289."""
        rpn_cls_loss = cls_loss(logits=rpn_score, labels=anchor_target_labels)
        rpn_reg_loss = smooth_l1_loss(rpn_bbox_txtytwth, anchor_target_bboxes_txtytwth, anchor_target_in_weights,
                                      anchor_target_out_weights, self._rpn_sigma, dim=[0, 1])
        return rpn_cls_loss, rpn_reg_loss

    def predict_rpns(self, image_shape, gt_bboxes, perm=None, seed=None, device=None):
        """:326-339: the anchors AnchorTarget labels as foreground."""
        all_anchors = self._get_anchors(image_shape, device=device)
        rpn_labels, _, _, _ = self._anchor_target((gt_bboxes, image_shape, all_anchors), True, perm=perm, seed=seed)
        return all_anchors[rpn_labels > 0]

    def predict_rois(self, preprocessed_img, gt_bboxes, gt_labels, training=True, perm=None, seed=None):
        """:341-362: the rois ProposalTarget samples for one training image."""
        _, _, _, _, _, rois, counts = self._front(preprocessed_img, training)
        k = int(counts[0].item())
        return self._proposal_target((rois[0, :k], gt_bboxes, gt_labels), True, perm=perm, seed=seed)[0]

    def im_detect(self, preprocessed_img, img_scale):
        """:364-390 -> (roi_score_softmax [K,C], roi_bboxes_txtytwth [K,4C], level-major rois / img_scale [K,4])."""
        image_shape, p_list, _, _, _, rois, counts = self._front(preprocessed_img, False)
        roi_score, roi_bboxes_txtytwth, final_rois, _ = self._eval_tail(p_list, rois, counts, image_shape)
        return torch.softmax(roi_score, dim=-1), roi_bboxes_txtytwth, final_rois[0] / float(img_scale)
