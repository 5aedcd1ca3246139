"""Mirror of the reference's `object_detection/model/losses.py` (same names, argument meaning) — "next" row f3.

The loss and its gradient w.r.t. the prediction come out of one library call; when the prediction is a torch tensor that
requires grad the result is wired into autograd, so `loss.backward()` works as `tape.gradient` does in
`scripts/train.py:99-103`."""
import torch

from . import ops

__all__ = ['cls_loss', 'smooth_l1_loss']


def _wants_grad(x):
    return isinstance(x, torch.Tensor) and x.requires_grad and torch.is_grad_enabled()


def cls_loss(logits, labels, weight=1):
    """model/losses.py:4-13.  logits [n,c]; labels [n] in [0,c) — rows with a negative label (AnchorTarget's "ignore")
    are skipped, which folds in the gather at base_faster_rcnn_model.py:204-206."""
    if _wants_grad(logits):
        return ops._LossFn.apply(logits, ops.cls_loss, (labels, weight))
    return ops.cls_loss(logits, labels, weight)


def smooth_l1_loss(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma=1.0, dim=[1]):  # noqa: B006
    """model/losses.py:16-28; `dim` is [1] (RoI head) or [0, 1] (RPN), the two forms the reference calls."""
    if _wants_grad(bbox_pred):
        return ops._LossFn.apply(bbox_pred, _sl1, (bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma,
                                                   tuple(dim)))
    return ops.smooth_l1_loss(bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma, tuple(dim))


def _sl1(pred, target, in_w, out_w, sigma, dim, with_grad=False):
    return ops.smooth_l1_loss(pred, target, in_w, out_w, sigma, dim, with_grad=with_grad)
