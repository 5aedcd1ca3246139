"""Functional layer over the C ABI: one Python function per `bx_*` entry point, torch CUDA tensors in and out.

Every function is asynchronous on torch's current stream and performs no host synchronisation; the class layer
(region_proposal.py, roi_pooling.py, ...) adds the reference's ragged Python views where its signatures need them."""
import ctypes
from ctypes import c_int, c_void_p

import torch

from . import _lib
from ._tensor import FLOAT32, INT32, Borrow, device_index_of, empty, stream_ptr, to_device

f32, i32 = torch.float32, torch.int32


def _ctx(t):
    dev = device_index_of(t)
    st = stream_ptr(dev)
    return dev, _lib.handle(dev, st.value), Borrow(dev), st, _lib.load()


def decode_clip(anchors, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), image_shape=None):
    """a1 (+a2 clip when image_shape is given).  anchors [n,4] or [b,n,4]; deltas [n,4] or [b,n,4]."""
    deltas = to_device(deltas, f32)
    anchors = to_device(anchors, f32, deltas.device)
    squeeze = deltas.dim() == 2
    d3 = deltas.unsqueeze(0) if squeeze else deltas
    b, n = d3.shape[0], d3.shape[1]
    batched_anchors = anchors.dim() == 3
    dev, h, bw, st, lib = _ctx(d3)
    out = empty((b, n, 4), f32, dev)
    if b * n:
        H, W = (int(image_shape[0]), int(image_shape[1])) if image_shape is not None else (0, 0)
        _lib.check(lib.bx_decode_clip(h, bw.ptr(anchors, FLOAT32, (b, n, 4) if batched_anchors else (n, 4), 16),
                                      1 if batched_anchors else 0, bw.ptr(d3, FLOAT32, (b, n, 4), 16), b, n,
                                      _lib.f4(means), _lib.f4(stds), H, W, bw.ptr(out, FLOAT32, (b, n, 4), 16), st))
    return out[0] if squeeze else out


def encode(src, dst, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """a12: utils/bbox_transform.py:4-29."""
    src = to_device(src, f32)
    dst = to_device(dst, f32, src.device)
    n = src.shape[0]
    dev, h, bw, st, lib = _ctx(src)
    out = empty((n, 4), f32, dev)
    if n:
        _lib.check(lib.bx_encode(h, bw.ptr(src, FLOAT32, (n, 4), 16), bw.ptr(dst, FLOAT32, (n, 4), 16), n,
                                 _lib.f4(means), _lib.f4(stds), bw.ptr(out, FLOAT32, (n, 4), 16), st))
    return out


def clip_filter(boxes, min_value, max_height, max_width, min_edge):
    """a2 with min_edge: returns (boxes_padded [n,4], idx_padded [n] int32, count [1] int32) — device-resident."""
    boxes = to_device(boxes, f32)
    n = boxes.shape[0]
    dev, h, bw, st, lib = _ctx(boxes)
    ob, oi, oc = empty((n, 4), f32, dev), empty((n,), i32, dev), torch.zeros((1,), dtype=i32, device=boxes.device)
    if n:
        _lib.check(lib.bx_clip_filter(h, bw.ptr(boxes, FLOAT32, (n, 4), 16), n, float(min_value), int(max_height),
                                      int(max_width), float(min_edge), bw.ptr(ob, FLOAT32, (n, 4), 16),
                                      bw.ptr(oi, INT32, (n,)), bw.ptr(oc, INT32, (1,)), st))
    return ob, oi, oc


def range_filter(anchors, max_height, max_width):
    """utils/bbox_tf.py:87-101: (idx_padded [n] int32, count [1])."""
    anchors = to_device(anchors, f32)
    n = anchors.shape[0]
    dev, h, bw, st, lib = _ctx(anchors)
    oi, oc = empty((n,), i32, dev), torch.zeros((1,), dtype=i32, device=anchors.device)
    if n:
        _lib.check(lib.bx_range_filter(h, bw.ptr(anchors, FLOAT32, (n, 4), 16), n, int(max_height), int(max_width),
                                       bw.ptr(oi, INT32, (n,)), bw.ptr(oc, INT32, (1,)), st))
    return oi, oc


def nms(boxes, scores, max_output_size, iou_threshold):
    """tf.image.non_max_suppression, batched: boxes [b,n,4], scores [b,n] -> (idx [b,max_out] int32 (-1 pad), count [b])."""
    boxes = to_device(boxes, f32)
    scores = to_device(scores, f32, boxes.device)
    squeeze = boxes.dim() == 2
    b3 = boxes.unsqueeze(0) if squeeze else boxes
    s2 = scores.unsqueeze(0) if squeeze else scores
    b, n = b3.shape[0], b3.shape[1]
    dev, h, bw, st, lib = _ctx(b3)
    oi, oc = empty((b, int(max_output_size)), i32, dev), empty((b,), i32, dev)
    if n == 0 or max_output_size == 0:
        oi.fill_(-1)
        oc.zero_()
        if not (0.0 <= iou_threshold <= 1.0):
            raise ValueError('iou_threshold must be in [0, 1]')
    else:
        _lib.check(lib.bx_nms(h, bw.ptr(b3, FLOAT32, (b, n, 4), 16), bw.ptr(s2, FLOAT32, (b, n)), b, n,
                              int(max_output_size), float(iou_threshold), bw.ptr(oi, INT32, (b, int(max_output_size))),
                              bw.ptr(oc, INT32, (b,)), st))
    return (oi[0], oc[0]) if squeeze else (oi, oc)


def proposal_params(image_shape, post_nms, iou_threshold=0.7, means=(0, 0, 0, 0), stds=(1, 1, 1, 1),
                    pre_nms_top_k=0, min_size=0.0):
    return _lib.ProposalParams(_lib.f4(means), _lib.f4(stds), int(image_shape[0]), int(image_shape[1]),
                               int(pre_nms_top_k), int(post_nms), float(iou_threshold), float(min_size))


def proposals(anchors, deltas, scores, image_shape, post_nms, iou_threshold=0.7, means=(0, 0, 0, 0),
              stds=(1, 1, 1, 1), pre_nms_top_k=0, min_size=0.0):
    """a3 batched: anchors [n,4]; deltas [b,n,4]; scores [b,n] -> (boxes [b,post,4], idx [b,post], count [b])."""
    deltas = to_device(deltas, f32)
    anchors = to_device(anchors, f32, deltas.device)
    scores = to_device(scores, f32, deltas.device)
    b, n = deltas.shape[0], deltas.shape[1]
    dev, h, bw, st, lib = _ctx(deltas)
    p = proposal_params(image_shape, post_nms, iou_threshold, means, stds, pre_nms_top_k, min_size)
    ob, oi, oc = empty((b, post_nms, 4), f32, dev), empty((b, post_nms), i32, dev), empty((b,), i32, dev)
    if n == 0:
        ob.zero_(); oi.fill_(-1); oc.zero_()
        return ob, oi, oc
    _lib.check(lib.bx_proposals(h, bw.ptr(anchors, FLOAT32, (n, 4), 16), bw.ptr(deltas, FLOAT32, (b, n, 4), 16),
                                bw.ptr(scores, FLOAT32, (b, n)), b, n, ctypes.byref(p),
                                bw.ptr(ob, FLOAT32, (b, post_nms, 4), 16), bw.ptr(oi, INT32, (b, post_nms)),
                                bw.ptr(oc, INT32, (b,)), st))
    return ob, oi, oc


def generate_anchors(feat_shapes, strides, offsets, device=None):
    """f2: anchors = (x*stride, y*stride, x*stride, y*stride) + offsets[level][k].  feat_shapes [(fh,fw)] per level;
    offsets [levels, A, 4] (host floats) -> [sum fh*fw*A, 4] on the device."""
    import numpy as np
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.float32))
    nl = len(feat_shapes)
    if off.ndim == 2:
        off = off[None]
    if off.shape[0] != nl or off.shape[2] != 4 or len(strides) != nl:
        raise ValueError('generate_anchors: offsets must be [levels, A, 4] with one (shape, stride) per level')
    a = off.shape[1]
    dev = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    total = sum(int(h_) * int(w_) for h_, w_ in feat_shapes) * a
    out = empty((total, 4), f32, dev)
    st = stream_ptr(dev)
    h = _lib.handle(dev, st.value)
    fh = (ctypes.c_int * nl)(*[int(s_[0]) for s_ in feat_shapes])
    fw = (ctypes.c_int * nl)(*[int(s_[1]) for s_ in feat_shapes])
    sd = (ctypes.c_float * nl)(*[float(v) for v in strides])
    _lib.check(_lib.load().bx_generate_anchors(h, nl, fh, fw, sd, a, off.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                               Borrow(dev).ptr(out, FLOAT32, (total, 4), 16) if total else 0, st))
    return out


def rpn_scores(logits, layout, anchors_per_cell=1):
    """f2: raw RPN logits -> foreground probability.  RPN_CAFFE: [b, cells, 2A] -> [b, cells*A]; RPN_PAIRS: [b, n, 2]
    -> [b, n]."""
    logits = to_device(logits, f32)
    if logits.dim() == 2:
        logits = logits.unsqueeze(0)
    b = logits.shape[0]
    n = logits.shape[1] * (anchors_per_cell if layout == _lib.RPN_CAFFE else 1)
    if logits.shape[2] != (2 * anchors_per_cell if layout == _lib.RPN_CAFFE else 2):
        raise ValueError('rpn_scores: logits last dimension does not match the layout')
    dev, h, bw, st, lib = _ctx(logits)
    out = empty((b, n), f32, dev)
    if b * n:
        _lib.check(lib.bx_rpn_scores(h, bw.ptr(logits, FLOAT32, tuple(logits.shape), 8), layout, anchors_per_cell, b, n,
                                     bw.ptr(out, FLOAT32, (b, n)), st))
    return out


def proposals_rpn(anchors, deltas, logits, layout, anchors_per_cell, image_shape, post_nms, iou_threshold=0.7,
                  means=(0, 0, 0, 0), stds=(1, 1, 1, 1), pre_nms_top_k=0, min_size=0.0, return_scores=False):
    """f2 + a3: proposals straight from the raw RPN logits (softmax fused into the key pass when n <= 24576)."""
    deltas = to_device(deltas, f32)
    anchors = to_device(anchors, f32, deltas.device)
    logits = to_device(logits, f32, deltas.device)
    b, n = deltas.shape[0], deltas.shape[1]
    if logits.numel() != 2 * b * n:
        raise ValueError('proposals_rpn: logits must hold 2 values per anchor')
    dev, h, bw, st, lib = _ctx(deltas)
    p = proposal_params(image_shape, post_nms, iou_threshold, means, stds, pre_nms_top_k, min_size)
    ob, oi, oc = empty((b, post_nms, 4), f32, dev), empty((b, post_nms), i32, dev), empty((b,), i32, dev)
    osc = empty((b, n), f32, dev) if return_scores else None
    if n == 0:
        ob.zero_(); oi.fill_(-1); oc.zero_()
    else:
        _lib.check(lib.bx_proposals_rpn(h, bw.ptr(anchors, FLOAT32, (n, 4), 16), bw.ptr(deltas, FLOAT32, (b, n, 4), 16),
                                        bw.ptr(logits, FLOAT32, tuple(logits.shape), 8), layout, anchors_per_cell, b, n,
                                        ctypes.byref(p), bw.ptr(ob, FLOAT32, (b, post_nms, 4), 16),
                                        bw.ptr(oi, INT32, (b, post_nms)), bw.ptr(oc, INT32, (b,)),
                                        bw.ptr(osc, FLOAT32, (b, n)) if return_scores else None, st))
    return (ob, oi, oc, osc) if return_scores else (ob, oi, oc)


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    """tf.image.crop_and_resize (bilinear): image [b,h,w,c]; boxes [r,4] (y1,x1,y2,x2) normalised; box_ind [r]."""
    image = to_device(image, f32)
    boxes = to_device(boxes, f32, image.device)
    box_ind = to_device(box_ind, i32, image.device)
    ch, cw = int(crop_size[0]), int(crop_size[1])
    b, ih, iw, c = image.shape
    r = boxes.shape[0]
    dev, h, bw, st, lib = _ctx(image)
    out = empty((r, ch, cw, c), f32, dev)
    if ch <= 0 or cw <= 0:
        raise ValueError('crop_size must be 2 positive ints')
    if r:
        _lib.check(lib.bx_crop_and_resize(h, bw.ptr(image, FLOAT32, (b, ih, iw, c)), b, ih, iw, c,
                                          bw.ptr(boxes, FLOAT32, (r, 4), 16), bw.ptr(box_ind, INT32, (r,)), r, ch, cw,
                                          float(extrapolation_value), bw.ptr(out, FLOAT32, (r, ch, cw, c)), st))
    return out


def roi_pool(mode, pool, pool_size, feat, rois, stride=16.0, image_shape=(0, 0), box_ind=None, roi_counts=None):
    """a4/a5/a8: feat [b,fh,fw,c]; rois [r,4] (or [b,k,4] with roi_counts [b]) -> [r,P,P,c]."""
    feat = to_device(feat, f32)
    rois = to_device(rois, f32, feat.device).reshape(-1, 4)
    b, fh, fw, c = feat.shape
    r = rois.shape[0]
    dev, h, bw, st, lib = _ctx(feat)
    out = empty((r, pool_size, pool_size, c), f32, dev)
    if r:
        bi = bw.ptr(to_device(box_ind, i32, feat.device), INT32, (r,)) if box_ind is not None else None
        rc = bw.ptr(to_device(roi_counts, i32, feat.device), INT32, (b,)) if roi_counts is not None else None
        _lib.check(lib.bx_roi_pool(h, mode, pool, int(pool_size), bw.ptr(feat, FLOAT32, (b, fh, fw, c)), b, fh, fw, c,
                                   bw.ptr(rois, FLOAT32, (r, 4), 16), bi, rc, r, float(stride), int(image_shape[0]),
                                   int(image_shape[1]), bw.ptr(out, FLOAT32, (r, pool_size, pool_size, c)), st))
    return out


def roi_pool_grad(mode, pool, pool_size, feat, rois, grad_out, stride=16.0, image_shape=(0, 0), box_ind=None,
                  roi_counts=None):
    """f3: gradient of roi_pool w.r.t. feat: grad_out [r,P,P,c] -> grad_feat [b,fh,fw,c].  Under
    torch.use_deterministic_algorithms(True) every mode runs on the row-owned, atomic-free kernel (bit-reproducible)."""
    feat = to_device(feat, f32)
    rois = to_device(rois, f32, feat.device).reshape(-1, 4)
    grad_out = to_device(grad_out, f32, feat.device)
    b, fh, fw, c = feat.shape
    r = rois.shape[0]
    dev, h, bw, st, lib = _ctx(feat)
    gf = empty((b, fh, fw, c), f32, dev)
    _lib.check(lib.bx_set_deterministic(h, int(torch.are_deterministic_algorithms_enabled())))
    bi = bw.ptr(to_device(box_ind, i32, feat.device), INT32, (r,)) if (box_ind is not None and r) else None
    rc = bw.ptr(to_device(roi_counts, i32, feat.device), INT32, (b,)) if roi_counts is not None else None
    _lib.check(lib.bx_roi_pool_grad(h, mode, pool, int(pool_size), bw.ptr(feat, FLOAT32, (b, fh, fw, c), 16), b, fh, fw, c,
                                    bw.ptr(rois, FLOAT32, (r, 4), 16) if r else 0, bi, rc, r, float(stride),
                                    int(image_shape[0]), int(image_shape[1]),
                                    bw.ptr(grad_out, FLOAT32, (r, pool_size, pool_size, c), 16) if r else 0,
                                    bw.ptr(gf, FLOAT32, (b, fh, fw, c), 16), st))
    return gf


class _RoiPoolFn(torch.autograd.Function):
    """roi_pool with a feature-map gradient (the boxes are constants, as under tf.stop_gradient in the reference)."""

    @staticmethod
    def forward(ctx, feat, rois, mode, pool, pool_size, stride, image_shape, box_ind, roi_counts):
        out = roi_pool(mode, pool, pool_size, feat, rois, stride, image_shape, box_ind, roi_counts)
        ctx.save_for_backward(feat.detach(), rois.detach())
        ctx.cfg = (mode, pool, pool_size, stride, image_shape, box_ind, roi_counts)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat, rois = ctx.saved_tensors
        mode, pool, pool_size, stride, image_shape, box_ind, roi_counts = ctx.cfg
        gf = roi_pool_grad(mode, pool, pool_size, feat, rois, grad_out.contiguous(), stride, image_shape, box_ind, roi_counts)
        return gf, None, None, None, None, None, None, None, None


def roi_pool_autograd(mode, pool, pool_size, feat, rois, stride=16.0, image_shape=(0, 0), box_ind=None, roi_counts=None):
    """roi_pool that records a backward pass when `feat` requires grad (training drop-in); plain roi_pool otherwise."""
    if isinstance(feat, torch.Tensor) and feat.requires_grad and torch.is_grad_enabled():
        rois_t = to_device(rois, f32, feat.device)
        return _RoiPoolFn.apply(feat, rois_t, mode, pool, pool_size, stride, image_shape, box_ind, roi_counts)
    return roi_pool(mode, pool, pool_size, feat, rois, stride, image_shape, box_ind, roi_counts)


def fpn_assign_levels(rois, min_level=2, max_level=5):
    """a6: rois [r,4] -> (level [r] int32, order [r] int32 level-major stable, counts [L] int32)."""
    rois = to_device(rois, f32)
    r = rois.shape[0]
    nl = max_level - min_level + 1
    dev, h, bw, st, lib = _ctx(rois)
    lv, order, counts = empty((r,), i32, dev), empty((r,), i32, dev), torch.zeros((nl,), dtype=i32, device=rois.device)
    if r:
        _lib.check(lib.bx_fpn_assign_levels(h, bw.ptr(rois, FLOAT32, (r, 4), 16), r, min_level, max_level,
                                            bw.ptr(lv, INT32, (r,)), bw.ptr(order, INT32, (r,)),
                                            bw.ptr(counts, INT32, (nl,)), st))
    return lv, order, counts


def fpn_roi_features(feats, rois, image_shape, pool_size=7, min_level=2, box_ind=None):
    """a6+a7 fused: feats = [P2..P5] each [b,fh,fw,c]; rois [r,4] -> (features [r,P,P,c] level-major, level, order, counts)."""
    feats = [to_device(f, f32) for f in feats]
    rois = to_device(rois, f32, feats[0].device)
    r = rois.shape[0]
    nl = len(feats)
    b, c = feats[0].shape[0], feats[0].shape[3]
    dev, h, bw, st, lib = _ctx(feats[0])
    out = empty((r, pool_size, pool_size, c), f32, dev)
    lv, order, counts = empty((r,), i32, dev), empty((r,), i32, dev), torch.zeros((nl,), dtype=i32, device=rois.device)
    if r:
        ptrs = (c_void_p * nl)(*[bw.ptr(f, FLOAT32, (b, f.shape[1], f.shape[2], c)) for f in feats])
        fh = (c_int * nl)(*[f.shape[1] for f in feats])
        fw = (c_int * nl)(*[f.shape[2] for f in feats])
        bi = bw.ptr(to_device(box_ind, i32, rois.device), INT32, (r,)) if box_ind is not None else None
        _lib.check(lib.bx_fpn_roi_features(h, ptrs, fh, fw, nl, min_level, b, c, bw.ptr(rois, FLOAT32, (r, 4), 16), bi,
                                           r, int(image_shape[0]), int(image_shape[1]), int(pool_size),
                                           bw.ptr(out, FLOAT32, (r, pool_size, pool_size, c)), bw.ptr(lv, INT32, (r,)),
                                           bw.ptr(order, INT32, (r,)), bw.ptr(counts, INT32, (nl,)), st))
    return out, lv, order, counts


def pairwise_iou(b1, b2):
    """a9: utils/bbox_tf.py:37-56 -> [n,m] fp32."""
    b1 = to_device(b1, f32)
    b2 = to_device(b2, f32, b1.device)
    n, m = b1.shape[0], b2.shape[0]
    dev, h, bw, st, lib = _ctx(b1)
    out = empty((n, m), f32, dev)
    if n * m:
        _lib.check(lib.bx_pairwise_iou(h, bw.ptr(b1, FLOAT32, (n, 4), 16), n, bw.ptr(b2, FLOAT32, (m, 4), 16), m,
                                       bw.ptr(out, FLOAT32, (n, m)), st))
    return out


def anchor_target(anchors, gt, perm, image_shape, pos_iou_threshold=0.7, neg_iou_threshold=0.3,
                  total_num_samples=256, max_pos_samples=128, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), gt_counts=None):
    """a10 batched: anchors [n,4]; gt [b,m,4]; perm [b,n] int32 -> labels [b,n], targets/in_w/out_w [b,n,4], counts [b,2]."""
    anchors = to_device(anchors, f32)
    gt = to_device(gt, f32, anchors.device)
    perm = to_device(perm, i32, anchors.device)
    n = anchors.shape[0]
    b, m = gt.shape[0], gt.shape[1]
    dev, h, bw, st, lib = _ctx(anchors)
    p = _lib.AnchorTargetParams(float(pos_iou_threshold), float(neg_iou_threshold), int(total_num_samples),
                                int(max_pos_samples), _lib.f4(means), _lib.f4(stds), int(image_shape[0]),
                                int(image_shape[1]))
    lab = empty((b, n), f32, dev)
    tg, iw, ow = empty((b, n, 4), f32, dev), empty((b, n, 4), f32, dev), empty((b, n, 4), f32, dev)
    cnt = empty((b, 2), i32, dev)
    gc = bw.ptr(to_device(gt_counts, i32, anchors.device), INT32, (b,)) if gt_counts is not None else None
    _lib.check(lib.bx_anchor_target(h, bw.ptr(anchors, FLOAT32, (n, 4), 16), n,
                                    bw.ptr(gt, FLOAT32, (b, m, 4), 16) if m else 0, gc, b, m,
                                    bw.ptr(perm, INT32, (b, n)), ctypes.byref(p), bw.ptr(lab, FLOAT32, (b, n)),
                                    bw.ptr(tg, FLOAT32, (b, n, 4), 16), bw.ptr(iw, FLOAT32, (b, n, 4), 16),
                                    bw.ptr(ow, FLOAT32, (b, n, 4), 16), bw.ptr(cnt, INT32, (b, 2)), st))
    return lab, tg, iw, ow, cnt


def proposal_target(rois, gt, gt_labels, perm, num_classes=21, pos_iou_threshold=0.5, neg_iou_threshold=0.5,
                    total_num_samples=128, max_pos_samples=32, means=(0, 0, 0, 0), stds=(1, 1, 1, 1),
                    roi_counts=None, gt_counts=None):
    """a11 batched: rois [b,k,4]; gt [b,m,4]; gt_labels [b,m] int32; perm [b,k] int32 ->
    (rois [b,S,4], labels [b,S] int32, targets, in_w, out_w [b,S,4C], keep [b,S] int32, counts [b,2])."""
    rois = to_device(rois, f32)
    gt = to_device(gt, f32, rois.device)
    gt_labels = to_device(gt_labels, i32, rois.device)
    perm = to_device(perm, i32, rois.device)
    b, k = rois.shape[0], rois.shape[1]
    m = gt.shape[1]
    S, C = int(total_num_samples), int(num_classes)
    dev, h, bw, st, lib = _ctx(rois)
    p = _lib.ProposalTargetParams(C, float(pos_iou_threshold), float(neg_iou_threshold), S, int(max_pos_samples),
                                  _lib.f4(means), _lib.f4(stds))
    o_rois, o_lab = empty((b, S, 4), f32, dev), empty((b, S), i32, dev)
    o_t, o_i, o_o = empty((b, S, 4 * C), f32, dev), empty((b, S, 4 * C), f32, dev), empty((b, S, 4 * C), f32, dev)
    o_keep, o_cnt = empty((b, S), i32, dev), empty((b, 2), i32, dev)
    rc = bw.ptr(to_device(roi_counts, i32, rois.device), INT32, (b,)) if roi_counts is not None else None
    gc = bw.ptr(to_device(gt_counts, i32, rois.device), INT32, (b,)) if gt_counts is not None else None
    _lib.check(lib.bx_proposal_target(h, bw.ptr(rois, FLOAT32, (b, k, 4), 16) if k else 0, rc, k,
                                      bw.ptr(gt, FLOAT32, (b, m, 4), 16) if m else 0,
                                      bw.ptr(gt_labels, INT32, (b, m)) if m else 0, gc, b, m,
                                      bw.ptr(perm, INT32, (b, k)) if k else 0, ctypes.byref(p),
                                      bw.ptr(o_rois, FLOAT32, (b, S, 4), 16), bw.ptr(o_lab, INT32, (b, S)),
                                      bw.ptr(o_t, FLOAT32, (b, S, 4 * C)), bw.ptr(o_i, FLOAT32, (b, S, 4 * C)),
                                      bw.ptr(o_o, FLOAT32, (b, S, 4 * C)), bw.ptr(o_keep, INT32, (b, S)),
                                      bw.ptr(o_cnt, INT32, (b, 2)), st))
    return o_rois, o_lab, o_t, o_i, o_o, o_keep, o_cnt


def smooth_l1_loss(pred, target, in_w, out_w, sigma=1.0, dim=(1,), with_grad=False):
    """f3: model/losses.py:16-28.  pred/target/in_w/out_w [n,d] -> loss [] (0-dim fp32), optionally d loss / d pred."""
    dim = tuple(int(v) for v in dim)
    if dim not in ((1,), (0, 1)):
        raise NotImplementedError('smooth_l1_loss: dim must be [1] or [0, 1] (the two forms the reference uses)')
    pred = to_device(pred, f32)
    n, d = pred.shape
    dev, h, bw, st, lib = _ctx(pred)
    args = [bw.ptr(to_device(t, f32, pred.device), FLOAT32, (n, d)) if n else 0 for t in (target, in_w, out_w)]
    loss = empty((1,), f32, dev)
    grad = empty((n, d), f32, dev) if with_grad else None
    _lib.check(lib.bx_smooth_l1_loss(h, bw.ptr(pred, FLOAT32, (n, d)) if n else 0, *args, n, d, float(sigma),
                                     int(dim == (0, 1)), bw.ptr(loss, FLOAT32, (1,)),
                                     bw.ptr(grad, FLOAT32, (n, d)) if (with_grad and n) else None, st))
    return (loss.reshape(()), grad) if with_grad else loss.reshape(())


def cls_loss(logits, labels, weight=1.0, with_grad=False):
    """f3: model/losses.py:4-13 with the `labels >= 0` gather folded in.  logits [n,c]; labels [n] -> loss []."""
    logits = to_device(logits, f32)
    n, c = logits.shape
    labels = to_device(labels, f32, logits.device).reshape(-1)
    dev, h, bw, st, lib = _ctx(logits)
    loss = empty((1,), f32, dev)
    grad = empty((n, c), f32, dev) if with_grad else None
    _lib.check(lib.bx_cls_loss(h, bw.ptr(logits, FLOAT32, (n, c)) if n else 0, bw.ptr(labels, FLOAT32, (n,)) if n else 0,
                               n, c, float(weight), bw.ptr(loss, FLOAT32, (1,)), None,
                               bw.ptr(grad, FLOAT32, (n, c)) if (with_grad and n) else None, st))
    return (loss.reshape(()), grad) if with_grad else loss.reshape(())


class _LossFn(torch.autograd.Function):
    """Loss + gradient from one library call; backward only scales the stored gradient."""

    @staticmethod
    def forward(ctx, x, fn, args):
        loss, grad = fn(x.detach(), *args, with_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def post_ops_prediction(scores, deltas, rois, image_shape, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), max_num_per_class=50,
                        max_num_per_image=150, nms_iou_threshold=0.3, score_threshold=0.05, extractor_stride=16,
                        roi_counts=None):
    """f1 batched: scores [b,r,C] softmax; deltas [b,r,C,4] (or [b,r,4C]); rois [b,r,4] ->
    (det [b,max_per_image,6] = (x1,y1,x2,y2,score,class) zero padded, count [b]).  No host sync."""
    scores = to_device(scores, f32)
    b, r, c = scores.shape
    deltas = to_device(deltas, f32, scores.device).reshape(b, r, c, 4)
    rois = to_device(rois, f32, scores.device)
    dev, h, bw, st, lib = _ctx(scores)
    p = _lib.PredictionParams(_lib.f4(means), _lib.f4(stds), int(image_shape[0]), int(image_shape[1]), int(c),
                              int(max_num_per_class), int(max_num_per_image), float(nms_iou_threshold),
                              float(score_threshold), float(extractor_stride if extractor_stride is not None else 0))
    det, cnt = empty((b, int(max_num_per_image), 6), f32, dev), empty((b,), i32, dev)
    rc = bw.ptr(to_device(roi_counts, i32, scores.device), INT32, (b,)) if roi_counts is not None else None
    _lib.check(lib.bx_post_ops_prediction(h, bw.ptr(scores, FLOAT32, (b, r, c)) if r else 0,
                                          bw.ptr(deltas, FLOAT32, (b, r, c, 4), 16) if r else 0,
                                          bw.ptr(rois, FLOAT32, (b, r, 4), 16) if r else 0, rc, b, r, ctypes.byref(p),
                                          bw.ptr(det, FLOAT32, (b, int(max_num_per_image), 6)), bw.ptr(cnt, INT32, (b,)), st))
    return det, cnt


def eval_detections(scores, deltas, rois, image_sizes=None, img_scale=None, image_shape=(0, 0), means=(0, 0, 0, 0),
                    stds=(1, 1, 1, 1), max_num_per_class=50, max_num_per_image=150, nms_iou_threshold=0.3,
                    score_threshold=0.05, min_size=10, cut=_lib.CUT_TOP_K, out_rows=None, roi_counts=None):
    """f1, evaluation-loop form (bx_eval_detections): as post_ops_prediction, with the rois divided by `img_scale` [b]
    first, boxes clipped to each image's own raw size `image_sizes` [b,2] = (raw_h, raw_w), and the per-image cut either
    top-k or the VOC loop's `score >= k-th largest` (ties kept, up to `out_rows` rows).  -> (det [b,out_rows,6], count [b])."""
    scores = to_device(scores, f32)
    b, r, c = scores.shape
    deltas = to_device(deltas, f32, scores.device).reshape(b, r, c, 4)
    rois = to_device(rois, f32, scores.device)
    dev, h, bw, st, lib = _ctx(scores)
    rows = int(max_num_per_image if out_rows is None else out_rows)
    p = _lib.PredictionParams(_lib.f4(means), _lib.f4(stds), int(image_shape[0]), int(image_shape[1]), int(c),
                              int(max_num_per_class), int(max_num_per_image), float(nms_iou_threshold),
                              float(score_threshold), float(min_size if min_size is not None else 0))
    det, cnt = empty((b, rows, 6), f32, dev), empty((b,), i32, dev)
    rc = bw.ptr(to_device(roi_counts, i32, scores.device), INT32, (b,)) if roi_counts is not None else None
    isz = bw.ptr(to_device(image_sizes, f32, scores.device), FLOAT32, (b, 2)) if image_sizes is not None else None
    isc = bw.ptr(to_device(img_scale, f32, scores.device), FLOAT32, (b,)) if img_scale is not None else None
    _lib.check(lib.bx_eval_detections(h, bw.ptr(scores, FLOAT32, (b, r, c)) if r else 0,
                                      bw.ptr(deltas, FLOAT32, (b, r, c, 4), 16) if r else 0,
                                      bw.ptr(rois, FLOAT32, (b, r, 4), 16) if r else 0, rc, isz, isc, b, r, ctypes.byref(p),
                                      int(cut), rows, bw.ptr(det, FLOAT32, (b, rows, 6)), bw.ptr(cnt, INT32, (b,)), st))
    return det, cnt


def c4_proposal_roi(anchors, deltas, scores, feat, image_shape, post_nms, stride=16.0, pool_size=7,
                    max_pooling_flag=False, iou_threshold=0.7, means=(0, 0, 0, 0), stds=(1, 1, 1, 1), pre_nms_top_k=0,
                    min_size=0.0, out=None):
    """Composite used by BaseFasterRcnn eval (base_faster_rcnn_model.py:153,182) and the benchmark, batched:
    -> (rois [b,post,4], idx [b,post], count [b], roi_features [b*post,P,P,c])."""
    deltas = to_device(deltas, f32)
    anchors = to_device(anchors, f32, deltas.device)
    scores = to_device(scores, f32, deltas.device)
    feat = to_device(feat, f32, deltas.device)
    b, n = deltas.shape[0], deltas.shape[1]
    _, fh, fw, c = feat.shape
    dev, h, bw, st, lib = _ctx(deltas)
    p = proposal_params(image_shape, post_nms, iou_threshold, means, stds, pre_nms_top_k, min_size)
    if out is None:
        out = (empty((b, post_nms, 4), f32, dev), empty((b, post_nms), i32, dev), empty((b,), i32, dev),
               empty((b * post_nms, pool_size, pool_size, c), f32, dev))
    ob, oi, oc, of = out
    _lib.check(lib.bx_c4_proposal_roi(h, bw.ptr(anchors, FLOAT32, (n, 4), 16), bw.ptr(deltas, FLOAT32, (b, n, 4), 16),
                                      bw.ptr(scores, FLOAT32, (b, n)), bw.ptr(feat, FLOAT32, (b, fh, fw, c)), b, n, fh,
                                      fw, c, ctypes.byref(p), float(stride), int(pool_size),
                                      _lib.POOL_MAX2 if max_pooling_flag else _lib.POOL_NONE,
                                      bw.ptr(ob, FLOAT32, (b, post_nms, 4), 16), bw.ptr(oi, INT32, (b, post_nms)),
                                      bw.ptr(oc, INT32, (b,)),
                                      bw.ptr(of, FLOAT32, (b * post_nms, pool_size, pool_size, c)), st))
    return ob, oi, oc, of


def c4_proposal_roi_host(anchors_dev, deltas_h, scores_h, feat_h, image_shape, post_nms, out_h, stride=16.0,
                         pool_size=7, max_pooling_flag=False, iou_threshold=0.7, means=(0, 0, 0, 0),
                         stds=(1, 1, 1, 1), pre_nms_top_k=0, min_size=0.0):
    """Same composite through HOST buffers (bx_c4_proposal_roi_host): deltas/scores/feat and the four outputs are CPU
    torch tensors (pinned for full PCIe rate); copies are enqueued on torch's current stream around the kernels.
    out_h = (rois [b,post,4] f32, idx [b,post] i32, count [b] i32, feat [b*post,P,P,c] f32) CPU tensors."""
    anchors_dev = to_device(anchors_dev, f32)
    for t in (deltas_h, scores_h, feat_h) + tuple(out_h):
        if t.is_cuda or not t.is_contiguous():
            raise ValueError('c4_proposal_roi_host expects contiguous CPU tensors')
    b, n = deltas_h.shape[0], deltas_h.shape[1]
    _, fh, fw, c = feat_h.shape
    dev, h, bw, st, lib = _ctx(anchors_dev)
    p = proposal_params(image_shape, post_nms, iou_threshold, means, stds, pre_nms_top_k, min_size)
    ob, oi, oc, of = out_h
    _lib.check(lib.bx_c4_proposal_roi_host(h, bw.ptr(anchors_dev, FLOAT32, (n, 4), 16), deltas_h.data_ptr(),
                                           scores_h.data_ptr(), feat_h.data_ptr(), b, n, fh, fw, c, ctypes.byref(p),
                                           float(stride), int(pool_size),
                                           _lib.POOL_MAX2 if max_pooling_flag else _lib.POOL_NONE, ob.data_ptr(),
                                           oi.data_ptr(), oc.data_ptr(), of.data_ptr(), st))
    return out_h
