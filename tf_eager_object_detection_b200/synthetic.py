"""Synthetic inputs for the box-processing path (host-side numpy; used by tests and bench.py).

Shapes and distributions follow SURVEY.md §8(d) / BASELINE.md "Synthetic inputs":
seed = 20260000 + 100*cfg + image_index, `np.random.default_rng(seed)`.

The two anchor generators restate what the reference's models feed the path with
(`object_detection/utils/anchor_generator.py:63-81,46-60` for C4, `:137-178` for FPN);
`oracle/make_golden.py` asserts them equal to the reference's own generators (and `tests/test_oracle_golden.py` pins the hashes).
"""
import math

import numpy as np

F = np.float32

# BASELINE.json configs -> (H, W)
IMAGE_600x1000 = (600, 1000)
IMAGE_800x1333 = (800, 1333)
FPN_STRIDES = (4, 8, 16, 32, 64)
FPN_BASE_SIZES = (32, 64, 128, 256, 512)


def seed_for(cfg, image_index=0):
    return 20260000 + 100 * int(cfg) + int(image_index)


# ----------------------------------------------------------------------------- anchors
def anchor_base(base_size=16, ratios=(0.5, 1.0, 2.0), scales=(8, 16, 32)):
    """py-faster-rcnn anchor table (reference: utils/anchor_generator.py:63-134).
    A window (0,0,base-1,base-1) is reshaped per ratio (rounded w/h), then scaled about its centre."""
    out = []
    w0 = h0 = float(base_size)
    ctr = 0.5 * (base_size - 1)
    for r in ratios:
        w = np.round(np.sqrt(w0 * h0 / r))
        h = np.round(w * r)
        for s in scales:
            ws, hs = w * s, h * s
            out.append([ctr - 0.5 * (ws - 1), ctr - 0.5 * (hs - 1), ctr + 0.5 * (ws - 1), ctr + 0.5 * (hs - 1)])
    return np.asarray(out, dtype=np.float64)


def c4_anchors(feat_h, feat_w, stride=16, ratios=(0.5, 1.0, 2.0), scales=(8, 16, 32)):
    """[feat_h*feat_w*A, 4] (x1,y1,x2,y2), cell-major / anchor-minor, shifts in (x,y,x,y) order
    (reference: utils/anchor_generator.py:46-60 as called at base_faster_rcnn_model.py:139-142)."""
    base = anchor_base(stride, ratios, scales).astype(F)
    sx = (np.arange(feat_w, dtype=np.int64) * stride).astype(F)
    sy = (np.arange(feat_h, dtype=np.int64) * stride).astype(F)
    gx, gy = np.meshgrid(sx, sy)
    shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)
    return (shifts[:, None, :] + base[None, :, :]).reshape(-1, 4).astype(F)


def fpn_level_anchors(base_size, feat_h, feat_w, stride, ratios=(0.5, 1.0, 2.0), scales=(1.0,)):
    """One FPN level (reference: utils/anchor_generator.py:137-178 `make_anchors`): centres at
    (j*stride, i*stride); per scale then per ratio r a box of width size*sqrt(r), height size/sqrt(r)
    (the reference's enum_ratios returns (hs, ws) swapped, :178, so ratio 0.5 is the TALL box)."""
    sq = np.sqrt(np.asarray(ratios, F))
    size = (F(base_size) * np.asarray(scales, F))
    ws = (size[None, :] * sq[:, None]).reshape(-1).astype(F)
    hs = (size[None, :] / sq[:, None]).reshape(-1).astype(F)
    cx = (np.arange(feat_w, dtype=F) * F(stride))
    cy = (np.arange(feat_h, dtype=F) * F(stride))
    gx, gy = np.meshgrid(cx, cy)
    gx = gx.reshape(-1, 1); gy = gy.reshape(-1, 1)
    x1 = gx - F(0.5) * ws[None, :]; x2 = gx + F(0.5) * ws[None, :]
    y1 = gy - F(0.5) * hs[None, :]; y2 = gy + F(0.5) * hs[None, :]
    return np.stack([x1, y1, x2, y2], axis=2).reshape(-1, 4).astype(F)


def fpn_feature_shapes(image_hw, strides=FPN_STRIDES):
    return [(math.ceil(image_hw[0] / s), math.ceil(image_hw[1] / s)) for s in strides]


def fpn_anchors(image_hw, strides=FPN_STRIDES, base_sizes=FPN_BASE_SIZES):
    """P2..P6 concatenation (reference: fpn/base_fpn_model.py:163-186)."""
    shapes = fpn_feature_shapes(image_hw, strides)
    return np.concatenate([fpn_level_anchors(b, h, w, s) for b, (h, w), s in zip(base_sizes, shapes, strides)], axis=0)


# ----------------------------------------------------------------------------- per-image tensors
def rpn_outputs(rng, n):
    """deltas [n,4] ~ N(0, (.2,.2,.3,.3)); scores [n] = fixed-seed permutation -> unique fp32 (n < 2^24)."""
    deltas = (rng.normal(0.0, 1.0, (n, 4)) * np.asarray([0.2, 0.2, 0.3, 0.3])).astype(F)
    scores = ((rng.permutation(n) + 1) / (n + 1)).astype(F)
    return deltas, scores


def features(rng, h, w, c):
    return rng.standard_normal((h, w, c), dtype=F)


def random_rois(rng, r, image_hw):
    """r boxes, log-uniform side 8..800 px, uniform centres, clipped to the image."""
    H, W = image_hw
    side_w = np.exp(rng.uniform(np.log(8.0), np.log(800.0), r))
    side_h = np.exp(rng.uniform(np.log(8.0), np.log(800.0), r))
    cx = rng.uniform(0, W - 1, r); cy = rng.uniform(0, H - 1, r)
    b = np.stack([cx - side_w / 2, cy - side_h / 2, cx + side_w / 2, cy + side_h / 2], axis=1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, W - 1); b[:, 1::2] = np.clip(b[:, 1::2], 0, H - 1)
    return b.astype(F)


def gt_boxes(rng, m, image_hw, num_classes=21):
    """m integer-cornered boxes with sides log-uniform 16..400 inside the image; labels in 1..C-1."""
    H, W = image_hw
    w = np.minimum(np.exp(rng.uniform(np.log(16.0), np.log(400.0), m)), W - 1).astype(np.int64)
    h = np.minimum(np.exp(rng.uniform(np.log(16.0), np.log(400.0), m)), H - 1).astype(np.int64)
    x1 = (rng.uniform(0, 1, m) * (W - w)).astype(np.int64)
    y1 = (rng.uniform(0, 1, m) * (H - h)).astype(np.int64)
    b = np.stack([x1, y1, x1 + w - 1, y1 + h - 1], axis=1).astype(F)
    labels = rng.integers(1, num_classes, m).astype(np.int32)
    return b, labels


def c4_image(cfg, image_index, image_hw=IMAGE_600x1000, stride=16, channels=1024, with_features=True):
    """One C4 image: dict(anchors, deltas, scores, feat[h,w,C])."""
    rng = np.random.default_rng(seed_for(cfg, image_index))
    fh, fw = math.ceil(image_hw[0] / stride), math.ceil(image_hw[1] / stride)
    anchors = c4_anchors(fh, fw, stride)
    deltas, scores = rpn_outputs(rng, anchors.shape[0])
    out = dict(anchors=anchors, deltas=deltas, scores=scores, image_shape=list(image_hw), feat_hw=(fh, fw))
    if with_features:
        out['feat'] = features(rng, fh, fw, channels)
    return out


def fpn_image(cfg, image_index, image_hw=IMAGE_600x1000, channels=256, with_features=True):
    """One FPN image: dict(anchors (P2..P6), deltas, scores, feats [P2..P5])."""
    rng = np.random.default_rng(seed_for(cfg, image_index))
    anchors = fpn_anchors(image_hw)
    deltas, scores = rpn_outputs(rng, anchors.shape[0])
    out = dict(anchors=anchors, deltas=deltas, scores=scores, image_shape=list(image_hw))
    if with_features:
        out['feats'] = [features(rng, h, w, channels) for (h, w) in fpn_feature_shapes(image_hw)[:4]]
    return out


def roi_head_outputs(rng, r, num_classes=21):
    """Synthetic RoI-head outputs for the post-head filtering stage: softmax scores [r,C] with a few confident
    foreground classes per roi (unique values), class-specific deltas [r,C,4] ~ N(0, 1)."""
    logits = rng.normal(0.0, 1.0, (r, num_classes)) + 4.0 * (rng.random((r, num_classes)) < 0.08)
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    scores = (e / e.sum(axis=1, keepdims=True)).astype(F)
    deltas = rng.normal(0.0, 1.0, (r, num_classes, 4)).astype(F)
    return scores, deltas
