"""CPU ORACLE (numpy) for the box-processing hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product package (`tf_eager_object_detection_b200`) never does.

Every function restates one reference function (citations relative to
`/root/reference/object_detection/`) in plain fp32 numpy, keeping the reference's op order so that
integer/index results are bit-exact and fp32 results differ only through `exp`/`log` ulps.

PARITY STATUS: **parity unpinned at the TensorFlow-kernel boundary.**  The reference ships no
tests, fixtures or golden vectors, and TensorFlow (inferred 1.13, unpinned) is not installable
here, so `nms_tf`, `crop_and_resize_tf`, `max_pool_2x2`, `avg_pool_2x2` restate the TF r1.13 CPU
kernels from SURVEY.md Appendix B.  What IS pinned:
  * the reference's own Python control flow — `oracle/make_golden.py` executes the reference's
    files unmodified on `oracle/tf_shim` and `tests/test_oracle_golden.py` checks this module
    against those outputs (committed under `tests/golden/`);
  * `pairwise_iou` against the reference's importable numpy twin `utils/bbox_np.py:42-55`;
  * NMS / IoU / bilinear sampling against independent witnesses (torchvision.ops.nms, box_iou,
    torch grid_sample) in `tests/test_oracle_witness.py`.
"""
import numpy as np

F = np.float32


# --------------------------------------------------------------------------- a1 / a12 codecs
def decode_bbox(anchors, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """utils/bbox_transform.py:32-55 `decode_bbox_with_mean_and_std` (note x2 = x1 + w, no -1)."""
    a = np.asarray(anchors, F)
    d = np.asarray(deltas, F) * np.asarray(stds, F) + np.asarray(means, F)      # :37
    w = a[:, 2] - a[:, 0] + F(1)                                                # :40
    h = a[:, 3] - a[:, 1] + F(1)                                                # :41
    cx = a[:, 0] + F(0.5) * w                                                   # :42
    cy = a[:, 1] + F(0.5) * h                                                   # :43
    cx = cx + d[:, 0] * w                                                       # :45
    cy = cy + d[:, 1] * h                                                       # :46
    w = w * np.exp(d[:, 2])                                                     # :47
    h = h * np.exp(d[:, 3])                                                     # :48
    x1 = cx - F(0.5) * w                                                        # :50
    y1 = cy - F(0.5) * h                                                        # :51
    return np.stack([x1, y1, x1 + w, y1 + h], axis=1).astype(F)                 # :52-54


def encode_bbox(src, dst, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """utils/bbox_transform.py:4-29 `encode_bbox_with_mean_and_std`."""
    b = np.asarray(src, F)
    g = np.asarray(dst, F)
    w = b[..., 2] - b[..., 0] + F(1)
    h = b[..., 3] - b[..., 1] + F(1)
    cx = b[..., 0] + F(0.5) * w
    cy = b[..., 1] + F(0.5) * h
    gw = g[..., 2] - g[..., 0] + F(1)
    gh = g[..., 3] - g[..., 1] + F(1)
    gcx = g[..., 0] + F(0.5) * gw
    gcy = g[..., 1] + F(0.5) * gh
    with np.errstate(divide='ignore', invalid='ignore'):
        d = np.stack([(gcx - cx) / w, (gcy - cy) / h, np.log(gw / w), np.log(gh / h)], axis=-1)
        return ((d - np.asarray(means, F)) / np.asarray(stds, F)).astype(F)


# --------------------------------------------------------------------------- a2 clip / filters
def bboxes_clip_filter(boxes, min_value, max_height, max_width, min_edge=None):
    """utils/bbox_tf.py:59-84.  Returns (boxes, idx); idx is int32 arange when min_edge is None
    (tf.range, :77-78) else int64 ascending kept indices (tf.where, :83)."""
    b = np.asarray(boxes, F).copy()
    lo = F(min_value)
    b[:, 0] = np.maximum(np.minimum(b[:, 0], F(max_width - 1)), lo)             # :71
    b[:, 1] = np.maximum(np.minimum(b[:, 1], F(max_height - 1)), lo)            # :72
    b[:, 2] = np.maximum(np.minimum(b[:, 2], F(max_width - 1)), lo)             # :73
    b[:, 3] = np.maximum(np.minimum(b[:, 3], F(max_height - 1)), lo)            # :74
    if min_edge is None:
        return b, np.arange(b.shape[0], dtype=np.int32)
    me = F(min_edge)
    keep = ((b[:, 2] - b[:, 0] + F(1)) >= me) & ((b[:, 3] - b[:, 1] + F(1)) >= me)   # :80-83
    idx = np.nonzero(keep)[0].astype(np.int64)
    return b[idx], idx


def bboxes_range_filter(anchors, max_height, max_width):
    """utils/bbox_tf.py:87-101: indices (int64) of anchors fully inside the image."""
    a = np.asarray(anchors, F)
    ok = (a[:, 0] >= 0) & (a[:, 1] >= 0) & (a[:, 2] <= F(max_width - 1)) & (a[:, 3] <= F(max_height - 1))
    return np.nonzero(ok)[0].astype(np.int64)


# --------------------------------------------------------------------------- TF kernels (App. B)
def nms_tf(boxes, scores, max_output_size, iou_threshold, return_examined=False):
    """`tf.image.non_max_suppression` (TF r1.13 CPU NonMaxSuppressionV3), SURVEY App. B.1.
    Call site: model/region_proposal.py:74-76.  Greedy, descending score (ties -> lower index),
    corners min/max-normalised, area WITHOUT +1, non-positive-area boxes neither suppress nor are
    suppressed, `iou = inter / (a_i + a_j - inter)` (fp32), strict `>`."""
    b = np.asarray(boxes, F)
    s = np.asarray(scores, F)
    thr = F(iou_threshold)
    lo0 = np.minimum(b[:, 0], b[:, 2]); hi0 = np.maximum(b[:, 0], b[:, 2])
    lo1 = np.minimum(b[:, 1], b[:, 3]); hi1 = np.maximum(b[:, 1], b[:, 3])
    area = (hi0 - lo0) * (hi1 - lo1)
    order = np.argsort(-s, kind='stable')
    sel = np.empty(min(int(max_output_size), b.shape[0]), dtype=np.int64)
    n = 0
    examined = 0
    for c in order:
        if n >= max_output_size:
            break
        examined += 1
        if n and area[c] > 0:
            k = sel[:n]
            i0 = np.maximum(F(0), np.minimum(hi0[c], hi0[k]) - np.maximum(lo0[c], lo0[k]))
            i1 = np.maximum(F(0), np.minimum(hi1[c], hi1[k]) - np.maximum(lo1[c], lo1[k]))
            inter = i0 * i1
            with np.errstate(divide='ignore', invalid='ignore'):
                iou = inter / (area[c] + area[k] - inter)
            if np.any((area[k] > 0) & (iou > thr)):
                continue
        sel[n] = c
        n += 1
    out = sel[:n].astype(np.int32)
    return (out, examined) if return_examined else out


def crop_and_resize_tf(image, boxes, box_ind, crop_h, crop_w, extrapolation_value=0.0):
    """`tf.image.crop_and_resize` bilinear (TF r1.13 CPU kernel), SURVEY App. B.2.
    image [b,h,w,c] NHWC, boxes [r,4]=(y1,x1,y2,x2) normalised, box_ind [r]."""
    img = np.asarray(image, F)
    bx = np.asarray(boxes, F)
    bi = np.asarray(box_ind).astype(np.int64)
    _, h, w, c = img.shape
    r = bx.shape[0]
    out = np.full((r, crop_h, crop_w, c), F(extrapolation_value), dtype=F)
    if r == 0:
        return out
    y1, x1, y2, x2 = bx[:, 0], bx[:, 1], bx[:, 2], bx[:, 3]
    hs = (y2 - y1) * F(h - 1) / F(crop_h - 1) if crop_h > 1 else np.zeros(r, F)
    ws = (x2 - x1) * F(w - 1) / F(crop_w - 1) if crop_w > 1 else np.zeros(r, F)
    for y in range(crop_h):
        in_y = (y1 * F(h - 1) + F(y) * hs) if crop_h > 1 else F(0.5) * (y1 + y2) * F(h - 1)
        in_y = in_y.astype(F)
        oky = ~((in_y < 0) | (in_y > F(h - 1)))
        top = np.floor(in_y); bot = np.ceil(in_y); ly = (in_y - top).astype(F)
        for x in range(crop_w):
            in_x = (x1 * F(w - 1) + F(x) * ws) if crop_w > 1 else F(0.5) * (x1 + x2) * F(w - 1)
            in_x = in_x.astype(F)
            ok = oky & ~((in_x < 0) | (in_x > F(w - 1)))
            k = np.nonzero(ok)[0]
            if k.size == 0:
                continue
            left = np.floor(in_x[k]); right = np.ceil(in_x[k]); lx = (in_x[k] - left).astype(F)[:, None]
            t_, b_ = top[k].astype(np.int64), bot[k].astype(np.int64)
            l_, r_ = left.astype(np.int64), right.astype(np.int64)
            n = bi[k]
            tl = img[n, t_, l_]; tr = img[n, t_, r_]; bl = img[n, b_, l_]; br = img[n, b_, r_]
            tp = tl + (tr - tl) * lx
            bt = bl + (br - bl) * lx
            out[k, y, x] = tp + (bt - tp) * ly[k, None]
    return out


def max_pool_2x2(x):
    """Keras `MaxPooling2D(padding='same')` defaults 2x2/stride 2 (model/roi_pooling.py:13,51); App. B.3."""
    x = np.asarray(x, F)
    n, h, w, c = x.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = np.empty((n, oh, ow, c), F)
    for i in range(oh):
        for j in range(ow):
            out[:, i, j] = x[:, 2 * i:2 * i + 2, 2 * j:2 * j + 2].reshape(n, -1, c).max(axis=1)
    return out


def avg_pool_2x2(x):
    """`tf.nn.avg_pool(ret,[1,2,2,1],[1,2,2,1],'SAME')` (model/roi_pooling.py:154).  Even inputs:
    sum the 4 window elements in row-major order then divide by 4 (Eigen mean reducer)."""
    x = np.asarray(x, F)
    n, h, w, c = x.shape
    assert h % 2 == 0 and w % 2 == 0
    s = ((x[:, 0::2, 0::2] + x[:, 0::2, 1::2]) + x[:, 1::2, 0::2]) + x[:, 1::2, 1::2]
    return (s / F(4)).astype(F)


# --------------------------------------------------------------------------- a3 region proposal
def region_proposal(deltas, anchors, scores, image_shape, post_nms, iou_threshold=0.7,
                    means=(0, 0, 0, 0), stds=(1, 1, 1, 1), pre_nms_top_k=0, min_size=0.0,
                    return_stats=False):
    """model/region_proposal.py:37-81.  Reference behaviour = pre_nms_top_k 0 (the top-k block at
    :65-69 is commented out) and min_size <= 0 (min_edge=None at :63).  The two extra knobs follow
    py-faster-rcnn order: min-size filter -> top-k by score (ties: lower index) -> NMS.
    Returns (rois [K,4] in selection order, idx [K] int32 into the anchor set)."""
    boxes = decode_bbox(anchors, deltas, means, stds)                                    # :59-60
    boxes, _ = bboxes_clip_filter(boxes, 0, image_shape[0], image_shape[1])              # :63
    s = np.asarray(scores, F)
    cand = np.arange(boxes.shape[0], dtype=np.int64)
    if min_size > 0:
        _, cand = bboxes_clip_filter(boxes, 0, image_shape[0], image_shape[1], min_edge=min_size)
    if pre_nms_top_k and pre_nms_top_k > 0 and cand.size > pre_nms_top_k:
        cand = cand[np.argsort(-s[cand], kind='stable')[:pre_nms_top_k]]
        cand = np.sort(cand)  # keep ascending-index order so NMS tie-breaking stays "lower index"
    sel, examined = nms_tf(boxes[cand], s[cand], post_nms, iou_threshold, return_examined=True)   # :74-76
    idx = cand[sel].astype(np.int32)
    if return_stats:
        return boxes[idx], idx, {'examined': examined}
    return boxes[idx], idx                                                               # :81


# --------------------------------------------------------------------------- a4 / a5 / a8 RoI pooling
def roi_pool_c4(feat, rois, stride, pool_size=7, max_pooling_flag=True, box_ind=None):
    """model/roi_pooling.py:53-90 `RoiPoolingCropAndResize.call` (C4 models)."""
    feat = np.asarray(feat, F)
    r = np.asarray(rois, F) / F(stride)                                                   # :64
    h, w = feat.shape[1:3]
    bi = np.zeros(r.shape[0], np.int32) if box_ind is None else box_ind                   # :66
    nb = np.stack([r[:, 1] / F(h - 1), r[:, 0] / F(w - 1),
                   r[:, 3] / F(h - 1), r[:, 2] / F(w - 1)], axis=1)                      # :69-74
    if max_pooling_flag:
        return max_pool_2x2(crop_and_resize_tf(feat, nb, bi, 2 * pool_size, 2 * pool_size))   # :75-84
    return crop_and_resize_tf(feat, nb, bi, pool_size, pool_size)                         # :85-90


def roi_pool_fpn(feat, rois, image_shape, pool_size=7, box_ind=None):
    """model/roi_pooling.py:15-42 `RoiPoolingCropAndResize2.call` (FPN): boxes normalised by IMAGE H, W."""
    feat = np.asarray(feat, F)
    r = np.asarray(rois, F)
    H, W = F(image_shape[0]), F(image_shape[1])                                           # :26
    bi = np.zeros(r.shape[0], np.int32) if box_ind is None else box_ind                   # :28
    nb = np.stack([r[:, 1] / H, r[:, 0] / W, r[:, 3] / H, r[:, 2] / W], axis=1)          # :30-35
    return max_pool_2x2(crop_and_resize_tf(feat, nb, bi, 2 * pool_size, 2 * pool_size))   # :36-42


def roi_align_pad(feat, rois, stride, pool_size=7, box_ind=None):
    """model/roi_pooling.py:93-176 `RoiPoolingRoiAlign` -> `roi_align` -> `crop_and_resize(pad_border=True)`
    (dormant in the reference; tensorpack-style RoIAlign)."""
    feat = np.asarray(feat, F)
    b = np.asarray(rois, F) / F(stride)                                                   # :175
    img = np.pad(feat, [[0, 0], [1, 1], [1, 1], [0, 0]], mode='symmetric')                # :100
    b = b + F(1)                                                                          # :101
    q = 2 * pool_size                                                                     # :153
    x0, y0, x1, y1 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    sw = (x1 - x0) / F(q)                                                                 # :120
    sh = (y1 - y0) / F(q)                                                                 # :121
    ih, iw = F(img.shape[1] - 1), F(img.shape[2] - 1)                                     # :123
    nx0 = (x0 + sw / F(2) - F(0.5)) / iw                                                  # :124
    ny0 = (y0 + sh / F(2) - F(0.5)) / ih                                                  # :125
    nw = sw * F(q - 1) / iw                                                               # :127
    nh = sh * F(q - 1) / ih                                                               # :128
    nb = np.stack([ny0, nx0, ny0 + nh, nx0 + nw], axis=1).astype(F)                       # :130
    bi = np.zeros(b.shape[0], np.int32) if box_ind is None else box_ind
    return avg_pool_2x2(crop_and_resize_tf(img, nb, bi, q, q))                            # :134-136,154


# --------------------------------------------------------------------------- a6 / a7 FPN routing
def assign_levels(rois, min_level=2, max_level=5):
    """model/fpn/base_fpn_model.py:303-324 `_assign_levels`.
    Returns (levels [R] int32, rois_list, order [R] int64 level-major/stable)."""
    r = np.asarray(rois, F)
    h = np.maximum(F(0), r[:, 3] - r[:, 1])                                               # :307
    w = np.maximum(F(0), r[:, 2] - r[:, 0])                                               # :308
    with np.errstate(divide='ignore'):
        lv = np.floor(F(4) + np.log(np.sqrt(w * h + F(1e-8)) / F(224.0)) / np.log(F(2.)))  # :309
    lv = np.minimum(np.maximum(lv, F(min_level)), F(max_level)).astype(F)                 # :312-313
    rois_list, idx_list = [], []
    for i in range(min_level, max_level + 1):                                             # :318-322
        k = np.nonzero(lv == F(i))[0].astype(np.int64)
        rois_list.append(r[k]); idx_list.append(k)
    return lv.astype(np.int32), rois_list, np.concatenate(idx_list)                       # :324


def level_margin(rois):
    """distance of the un-floored level value to the nearest integer (pre-screen helper)."""
    r = np.asarray(rois, F)
    h = np.maximum(F(0), r[:, 3] - r[:, 1]); w = np.maximum(F(0), r[:, 2] - r[:, 0])
    with np.errstate(divide='ignore'):
        v = (F(4) + np.log(np.sqrt(w * h + F(1e-8)) / F(224.0)) / np.log(F(2.))).astype(np.float64)
    return np.abs(v - np.round(v))


def fpn_roi_features(rois_list, p_list, image_shape, pool_size=7):
    """model/fpn/base_fpn_model.py:152-161 `_get_roi_features`: per non-empty level, concat axis 0."""
    outs = [roi_pool_fpn(p, r, image_shape, pool_size) for r, p in zip(rois_list, p_list) if r.shape[0] > 0]
    return np.concatenate(outs, axis=0)


# --------------------------------------------------------------------------- a9 IoU
def pairwise_iou(b1, b2):
    """utils/bbox_tf.py:37-56 (+area :7-16, pairwise_intersection :19-34): "+1" convention,
    iou = 0 where inter == 0 else inter / (a1 + a2 - inter)."""
    a = np.asarray(b1, F); b = np.asarray(b2, F)
    ih = np.maximum(F(0), np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]) + F(1))
    iw = np.maximum(F(0), np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]) + F(1))
    inter = ih * iw
    a1 = (a[:, 3] - a[:, 1] + F(1)) * (a[:, 2] - a[:, 0] + F(1))
    a2 = (b[:, 3] - b[:, 1] + F(1)) * (b[:, 2] - b[:, 0] + F(1))
    union = a1[:, None] + a2[None, :] - inter
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(inter == 0, F(0), inter / union).astype(F)


# --------------------------------------------------------------------------- sampling convention
def shuffle_by_perm(idx, prio):
    """The injected replacement for the reference's unseeded `tf.random_shuffle`:
    shuffle(idx) := idx sorted ascending by prio (prio = perm[...] of each element, all distinct)."""
    idx = np.asarray(idx)
    return idx[np.argsort(np.asarray(prio), kind='stable')]


# --------------------------------------------------------------------------- a10 anchor target
def anchor_target(gt_bboxes, image_shape, all_anchors, perm, pos_iou_threshold=0.7, neg_iou_threshold=0.3,
                  total_num_samples=256, max_pos_samples=128, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """model/anchor_target.py:29-107 (+_unmap :110-125).  `perm` [N_all] int: sampling priority of each
    ORIGINAL anchor index (lower = earlier in the shuffled order)."""
    anc_all = np.asarray(all_anchors, F)
    gt = np.asarray(gt_bboxes, F)
    n_all = anc_all.shape[0]
    inside = bboxes_range_filter(anc_all, image_shape[0], image_shape[1])                 # :54
    anc = anc_all[inside]                                                                 # :55
    labels = -np.ones(anc.shape[0], np.int32)                                             # :59
    ov = pairwise_iou(anc, gt)                                                            # :60
    argmax_r = np.argmax(ov, axis=1).astype(np.int32)                                     # :61
    max_r = np.max(ov, axis=1)                                                            # :62
    max_c = np.max(ov, axis=0)                                                            # :63
    gt_arg = np.argwhere(ov == max_c)[:, 0]                                               # :64
    labels[max_r < F(neg_iou_threshold)] = 0                                              # :67
    labels[gt_arg] = 1                                                                    # :68
    labels[max_r >= F(pos_iou_threshold)] = 1                                             # :69
    fg = np.nonzero(labels == 1)[0]                                                       # :72
    if fg.size > max_pos_samples:                                                         # :73
        fg = shuffle_by_perm(fg, np.asarray(perm)[inside[fg]])                            # :74
        labels[fg[max_pos_samples:]] = -1                                                 # :75-77
        fg = fg[:max_pos_samples]
    num_bg = total_num_samples - int(np.sum(labels == 1))                                 # :78
    bg = np.nonzero(labels == 0)[0]                                                       # :79
    if bg.size > num_bg:                                                                  # :80
        bg = shuffle_by_perm(bg, np.asarray(perm)[inside[bg]])                            # :81
        labels[bg[num_bg:]] = -1                                                          # :82-84
        bg = bg[:num_bg]
    targets = encode_bbox(anc, gt[argmax_r], means, stds)                                 # :88-90
    in_w = np.zeros((anc.shape[0], 4), F); in_w[labels == 1] = 1                          # :93-95
    out_w = np.zeros((anc.shape[0], 4), F)
    num_examples = F(np.sum(labels >= 0))                                                 # :99
    out_w[labels >= 0] = F(1.0) / num_examples                                            # :100-101

    def unmap(data, fill):                                                                # :110-125
        ret = np.full((n_all,) + data.shape[1:], F(fill), F)
        ret[inside] = data.astype(F)
        return ret
    return (unmap(labels, -1), unmap(targets, 0), unmap(in_w, 0), unmap(out_w, 0),
            {'inside': inside, 'fg': fg, 'bg': bg})


# --------------------------------------------------------------------------- a11 proposal target
def proposal_target(rois, gt_bboxes, gt_labels, perm, num_classes=21, pos_iou_threshold=0.5,
                    neg_iou_threshold=0.5, total_num_samples=128, max_pos_samples=32,
                    means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """model/proposal_target.py:32-124.  `perm` [K] int: sampling priority per roi index.  Background
    padding (`np.random.choice(replace=True)`, :77) := cycle through shuffle(bg).
    Keeps the reference's `labels[idx]` quirk (:99,117: label of roi #idx, not of fg_inds[idx])."""
    rois = np.asarray(rois, F); gt = np.asarray(gt_bboxes, F); gl = np.asarray(gt_labels)
    perm = np.asarray(perm)
    iou = pairwise_iou(rois, gt)                                                          # :56
    max_r = np.max(iou, axis=1)                                                           # :57
    assign = np.argmax(iou, axis=1).astype(np.int64)                                      # :58
    labels = gl[assign]                                                                   # :59
    fg = np.nonzero(max_r >= F(pos_iou_threshold))[0]                                     # :62
    bg = np.nonzero((max_r < F(pos_iou_threshold)) & (max_r >= F(neg_iou_threshold)))[0]  # :63-64
    if fg.size > max_pos_samples:                                                         # :67
        fg = shuffle_by_perm(fg, perm[fg])[:max_pos_samples]                              # :68
    want = total_num_samples - fg.size
    if bg.size > want:                                                                    # :69
        bg = shuffle_by_perm(bg, perm[bg])[:want]                                         # :71
    elif bg.size < want:                                                                  # :74-77
        if bg.size == 0:
            raise ValueError("'a' cannot be empty unless no samples are taken")           # np.random.choice
        sb = shuffle_by_perm(bg, perm[bg])
        bg = sb[np.arange(want) % sb.size]
    keep = np.concatenate([fg, bg]).astype(np.int64)                                      # :81
    final_rois = rois[keep]                                                               # :82
    final_labels = labels[keep].copy()                                                    # :83
    final_labels[fg.size:] = 0                                                            # :85-86
    s = keep.size
    in_w = np.zeros((s, num_classes, 4), F)                                               # :89
    tg = np.zeros((s, num_classes, 4), F)                                                 # :103
    if fg.size > 0:
        bt = encode_bbox(final_rois[:fg.size], gt[assign[fg]], means, stds)               # :105-108
        for i in range(fg.size):
            in_w[i, labels[i]] = 1                                                        # :98-99  (labels[idx] quirk)
            tg[i, labels[i]] = bt[i]                                                      # :116-117
    return (final_rois, final_labels, tg.reshape(s, num_classes * 4), in_w.reshape(s, num_classes * 4),
            np.ones((s, num_classes * 4), F), {'keep': keep, 'num_fg': fg.size})         # :122-124


# --------------------------------------------------------------------------- f1 post-head detection filtering
def post_ops_prediction(roi_scores_softmax, roi_txtytwth, rois, image_shape, means=(0, 0, 0, 0), stds=(1, 1, 1, 1),
                        max_num_per_class=50, max_num_per_image=150, nms_iou_threshold=0.3, score_threshold=0.05,
                        extractor_stride=16, num_classes=21):
    """model/prediction.py:103-163 `post_ops_prediction`.  Per foreground class: score threshold
    (strict >, :135) -> decode with the roi-head stds (:137-139) -> clip + min-edge filter with min_edge =
    extractor_stride (:140-142) -> NMS (:145) -> concat (:154-156) -> top max_num_per_image by score (:158-159;
    `tf.nn.top_k(sorted=False)` order is unspecified in TF — fixed here as descending score, ties to the lower
    concatenated index).  Returns (boxes [n,4], classes [n] int32, scores [n]) or (None, None, None)."""
    s = np.asarray(roi_scores_softmax, F); d = np.asarray(roi_txtytwth, F).reshape(s.shape[0], s.shape[1], 4)
    rois = np.asarray(rois, F)
    res_s, res_b, res_c = [], [], []
    for i in range(1, num_classes):
        inds = np.nonzero(s[:, i] > F(score_threshold))[0]
        cls_score = s[inds, i]
        boxes = decode_bbox(rois[inds], d[inds, i, :], means, stds)
        boxes, sel = bboxes_clip_filter(boxes, 0, image_shape[0], image_shape[1], min_edge=extractor_stride)
        cls_score = cls_score[sel]
        keep = nms_tf(boxes, cls_score, max_num_per_class, nms_iou_threshold)
        if keep.size == 0:
            continue
        res_s.append(cls_score[keep]); res_b.append(boxes[keep]); res_c.append(np.full(keep.size, i, np.int32))
    if not res_s:
        return None, None, None
    sc = np.concatenate(res_s); bb = np.concatenate(res_b); cc = np.concatenate(res_c)
    order = np.argsort(-sc, kind='stable')[:min(max_num_per_image, sc.size)]
    return bb[order], cc[order], sc[order]


def eval_loop_detections(scores, roi_txtytwth, rois_net, img_scale, raw_h, raw_w, means=(0, 0, 0, 0), stds=(0.1, 0.1, 0.2, 0.2),
                         score_threshold=0.05, iou_threshold=0.3, max_objects_per_class=50, max_objects_per_image=50,
                         min_size=10, loop='voc'):
    """The per-image body of the reference's evaluation loops.  loop='voc': evaluation/pascal_eval_files_utils.py:76-106
    (per class: score > thr :82, decode :84-86, clip to the RAW image + min_size filter :87, NMS :89-90; per image:
    `image_thresh = np.sort(scores)[-max]`, keep `score >= image_thresh` :98-106 — ties survive).  loop='coco':
    scripts/eval_coco.py:127-153 (same per class; `tf.nn.top_k` :148-150).  `rois_net / img_scale` is im_detect's last
    statement (faster_rcnn/base_faster_rcnn_model.py:304).  Returns per-class lists [(boxes [k,4], scores [k])] indexed by
    class (entry 0 unused), restricted by the per-image cut."""
    s = np.asarray(scores, F)
    num_classes = s.shape[1]
    d = np.asarray(roi_txtytwth, F).reshape(s.shape[0], num_classes, 4)
    rois = (np.asarray(rois_net, F) / F(img_scale)).astype(F)                              # base_faster_rcnn_model.py:304
    raw_h, raw_w = F(raw_h), F(raw_w)                                                      # tf.to_float, :77-78
    per_class = [None] * num_classes
    for j in range(1, num_classes):
        inds = np.nonzero(s[:, j] > F(score_threshold))[0]                                 # :82
        cls_scores = s[inds, j]
        boxes = decode_bbox(rois[inds], d[inds, j, :], means, stds)                        # :84-86
        boxes, sel = bboxes_clip_filter(boxes, 0, raw_h, raw_w, min_edge=min_size)         # :87
        cls_scores = cls_scores[sel]
        keep = nms_tf(boxes, cls_scores, max_objects_per_class, iou_threshold)             # :89-90
        per_class[j] = (boxes[keep], cls_scores[keep])
    if max_objects_per_image > 0:
        all_scores = np.concatenate([per_class[j][1] for j in range(1, num_classes)])      # :99-100
        if all_scores.size > max_objects_per_image:
            if loop == 'voc':
                thresh = np.sort(all_scores)[-max_objects_per_image]                       # :102
                for j in range(1, num_classes):
                    k = np.nonzero(per_class[j][1] >= thresh)[0]                           # :104-105
                    per_class[j] = (per_class[j][0][k], per_class[j][1][k])
            else:                                                                          # eval_coco.py:148-150
                cls_of = np.concatenate([np.full(per_class[j][1].size, j) for j in range(1, num_classes)])
                order = np.argsort(-all_scores, kind='stable')[:max_objects_per_image]
                chosen = np.zeros(all_scores.size, bool); chosen[order] = True
                pos = 0
                for j in range(1, num_classes):
                    n = per_class[j][1].size
                    k = np.nonzero(chosen[pos:pos + n])[0]
                    per_class[j] = (per_class[j][0][k], per_class[j][1][k])
                    pos += n
                del cls_of
    return per_class


# --------------------------------------------------------------------------- f3 RoI-pooling backward (w.r.t. features)
def crop_and_resize_grad_image(image_shape, boxes, box_ind, grad_crops):
    """TF r1.13 `CropAndResizeGradImage` (the gradient TF applies for tf.image.crop_and_resize w.r.t. `image`,
    recalled from core/kernels/crop_and_resize_op.cc like App. B.2): every in-range sample scatters
    dtop = (1-ly) g, dbottom = ly g, then (1-lx) / lx of each to its left / right tap."""
    b, h, w, c = image_shape
    g = np.asarray(grad_crops, F)
    bx = np.asarray(boxes, F); bi = np.asarray(box_ind).astype(np.int64)
    r, ch, cw, _ = g.shape
    out = np.zeros((b, h, w, c), F)
    if r == 0:
        return out
    y1, x1, y2, x2 = bx[:, 0], bx[:, 1], bx[:, 2], bx[:, 3]
    hs = (y2 - y1) * F(h - 1) / F(ch - 1) if ch > 1 else np.zeros(r, F)
    ws = (x2 - x1) * F(w - 1) / F(cw - 1) if cw > 1 else np.zeros(r, F)
    for y in range(ch):
        in_y = ((y1 * F(h - 1) + F(y) * hs) if ch > 1 else F(0.5) * (y1 + y2) * F(h - 1)).astype(F)
        oky = ~((in_y < 0) | (in_y > F(h - 1)))
        top = np.floor(in_y); bot = np.ceil(in_y); ly = (in_y - top).astype(F)
        for x in range(cw):
            in_x = ((x1 * F(w - 1) + F(x) * ws) if cw > 1 else F(0.5) * (x1 + x2) * F(w - 1)).astype(F)
            ok = oky & ~((in_x < 0) | (in_x > F(w - 1)))
            k = np.nonzero(ok)[0]
            if k.size == 0:
                continue
            left = np.floor(in_x[k]); right = np.ceil(in_x[k]); lx = (in_x[k] - left).astype(F)[:, None]
            t_, b_ = top[k].astype(np.int64), bot[k].astype(np.int64)
            l_, r_ = left.astype(np.int64), right.astype(np.int64)
            gk = g[k, y, x]
            dtop = (F(1) - ly[k, None]) * gk
            dbot = ly[k, None] * gk
            np.add.at(out, (bi[k], t_, l_), (F(1) - lx) * dtop)
            np.add.at(out, (bi[k], t_, r_), lx * dtop)
            np.add.at(out, (bi[k], b_, l_), (F(1) - lx) * dbot)
            np.add.at(out, (bi[k], b_, r_), lx * dbot)
    return out


def roi_pool_c4_grad(feat, rois, stride, grad_out, pool_size=7, max_pooling_flag=True, box_ind=None):
    """Gradient of roi_pool_c4 w.r.t. feat: MaxPooling2D gradient (all of it to the first maximal element of each 2x2
    window, row-major — TF MaxPoolGrad) followed by crop_and_resize_grad_image."""
    feat = np.asarray(feat, F); g = np.asarray(grad_out, F)
    r = np.asarray(rois, F) / F(stride)
    h, w = feat.shape[1:3]
    bi = np.zeros(r.shape[0], np.int32) if box_ind is None else box_ind
    nb = np.stack([r[:, 1] / F(h - 1), r[:, 0] / F(w - 1), r[:, 3] / F(h - 1), r[:, 2] / F(w - 1)], axis=1)
    if not max_pooling_flag:
        return crop_and_resize_grad_image(feat.shape, nb, bi, g)
    q = 2 * pool_size
    crops = crop_and_resize_tf(feat, nb, bi, q, q)
    return crop_and_resize_grad_image(feat.shape, nb, bi, _max_pool_grad_2x2(crops, g, pool_size))


def _max_pool_grad_2x2(crops, g, pool_size):
    """TF MaxPoolGrad for the 2x2 / stride 2 window: all of it to the first maximal element (row-major)."""
    gc = np.zeros_like(crops)
    n, _, _, c = crops.shape
    for i in range(pool_size):
        for j in range(pool_size):
            win = crops[:, 2 * i:2 * i + 2, 2 * j:2 * j + 2].reshape(n, 4, c)
            arg = np.argmax(win, axis=1)
            for s in range(4):
                gc[:, 2 * i + s // 2, 2 * j + s % 2] = np.where(arg == s, g[:, i, j], F(0))
    return gc


def roi_pool_fpn_grad(feat, rois, image_shape, grad_out, pool_size=7, box_ind=None):
    """Gradient of roi_pool_fpn (model/roi_pooling.py:15-42: image-normalised boxes, 14x14 crop, 2x2 max pool) w.r.t.
    feat — what scripts/train.py:99-103 back-propagates through the FPN extractor."""
    feat = np.asarray(feat, F); g = np.asarray(grad_out, F)
    r = np.asarray(rois, F)
    H, W = F(image_shape[0]), F(image_shape[1])
    bi = np.zeros(r.shape[0], np.int32) if box_ind is None else box_ind
    nb = np.stack([r[:, 1] / H, r[:, 0] / W, r[:, 3] / H, r[:, 2] / W], axis=1)
    q = 2 * pool_size
    crops = crop_and_resize_tf(feat, nb, bi, q, q)
    return crop_and_resize_grad_image(feat.shape, nb, bi, _max_pool_grad_2x2(crops, g, pool_size))


def roi_align_pad_grad(feat, rois, stride, grad_out, pool_size=7, box_ind=None):
    """Gradient of roi_align_pad (model/roi_pooling.py:93-176) w.r.t. feat: avg_pool gradient (a quarter to each sample),
    CropAndResizeGradImage on the padded map, then the gradient of tf.pad(SYMMETRIC, 1): each border ring element adds
    to the element it mirrors (rows first, then columns, the reverse of the padding order)."""
    feat = np.asarray(feat, F); g = np.asarray(grad_out, F)
    b = np.asarray(rois, F) / F(stride)
    b = b + F(1)
    q = 2 * pool_size
    x0, y0, x1, y1 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    sw = (x1 - x0) / F(q); sh = (y1 - y0) / F(q)
    ph, pw = feat.shape[1] + 2, feat.shape[2] + 2
    ih, iw = F(ph - 1), F(pw - 1)
    nx0 = (x0 + sw / F(2) - F(0.5)) / iw; ny0 = (y0 + sh / F(2) - F(0.5)) / ih
    nb = np.stack([ny0, nx0, ny0 + sh * F(q - 1) / ih, nx0 + sw * F(q - 1) / iw], axis=1).astype(F)
    bi = np.zeros(b.shape[0], np.int32) if box_ind is None else box_ind
    gq = (g / F(4))
    gc = np.repeat(np.repeat(gq, 2, axis=1), 2, axis=2)
    gp = crop_and_resize_grad_image((feat.shape[0], ph, pw, feat.shape[3]), nb, bi, gc)
    gp[:, :, 1] += gp[:, :, 0]; gp[:, :, -2] += gp[:, :, -1]
    gp = gp[:, :, 1:-1]
    gp[:, 1] += gp[:, 0]; gp[:, -2] += gp[:, -1]
    return np.ascontiguousarray(gp[:, 1:-1])


# --------------------------------------------------------------------------- f3 losses (model/losses.py)
def smooth_l1_loss(pred, target, in_w, out_w, sigma=1.0, dim=(1,)):
    """model/losses.py:16-28.  dim=[1]: mean over rows of the row sums; dim=[0,1]: the total sum."""
    s2 = F(sigma) * F(sigma)
    d = np.asarray(in_w, F) * (np.asarray(pred, F) - np.asarray(target, F))
    a = np.abs(d)
    sign = (a < F(1.0) / s2).astype(F)
    per = (d * d) * (s2 / F(2)) * sign + (a - F(0.5) / s2) * (F(1) - sign)
    per = np.asarray(out_w, F) * per
    return F(np.mean(np.sum(per, axis=tuple(dim), dtype=F), dtype=F))


def smooth_l1_loss_grad(pred, target, in_w, out_w, sigma=1.0, dim=(1,)):
    """d loss / d pred (sign is under stop_gradient, losses.py:21)."""
    s2 = F(sigma) * F(sigma)
    in_w = np.asarray(in_w, F)
    d = in_w * (np.asarray(pred, F) - np.asarray(target, F))
    a = np.abs(d)
    sign = a < F(1.0) / s2
    dd = np.where(sign, d * s2, np.sign(d)).astype(F)
    denom = F(1) if tuple(dim) == (0, 1) else F(np.asarray(pred).shape[0])
    return (np.asarray(out_w, F) * dd * in_w / denom).astype(F)


def cls_loss(logits, labels, weight=1.0):
    """model/losses.py:4-13 (tf.losses.sparse_softmax_cross_entropy, SUM_BY_NONZERO_WEIGHTS) with the caller's
    `labels >= 0` gather (base_faster_rcnn_model.py:204-206) folded in: rows with a negative label are skipped."""
    x = np.asarray(logits, F); lab = np.asarray(labels)
    sel = np.nonzero(lab >= 0)[0]
    if sel.size == 0 or weight == 0:
        return F(0)
    x = x[sel]; li = lab[sel].astype(np.int64)
    z = x - x.max(axis=1, keepdims=True)
    per = np.log(np.exp(z).sum(axis=1, dtype=F)) - z[np.arange(li.size), li]
    return F(np.sum(per.astype(F) * F(weight), dtype=F) / F(sel.size))


def cls_loss_grad(logits, labels, weight=1.0):
    x = np.asarray(logits, F); lab = np.asarray(labels)
    g = np.zeros_like(x)
    sel = np.nonzero(lab >= 0)[0]
    if sel.size == 0 or weight == 0:
        return g
    z = x[sel] - x[sel].max(axis=1, keepdims=True)
    e = np.exp(z)
    p = e / e.sum(axis=1, keepdims=True, dtype=F)
    p[np.arange(sel.size), lab[sel].astype(np.int64)] -= F(1)
    g[sel] = p * (F(weight) / F(sel.size))
    return g


# --------------------------------------------------------------------------- f2 RPN score layout
def rpn_fg_scores(logits, layout, anchors_per_cell=1):
    """Foreground probability per anchor from the raw RPN logits.
    'caffe': faster_rcnn/base_faster_rcnn_model.py:149-152 — rows [cells, 2A] = [bg x A | fg x A], the four
    reshape/transpose steps reduce to softmax over (row[k], row[A+k]) read out cell-major, anchor-minor.
    'pairs': fpn/base_fpn_model.py:223 — softmax(all_fpn_scores)[:, 1]."""
    x = np.asarray(logits, F)
    if layout == 'caffe':
        a = anchors_per_cell
        x = x.reshape(-1, 2, a)
        pair = np.stack([x[:, 0, :], x[:, 1, :]], axis=-1).reshape(-1, 2)
    else:
        pair = x.reshape(-1, 2)
    z = pair - pair.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e[:, 1] / (e[:, 0] + e[:, 1])).astype(F)
