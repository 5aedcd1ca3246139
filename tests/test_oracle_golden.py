"""The standalone numpy oracle (oracle/boxpath_oracle.py) against the committed golden vectors that
`oracle/make_golden.py` produced by running the reference's own files on the numpy TF shim."""
import hashlib

import numpy as np
import pytest

from oracle import boxpath_oracle as orc
from tf_eager_object_detection_b200 import synthetic as syn


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.fixture(scope='module')
def c4():
    return syn.c4_image(1, 0, channels=8)


@pytest.fixture(scope='module')
def fpn():
    return syn.fpn_image(3, 0, channels=8)


def test_anchor_generators_match_reference(golden):
    assert np.array_equal(syn.anchor_base(16).astype(np.float32), golden['anchor_base'])
    assert np.array_equal(sha(syn.c4_anchors(38, 63, 16)), golden['c4_anchors_sha'])
    assert np.array_equal(sha(syn.fpn_anchors((600, 1000))), golden['fpn_anchors_sha'])
    assert syn.c4_anchors(38, 63).shape[0] == 21546
    assert syn.fpn_anchors((600, 1000)).shape[0] == 150111
    assert syn.fpn_anchors((800, 1333)).shape[0] == 267069
    a = syn.c4_anchors(38, 63)
    assert orc.bboxes_range_filter(a, 600, 1000).shape[0] == int(golden['c4_inside_count']) == 8151


def test_decode_clip(golden, c4):
    dec = orc.decode_bbox(c4['anchors'], c4['deltas'])
    dec, idx = orc.bboxes_clip_filter(dec, 0, 600, 1000)
    assert np.array_equal(dec[:2048], golden['c4_decoded_clipped_head'])
    assert np.array_equal(sha(dec), golden['c4_decoded_clipped_sha'])
    assert idx.dtype == np.int32 and np.array_equal(idx, np.arange(21546))


@pytest.mark.parametrize('mode,post', [('eval', 300), ('train', 2000)])
def test_c4_region_proposal(golden, c4, mode, post):
    rois, idx = orc.region_proposal(c4['deltas'], c4['anchors'], c4['scores'], c4['image_shape'], post)
    assert np.array_equal(idx, golden['c4_%s_idx' % mode])
    assert np.array_equal(rois, golden['c4_%s_rois' % mode])


def test_pre_nms_top_k_is_a_noop_when_quota_fills(golden, c4):
    """with unique scores, top-k(6000) + NMS == NMS over all anchors if 300 are reached."""
    rois, idx, st = orc.region_proposal(c4['deltas'], c4['anchors'], c4['scores'], c4['image_shape'], 300,
                                        pre_nms_top_k=6000, return_stats=True)
    assert st['examined'] <= 6000
    assert np.array_equal(idx, golden['c4_eval_idx'])


def test_c4_roi_pooling(golden, c4):
    rois = golden['c4_eval_rois'][:64]
    feat = c4['feat'][None]
    assert np.array_equal(orc.roi_pool_c4(feat, rois, 16, 7, False), golden['c4_pool_nomax'])
    assert np.array_equal(orc.roi_pool_c4(feat, rois, 16, 7, True), golden['c4_pool_max'])
    assert np.array_equal(orc.roi_align_pad(feat, rois, 16, 7), golden['c4_roialign'])


def test_fpn_proposals_levels_features(golden, fpn):
    rois, idx = orc.region_proposal(fpn['deltas'], fpn['anchors'], fpn['scores'], fpn['image_shape'], 1000)
    assert np.array_equal(idx, golden['fpn_eval_idx'])
    assert np.array_equal(rois, golden['fpn_eval_rois'])
    lv, rois_list, order = orc.assign_levels(rois)
    assert np.array_equal([r.shape[0] for r in rois_list], golden['fpn_level_counts'])
    assert np.array_equal(order, golden['fpn_level_order'])
    feats = [f[None] for f in fpn['feats']]
    out = orc.fpn_roi_features(rois_list, feats, fpn['image_shape'])
    assert out.shape == (1000, 7, 7, 8)
    assert np.array_equal(out[:128], golden['fpn_roi_features_head'])


def test_fpn_random_rois_cover_levels_and_borders(golden, fpn):
    rr = syn.random_rois(np.random.default_rng(syn.seed_for(3, 50)), 256, (600, 1000))
    lv, rois_list, order = orc.assign_levels(rr)
    assert np.array_equal([r.shape[0] for r in rois_list], golden['rand_level_counts'])
    assert (golden['rand_level_counts'] > 0).all()
    assert np.array_equal(order, golden['rand_level_order'])
    out = orc.fpn_roi_features(rois_list, [f[None] for f in fpn['feats']], fpn['image_shape'])
    assert np.array_equal(out, golden['rand_roi_features'])


def _targets_inputs():
    rng = np.random.default_rng(syn.seed_for(4, 0))
    gt, gl = syn.gt_boxes(rng, 100, (600, 1000))
    anchors = syn.c4_anchors(38, 63)
    perm = rng.permutation(anchors.shape[0])
    return rng, gt, gl, anchors, perm


def test_pairwise_iou(golden):
    _, gt, _, anchors, _ = _targets_inputs()
    assert np.array_equal(orc.pairwise_iou(anchors[:4096], gt), golden['iou_anchors4096_gt100'])
    got = orc.pairwise_iou(np.float32([[0, 0, 9, 9], [5, 5, 14, 14]]), np.float32([[0, 0, 9, 9]]))
    assert got[0, 0] == 1.0 and got[1, 0] == np.float32(25.0) / np.float32(175.0)


@pytest.mark.parametrize('name,m', [('at', 100), ('at3', 3)])
def test_anchor_target(golden, name, m):
    _, gt, _, anchors, perm = _targets_inputs()
    lab, tg, iw, ow, info = orc.anchor_target(gt[:m], [600, 1000], anchors, perm)
    assert np.array_equal(lab, golden[name + '_labels'])
    assert np.array_equal(iw, golden[name + '_in_w'])
    assert np.array_equal(ow, golden[name + '_out_w'])
    assert np.array_equal(tg, golden[name + '_targets'])
    assert (lab == 1).sum() <= 128 and (lab >= 0).sum() <= 256


@pytest.mark.parametrize('name,kw', [
    ('pt', dict(total_num_samples=128, max_pos_samples=32, neg_iou_threshold=0.0)),
    ('pt_fpn', dict(total_num_samples=256, max_pos_samples=64, neg_iou_threshold=0.0)),
    ('pt_pad', dict(total_num_samples=2048, max_pos_samples=512, neg_iou_threshold=0.1)),
])
def test_proposal_target(golden, name, kw):
    rng, gt, gl, anchors, _ = _targets_inputs()
    rois = golden['c4_train_rois']
    perm_r = rng.permutation(rois.shape[0])
    out = orc.proposal_target(rois, gt, gl, perm_r, num_classes=21, pos_iou_threshold=0.5,
                              stds=(0.1, 0.1, 0.2, 0.2), **kw)
    for k, v in zip(('rois', 'labels', 'targets', 'in_w', 'out_w'), out[:5]):
        assert np.array_equal(v, golden['%s_%s' % (name, k)]), k
    if name == 'pt_pad':   # the np.random.choice(replace=True) branch really ran
        assert len(np.unique(out[5]['keep'])) < 2048


def test_post_ops_prediction(golden):
    """model/prediction.py:103-163 on synthetic roi-head outputs over the 300 eval rois."""
    hs, hd = syn.roi_head_outputs(np.random.default_rng(syn.seed_for(1, 77)), 300, 21)
    b, c, s = orc.post_ops_prediction(hs, hd, golden['c4_eval_rois'], (600, 1000), stds=(0.1, 0.1, 0.2, 0.2))
    assert np.array_equal(b, golden['post_boxes']) and np.array_equal(c, golden['post_classes'])
    assert np.array_equal(s, golden['post_scores']) and b.shape == (150, 4)
    assert (np.diff(s) <= 0).all()
    assert orc.post_ops_prediction(hs, hd, golden['c4_eval_rois'], (600, 1000), score_threshold=2.0) == (None, None, None)


# ------------------------------------------------------------------------------------------------ f3 losses + RoI backward
def test_losses_match_reference_on_shim(golden):
    g = golden
    assert np.isclose(orc.smooth_l1_loss(g['loss_rpn_pred'], g['at_targets'], g['at_in_w'], g['at_out_w'], 3.0, (0, 1)),
                      g['loss_rpn_reg'], rtol=1e-5)
    assert np.isclose(orc.smooth_l1_loss(g['loss_roi_pred'], g['pt_targets'], g['pt_in_w'], g['pt_out_w'], 1.0, (1,)),
                      g['loss_roi_reg'], rtol=1e-5)
    assert np.isclose(orc.cls_loss(g['loss_rpn_logits'], g['at_labels']), g['loss_rpn_cls'], rtol=1e-5)
    assert np.isclose(orc.cls_loss(g['loss_roi_logits'], g['pt_labels']), g['loss_roi_cls'], rtol=1e-5)


def test_loss_and_roi_gradients_against_torch_autograd(golden):
    """Independent witness for the oracle's hand-written gradients: torch autograd through a torch restatement."""
    import torch
    import torch.nn.functional as tnf
    g = golden
    p = torch.tensor(g['loss_roi_pred'], requires_grad=True)
    d = torch.tensor(g['pt_in_w']) * (p - torch.tensor(g['pt_targets']))
    per = torch.where(d.abs() < 1.0, 0.5 * d * d, d.abs() - 0.5) * torch.tensor(g['pt_out_w'])
    per.sum(1).mean().backward()
    np.testing.assert_allclose(orc.smooth_l1_loss_grad(g['loss_roi_pred'], g['pt_targets'], g['pt_in_w'], g['pt_out_w']),
                               p.grad.numpy(), rtol=1e-5, atol=1e-8)
    x = torch.tensor(g['loss_rpn_logits'], requires_grad=True)
    lab = torch.tensor(g['at_labels']).long()
    tnf.cross_entropy(x, lab, ignore_index=-1).backward()
    np.testing.assert_allclose(orc.cls_loss_grad(g['loss_rpn_logits'], g['at_labels']), x.grad.numpy(), rtol=1e-4,
                               atol=1e-9)
    assert np.isclose(orc.cls_loss(g['loss_rpn_logits'], g['at_labels']),
                      float(tnf.cross_entropy(x.detach(), lab, ignore_index=-1)), rtol=1e-5)
    # RoI-pooling backward: autograd through a dense bilinear-weight restatement of the C4 extractor
    rng = np.random.default_rng(5)
    feat = rng.standard_normal((1, 9, 11, 4), dtype=np.float32)
    rois = syn.random_rois(rng, 6, (144, 176))
    go = rng.standard_normal((6, 3, 3, 4), dtype=np.float32)
    for max_flag in (False, True):
        ft = torch.tensor(feat, dtype=torch.float64, requires_grad=True)
        q = 6 if max_flag else 3
        r = rois.astype(np.float64) / 16.0
        crops = []
        for k in range(6):
            ys = r[k, 1] + np.arange(q) * ((r[k, 3] - r[k, 1]) / (q - 1))
            xs = r[k, 0] + np.arange(q) * ((r[k, 2] - r[k, 0]) / (q - 1))
            wy = torch.zeros(q, 9, dtype=torch.float64); wx = torch.zeros(q, 11, dtype=torch.float64)
            for i, y in enumerate(ys):
                if 0 <= y <= 8:
                    t = int(np.floor(y)); wy[i, t] += 1 - (y - t); wy[i, int(np.ceil(y))] += y - t
            for i, xx in enumerate(xs):
                if 0 <= xx <= 10:
                    t = int(np.floor(xx)); wx[i, t] += 1 - (xx - t); wx[i, int(np.ceil(xx))] += xx - t
            crops.append(torch.einsum('yh,xw,hwc->yxc', wy, wx, ft[0]))
        out = torch.stack(crops)
        if max_flag:
            out = tnf.max_pool2d(out.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
        out.backward(torch.tensor(go, dtype=torch.float64))
        got = orc.roi_pool_c4_grad(feat, rois, 16, go, 3, max_flag)
        np.testing.assert_allclose(got, ft.grad.numpy(), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------ f2 RPN score layout
def test_rpn_score_layouts_match_reference_statements(golden):
    g = golden
    assert np.array_equal(orc.rpn_fg_scores(g['rpn_caffe_logits'], 'caffe', 9), g['rpn_caffe_scores'])
    assert np.array_equal(orc.rpn_fg_scores(g['rpn_pairs_logits'], 'pairs'), g['rpn_pairs_scores'])


def test_anchor_generator_mirror_host_tables(golden):
    from tf_eager_object_detection_b200 import anchor_generator as ag
    assert np.array_equal(ag.generate_anchor_base(16).astype(np.float32), golden['anchor_base'])
    off = ag._offsets(32, (1.0,), (0.5, 1.0, 2.0))
    lvl = syn.fpn_level_anchors(32, 1, 1, 4)                    # one cell at the origin = the offsets themselves
    assert np.array_equal(off, lvl)


def _dense_crop(ft, ys, xs, h, w):
    """Differentiable crop_and_resize of one box by dense bilinear weight matrices (float64): ft [h,w,c] torch tensor,
    ys / xs = sample coordinates in pixels; samples outside [0,h-1] x [0,w-1] contribute 0."""
    import torch
    wy = torch.zeros(len(ys), h, dtype=torch.float64); wx = torch.zeros(len(xs), w, dtype=torch.float64)
    for i, y in enumerate(ys):
        if 0 <= y <= h - 1:
            t = int(np.floor(y)); wy[i, t] += 1 - (y - t); wy[i, int(np.ceil(y))] += y - t
    for i, x in enumerate(xs):
        if 0 <= x <= w - 1:
            t = int(np.floor(x)); wx[i, t] += 1 - (x - t); wx[i, int(np.ceil(x))] += x - t
    return torch.einsum('yh,xw,hwc->yxc', wy, wx, ft)


def test_fpn_and_roialign_gradients_against_torch_autograd():
    """The two other extractor gradients (FPN: image-normalised boxes + 2x2 max; dormant RoIAlign: symmetric pad + 2x2
    mean) against torch autograd through a dense restatement — the oracle functions the GPU backward tests compare with."""
    import torch
    import torch.nn.functional as tnf
    rng = np.random.default_rng(6)
    fh, fw, c, P = 10, 13, 3, 3
    feat = rng.standard_normal((1, fh, fw, c), dtype=np.float32)
    H, W = 160, 208
    rois = syn.random_rois(rng, 7, (H, W))
    go = rng.standard_normal((7, P, P, c), dtype=np.float32)
    q = 2 * P
    # FPN extractor
    ft = torch.tensor(feat, dtype=torch.float64, requires_grad=True)
    crops = []
    for k in range(7):
        r = rois[k].astype(np.float32)
        y1, y2 = np.float64(r[1] / np.float32(H)), np.float64(r[3] / np.float32(H))
        x1, x2 = np.float64(r[0] / np.float32(W)), np.float64(r[2] / np.float32(W))
        ys = y1 * (fh - 1) + np.arange(q) * ((y2 - y1) * (fh - 1) / (q - 1))
        xs = x1 * (fw - 1) + np.arange(q) * ((x2 - x1) * (fw - 1) / (q - 1))
        crops.append(_dense_crop(ft[0], ys, xs, fh, fw))
    out = tnf.max_pool2d(torch.stack(crops).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    np.testing.assert_allclose(orc.roi_pool_fpn(feat, rois, (H, W), P), out.detach().numpy(), rtol=1e-4, atol=1e-5)
    out.backward(torch.tensor(go, dtype=torch.float64))
    np.testing.assert_allclose(orc.roi_pool_fpn_grad(feat, rois, (H, W), go, P), ft.grad.numpy(), rtol=1e-4, atol=1e-5)
    # RoIAlign variant: pad (replicate == SYMMETRIC for a 1-pixel border), tensorpack sample centres, mean pool
    ft = torch.tensor(feat, dtype=torch.float64, requires_grad=True)
    padded = tnf.pad(ft[0].permute(2, 0, 1)[None], (1, 1, 1, 1), mode='replicate')[0].permute(1, 2, 0)
    crops = []
    for k in range(7):
        b = rois[k].astype(np.float64) / 16.0 + 1.0
        sw, sh = (b[2] - b[0]) / q, (b[3] - b[1]) / q
        xs = b[0] + sw / 2 - 0.5 + np.arange(q) * sw
        ys = b[1] + sh / 2 - 0.5 + np.arange(q) * sh
        crops.append(_dense_crop(padded, ys, xs, fh + 2, fw + 2))
    out = tnf.avg_pool2d(torch.stack(crops).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    np.testing.assert_allclose(orc.roi_align_pad(feat, rois, 16, P), out.detach().numpy(), rtol=1e-4, atol=1e-5)
    out.backward(torch.tensor(go, dtype=torch.float64))
    np.testing.assert_allclose(orc.roi_align_pad_grad(feat, rois, 16, go, P), ft.grad.numpy(), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------ f1 evaluation loops
def _voc_lines_from_per_class(names, per_image):
    """The reference's file format (pascal_eval_files_utils.py:109-122) from the oracle's per-class results."""
    out = []
    for j in range(1, 21):
        for name, pc in zip(names, per_image):
            boxes, scores = pc[j]
            for k in range(boxes.shape[0]):
                out.append('{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(name, scores[k], boxes[k, 0] + 1, boxes[k, 1] + 1,
                                                                          boxes[k, 2] + 1, boxes[k, 3] + 1))
    return ''.join(out)


def test_eval_loop_oracle_reproduces_the_reference_result_files(golden):
    """oracle.eval_loop_detections against the text of the VOC result files the reference's own get_prediction_files wrote
    (run unmodified on the shim by make_golden.py): per-image raw-size clip, rois / img_scale, the `>= image_thresh` cut
    with 26 tied detections surviving on image 3, and the no-cut variant."""
    from oracle.voc_fixture import eval_loop_inputs
    imgs = eval_loop_inputs()
    names = ['%06d' % (i + 1) for i in range(len(imgs))]
    for tag, max_img in (('eval_voc', 50), ('eval_voc_nocut', 0)):
        per_image = [orc.eval_loop_detections(im['scores'], im['deltas'], im['rois'], im['scale'], im['raw_h'], im['raw_w'],
                                              max_objects_per_image=max_img, loop='voc') for im in imgs]
        assert _voc_lines_from_per_class(names, per_image) == bytes(golden[tag + '_files']).decode()
    counts = [sum(pc[j][1].size for j in range(1, 21)) for pc in per_image]
    assert counts[0] > 50 and counts[2] > 50                                   # the no-cut run keeps everything
    cut = [orc.eval_loop_detections(im['scores'], im['deltas'], im['rois'], im['scale'], im['raw_h'], im['raw_w'], loop='voc')
           for im in imgs]
    assert [sum(pc[j][1].size for j in range(1, 21)) for pc in cut] == [50, 50, 76]
    coco = [orc.eval_loop_detections(im['scores'], im['deltas'], im['rois'], im['scale'], im['raw_h'], im['raw_w'], loop='coco')
            for im in imgs]
    assert [sum(pc[j][1].size for j in range(1, 21)) for pc in coco] == [50, 50, 50]
