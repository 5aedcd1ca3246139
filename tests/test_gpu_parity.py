"""GPU parity: the CUDA path (through the Python mirror -> ctypes -> C ABI) against the numpy oracle and the committed
golden vectors.  Bit-exact for indices / labels / IEEE-only arithmetic; <= 1e-5 relative where exp/log differ by ulps.
Nothing here reads /root/reference."""
import hashlib
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import boxpath_oracle as orc  # noqa: E402
from tf_eager_object_detection_b200 import synthetic as syn  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # BASELINE.json north_star: "RoI features and decoded boxes within 1e-5 relative (fp32)"


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def close(a, b, scale=1.0):
    """|a-b| <= 1e-5 * max(|a|,|b|, scale): relative, with `scale` as the magnitude floor of the tensor's own units."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    tol = RTOL * np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    bad = np.abs(a - b) > tol
    assert not bad.any(), 'max err %g at %s' % (np.abs(a - b).max(), np.argwhere(bad)[:3].tolist())


def box_scale(anchors, deltas, means=(0, 0, 0, 0), stds=(1, 1, 1, 1)):
    """Per-box magnitude of the operands of the final `cx -/+ 0.5 w` (the UNCLIPPED corners): the quantity a 1e-5
    relative fp32 error is relative to — clipped corners near 0 are differences of numbers this large."""
    return np.abs(orc.decode_bbox(anchors, deltas, means, stds)).max(axis=1, keepdims=True).astype(np.float64)


@pytest.fixture(scope='module')
def bx():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import tf_eager_object_detection_b200.ops as ops
    return ops


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope='module')
def c4():
    return syn.c4_image(1, 0, channels=8)


@pytest.fixture(scope='module')
def fpn():
    return syn.fpn_image(3, 0, channels=8)


# ------------------------------------------------------------------------------------------------ a1 / a2 / a12
def test_decode_clip(bx, golden, c4):
    out = bx.decode_clip(cu(c4['anchors']), cu(c4['deltas']), image_shape=(600, 1000)).cpu().numpy()
    ref, _ = orc.bboxes_clip_filter(orc.decode_bbox(c4['anchors'], c4['deltas']), 0, 600, 1000)
    sc = box_scale(c4['anchors'], c4['deltas'])
    close(out, ref, scale=sc)
    close(out[:2048], golden['c4_decoded_clipped_head'], scale=sc[:2048])
    # known answer: zero delta on the first base anchor
    one = bx.decode_clip(cu(np.float32([[-84, -40, 99, 55]])), cu(np.zeros((1, 4), np.float32)), image_shape=(600, 1000))
    assert one.cpu().numpy().tolist() == [[0.0, 0.0, 100.0, 56.0]]
    raw = bx.decode_clip(cu(np.float32([[-84, -40, 99, 55]])), cu(np.zeros((1, 4), np.float32)))
    assert raw.cpu().numpy().tolist() == [[-84.0, -40.0, 100.0, 56.0]]


def test_decode_with_stds_and_encode_roundtrip(bx):
    rng = np.random.default_rng(5)
    rois = syn.random_rois(rng, 1000, (600, 1000))
    deltas = (rng.normal(0, 1, (1000, 4))).astype(np.float32)
    means, stds = (0.0, 0.0, 0.0, 0.0), (0.1, 0.1, 0.2, 0.2)
    dec = bx.decode_clip(cu(rois), cu(deltas), means, stds).cpu().numpy()
    close(dec, orc.decode_bbox(rois, deltas, means, stds), scale=box_scale(rois, deltas, means, stds))
    gt, _ = syn.gt_boxes(rng, 1000, (600, 1000))
    enc = bx.encode(cu(rois), cu(gt), means, stds).cpu().numpy()
    close(enc, orc.encode_bbox(rois, gt, means, stds), scale=1.0)


def test_clip_and_range_filters(bx, c4):
    from tf_eager_object_detection_b200 import bbox_tf
    dec = orc.decode_bbox(c4['anchors'], c4['deltas'])
    b, idx = bbox_tf.bboxes_clip_filter(cu(dec), 0, 600, 1000, min_edge=16)
    rb, ridx = orc.bboxes_clip_filter(dec, 0, 600, 1000, min_edge=16)
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(b.cpu().numpy(), rb)
    b2, idx2 = bbox_tf.bboxes_clip_filter(cu(dec), 0, 600, 1000)
    assert np.array_equal(b2.cpu().numpy(), orc.bboxes_clip_filter(dec, 0, 600, 1000)[0])
    assert idx2.dtype == torch.int32 and np.array_equal(idx2.cpu().numpy(), np.arange(dec.shape[0]))
    inside = bbox_tf.bboxes_range_filter(cu(c4['anchors']), 600, 1000)
    assert inside.shape[0] == 8151 and np.array_equal(inside.cpu().numpy(), orc.bboxes_range_filter(c4['anchors'], 600, 1000))


# ------------------------------------------------------------------------------------------------ NMS (stage-wise)
def test_nms_on_oracle_boxes_is_bit_exact(bx, golden, c4):
    """SURVEY §7 mitigation (i): fed the oracle's decoded boxes, the kept indices must match without pre-screening."""
    dec, _ = orc.bboxes_clip_filter(orc.decode_bbox(c4['anchors'], c4['deltas']), 0, 600, 1000)
    for post, key in ((300, 'c4_eval_idx'), (2000, 'c4_train_idx')):
        idx, cnt = bx.nms(cu(dec), cu(c4['scores']), post, 0.7)
        assert int(cnt) == post
        assert np.array_equal(idx.cpu().numpy(), golden[key])


@pytest.mark.parametrize('n,max_out,thr', [(1, 5, 0.5), (37, 10, 0.3), (64, 64, 0.7), (65, 100, 0.5), (500, 500, 0.0),
                                           (2049, 2048, 0.9), (5000, 300, 0.5), (30000, 2000, 0.6)])
def test_nms_random_shapes(bx, n, max_out, thr):
    rng = np.random.default_rng(n)
    b = syn.random_rois(rng, n, (600, 1000))
    s = ((rng.permutation(n) + 1) / (n + 1)).astype(np.float32)
    idx, cnt = bx.nms(cu(b), cu(s), max_out, thr)
    ref = orc.nms_tf(b, s, max_out, thr)
    assert int(cnt) == ref.size
    assert np.array_equal(idx.cpu().numpy()[:ref.size], ref)
    assert (idx.cpu().numpy()[ref.size:] == -1).all()


def test_nms_ties_degenerate_boxes_and_batch(bx):
    rng = np.random.default_rng(11)
    n = 3000
    b = syn.random_rois(rng, n, (600, 1000))
    b[::7, 2] = b[::7, 0]                       # zero-area boxes: never suppressed, never suppress
    b[1::50] = b[1::50][:, [2, 3, 0, 1]]        # flipped corners: TF normalises with min/max
    s = rng.integers(0, 40, n).astype(np.float32) / 40.0   # heavy score ties -> lower index first
    s[5] = -0.0; s[6] = 0.0
    batch_b = np.stack([b, b[::-1].copy(), np.roll(b, 17, axis=0)])
    batch_s = np.stack([s, s[::-1].copy(), np.roll(s, 17)])
    idx, cnt = bx.nms(cu(batch_b), cu(batch_s), 400, 0.5)
    for i in range(3):
        ref = orc.nms_tf(batch_b[i], batch_s[i], 400, 0.5)
        assert int(cnt[i]) == ref.size
        assert np.array_equal(idx[i].cpu().numpy()[:ref.size], ref)
    # all scores equal: more ties than one selection chunk holds (exercises the index refinement of the select)
    s_eq = np.full(n, 0.25, np.float32)
    idx, cnt = bx.nms(cu(b), cu(s_eq), 2048, 0.3)
    ref = orc.nms_tf(b, s_eq, 2048, 0.3)
    assert int(cnt) == ref.size and np.array_equal(idx.cpu().numpy()[:ref.size], ref)


def test_nms_argument_errors(bx):
    b = cu(np.zeros((4, 4), np.float32)); s = cu(np.zeros(4, np.float32))
    with pytest.raises(ValueError):
        bx.nms(b, s, 10, 1.5)            # TF: iou_threshold must be in [0, 1]
    with pytest.raises(NotImplementedError):
        bx.nms(b, s, 5000, 0.5)          # documented limit: max_output_size <= 2048
    from tf_eager_object_detection_b200._tensor import FLOAT32, INT32, Borrow
    for mode in ('0', '1'):              # torch metadata borrow, then the DLPack capsule path (bx_dlpack_data)
        os.environ['BX_FORCE_DLPACK'] = mode
        try:
            with pytest.raises(TypeError):       # wrong dtype is rejected, never silently converted
                Borrow(0).ptr(b.to(torch.float64), FLOAT32, (4, 4))
            with pytest.raises(TypeError):       # wrong shape
                Borrow(0).ptr(b, FLOAT32, (4, 5))
            with pytest.raises(TypeError):       # host memory is not accepted: there is no CPU path
                Borrow(0).ptr(torch.zeros(4, 4), FLOAT32, (4, 4))
            with pytest.raises(TypeError):       # misaligned box tensor
                Borrow(0).ptr(cu(np.zeros(17, np.float32))[1:].reshape(4, 4), FLOAT32, (4, 4), 16)
            si = s.to(torch.int32)
            assert Borrow(0).ptr(si, INT32, (-1,)) == si.data_ptr()
            assert Borrow(0).ptr(b, FLOAT32, (-1, 4), 16) == b.data_ptr()
        finally:
            os.environ.pop('BX_FORCE_DLPACK', None)


def test_dlpack_capsule_path_end_to_end(bx, golden, c4, monkeypatch):
    """Every tensor of a whole call travels as a DLPack capsule (the protocol a non-torch host uses)."""
    monkeypatch.setenv('BX_FORCE_DLPACK', '1')
    ob, oi, oc = bx.proposals(cu(c4['anchors']), cu(c4['deltas'])[None], cu(c4['scores'])[None], (600, 1000), 300)
    assert np.array_equal(oi[0].cpu().numpy(), golden['c4_eval_idx'])


# ------------------------------------------------------------------------------------------------ a3 proposals
def _margin(dec, scores, idx, thr):
    """min |IoU - thr| over the pairs the greedy sweep compared decisively (kept x kept): pre-screen witness."""
    k = dec[idx]
    area = (k[:, 2] - k[:, 0]) * (k[:, 3] - k[:, 1])
    iw = np.maximum(0, np.minimum(k[:, None, 2], k[None, :, 2]) - np.maximum(k[:, None, 0], k[None, :, 0]))
    ih = np.maximum(0, np.minimum(k[:, None, 3], k[None, :, 3]) - np.maximum(k[:, None, 1], k[None, :, 1]))
    inter = iw * ih
    union = area[:, None] + area[None, :] - inter
    iou = np.where(union > 0, inter / np.where(union > 0, union, 1), 0)   # zero-area pairs are never compared by TF
    np.fill_diagonal(iou, 0)
    return np.abs(iou - thr).min()


@pytest.mark.parametrize('mode,post', [('eval', 300), ('train', 2000)])
def test_c4_region_proposal(bx, golden, c4, mode, post):
    from tf_eager_object_detection_b200.region_proposal import RegionProposal
    rp = RegionProposal()
    rois = rp((cu(c4['deltas']), cu(c4['anchors']), cu(c4['scores']), c4['image_shape']), training=(mode == 'train'))
    assert rois.shape == (post, 4)
    sc = box_scale(c4['anchors'], c4['deltas'])
    close(rois.cpu().numpy(), golden['c4_%s_rois' % mode], scale=sc[golden['c4_%s_idx' % mode]])
    ob, oi, oc = rp.call_batched((cu(c4['deltas'])[None], cu(c4['anchors']), cu(c4['scores'])[None], c4['image_shape']),
                                 training=(mode == 'train'))
    assert int(oc[0]) == post
    assert np.array_equal(oi[0].cpu().numpy(), golden['c4_%s_idx' % mode])     # kept indices bit-exact
    dec, _ = orc.bboxes_clip_filter(orc.decode_bbox(c4['anchors'], c4['deltas']), 0, 600, 1000)
    assert _margin(dec, c4['scores'], golden['c4_%s_idx' % mode], 0.7) > 1e-5  # seed is decisively pre-screened


def test_pre_nms_top_k_and_min_size(bx, golden, c4):
    ob, oi, oc = bx.proposals(cu(c4['anchors']), cu(c4['deltas'])[None], cu(c4['scores'])[None], (600, 1000), 300,
                              pre_nms_top_k=6000)
    assert np.array_equal(oi[0].cpu().numpy(), golden['c4_eval_idx'])
    # a top-k small enough to bite, and a min-size filter, against the oracle
    for kw in (dict(pre_nms_top_k=200), dict(min_size=64.0), dict(pre_nms_top_k=3000, min_size=100.0)):
        ob, oi, oc = bx.proposals(cu(c4['anchors']), cu(c4['deltas'])[None], cu(c4['scores'])[None], (600, 1000), 300, **kw)
        rois, idx = orc.region_proposal(c4['deltas'], c4['anchors'], c4['scores'], (600, 1000), 300, **kw)
        assert int(oc[0]) == idx.size
        assert np.array_equal(oi[0].cpu().numpy()[:idx.size], idx)
        close(ob[0].cpu().numpy()[:idx.size], rois, scale=box_scale(c4['anchors'], c4['deltas'])[idx])
        assert (oi[0].cpu().numpy()[idx.size:] == -1).all() and (ob[0].cpu().numpy()[idx.size:] == 0).all()


def test_fpn_global_proposals(bx, golden, fpn):
    """150 111 anchors (P2-P6), one global NMS, post 1000 — keys stream from global memory (no smem cache)."""
    ob, oi, oc = bx.proposals(cu(fpn['anchors']), cu(fpn['deltas'])[None], cu(fpn['scores'])[None], (600, 1000), 1000)
    assert int(oc[0]) == 1000
    assert np.array_equal(oi[0].cpu().numpy(), golden['fpn_eval_idx'])
    close(ob[0].cpu().numpy(), golden['fpn_eval_rois'], scale=box_scale(fpn['anchors'], fpn['deltas'])[golden['fpn_eval_idx']])


def test_proposals_batch_of_images(bx):
    imgs = [syn.c4_image(2, i, with_features=False) for i in range(4)]
    deltas = np.stack([im['deltas'] for im in imgs]); scores = np.stack([im['scores'] for im in imgs])
    ob, oi, oc = bx.proposals(cu(imgs[0]['anchors']), cu(deltas), cu(scores), (600, 1000), 300, pre_nms_top_k=6000)
    for i, im in enumerate(imgs):
        rois, idx = orc.region_proposal(im['deltas'], im['anchors'], im['scores'], (600, 1000), 300, pre_nms_top_k=6000)
        assert int(oc[i]) == 300 and np.array_equal(oi[i].cpu().numpy(), idx)
        close(ob[i].cpu().numpy(), rois, scale=box_scale(im['anchors'], im['deltas'])[idx])


def test_proposals_quota_not_filled(bx):
    """few anchors, strong overlap: NMS exhausts every candidate before post_nms (multi-round path, padded output)."""
    rng = np.random.default_rng(3)
    anchors = syn.c4_anchors(6, 8, 16)
    n = anchors.shape[0]
    deltas, scores = syn.rpn_outputs(rng, n)
    ob, oi, oc = bx.proposals(cu(anchors), cu(deltas)[None], cu(scores)[None], (96, 128), 300, iou_threshold=0.3)
    rois, idx = orc.region_proposal(deltas, anchors, scores, (96, 128), 300, iou_threshold=0.3)
    assert idx.size < 300 and int(oc[0]) == idx.size
    assert np.array_equal(oi[0].cpu().numpy()[:idx.size], idx)


# ------------------------------------------------------------------------------------------------ a4 / a5 / a8 RoI pooling
def test_c4_roi_pooling_golden(bx, golden, c4):
    from tf_eager_object_detection_b200.roi_pooling import RoiPoolingCropAndResize, RoiPoolingRoiAlign
    rois = golden['c4_eval_rois'][:64]
    feat = cu(c4['feat'][None])
    out = RoiPoolingCropAndResize(7, False)((feat, cu(rois), 16)).cpu().numpy()
    assert np.array_equal(out, golden['c4_pool_nomax'])          # IEEE-only arithmetic, same op order: bit-exact
    out = RoiPoolingCropAndResize(7, True)((feat, cu(rois), 16)).cpu().numpy()
    assert np.array_equal(out, golden['c4_pool_max'])
    out = RoiPoolingRoiAlign(7)((feat, cu(rois), 16)).cpu().numpy()
    close(out, golden['c4_roialign'], scale=1.0)


def test_roi_pool_scalar_channel_path_and_box_ind(bx):
    rng = np.random.default_rng(8)
    feat = rng.standard_normal((3, 20, 30, 6), dtype=np.float32)       # C=6: not a multiple of 4 -> scalar kernel
    rois = syn.random_rois(rng, 50, (320, 480))
    bi = rng.integers(0, 3, 50).astype(np.int32)
    from tf_eager_object_detection_b200 import _lib
    out = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_MAX2, 7, cu(feat), cu(rois), stride=16.0, box_ind=cu(bi)).cpu().numpy()
    assert np.array_equal(out, orc.roi_pool_c4(feat, rois, 16, 7, True, box_ind=bi))
    out = bx.roi_pool(_lib.ROI_IMAGE_NORM, _lib.POOL_MAX2, 7, cu(feat), cu(rois), image_shape=(320, 480), box_ind=cu(bi)).cpu().numpy()
    assert np.array_equal(out, orc.roi_pool_fpn(feat, rois, (320, 480), 7, box_ind=bi))


def test_crop_and_resize_raw_op(bx):
    rng = np.random.default_rng(9)
    img = rng.standard_normal((2, 11, 13, 4), dtype=np.float32)
    boxes = np.float32([[0, 0, 1, 1], [0.1, 0.2, 1.3, 0.9], [-0.2, 0.0, 0.5, 0.5], [0.6, 0.6, 0.2, 0.1]])
    bi = np.int32([0, 1, 1, 0])
    from tf_eager_object_detection_b200.roi_pooling import crop_and_resize
    out = crop_and_resize(cu(img), cu(boxes), cu(bi), [11, 11]).cpu().numpy()
    ref = orc.crop_and_resize_tf(img, boxes, bi, 11, 11)
    assert np.array_equal(out, ref)
    # box [0,0,1,1] with crop == (square) feature size returns the map; y2n > 1 -> trailing rows exactly 0
    sq = rng.standard_normal((1, 9, 9, 4), dtype=np.float32)
    out = crop_and_resize(cu(sq), cu(np.float32([[0, 0, 1, 1], [0, 0, 1.5, 1]])), cu(np.int32([0, 0])), [9, 9]).cpu().numpy()
    assert np.array_equal(out[0], sq[0])
    assert (out[1][6:] == 0).all() and (out[1][:6] != 0).any()
    with pytest.raises(ValueError):
        crop_and_resize(cu(sq), cu(np.float32([[0, 0, 1, 1]])), cu(np.int32([0])), [0, 9])


def test_fpn_levels_and_features(bx, golden, fpn):
    from tf_eager_object_detection_b200 import fpn as fpn_mod
    feats = [cu(f[None]) for f in fpn['feats']]
    rois = golden['fpn_eval_rois']
    assert orc.level_margin(rois).min() > 1e-5
    rois_list, order = fpn_mod.assign_levels(cu(rois))
    assert [r.shape[0] for r in rois_list] == golden['fpn_level_counts'].tolist()
    assert order.dtype == torch.int64 and np.array_equal(order.cpu().numpy(), golden['fpn_level_order'])
    out = fpn_mod.get_roi_features(rois_list, feats, fpn['image_shape'])
    assert np.array_equal(out[:128].cpu().numpy(), golden['fpn_roi_features_head'])
    fused, order2, lv, counts = fpn_mod.fpn_roi_features(cu(rois), feats, fpn['image_shape'])
    assert torch.equal(fused, out) and np.array_equal(order2.cpu().numpy(), golden['fpn_level_order'])
    assert counts.cpu().numpy().tolist() == golden['fpn_level_counts'].tolist()
    # random rois: every level populated, borders extrapolated
    rr = syn.random_rois(np.random.default_rng(syn.seed_for(3, 50)), 256, (600, 1000))
    fused, order, lv, counts = fpn_mod.fpn_roi_features(cu(rr), feats, fpn['image_shape'])
    assert np.array_equal(order.cpu().numpy(), golden['rand_level_order'])
    assert np.array_equal(fused.cpu().numpy(), golden['rand_roi_features'])
    # known answers, SURVEY A.8
    sides = np.float32([300, 150, 500, 60, 20, 900])
    b = np.stack([np.zeros(6, np.float32), np.zeros(6, np.float32), sides, sides], 1)
    lv, _, _ = bx.fpn_assign_levels(cu(b))
    assert lv.cpu().numpy().tolist() == [4, 3, 5, 2, 2, 5]


# ------------------------------------------------------------------------------------------------ a9 - a11 targets
def _targets_inputs():
    rng = np.random.default_rng(syn.seed_for(4, 0))
    gt, gl = syn.gt_boxes(rng, 100, (600, 1000))
    anchors = syn.c4_anchors(38, 63)
    perm = rng.permutation(anchors.shape[0])
    return rng, gt, gl, anchors, perm


def test_pairwise_iou(bx, golden):
    from tf_eager_object_detection_b200.bbox_tf import pairwise_iou
    _, gt, _, anchors, _ = _targets_inputs()
    out = pairwise_iou(cu(anchors[:4096]), cu(gt)).cpu().numpy()
    assert np.array_equal(out, golden['iou_anchors4096_gt100'])
    full = pairwise_iou(cu(anchors), cu(gt)).cpu().numpy()           # BASELINE cfg 4: 21 546 x 100
    assert np.array_equal(full, orc.pairwise_iou(anchors, gt))
    ka = pairwise_iou(cu(np.float32([[0, 0, 9, 9], [5, 5, 14, 14]])), cu(np.float32([[0, 0, 9, 9]]))).cpu().numpy()
    assert ka[0, 0] == 1.0 and ka[1, 0] == np.float32(25.0) / np.float32(175.0)


@pytest.mark.parametrize('name,m', [('at', 100), ('at3', 3)])
def test_anchor_target(bx, golden, name, m):
    from tf_eager_object_detection_b200.anchor_target import AnchorTarget
    _, gt, _, anchors, perm = _targets_inputs()
    lab, tg, iw, ow = AnchorTarget()((cu(gt[:m]), [600, 1000], cu(anchors)), perm=perm.astype(np.int32))
    assert np.array_equal(lab.cpu().numpy(), golden[name + '_labels'])        # sampled indices bit-exact
    assert np.array_equal(iw.cpu().numpy(), golden[name + '_in_w'])
    assert np.array_equal(ow.cpu().numpy(), golden[name + '_out_w'])
    close(tg.cpu().numpy(), golden[name + '_targets'], scale=1.0)


def test_anchor_target_batched_and_all_positive_quirk(bx):
    from tf_eager_object_detection_b200.anchor_target import AnchorTarget
    rng = np.random.default_rng(21)
    anchors = syn.c4_anchors(38, 63)
    gts, perms = [], []
    for i in range(3):
        g, _ = syn.gt_boxes(rng, 20, (600, 1000))
        gts.append(g); perms.append(rng.permutation(anchors.shape[0]).astype(np.int32))
    # image 2: one gt box no inside anchor overlaps -> column max 0 -> every inside anchor becomes fg (py-faster-rcnn quirk)
    gts[2][0] = np.float32([0, 0, 1, 1])
    lab, tg, iw, ow, cnt = AnchorTarget().call_batched((cu(np.stack(gts)), [600, 1000], cu(anchors)), perm=np.stack(perms))
    for i in range(3):
        rl, rt, ri, ro, info = orc.anchor_target(gts[i], [600, 1000], anchors, perms[i])
        assert np.array_equal(lab[i].cpu().numpy(), rl), i
        assert np.array_equal(iw[i].cpu().numpy(), ri) and np.array_equal(ow[i].cpu().numpy(), ro)
        close(tg[i].cpu().numpy(), rt, scale=1.0)
        assert cnt[i].cpu().numpy().tolist() == [int((rl == 1).sum()), int((rl == 0).sum())]


@pytest.mark.parametrize('name,kw', [
    ('pt', dict(total_num_samples=128, max_pos_samples=32, neg_iou_threshold=0.0)),
    ('pt_fpn', dict(total_num_samples=256, max_pos_samples=64, neg_iou_threshold=0.0)),
    ('pt_pad', dict(total_num_samples=2048, max_pos_samples=512, neg_iou_threshold=0.1)),
])
def test_proposal_target(bx, golden, name, kw):
    from tf_eager_object_detection_b200.proposal_target import ProposalTarget
    rng, gt, gl, _, _ = _targets_inputs()
    rois = golden['c4_train_rois']
    perm_r = rng.permutation(rois.shape[0]).astype(np.int32)
    pt = ProposalTarget(num_classes=21, pos_iou_threshold=0.5, target_stds=[0.1, 0.1, 0.2, 0.2], **kw)
    out = pt((cu(rois), cu(gt), cu(gl)), perm=perm_r)
    names = ('rois', 'labels', 'targets', 'in_w', 'out_w')
    for k, v in zip(names, out):
        g = golden['%s_%s' % (name, k)]
        if k == 'targets':
            close(v.cpu().numpy(), g, scale=1.0)
        else:
            assert np.array_equal(v.cpu().numpy(), g), k


def test_proposal_target_empty_background_raises(bx):
    from tf_eager_object_detection_b200.proposal_target import ProposalTarget
    gt = np.float32([[10, 10, 100, 100]])
    rois = np.float32([[10, 10, 100, 100], [12, 11, 101, 99]])      # both fg, no bg: np.random.choice would raise
    with pytest.raises(ValueError):
        ProposalTarget(neg_iou_threshold=0.0)((cu(rois), cu(gt), cu(np.int32([3]))), perm=np.int32([0, 1]))


# ------------------------------------------------------------------------------------------------ composite + host path
def test_c4_composite_and_host_entry(bx):
    imgs = [syn.c4_image(2, i, channels=32) for i in range(2)]
    anchors = imgs[0]['anchors']
    deltas = np.stack([im['deltas'] for im in imgs]); scores = np.stack([im['scores'] for im in imgs])
    feat = np.stack([im['feat'] for im in imgs])
    ob, oi, oc, of = bx.c4_proposal_roi(cu(anchors), cu(deltas), cu(scores), cu(feat), (600, 1000), 300,
                                        pre_nms_top_k=6000)
    of = of.reshape(2, 300, 7, 7, 32)
    for i, im in enumerate(imgs):
        rois, idx = orc.region_proposal(im['deltas'], anchors, im['scores'], (600, 1000), 300, pre_nms_top_k=6000)
        assert np.array_equal(oi[i].cpu().numpy(), idx)
        # stage-wise exactness: pooling the GPU's own rois with the oracle reproduces the GPU features bit for bit
        assert np.array_equal(of[i].cpu().numpy(), orc.roi_pool_c4(im['feat'][None], ob[i].cpu().numpy(), 16, 7, False))
        # end to end: within 1e-5 relative of the oracle chain (decode differs by exp ulps)
        # (relative to the magnitude of the bilinear taps: a 1-ulp roi difference moves a sample by ~1e-5 feature px)
        close(of[i].cpu().numpy(), orc.roi_pool_c4(im['feat'][None], rois, 16, 7, False), scale=np.abs(im['feat']).max())
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    out_h = (torch.empty((2, 300, 4)).pin_memory(), torch.empty((2, 300), dtype=torch.int32).pin_memory(),
             torch.empty((2,), dtype=torch.int32).pin_memory(), torch.empty((600, 7, 7, 32)).pin_memory())
    bx.c4_proposal_roi_host(cu(anchors), pin(deltas), pin(scores), pin(feat), (600, 1000), 300, out_h, pre_nms_top_k=6000)
    torch.cuda.synchronize()
    assert torch.equal(out_h[0], ob.cpu()) and torch.equal(out_h[1], oi.cpu()) and torch.equal(out_h[2], oc.cpu())
    assert torch.equal(out_h[3], of.reshape(600, 7, 7, 32).cpu())


# ------------------------------------------------------------------------------------------------ full BASELINE sizes
def test_full_size_cfg2_properties(bx):
    """BASELINE cfg 2 at full size (batch 8, 38x63x1024, 6000 -> 300, 7x7x1024): size-independent properties."""
    B, C = 8, 1024
    imgs = [syn.c4_image(2, i, with_features=False) for i in range(B)]
    anchors = cu(imgs[0]['anchors'])
    deltas = cu(np.stack([im['deltas'] for im in imgs])); scores = cu(np.stack([im['scores'] for im in imgs]))
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    f1 = torch.randn((B, 38, 63, C), device='cuda', generator=g)
    f2 = torch.randn((B, 38, 63, C), device='cuda', generator=g)
    ob, oi, oc, o1 = bx.c4_proposal_roi(anchors, deltas, scores, f1, (600, 1000), 300, pre_nms_top_k=6000)
    assert (oc == 300).all()
    s_kept = torch.gather(scores, 1, oi.long())
    assert (s_kept[:, 1:] < s_kept[:, :-1]).all()                        # selection order = strictly descending score
    # idempotence: NMS over the kept boxes keeps all of them, in order
    idx2, cnt2 = bx.nms(ob, s_kept, 300, 0.7)
    assert (cnt2 == 300).all() and (idx2 == torch.arange(300, device='cuda', dtype=torch.int32)[None]).all()
    # kept indices match the oracle on every image
    for i, im in enumerate(imgs):
        _, idx = orc.region_proposal(im['deltas'], im['anchors'], im['scores'], (600, 1000), 300, pre_nms_top_k=6000)
        assert np.array_equal(oi[i].cpu().numpy(), idx)
    # linearity of crop_and_resize in the feature map: pool(f1 + f2) == pool(f1) + pool(f2) within fp32 rounding
    _, _, _, o2 = bx.c4_proposal_roi(anchors, deltas, scores, f2, (600, 1000), 300, pre_nms_top_k=6000)
    _, _, _, o12 = bx.c4_proposal_roi(anchors, deltas, scores, f1 + f2, (600, 1000), 300, pre_nms_top_k=6000)
    assert torch.allclose(o12, o1 + o2, rtol=1e-5, atol=2e-5)
    # spot-check rows against the oracle at full channel count
    sel = [0, 299, 1200, 2399]
    rows = o1.reshape(B, 300, 7, 7, C)
    for j in sel:
        i, r = divmod(j, 300)
        ref = orc.roi_pool_c4(f1[i:i + 1].cpu().numpy(), ob[i, r:r + 1].cpu().numpy(), 16, 7, False)
        assert np.array_equal(rows[i, r].cpu().numpy(), ref[0])


# ------------------------------------------------------------------------------------------------ TMA band kernel
@pytest.mark.parametrize('mode,pool,C,fh,fw,R,B', [
    ('stride', 'none', 32, 38, 63, 300, 1),      # cfg2 geometry, one slice, two bands
    ('stride', 'max', 64, 38, 63, 64, 2),        # VGG16 path: 14x14 + max pool, pairs of sample rows straddle bands
    ('align', 'avg', 32, 20, 30, 100, 1),        # dormant RoIAlign variant (symmetric pad)
    ('image', 'max', 32, 75, 125, 200, 1),       # FPN P3-sized map: narrow bands (many CTAs), image-normalised boxes
    ('stride', 'none', 96, 12, 17, 700, 3),      # more rois than one staging chunk, 3 images via box_ind
])
def test_band_kernel_matches_direct_kernel_and_oracle(bx, monkeypatch, mode, pool, C, fh, fw, R, B):
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(C + fh + R)
    H, W = fh * 16, fw * 16
    feat = rng.standard_normal((B, fh, fw, C), dtype=np.float32)
    rois = syn.random_rois(rng, R, (H, W))
    rois[:3] = np.float32([[0, 0, W - 1, H - 1], [5, 5, 5, 5], [W - 1, H - 1, W - 1, H - 1]])   # full image, point rois
    bi = rng.integers(0, B, R).astype(np.int32) if B > 1 else None
    m = {'stride': _lib.ROI_STRIDE_NORM, 'image': _lib.ROI_IMAGE_NORM, 'align': _lib.ROI_ALIGN_PAD}[mode]
    p = {'none': _lib.POOL_NONE, 'max': _lib.POOL_MAX2, 'avg': _lib.POOL_AVG2}[pool]
    kw = dict(stride=16.0, image_shape=(H, W), box_ind=None if bi is None else cu(bi))
    for k in ('BX_ROI_DIRECT', 'BX_ROI_NO_POOL2'):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv('BX_ROI_BAND_POOLED', '1')                 # pooled crops default to the roi-stationary kernel
    band = bx.roi_pool(m, p, 7, cu(feat), cu(rois), **kw).cpu().numpy()
    monkeypatch.delenv('BX_ROI_BAND_POOLED', raising=False)
    default = bx.roi_pool(m, p, 7, cu(feat), cu(rois), **kw).cpu().numpy()       # band (plain) / roi_pool2 (pooled)
    monkeypatch.setenv('BX_ROI_DIRECT', '1')
    monkeypatch.setenv('BX_ROI_NO_POOL2', '1')
    direct = bx.roi_pool(m, p, 7, cu(feat), cu(rois), **kw).cpu().numpy()        # generic gather kernel
    monkeypatch.delenv('BX_ROI_DIRECT', raising=False)
    monkeypatch.delenv('BX_ROI_NO_POOL2', raising=False)
    assert np.array_equal(band, default)
    assert np.array_equal(band, direct)
    if mode == 'stride' and pool != 'avg':
        ref = orc.roi_pool_c4(feat, rois, 16, 7, pool == 'max', box_ind=bi)
        assert np.array_equal(band, ref)
    elif mode == 'image':
        assert np.array_equal(band, orc.roi_pool_fpn(feat, rois, (H, W), 7, box_ind=bi))
    else:
        close(band, orc.roi_align_pad(feat, rois, 16, 7), scale=1.0)


def test_band_kernel_padded_rois_are_zero_filled(bx):
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(77)
    feat = rng.standard_normal((2, 38, 63, 64), dtype=np.float32)
    rois = syn.random_rois(rng, 40, (600, 1000)).reshape(2, 20, 4)
    counts = np.int32([20, 7])
    out = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, cu(feat), cu(rois), stride=16.0,
                      roi_counts=cu(counts)).cpu().numpy().reshape(2, 20, 7, 7, 64)
    for i in range(2):
        ref = orc.roi_pool_c4(feat[i:i + 1], rois[i, :counts[i]], 16, 7, False)
        assert np.array_equal(out[i, :counts[i]], ref)
        assert (out[i, counts[i]:] == 0).all()


def test_fpn_band_path_matches_oracle(bx):
    """FPN level routing through the per-level band launches (C a multiple of 32), batched with box_ind."""
    from tf_eager_object_detection_b200 import fpn as fpn_mod
    rng = np.random.default_rng(99)
    B, C = 3, 64
    shapes = syn.fpn_feature_shapes((600, 1000))[:4]
    feats = [rng.standard_normal((B, h, w, C), dtype=np.float32) for (h, w) in shapes]
    rois = np.concatenate([syn.random_rois(rng, 300, (600, 1000)), np.float32([[0, 0, 999, 599], [10, 10, 10, 10]])])
    R = rois.shape[0]
    bi = rng.integers(0, B, R).astype(np.int32)
    assert orc.level_margin(rois).min() > 1e-5
    fused, order, lv, counts = fpn_mod.fpn_roi_features(cu(rois), [cu(f) for f in feats], (600, 1000), box_ind=cu(bi))
    _, rois_list, ref_order = orc.assign_levels(rois)
    assert np.array_equal(order.cpu().numpy(), ref_order)
    assert counts.cpu().numpy().tolist() == [r_.shape[0] for r_ in rois_list]
    ref = np.concatenate([orc.roi_pool_fpn(f, rois[idx], (600, 1000), 7, box_ind=bi[idx])
                          for f, idx in zip(feats, np.split(ref_order, np.cumsum([r_.shape[0] for r_ in rois_list])[:-1]))
                          if idx.size])
    assert np.array_equal(fused.cpu().numpy(), ref)
    # single image, no box_ind (the reference's own call pattern)
    fused1, order1, _, _ = fpn_mod.fpn_roi_features(cu(rois), [cu(f[:1]) for f in feats], (600, 1000))
    ref1 = np.concatenate([orc.roi_pool_fpn(f[:1], rois[idx], (600, 1000), 7)
                           for f, idx in zip(feats, np.split(ref_order, np.cumsum([r_.shape[0] for r_ in rois_list])[:-1]))
                           if idx.size])
    assert np.array_equal(fused1.cpu().numpy(), ref1)


# ------------------------------------------------------------------------------------------------ other BASELINE configs
def test_full_size_cfg3_fpn_batch16(bx):
    """BASELINE cfg 3 (ResNet-101 FPN, 150 111 anchors P2-P6, batch 16): every image's global NMS matches the oracle;
    level routing and pooled features (C=256) match on a sampled image; size-independent properties on the whole batch."""
    from tf_eager_object_detection_b200 import fpn as fpn_mod
    B = 16
    imgs = [syn.fpn_image(3, i, with_features=False) for i in range(B)]
    anchors = cu(imgs[0]['anchors'])
    deltas = cu(np.stack([im['deltas'] for im in imgs])); scores = cu(np.stack([im['scores'] for im in imgs]))
    ob, oi, oc = bx.proposals(anchors, deltas, scores, (600, 1000), 1000)
    assert (oc == 1000).all()
    s_kept = torch.gather(scores, 1, oi.long())
    assert (s_kept[:, 1:] < s_kept[:, :-1]).all()
    for i in (0, 7, 15):
        _, idx = orc.region_proposal(imgs[i]['deltas'], imgs[i]['anchors'], imgs[i]['scores'], (600, 1000), 1000)
        assert np.array_equal(oi[i].cpu().numpy(), idx)
    g = torch.Generator(device='cuda'); g.manual_seed(3)
    feats = [torch.randn((B, h, w, 256), device='cuda', generator=g) for (h, w) in syn.fpn_feature_shapes((600, 1000))[:4]]
    rois = ob.reshape(-1, 4)
    bi = torch.arange(B, device='cuda', dtype=torch.int32).repeat_interleave(1000)
    fused, order, lv, counts = fpn_mod.fpn_roi_features(rois, feats, (600, 1000), box_ind=bi)
    assert int(counts.sum()) == B * 1000 and fused.shape == (B * 1000, 7, 7, 256)
    o = order.cpu().numpy()
    assert np.array_equal(np.sort(o), np.arange(B * 1000))                       # a permutation
    lvl = lv.cpu().numpy()
    assert (np.diff(lvl[o]) >= 0).all()                                           # level-major
    for l_ in range(2, 6):                                                        # stable inside a level
        assert (np.diff(o[lvl[o] == l_]) > 0).all()
    # oracle on image 5: its rows of the fused output, whatever their position
    rows = np.nonzero((o >= 5000) & (o < 6000))[0]
    r5 = ob[5].cpu().numpy()
    assert orc.level_margin(r5).min() > 1e-6
    _, rl, ro = orc.assign_levels(r5)
    ref = orc.fpn_roi_features(rl, [f[5:6].cpu().numpy() for f in feats], (600, 1000))
    assert np.array_equal(o[rows] - 5000, ro)
    assert np.array_equal(fused[torch.as_tensor(rows, device='cuda')].cpu().numpy(), ref)


def test_full_size_cfg4_targets_batch16(bx):
    """BASELINE cfg 4: anchor_target + proposal_target at batch 16 (21 546 anchors x 100 gt, 2000 proposals)."""
    from tf_eager_object_detection_b200.anchor_target import AnchorTarget
    from tf_eager_object_detection_b200.proposal_target import ProposalTarget
    B = 16
    rng = np.random.default_rng(syn.seed_for(4, 1))
    anchors = syn.c4_anchors(38, 63)
    gts, gls, perms = [], [], []
    for _ in range(B):
        g_, l_ = syn.gt_boxes(rng, 100, (600, 1000))
        gts.append(g_); gls.append(l_); perms.append(rng.permutation(anchors.shape[0]).astype(np.int32))
    lab, tg, iw, ow, cnt = AnchorTarget().call_batched((cu(np.stack(gts)), [600, 1000], cu(anchors)), perm=np.stack(perms))
    labn = lab.cpu().numpy()
    assert ((labn == 1).sum(1) <= 128).all() and ((labn >= 0).sum(1) <= 256).all()
    for i in (0, 9, 15):
        rl, rt, ri, ro, _ = orc.anchor_target(gts[i], [600, 1000], anchors, perms[i])
        assert np.array_equal(labn[i], rl) and np.array_equal(iw[i].cpu().numpy(), ri)
        assert np.array_equal(ow[i].cpu().numpy(), ro)
        close(tg[i].cpu().numpy(), rt, scale=1.0)
    imgs = [syn.c4_image(4, i, with_features=False) for i in range(B)]
    deltas = cu(np.stack([im['deltas'] for im in imgs])); scores = cu(np.stack([im['scores'] for im in imgs]))
    rois, _, rc = bx.proposals(cu(anchors), deltas, scores, (600, 1000), 2000)
    perm_r = np.stack([rng.permutation(2000) for _ in range(B)]).astype(np.int32)
    pt = ProposalTarget(num_classes=21, pos_iou_threshold=0.5, neg_iou_threshold=0.0, total_num_samples=128,
                        max_pos_samples=32, target_stds=[0.1, 0.1, 0.2, 0.2])
    o_r, o_l, o_t, o_i, o_o, o_k, o_c = pt.call_batched((rois, cu(np.stack(gts)), cu(np.stack(gls))), perm=perm_r, roi_counts=rc)
    assert (o_c[:, 1] == 0).all()
    for i in (0, 8, 15):
        ref = orc.proposal_target(rois[i].cpu().numpy(), gts[i], gls[i], perm_r[i], 21, 0.5, 0.0, 128, 32, stds=(0.1, 0.1, 0.2, 0.2))
        assert np.array_equal(o_k[i].cpu().numpy(), ref[5]['keep'])                 # sampled indices bit-exact
        assert np.array_equal(o_l[i].cpu().numpy(), ref[1]) and np.array_equal(o_i[i].cpu().numpy(), ref[3])
        close(o_t[i].cpu().numpy(), ref[2], scale=1.0)


def test_cfg5_shape_shard_equivalence(bx):
    """BASELINE cfg 5 (800x1333 FPN, 267 069 anchors): processing a batch in shards (as the GPUs of a node do) gives
    byte-identical per-image results to processing it whole — the property the multi-GPU all-gather relies on."""
    from tf_eager_object_detection_b200 import distributed as bxd
    B = 8
    imgs = [syn.fpn_image(5, i, (800, 1333), with_features=False) for i in range(B)]
    anchors = cu(imgs[0]['anchors'])
    assert anchors.shape[0] == 267069
    deltas = cu(np.stack([im['deltas'] for im in imgs])); scores = cu(np.stack([im['scores'] for im in imgs]))
    ob, oi, oc = bx.proposals(anchors, deltas, scores, (800, 1333), 1000)
    parts = []
    for rank in range(4):
        lo, hi = bxd.shard_bounds(B, rank, 4)
        parts.append(bx.proposals(anchors, deltas[lo:hi], scores[lo:hi], (800, 1333), 1000))
    assert torch.equal(torch.cat([p[0] for p in parts]), ob)
    assert torch.equal(torch.cat([p[1] for p in parts]), oi) and torch.equal(torch.cat([p[2] for p in parts]), oc)
    _, idx = orc.region_proposal(imgs[3]['deltas'], imgs[3]['anchors'], imgs[3]['scores'], (800, 1333), 1000)
    assert np.array_equal(oi[3].cpu().numpy(), idx)


# ------------------------------------------------------------------------------------------------ f1 post-head filtering
def test_post_ops_prediction(bx, golden):
    from tf_eager_object_detection_b200.prediction import post_ops_prediction, post_ops_prediction_batched
    hs, hd = syn.roi_head_outputs(np.random.default_rng(syn.seed_for(1, 77)), 300, 21)
    rois = golden['c4_eval_rois']
    b, c, s = post_ops_prediction(cu(hs), cu(hd), cu(rois), [600, 1000], [0, 0, 0, 0], [0.1, 0.1, 0.2, 0.2])
    assert c.dtype == torch.int32 and np.array_equal(c.cpu().numpy(), golden['post_classes'])     # kept set + order bit-exact
    assert np.array_equal(s.cpu().numpy(), golden['post_scores'])
    close(b.cpu().numpy(), golden['post_boxes'], scale=1000.0)
    assert post_ops_prediction(cu(hs), cu(hd), cu(rois), [600, 1000], None, None, score_threshold=2.0) == (None, None, None)
    # batched, other limits, roi_counts padding; COCO-sized class count
    rng = np.random.default_rng(5)
    B, R, C = 3, 200, 81
    sc, dl = zip(*[syn.roi_head_outputs(rng, R, C) for _ in range(B)])
    rr = np.stack([syn.random_rois(rng, R, (600, 1000)) for _ in range(B)])
    counts = np.int32([200, 120, 0])
    det, cnt = post_ops_prediction_batched(cu(np.stack(sc)), cu(np.stack(dl)), cu(rr), [600, 1000], None, [0.1, 0.1, 0.2, 0.2],
                                           max_num_per_class=100, max_num_per_image=100, score_threshold=0.02,
                                           roi_counts=cu(counts))
    for i in range(B):
        n = counts[i]
        ref = orc.post_ops_prediction(sc[i][:n], dl[i][:n], rr[i][:n], (600, 1000), stds=(0.1, 0.1, 0.2, 0.2),
                                      max_num_per_class=100, max_num_per_image=100, score_threshold=0.02, num_classes=C)
        k = int(cnt[i])
        if ref[0] is None:
            assert k == 0
            continue
        assert k == ref[0].shape[0]
        d = det[i, :k].cpu().numpy()
        assert np.array_equal(d[:, 5].astype(np.int32), ref[1]) and np.array_equal(d[:, 4], ref[2])
        close(d[:, :4], ref[0], scale=1000.0)
        assert (det[i, k:] == 0).all()


# ------------------------------------------------------------------------------------------------ f3 RoI pooling backward
@pytest.mark.parametrize('max_flag', [False, True])
def test_roi_pool_backward(bx, max_flag):
    """Gradient of the C4 extractor w.r.t. the feature map (what TF back-propagates through crop_and_resize + max pool
    in scripts/train.py:99-103), through torch autograd on the mirror class."""
    from tf_eager_object_detection_b200.roi_pooling import RoiPoolingCropAndResize
    rng = np.random.default_rng(31 + max_flag)
    feat = rng.standard_normal((2, 20, 30, 32), dtype=np.float32)
    rois = syn.random_rois(rng, 40, (320, 480))
    g = rng.standard_normal((40, 7, 7, 32), dtype=np.float32)
    ft = cu(feat[:1]).requires_grad_(True)
    out = RoiPoolingCropAndResize(7, max_flag)((ft, cu(rois), 16))
    assert out.requires_grad
    out.backward(cu(g))
    ref = orc.roi_pool_c4_grad(feat[:1], rois, 16, g, 7, max_flag)
    got = ft.grad.cpu().numpy()
    close(got, ref, scale=np.abs(ref).max())            # fp32 atomics: summation order differs from the oracle's
    assert np.array_equal(out.detach().cpu().numpy(), orc.roi_pool_c4(feat[:1], rois, 16, 7, max_flag))
    # batched with box_ind through the functional op; linear in grad_out
    from tf_eager_object_detection_b200 import _lib
    bi = rng.integers(0, 2, 40).astype(np.int32)
    pool = _lib.POOL_MAX2 if max_flag else _lib.POOL_NONE
    g1 = bx.roi_pool_grad(_lib.ROI_STRIDE_NORM, pool, 7, cu(feat), cu(rois), cu(g), stride=16.0, box_ind=cu(bi))
    ref2 = orc.roi_pool_c4_grad(feat, rois, 16, g, 7, max_flag, box_ind=bi)
    close(g1.cpu().numpy(), ref2, scale=np.abs(ref2).max())
    g2 = bx.roi_pool_grad(_lib.ROI_STRIDE_NORM, pool, 7, cu(feat), cu(rois), cu(2 * g), stride=16.0, box_ind=cu(bi))
    assert torch.allclose(g2, 2 * g1, rtol=1e-5, atol=1e-5)


def test_losses(bx, golden):
    """model/losses.py through the mirror: values against the reference-on-shim goldens and the oracle, gradients
    (through torch autograd) against the oracle's."""
    from tf_eager_object_detection_b200.losses import cls_loss, smooth_l1_loss
    g = golden
    cases = [(g['loss_rpn_pred'], g['at_targets'], g['at_in_w'], g['at_out_w'], 3.0, [0, 1], g['loss_rpn_reg']),
             (g['loss_roi_pred'], g['pt_targets'], g['pt_in_w'], g['pt_out_w'], 1.0, [1], g['loss_roi_reg'])]
    for pred, tgt, iw, ow, sigma, dim, want in cases:
        p = cu(pred).requires_grad_(True)
        loss = smooth_l1_loss(p, cu(tgt), cu(iw), cu(ow), sigma, dim)
        assert loss.shape == () and np.isclose(float(loss.detach()), float(want), rtol=1e-5)
        (3.0 * loss).backward()
        ref = 3.0 * orc.smooth_l1_loss_grad(pred, tgt, iw, ow, sigma, tuple(dim))
        close(p.grad.cpu().numpy(), ref, scale=np.abs(ref).max())
        again = smooth_l1_loss(cu(pred), cu(tgt), cu(iw), cu(ow), sigma, dim)
        assert float(again) == float(loss.detach())                       # deterministic reduction
    for logits, labels, want in ((g['loss_rpn_logits'], g['at_labels'], g['loss_rpn_cls']),
                                 (g['loss_roi_logits'], g['pt_labels'], g['loss_roi_cls'])):
        x = cu(logits).requires_grad_(True)
        loss = cls_loss(x, cu(labels))
        assert np.isclose(float(loss.detach()), float(want), rtol=1e-5)
        loss.backward()
        ref = orc.cls_loss_grad(logits, labels)
        close(x.grad.cpu().numpy(), ref, scale=np.abs(ref).max())
    # edge cases: nothing selected, empty input, weight
    none = cls_loss(cu(g['loss_roi_logits']), cu(np.full(128, -1, np.float32)))
    assert float(none) == 0.0
    assert float(bx.cls_loss(cu(np.zeros((0, 21), np.float32)), cu(np.zeros((0,), np.float32)))) == 0.0
    w = bx.cls_loss(cu(g['loss_roi_logits']), cu(g['pt_labels']), weight=0.5)
    assert np.isclose(float(w), 0.5 * float(g['loss_roi_cls']), rtol=1e-5)
    assert float(bx.smooth_l1_loss(*(cu(np.zeros((0, 4), np.float32)),) * 4)) == 0.0
    with pytest.raises(NotImplementedError):
        smooth_l1_loss(cu(g['loss_roi_pred']), cu(g['pt_targets']), cu(g['pt_in_w']), cu(g['pt_out_w']), dim=[0])


# ------------------------------------------------------------------------------------------------ f2 inputs of the path
def test_generate_anchors(bx, golden):
    from tf_eager_object_detection_b200 import anchor_generator as ag
    a = ag.generate_by_anchor_base_tf(ag.generate_anchor_base(16), 16, 38, 63)
    assert np.array_equal(a.cpu().numpy(), syn.c4_anchors(38, 63, 16))         # == the reference's (sha in the goldens)
    assert np.array_equal(sha(a.cpu().numpy()), golden['c4_anchors_sha'])
    for hw in ((600, 1000), (800, 1333)):
        f = ag.make_fpn_anchors(hw)
        assert np.array_equal(f.cpu().numpy(), syn.fpn_anchors(hw))
    one = ag.make_anchors(64, (1.0,), (0.5, 1.0, 2.0), 75, 125, 8)
    assert np.array_equal(one.cpu().numpy(), syn.fpn_level_anchors(64, 75, 125, 8))
    assert ag.make_anchors(64, (1.0,), (0.5, 1.0, 2.0), 0, 125, 8).shape == (0, 4)
    with pytest.raises(NotImplementedError):
        bx.generate_anchors([(2, 2)], [16.0], np.zeros((1, 33, 4), np.float32))


def test_rpn_scores_and_fused_proposals(bx, golden):
    from tf_eager_object_detection_b200 import _lib
    g = golden
    sc = bx.rpn_scores(cu(g['rpn_caffe_logits']), _lib.RPN_CAFFE, 9)
    close(sc[0].cpu().numpy(), g['rpn_caffe_scores'], scale=0.0)                # expf ulps only
    sp = bx.rpn_scores(cu(g['rpn_pairs_logits']), _lib.RPN_PAIRS)
    close(sp[0].cpu().numpy(), g['rpn_pairs_scores'], scale=0.0)
    # C4 (cached keys: softmax inside the proposal kernel), batch of 3 with different logits
    im = syn.c4_image(1, 0, with_features=False)
    rng = np.random.default_rng(91)
    logits = rng.normal(0, 3, (3, 38 * 63, 18)).astype(np.float32)
    deltas = np.stack([syn.rpn_outputs(np.random.default_rng(92 + i), 21546)[0] for i in range(3)])
    ob, oi, oc, osc = bx.proposals_rpn(cu(im['anchors']), cu(deltas), cu(logits), _lib.RPN_CAFFE, 9, (600, 1000), 300,
                                       return_scores=True)
    assert torch.equal(osc, bx.rpn_scores(cu(logits), _lib.RPN_CAFFE, 9))       # same bits fused and stand-alone
    rb, ri, rc = bx.proposals(cu(im['anchors']), cu(deltas), osc, (600, 1000), 300)
    assert torch.equal(oi, ri) and torch.equal(oc, rc) and torch.equal(ob, rb)
    nb, ni, nc = bx.proposals_rpn(cu(im['anchors']), cu(deltas), cu(logits), _lib.RPN_CAFFE, 9, (600, 1000), 300)
    assert torch.equal(ni, oi) and torch.equal(nb, ob)                          # without the score output
    for i in range(3):                                                          # oracle on the device's scores
        boxes, idx = orc.region_proposal(deltas[i], im['anchors'], osc[i].cpu().numpy(), (600, 1000), 300, 0.7)
        k = int(oc[i])
        assert k == idx.shape[0] and np.array_equal(oi[i, :k].cpu().numpy(), idx)
        close(osc[i].cpu().numpy(), orc.rpn_fg_scores(logits[i], 'caffe', 9), scale=0.0)
    # min_size path and the streaming (n > 24576) path go through the stand-alone score kernel
    mb, mi, mc = bx.proposals_rpn(cu(im['anchors']), cu(deltas), cu(logits), _lib.RPN_CAFFE, 9, (600, 1000), 300,
                                  min_size=16.0)
    eb, ei, ec = bx.proposals(cu(im['anchors']), cu(deltas), osc, (600, 1000), 300, min_size=16.0)
    assert torch.equal(mi, ei) and torch.equal(mc, ec)
    f = syn.fpn_image(3, 0, with_features=False)
    n = f['anchors'].shape[0]
    fl = rng.normal(0, 3, (2, n, 2)).astype(np.float32)
    fd = np.stack([f['deltas'], f['deltas'][::-1].copy()])
    pb, pi, pc, psc = bx.proposals_rpn(cu(f['anchors']), cu(fd), cu(fl), _lib.RPN_PAIRS, 1, (600, 1000), 1000,
                                       return_scores=True)
    qb, qi, qc = bx.proposals(cu(f['anchors']), cu(fd), psc, (600, 1000), 1000)
    assert torch.equal(pi, qi) and torch.equal(pc, qc) and torch.equal(pb, qb)
    wb, wi, wc = bx.proposals_rpn(cu(f['anchors']), cu(fd), cu(fl), _lib.RPN_PAIRS, 1, (600, 1000), 1000)
    assert torch.equal(wi, qi)                                                  # scores staged in the workspace
    with pytest.raises(ValueError):
        bx.proposals_rpn(cu(im['anchors']), cu(deltas), cu(logits), _lib.RPN_CAFFE, 8, (600, 1000), 300)


def test_detection_records_to_voc_lines(bx, golden):
    """f4 on the device's own records: [B,K,6] from post_ops_prediction_batched -> VOC result lines equal the lines
    built from the reference-on-shim goldens."""
    from tf_eager_object_detection_b200 import evaluation as ev
    from tf_eager_object_detection_b200.prediction import post_ops_prediction_batched
    hs, hd = syn.roi_head_outputs(np.random.default_rng(syn.seed_for(1, 77)), 300, 21)
    rois = golden['c4_eval_rois']
    det, cnt = post_ops_prediction_batched(cu(hs)[None], cu(hd)[None], cu(rois)[None], [600, 1000], [0, 0, 0, 0],
                                           [0.1, 0.1, 0.2, 0.2])
    lines = ev.voc_result_lines(['000001'], det, cnt)
    gb, gc, gs = golden['post_boxes'], golden['post_classes'], golden['post_scores']
    ref = np.concatenate([gb, gs[:, None], gc[:, None].astype(np.float32)], axis=1)[None]
    want = ev.voc_result_lines(['000001'], ref, np.array([gb.shape[0]]))
    assert lines == want and sum(len(v) for v in lines.values()) == gb.shape[0]
    coco = ev.coco_results([7], det, cnt)
    assert len(coco) == gb.shape[0] and coco[0]['category_id'] == ev.coco_category_ids()[int(gc[0])]


# ------------------------------------------------------------------------------------------------ top-set prefilter
def _topset_case(bx, anchors, deltas, scores, post, **kw):
    ob, oi, oc = bx.proposals(cu(anchors), cu(deltas)[None], cu(scores)[None], (600, 1000), post, **kw)
    boxes, idx = orc.region_proposal(deltas, anchors, scores, (600, 1000), post, 0.7, **kw)
    k = int(oc[0])
    assert k == idx.shape[0], (k, idx.shape[0])
    assert np.array_equal(oi[0, :k].cpu().numpy(), idx)
    assert (oi[0, k:] == -1).all()
    return k


def test_topset_prefilter_paths(bx):
    """n > 24576 goes through the multi-CTA radix select + compaction; each branch of it against the oracle:
    one-level thresholds, two- and three-level refinement (clustered scores, heavy ties), more exact ties than the
    budget (threshold above them, empty list -> full-length redo), a list that runs dry (everything suppressed),
    and the pre-NMS cut."""
    rng = np.random.default_rng(123)
    n = 30000
    anchors = syn.random_rois(rng, n, (600, 1000))
    deltas = (rng.normal(0, 1, (n, 4)) * [0.2, 0.2, 0.3, 0.3]).astype(np.float32)
    uniq = ((rng.permutation(n) + 1) / (n + 1)).astype(np.float32)
    assert _topset_case(bx, anchors, deltas, uniq, 300) == 300                       # level 0 resolves
    assert _topset_case(bx, anchors, deltas, uniq, 300, pre_nms_top_k=6000) == 300
    assert _topset_case(bx, anchors, deltas, uniq, 2000, pre_nms_top_k=12000) == 2000
    clustered = (1.0 - 1e-4 * rng.random(n)).astype(np.float32)                      # ~1700 distinct keys, all levels
    assert _topset_case(bx, anchors, deltas, clustered, 300) == 300
    mid = (0.5 + 0.01 * rng.random(n)).astype(np.float32)                            # two levels
    assert _topset_case(bx, anchors, deltas, mid, 300) == 300
    ties = uniq * 0.5
    ties[rng.permutation(n)[:26000]] = 0.75                                          # 26000 > 24576 exact ties on top
    assert _topset_case(bx, anchors, deltas, ties, 300) == 300
    ties2 = uniq * 0.5
    ties2[rng.permutation(n)[:100]] = 0.9                                            # a few above the tie wall
    ties2[rng.permutation(n)[100:26100]] = 0.75
    assert _topset_case(bx, anchors, deltas, ties2, 300) == 300
    same = np.tile(np.array([[100, 100, 300, 300]], np.float32), (n, 1))             # everything suppressed by the first
    assert _topset_case(bx, same, np.zeros((n, 4), np.float32), uniq, 50) == 1
    few = np.zeros(n, np.float32); few[:10] = uniq[:10]                              # scores 0 stay valid candidates
    assert _topset_case(bx, anchors, deltas, few, 300) == 300
    # stand-alone NMS entry over the same size
    boxes = orc.decode_bbox(anchors, deltas)
    idx_g, cnt_g = bx.nms(cu(boxes)[None], cu(uniq)[None], 300, 0.7)
    ref = orc.nms_tf(boxes, uniq, 300, 0.7)
    assert int(cnt_g[0]) == ref.shape[0] and np.array_equal(idx_g[0, :ref.shape[0]].cpu().numpy(), ref)


def test_native_allgather_on_two_ranks():
    """bx_allgather_detections on torch's own ncclComm_t against torch.distributed collectives (uneven shards), two
    processes on two GPUs; skipped on single-GPU boxes (the gloo tests cover the host logic)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_allgather_ranks.py')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29541', script], capture_output=True, text=True,
                         timeout=240)
    assert out.returncode == 0 and 'allgather ok' in out.stdout and 'prediction files ok' in out.stdout, \
        out.stdout[-2000:] + out.stderr[-2000:]


def test_nms_long_suppression_chains_and_cluster_sweep(bx):
    """Adversarial tiles for the fixed-point resolve and the cluster sweep: (a) a chain of boxes each suppressing only
    its successor (greedy keeps every other one: the dependency chain spans whole 64-candidate tiles), (b) heavy
    near-duplicate suppression with a large quota, so that many tiles run against a long, cluster-distributed kept
    list and the list runs dry before the quota."""
    w = 100.0
    k = np.arange(1500, dtype=np.float32)
    chain = np.stack([k * 12.0, np.zeros_like(k), k * 12.0 + w, np.full_like(k, 80.0)], axis=1)   # IoU(i,i+1)=.786, (i,i+2)=.61
    sc = (1.0 - k / 2000.0).astype(np.float32)
    for post in (300, 700):                                       # single CTA / cluster of 8
        idx, cnt = bx.nms(cu(chain)[None], cu(sc)[None], post, 0.7)
        ref = orc.nms_tf(chain, sc, post, 0.7)
        assert int(cnt[0]) == ref.shape[0] == post and np.array_equal(idx[0, :post].cpu().numpy(), ref)
        assert np.array_equal(ref, 2 * np.arange(post))           # every other box
    rng = np.random.default_rng(77)
    centers = rng.uniform(50, 900, (400, 2)).astype(np.float32)
    pick = rng.integers(0, 400, 20000)
    jit = rng.normal(0, 2.0, (20000, 4)).astype(np.float32)
    half = rng.uniform(20, 60, (400, 2)).astype(np.float32)
    boxes = np.concatenate([centers[pick] - half[pick], centers[pick] + half[pick]], axis=1) + jit
    scores = ((rng.permutation(20000) + 1) / 20001.0).astype(np.float32)
    for post in (600, 2000):
        idx, cnt = bx.nms(cu(boxes)[None], cu(scores)[None], post, 0.7)
        ref = orc.nms_tf(boxes, scores, post, 0.7)
        n = int(cnt[0])
        assert n == ref.shape[0] and n < post                      # the list runs dry: every candidate was visited
        assert np.array_equal(idx[0, :n].cpu().numpy(), ref) and (idx[0, n:] == -1).all()
    # a cluster launch with nothing to do: 30 000 equal scores overflow the top-set budget (empty compacted list), the
    # helpers must be released at once and the full-length redo decides by index order
    big = np.concatenate([boxes, boxes[:10000] + 1.0])
    same = np.full(30000, 0.5, np.float32)
    idx, cnt = bx.nms(cu(big)[None], cu(same)[None], 600, 0.7)
    ref = orc.nms_tf(big, same, 600, 0.7)
    n = int(cnt[0])
    assert n == ref.shape[0] and np.array_equal(idx[0, :n].cpu().numpy(), ref)


# ------------------------------------------------------------------------------------------------ round-2 parity holes
def test_nms_iou_threshold_guard_band(bx):
    """iou_gt() screens `inter / union > thr` without the division when the pair is >= 2e-6 (relative) away from the
    threshold and falls back to TF's exact fp32 division otherwise.  Pairs whose quotient is exactly the threshold, one
    ulp above and one ulp below (integer geometry, scaled down to denormal areas) must be decided like the oracle on
    every code path: the intra-tile pair test (2 boxes), the kept-list test of the single-CTA kernel (64-candidate
    tiles) and of the cluster kernel (128-candidate tiles, helpers over DSMEM)."""
    from _nms_cases import near_threshold_cases
    cases = near_threshold_cases()
    # 200 disjoint unit fillers score between the two boxes, so box B meets box A through the KEPT list, two tiles later
    k = np.arange(200, dtype=np.float32)
    filler = np.stack([1e6 + 3.0 * k, np.full_like(k, 1e6), 1e6 + 1.0 + 3.0 * k, np.full_like(k, 1e6 + 1.0)], axis=1)   # far from every pair
    fs = (0.8 - k / 1000.0).astype(np.float32)
    groups = {}
    for bxs, thr, sup in cases:
        groups.setdefault(float(thr), []).append((bxs, sup))
    n_checked = 0
    for thr, items in groups.items():
        pair = np.stack([b for b, _ in items]).astype(np.float32)                       # [g,2,4]
        sc = np.tile(np.float32([0.9, 0.1]), (len(items), 1))
        idx, cnt = bx.nms(cu(pair), cu(sc), 2, thr)
        want = np.array([1 if s else 2 for _, s in items])
        assert np.array_equal(cnt.cpu().numpy(), want), ('intra-tile', thr)
        big = np.stack([np.concatenate([b[:1], filler, b[1:]]) for b, _ in items]).astype(np.float32)   # [g,202,4]
        bsc = np.tile(np.concatenate([np.float32([0.9]), fs, np.float32([0.1])]), (len(items), 1))
        for post in (202, 600):                                                         # single CTA / cluster sweep
            idx, cnt = bx.nms(cu(big), cu(bsc), post, thr)
            assert np.array_equal(cnt.cpu().numpy(), want + 200), ('kept list', post, thr)
            last = idx.cpu().numpy()[np.arange(len(items)), want + 200 - 1]
            assert np.array_equal(last, np.where(want == 2, 201, 200))
        for (b, s) in items[:3]:                                                        # and the oracle agrees with `want`
            assert orc.nms_tf(b, np.float32([0.9, 0.1]), 2, np.float32(thr)).size == (1 if s else 2)
        n_checked += len(items)
    assert n_checked == len(cases) > 800


def test_cfg5_roi_features_one_image_all_rois(bx):
    """BASELINE cfg 5 geometry at full channel count: 800x1333, P2 = 200x334 ... P5 = 25x42, C = 256, all 1000 proposals of
    one image through bx_fpn_roi_features against the oracle's per-level crop + max pool (bit-exact given the rois)."""
    img = syn.fpn_image(5, 0, (800, 1333), channels=256)
    ob, oi, oc = bx.proposals(cu(img['anchors']), cu(img['deltas'])[None], cu(img['scores'])[None], (800, 1333), 1000)
    assert int(oc[0]) == 1000
    _, idx = orc.region_proposal(img['deltas'], img['anchors'], img['scores'], (800, 1333), 1000)
    assert np.array_equal(oi[0].cpu().numpy(), idx)
    feats = [cu(f)[None] for f in img['feats']]
    assert [tuple(f.shape[1:3]) for f in feats] == [(200, 334), (100, 167), (50, 84), (25, 42)]
    out, lv, order, counts = bx.fpn_roi_features(feats, ob[0], (800, 1333))
    rois = ob[0].cpu().numpy()
    olv, _, oorder = orc.assign_levels(rois)
    assert np.array_equal(order.cpu().numpy(), oorder) and np.array_equal(lv.cpu().numpy(), olv)
    got = out.cpu().numpy()
    pos = 0
    for l in range(4):
        k = oorder[olv[oorder] == l + 2]
        assert int(counts[l]) == k.size
        if k.size:
            ref = orc.roi_pool_fpn(img['feats'][l][None], rois[k], (800, 1333), 7)
            assert np.array_equal(got[pos:pos + k.size], ref), 'level P%d' % (l + 2)
        pos += k.size
    assert pos == 1000 and (counts > 0).sum() >= 3


def test_cfg2_one_image_all_300_rois_at_c1024(bx):
    """BASELINE cfg 2 at full channel count: every one of the 300 RoI feature blocks [7,7,1024] of one image, through the
    composite entry point (TMA band kernel), bit-exact against the oracle."""
    from tf_eager_object_detection_b200 import _lib
    img = syn.c4_image(2, 3)
    s0 = _lib.total_stats()
    before, fb0 = s0.get('band_launches', 0), s0.get('band_fallbacks', 0)
    ob, oi, oc, of = bx.c4_proposal_roi(cu(img['anchors']), cu(img['deltas'])[None], cu(img['scores'])[None],
                                        cu(img['feat'])[None], (600, 1000), 300, pre_nms_top_k=6000)
    assert int(oc[0]) == 300
    st = _lib.total_stats()
    assert st['band_launches'] == before + 1 and st['band_fallbacks'] == fb0        # served by the band kernel, no fallback
    ref = orc.roi_pool_c4(img['feat'][None], ob[0].cpu().numpy(), 16, 7, False)
    assert ref.shape == (300, 7, 7, 1024)
    assert np.array_equal(of.cpu().numpy(), ref)


def test_band_fallback_is_counted(bx):
    """A plain crop outside the band kernel's limits (C % 32 != 0) runs on the gather kernel and is reported by bx_stats."""
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(9)
    feat = rng.standard_normal((1, 20, 30, 20), dtype=np.float32)
    rois = syn.random_rois(rng, 50, (320, 480))
    s0 = _lib.total_stats()
    out = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, cu(feat), cu(rois))
    s1 = _lib.total_stats()
    assert s1['band_fallbacks'] == s0.get('band_fallbacks', 0) + 1 and s1['band_launches'] == s0.get('band_launches', 0)
    assert np.array_equal(out.cpu().numpy(), orc.roi_pool_c4(feat, rois, 16, 7, False))


@pytest.mark.parametrize('which', ['image_max', 'align_avg'])
def test_roi_pool_backward_fpn_and_roialign(bx, which):
    """bx_roi_pool_grad for the two other extractors: IMAGE_NORM + 2x2 max (the FPN training path,
    model/roi_pooling.py:15-42) and ALIGN_PAD + 2x2 mean (dormant RoIAlign, :93-176), functional op and autograd hook."""
    from tf_eager_object_detection_b200 import _lib
    from tf_eager_object_detection_b200.roi_pooling import RoiPoolingCropAndResize2, RoiPoolingRoiAlign
    rng = np.random.default_rng(41 + len(which))
    feat = rng.standard_normal((2, 25, 38, 32), dtype=np.float32)
    H, W = 400, 608
    rois = syn.random_rois(rng, 60, (H, W))
    bi = rng.integers(0, 2, 60).astype(np.int32)
    g = rng.standard_normal((60, 7, 7, 32), dtype=np.float32)
    if which == 'image_max':
        got = bx.roi_pool_grad(_lib.ROI_IMAGE_NORM, _lib.POOL_MAX2, 7, cu(feat), cu(rois), cu(g), image_shape=(H, W), box_ind=cu(bi))
        ref = orc.roi_pool_fpn_grad(feat, rois, (H, W), g, 7, box_ind=bi)
        ft = cu(feat[:1]).requires_grad_(True)
        out = RoiPoolingCropAndResize2(7)((ft, cu(rois), (H, W)))
        fwd, ref1 = orc.roi_pool_fpn(feat[:1], rois, (H, W), 7), orc.roi_pool_fpn_grad(feat[:1], rois, (H, W), g, 7)
    else:
        got = bx.roi_pool_grad(_lib.ROI_ALIGN_PAD, _lib.POOL_AVG2, 7, cu(feat), cu(rois), cu(g), stride=16.0, box_ind=cu(bi))
        ref = orc.roi_align_pad_grad(feat, rois, 16, g, 7, box_ind=bi)
        ft = cu(feat[:1]).requires_grad_(True)
        out = RoiPoolingRoiAlign(7)((ft, cu(rois), 16))
        fwd, ref1 = orc.roi_align_pad(feat[:1], rois, 16, 7), orc.roi_align_pad_grad(feat[:1], rois, 16, g, 7)
    close(got.cpu().numpy(), ref, scale=np.abs(ref).max())          # fp32 accumulation order differs from the oracle's
    close(out.detach().cpu().numpy(), fwd, scale=np.abs(fwd).max())
    out.backward(cu(g))
    close(ft.grad.cpu().numpy(), ref1, scale=np.abs(ref1).max())


def test_targets_without_ground_truth(bx):
    """An image with no ground-truth boxes: max_gt == 0 (NULL gt pointer) must behave like gt_counts == 0 — all inside
    anchors background (label 0), 256 of them sampled, zero targets / inside weights; ProposalTarget pads with background."""
    from tf_eager_object_detection_b200.anchor_target import AnchorTarget
    img = syn.c4_image(4, 0, with_features=False)
    n = img['anchors'].shape[0]
    perm = np.random.default_rng(3).permutation(n).astype(np.int32)[None]
    at = AnchorTarget()
    empty = torch.zeros((1, 0, 4), device='cuda')
    lab0, tg0, iw0, ow0, c0 = at.call_batched((empty, [600, 1000], cu(img['anchors'])), perm=perm)
    dummy = torch.zeros((1, 5, 4), device='cuda')
    lab1, tg1, iw1, ow1, c1 = at.call_batched((dummy, [600, 1000], cu(img['anchors'])), perm=perm,
                                              gt_counts=torch.zeros(1, dtype=torch.int32, device='cuda'))
    for x, y in ((lab0, lab1), (tg0, tg1), (iw0, iw1), (ow0, ow1), (c0, c1)):
        assert torch.equal(x, y)
    lab = lab0[0].cpu().numpy()
    assert set(np.unique(lab).tolist()) == {-1.0, 0.0} and int((lab == 0).sum()) == 256
    assert c0[0].tolist() == [0, 256] and float(tg0.abs().max()) == 0.0 and float(iw0.abs().max()) == 0.0
    rois = cu(syn.random_rois(np.random.default_rng(4), 300, (600, 1000)))[None]
    permr = np.random.default_rng(5).permutation(300).astype(np.int32)[None]
    o = bx.proposal_target(rois, torch.zeros((1, 0, 4), device='cuda'), torch.zeros((1, 0), dtype=torch.int32, device='cuda'),
                           permr, neg_iou_threshold=0.0)
    assert int(o[6][0, 0]) == 0 and int(o[6][0, 1]) == 0 and (o[1] == 0).all() and float(o[3].abs().max()) == 0.0


def test_cls_loss_rejects_out_of_range_labels(bx):
    """tf.losses.sparse_softmax_cross_entropy raises (CPU) / yields NaN (GPU) for a label >= num_classes; the library
    yields NaN — never a silently clamped class."""
    logits = torch.randn((16, 5), device='cuda')
    labels = torch.tensor([0, 1, 2, 3, 4, 5, -1, 0] * 2, device='cuda', dtype=torch.float32)
    loss, grad = bx.cls_loss(logits, labels, with_grad=True)
    assert torch.isnan(loss) and torch.isnan(grad[5]).all() and not torch.isnan(grad[4]).any()
    ok = labels.clone(); ok[ok == 5] = 4
    assert torch.isfinite(bx.cls_loss(logits, ok))
    from tf_eager_object_detection_b200.prediction import post_ops_prediction
    with pytest.raises(ValueError):
        post_ops_prediction(torch.rand((4, 21), device='cuda'), torch.zeros((4, 21, 4), device='cuda'),
                            torch.zeros((4, 4), device='cuda'), [600, 1000], None, None, num_classes=20)


def test_no_hidden_conversions_on_the_device_path(bx, monkeypatch):
    """Tensor hand-off rules (_tensor.py): well-formed CUDA tensors are borrowed zero-copy — the conversion counters stay
    at zero — and BX_STRICT=1 turns every upload / cast / compaction into a TypeError."""
    from tf_eager_object_detection_b200 import _tensor
    img = syn.c4_image(2, 0, channels=32)
    a, d, s, f = cu(img['anchors']), cu(img['deltas'])[None], cu(img['scores'])[None], cu(img['feat'])[None]
    before = dict(_tensor.CONVERSIONS)
    bx.c4_proposal_roi(a, d, s, f, (600, 1000), 300, pre_nms_top_k=6000)
    assert _tensor.CONVERSIONS == before
    bx.nms(img['anchors'], img['scores'], 10, 0.7)                    # numpy inputs: uploaded, and counted
    assert _tensor.CONVERSIONS['uploads'] == before['uploads'] + 2
    monkeypatch.setenv('BX_STRICT', '1')
    with pytest.raises(TypeError):
        bx.nms(img['anchors'], img['scores'], 10, 0.7)
    with pytest.raises(TypeError):
        bx.nms(a.double(), s[0], 10, 0.7)
    with pytest.raises(TypeError):
        bx.roi_pool(0, 0, 7, f.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2), a[:8].clone())
    bx.c4_proposal_roi(a, d, s, f, (600, 1000), 300, pre_nms_top_k=6000)   # the well-formed call still passes


def test_calls_run_on_the_handles_device_not_the_current_one():
    """A handle created for cuda:1 launches on cuda:1 (and allocates its workspace there) while cuda:0 is current, and
    neither bx_create nor the calls change the caller's current device.  Needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import tf_eager_object_detection_b200.ops as ops
    torch.cuda.set_device(0)
    img = syn.fpn_image(3, 1, with_features=False)
    dev1 = torch.device('cuda', 1)
    a = torch.as_tensor(img['anchors']).to(dev1); d = torch.as_tensor(img['deltas']).to(dev1)[None]
    s = torch.as_tensor(img['scores']).to(dev1)[None]
    ob, oi, oc = ops.proposals(a, d, s, (600, 1000), 1000)            # top-set path: workspace allocation + memsets + 8 launches
    assert torch.cuda.current_device() == 0 and ob.device == dev1
    _, idx = orc.region_proposal(img['deltas'], img['anchors'], img['scores'], (600, 1000), 1000)
    assert np.array_equal(oi[0].cpu().numpy(), idx)
    x = torch.empty(4, device='cuda')
    assert x.device.index == 0


def test_tma_staged_extractor_matches_default_and_oracle(bx, monkeypatch):
    """bx_roi_stage.cu (opt-in, BX_ROI_STAGE=1): the persistent TMA-staged FPN extractor — producer warp, mbarrier ring,
    pixel-column walker — must reproduce the default kernel and the oracle bit for bit, including rois cut into several
    strips, a pooled row too large for a stage (global taps), inverted boxes (non-monotone walker) and padded rois."""
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(123)
    hw = (600, 1000)
    shapes = syn.fpn_feature_shapes(hw)[:4]
    B, C = 2, 128
    feats = [rng.standard_normal((B, h, w, C), dtype=np.float32) for h, w in shapes]
    rois = syn.random_rois(rng, 400, hw)
    rois[:8] = np.float32([[10, 5, 40, 590], [3, 3, 990, 30], [0, 0, 999, 599], [500, 300, 500, 300],     # tall / wide / whole / point
                           [200, 400, 100, 100], [-50, -50, 20, 20], [980, 580, 1100, 700], [300, 200, 310, 212]])   # inverted, outside
    bi = rng.integers(0, B, 400).astype(np.int32)
    fc = [cu(f) for f in feats]
    monkeypatch.delenv('BX_ROI_STAGE', raising=False)
    base, lv, order, counts = bx.fpn_roi_features(fc, cu(rois), hw, box_ind=cu(bi))
    monkeypatch.setenv('BX_ROI_STAGE', '1')
    for kb in ('', '12'):                                   # default stage size, and one small enough to force many strips
        if kb:
            monkeypatch.setenv('BX_ROI_STAGE_KB', kb)
        got, lv2, order2, _ = bx.fpn_roi_features(fc, cu(rois), hw, box_ind=cu(bi))
        assert torch.equal(order2, order) and torch.equal(got, base), 'stage budget %r' % kb
    olv, _, oorder = orc.assign_levels(rois)
    assert np.array_equal(order.cpu().numpy(), oorder)
    got = base.cpu().numpy()
    pos = 0
    for l in range(4):
        k = oorder[olv[oorder] == l + 2]
        if k.size:
            assert np.array_equal(got[pos:pos + k.size], orc.roi_pool_fpn(feats[l], rois[k], hw, 7, box_ind=bi[k])), l
        pos += k.size
    # C4 pooled extractors through the same kernel: VGG16 14x14 + max (stride norm) and RoIAlign (pad + mean)
    feat = rng.standard_normal((1, 38, 63, 64), dtype=np.float32)
    r2 = syn.random_rois(rng, 100, hw)
    a = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_MAX2, 7, cu(feat), cu(r2), stride=16.0)
    b = bx.roi_pool(_lib.ROI_ALIGN_PAD, _lib.POOL_AVG2, 7, cu(feat), cu(r2), stride=16.0)
    assert np.array_equal(a.cpu().numpy(), orc.roi_pool_c4(feat, r2, 16, 7, True))
    close(b.cpu().numpy(), orc.roi_align_pad(feat, r2, 16, 7), scale=1.0)


def test_eval_loop_detections(bx, golden):
    """f1, evaluation-loop form (bx_eval_detections): three images of different raw size and resize factor in ONE batched
    call — rois / img_scale, per-image clip, per-class NMS, and the VOC loop's `score >= k-th largest` cut with 26 tied
    detections surviving on image 3 — against the oracle (sets and order exact) and against the text of the result files
    the reference's own get_prediction_files wrote (golden), through the package's file writer."""
    from oracle.voc_fixture import eval_loop_inputs
    from tf_eager_object_detection_b200 import evaluation as ev
    from tf_eager_object_detection_b200.prediction import eval_loop_detections
    imgs = eval_loop_inputs()
    names = ['%06d' % (i + 1) for i in range(len(imgs))]
    scores = cu(np.stack([im['scores'] for im in imgs])); deltas = cu(np.stack([im['deltas'] for im in imgs]))
    rois = cu(np.stack([im['rois'] for im in imgs]))
    scale = cu(np.float32([im['scale'] for im in imgs])); raw = cu(np.float32([[im['raw_h'], im['raw_w']] for im in imgs]))
    for loop, max_img, rows, want in (('voc', 50, 128, [50, 50, 76]), ('coco', 50, None, [50, 50, 50]), ('voc', 0, None, None)):
        det, cnt = eval_loop_detections(scores, deltas, rois, scale, raw, score_threshold=0.05, iou_threshold=0.3,
                                        max_objects_per_class=50, max_objects_per_image=max_img, min_size=10, loop=loop,
                                        out_rows=rows)
        d, c = det.cpu().numpy(), cnt.cpu().numpy()
        if want is not None:
            assert c.tolist() == want, (loop, c.tolist())
        for i, im in enumerate(imgs):
            ref = orc.eval_loop_detections(im['scores'], im['deltas'], im['rois'], im['scale'], im['raw_h'], im['raw_w'],
                                           max_objects_per_image=max_img, loop=loop)
            rec = d[i, :c[i]]
            assert (d[i, c[i]:] == 0).all()
            assert (np.diff(rec[:, 4]) <= 0).all()                                  # records in descending score order
            for j in range(1, 21):
                mine = rec[rec[:, 5] == j]
                assert mine.shape[0] == ref[j][1].size, (loop, i, j)
                assert np.array_equal(mine[:, 4], ref[j][1])                        # scores exact, NMS order within the class
                close(mine[:, :4], ref[j][0], scale=1000.0)
        if loop == 'voc':
            lines = ev.voc_result_lines(names, d, c)
            got = ''.join(''.join(lines[j]) for j in range(1, 21)).splitlines()
            ref_lines = bytes(golden['eval_voc_files' if max_img else 'eval_voc_nocut_files']).decode().splitlines()
            assert len(got) == len(ref_lines)
            for x, y in zip(got, ref_lines):          # same image, same printed score; corners printed to 0.1 px: an expf ulp
                gx, gy = x.split(), y.split()         # may move a corner across a rounding boundary, never further
                assert gx[:2] == gy[:2] and all(abs(float(p) - float(q)) <= 0.1001 for p, q in zip(gx[2:], gy[2:])), (x, y)
    with pytest.raises(ValueError):
        bx.eval_detections(scores, deltas, rois, raw, scale, max_num_per_image=50, out_rows=10)     # out_rows < max_per_image


def test_cuda_graph_capture_and_stream_ordered_workspaces(bx):
    """The calls are capturable into a CUDA graph (what bench.py times): a capture that would have to GROW a workspace is
    refused with NotImplementedError (BX_ERR_UNSUPPORTED), after one warm-up call (or bx_reserve) the same capture works and
    its replay reproduces the eager result; growth itself is stream-ordered — no device-wide synchronisation — and the
    caller's current stream / device are untouched."""
    import ctypes
    from tf_eager_object_detection_b200 import _lib
    img = syn.fpn_image(3, 2, with_features=False)                       # 150 111 anchors: top-set workspace needed
    a, d, s = cu(img['anchors']), cu(img['deltas'])[None], cu(img['scores'])[None]
    side = torch.cuda.Stream()
    lib = _lib.load()
    with torch.cuda.stream(side):
        h = _lib.handle(0, side.cuda_stream)
        assert _lib.stats(h)['ws_bytes'] == 0
        g = torch.cuda.CUDAGraph()
        with pytest.raises(NotImplementedError):
            with torch.cuda.graph(g, stream=side, capture_error_mode='thread_local'):
                bx.proposals(a, d, s, (600, 1000), 1000)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        ob, oi, oc = bx.proposals(a, d, s, (600, 1000), 1000)            # warm-up: the workspace grows here (cudaMallocAsync)
        assert _lib.stats(h)['ws_bytes'] > 0
        out = [torch.empty_like(ob), torch.empty_like(oi), torch.empty_like(oc)]
        p = bx.proposal_params((600, 1000), 1000)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode='thread_local'):
            _lib.check(lib.bx_proposals(h, a.data_ptr(), d.data_ptr(), s.data_ptr(), 1, a.shape[0], ctypes.byref(p),
                                        out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), ctypes.c_void_p(side.cuda_stream)))
        for t in out:
            t.zero_()
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], ob) and torch.equal(out[1], oi) and torch.equal(out[2], oc)
    _, idx = orc.region_proposal(img['deltas'], img['anchors'], img['scores'], (600, 1000), 1000)
    assert np.array_equal(out[1][0].cpu().numpy(), idx)
    # bx_reserve sizes a fresh handle up front: the first call of that handle can then be captured directly
    hh = ctypes.c_void_p()
    _lib.check(lib.bx_create(0, ctypes.byref(hh)))
    st = _lib.stats(h)
    _lib.check(lib.bx_reserve(hh, st['ws_bytes'], st['plan_bytes'], 0, ctypes.c_void_p(side.cuda_stream)))
    with torch.cuda.stream(side):
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, stream=side, capture_error_mode='thread_local'):
            _lib.check(lib.bx_proposals(hh, a.data_ptr(), d.data_ptr(), s.data_ptr(), 1, a.shape[0], ctypes.byref(p),
                                        out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), ctypes.c_void_p(side.cuda_stream)))
        out[1].zero_()
        g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[1], oi)
    del g2
    _lib.check(lib.bx_destroy(hh))


def test_band_height_does_not_change_the_crops(bx, monkeypatch):
    """BX_ROI_BAND_ROWS caps the feature rows a band owns (measurement switch): more, shorter bands — every roi split over
    more (image, band) units, more halo rows — must give the same bits."""
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(78)
    feat = cu(rng.standard_normal((2, 38, 63, 64), dtype=np.float32))
    rois = np.stack([syn.random_rois(rng, 300, (600, 1000)) for _ in range(2)])
    counts = cu(np.int32([300, 123]))
    monkeypatch.delenv('BX_ROI_BAND_ROWS', raising=False)
    base = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, cu(rois), stride=16.0, roi_counts=counts)
    for rows in ('2', '5', '8'):
        monkeypatch.setenv('BX_ROI_BAND_ROWS', rows)
        got = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, cu(rois), stride=16.0, roi_counts=counts)
        assert torch.equal(got, base), rows


def test_persistent_band_kernel_matches_default(bx, monkeypatch):
    """BX_ROI_BAND_PERSIST=1: the persistent, double-buffered form of the TMA band kernel (one CTA per SM, units drawn from
    a global counter, next band + plan block prefetched) must reproduce the per-unit kernel bit for bit — plain crops at the
    cfg2 geometry with padded rois, several plan blocks per unit, and a pooled crop through the band path."""
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(77)
    B, C = 3, 96
    feat = cu(rng.standard_normal((B, 38, 63, C), dtype=np.float32))
    rois = np.stack([syn.random_rois(rng, 400, (600, 1000)) for _ in range(B)])
    counts = cu(np.int32([400, 250, 0]))
    monkeypatch.delenv('BX_ROI_BAND_PERSIST', raising=False)
    base = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, cu(rois), stride=16.0, roi_counts=counts)
    monkeypatch.setenv('BX_ROI_BAND_PERSIST', '1')
    got = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, cu(rois), stride=16.0, roi_counts=counts)
    assert torch.equal(got, base)
    ref = orc.roi_pool_c4(feat[:1].cpu().numpy(), rois[0], 16, 7, False)
    assert np.array_equal(got[:400].cpu().numpy(), ref) and (got[650:] == 0).all()
    monkeypatch.setenv('BX_ROI_BAND_POOLED', '1')
    pooled = bx.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_MAX2, 7, feat[:1], cu(rois[0]), stride=16.0)
    assert np.array_equal(pooled.cpu().numpy(), orc.roi_pool_c4(feat[:1].cpu().numpy(), rois[0], 16, 7, True))


@pytest.mark.parametrize('case', ['plain7', 'plain10_wide', 'max5', 'avg7_pad', 'counts_c96'])
def test_roi_pool_backward_row_owned_kernel(bx, monkeypatch, case):
    """The row-owned, atomic-free backward kernel (csrc/bx_roi_grad.cu; BX_ROI_GRAD_DETERMINISTIC=1 or
    bx_set_deterministic): same gradients as the oracle within the fp32 tolerance, identical bits run to run, with the warps
    of a CTA on different channel slices (SPLIT=0) and on one slice (SPLIT=1), for every extractor, pooled sizes with one and
    two chunks of output pixels, a map wider than one 64-pixel segment, channel counts below / not a multiple of a slice,
    and padded rois behind roi_counts.  The scatter kernel (BX_ROI_GRAD_ATOMIC=1) must agree within the same tolerance."""
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(hash(case) % 1000)
    kw = {}
    if case == 'plain7':
        feat = rng.standard_normal((2, 38, 63, 64), dtype=np.float32); P = 7; pool = _lib.POOL_NONE; mode = _lib.ROI_STRIDE_NORM
        rois = syn.random_rois(rng, 90, (600, 1000)); bi = rng.integers(0, 2, 90).astype(np.int32)
        ref = orc.roi_pool_c4_grad
        args = lambda rr, g: (feat, rr, 16, g, P, False)
        kw = dict(stride=16.0)
    elif case == 'plain10_wide':
        feat = rng.standard_normal((1, 20, 150, 8), dtype=np.float32); P = 10; pool = _lib.POOL_NONE; mode = _lib.ROI_STRIDE_NORM
        rois = syn.random_rois(rng, 70, (80, 600)); bi = np.zeros(70, np.int32)
        ref = orc.roi_pool_c4_grad
        args = lambda rr, g: (feat, rr, 4, g, P, False)
        kw = dict(stride=4.0)
    elif case == 'max5':
        feat = rng.standard_normal((2, 25, 38, 32), dtype=np.float32); P = 5; pool = _lib.POOL_MAX2; mode = _lib.ROI_STRIDE_NORM
        rois = syn.random_rois(rng, 60, (400, 608)); bi = rng.integers(0, 2, 60).astype(np.int32)
        ref = orc.roi_pool_c4_grad
        args = lambda rr, g: (feat, rr, 16, g, P, True)
        kw = dict(stride=16.0)
    elif case == 'avg7_pad':
        feat = rng.standard_normal((2, 25, 38, 32), dtype=np.float32); P = 7; pool = _lib.POOL_AVG2; mode = _lib.ROI_ALIGN_PAD
        rois = syn.random_rois(rng, 60, (400, 608)); bi = rng.integers(0, 2, 60).astype(np.int32)
        ref = orc.roi_align_pad_grad
        args = lambda rr, g: (feat, rr, 16, g, P)
        kw = dict(stride=16.0)
    else:
        feat = rng.standard_normal((3, 38, 63, 96), dtype=np.float32); P = 7; pool = _lib.POOL_NONE; mode = _lib.ROI_STRIDE_NORM
        rois = np.stack([syn.random_rois(rng, 40, (600, 1000)) for _ in range(3)]).reshape(-1, 4)
        counts = np.int32([40, 0, 17])
        bi = np.repeat(np.arange(3, dtype=np.int32), 40)
        bi[40:80] = -1; bi[80 + 17:] = -1                            # rois behind the counts: left out of the oracle's input below
        ref = orc.roi_pool_c4_grad
        args = lambda rr, g: (feat, rr, 16, g, P, False)
        kw = dict(stride=16.0, roi_counts=cu(counts))
    g = rng.standard_normal((rois.shape[0], P, P, feat.shape[3]), dtype=np.float32)
    keep = bi >= 0
    want = ref(*args(rois[keep], g[keep]), box_ind=bi[keep])
    scale = np.abs(want).max()
    if 'roi_counts' not in kw:
        kw['box_ind'] = cu(bi)
    call = lambda: bx.roi_pool_grad(mode, pool, P, cu(feat), cu(rois), cu(g), **kw)
    monkeypatch.setenv('BX_ROI_GRAD_ATOMIC', '1')
    close(call().cpu().numpy(), want, scale=scale)
    monkeypatch.setenv('BX_ROI_GRAD_ATOMIC', '0')
    monkeypatch.setenv('BX_ROI_GRAD_DETERMINISTIC', '1')
    for split in ('0', '1'):
        monkeypatch.setenv('BX_ROI_GRAD_SPLIT', split)
        a1, a2 = call(), call()
        assert torch.equal(a1, a2)
        close(a1.cpu().numpy(), want, scale=scale)
    monkeypatch.delenv('BX_ROI_GRAD_SPLIT')
    monkeypatch.delenv('BX_ROI_GRAD_DETERMINISTIC')
    # torch's switch reaches the handle (bx_set_deterministic): the 2x2 max, whose default is the scatter kernel, becomes
    # reproducible too
    before = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True)
    try:
        b1, b2 = call(), call()
    finally:
        torch.use_deterministic_algorithms(before)
    assert torch.equal(b1, b2)
    close(b1.cpu().numpy(), want, scale=scale)


def test_roi_pool_backward_kernels_agree_on_random_shapes(bx, monkeypatch):
    """Randomised cross-check of the two backward kernels (and the oracle on the smaller cases): pooled sizes 1..12, all three
    extractors, 4..100 channels, maps up to 90 x 150 (up to three 64-pixel segments), boxes that are inverted, degenerate,
    partly or wholly outside the image, images without rois."""
    from tf_eager_object_detection_b200 import _lib
    rng = np.random.default_rng(20260)
    for trial in range(24):
        P = int(rng.integers(1, 13))
        pool = int(rng.integers(0, 3))
        mode = [_lib.ROI_STRIDE_NORM, _lib.ROI_IMAGE_NORM, _lib.ROI_ALIGN_PAD][int(rng.integers(0, 3))]
        if P == 1 and pool == _lib.POOL_NONE and mode == _lib.ROI_ALIGN_PAD:
            P = 2
        b = int(rng.integers(1, 4))
        fh, fw = int(rng.integers(3, 91)), int(rng.integers(3, 151))
        c = 4 * int(rng.integers(1, 26))
        stride = float(rng.choice([4.0, 8.0, 16.0]))
        H, W = int(fh * stride), int(fw * stride)
        r = int(rng.integers(1, 60))
        rois = syn.random_rois(rng, r, (H, W)).astype(np.float32)
        k = rng.integers(0, 6, r)
        rois[k == 1] = rois[k == 1][:, [2, 3, 0, 1]]                      # inverted
        rois[k == 2, 2:] = rois[k == 2, :2]                               # zero size
        rois[k == 3] += np.float32([W, H, W, H]) * np.float32(0.6)        # partly / wholly outside
        rois[k == 4] -= np.float32([W, H, W, H]) * np.float32(0.4)
        bi = rng.integers(0, b, r).astype(np.int32)
        if b > 1:
            bi[bi == b - 1] = 0                                           # the last image has no rois
        feat = rng.standard_normal((b, fh, fw, c), dtype=np.float32)
        g = rng.standard_normal((r, P, P, c), dtype=np.float32)
        kw = dict(stride=stride, image_shape=(H, W), box_ind=cu(bi))
        call = lambda: bx.roi_pool_grad(mode, pool, P, cu(feat), cu(rois), cu(g), **kw)
        monkeypatch.setenv('BX_ROI_GRAD_ATOMIC', '1')
        ref = call()
        monkeypatch.setenv('BX_ROI_GRAD_ATOMIC', '0')
        monkeypatch.setenv('BX_ROI_GRAD_DETERMINISTIC', '1')
        monkeypatch.setenv('BX_ROI_GRAD_SPLIT', str(trial & 1))
        a1, a2 = call(), call()
        monkeypatch.delenv('BX_ROI_GRAD_DETERMINISTIC')
        monkeypatch.delenv('BX_ROI_GRAD_SPLIT')
        tag = 'trial %d: P=%d pool=%d mode=%d b=%d map=%dx%d c=%d r=%d' % (trial, P, pool, mode, b, fh, fw, c, r)
        assert torch.equal(a1, a2), tag
        scale = max(float(ref.abs().max()), 1e-6)
        assert float((a1 - ref).abs().max()) <= 2e-5 * scale, tag
        if mode == _lib.ROI_STRIDE_NORM and pool != _lib.POOL_AVG2 and trial % 3 == 0:
            want = orc.roi_pool_c4_grad(feat, rois, stride, g, P, pool == _lib.POOL_MAX2, box_ind=bi)
            close(a1.cpu().numpy(), want, scale=max(np.abs(want).max(), 1e-6))
