"""Independent witnesses for the oracle's restated TensorFlow kernels (parity is unpinned at that boundary):
torchvision NMS / box_iou, torch grid_sample, the oracle's C twin, hand-derived known answers, and
hypothesis properties of greedy NMS."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import boxpath_oracle as orc
from tf_eager_object_detection_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nms_matches_torchvision_on_positive_area_unique_scores():
    tv = pytest.importorskip('torchvision')
    rng = np.random.default_rng(0)
    for n, thr in ((50, 0.5), (400, 0.7), (2000, 0.3)):
        b = syn.random_rois(rng, n, (600, 1000))
        b = b[(b[:, 2] > b[:, 0]) & (b[:, 3] > b[:, 1])]
        s = ((rng.permutation(len(b)) + 1) / (len(b) + 1)).astype(np.float32)
        ref = tv.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
        got = orc.nms_tf(b, s, len(b), thr)
        assert np.array_equal(got, ref)
    assert orc.nms_tf(np.float32([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30]]), np.float32([.9, .8, .7]), 3, 0.5).tolist() == [0, 2]


def test_nms_iou_matches_torchvision_box_iou():
    tv = pytest.importorskip('torchvision')
    iou = tv.ops.box_iou(torch.tensor([[0., 0., 10., 10.]]), torch.tensor([[1., 1., 11., 11.]])).item()
    assert abs(iou - 81.0 / 119.0) < 1e-6                      # no +1 in the NMS IoU
    # threshold just below / above that IoU flips the decision
    b = np.float32([[0, 0, 10, 10], [1, 1, 11, 11]]); s = np.float32([0.9, 0.8])
    assert orc.nms_tf(b, s, 2, 0.68).tolist() == [0] and orc.nms_tf(b, s, 2, 0.69).tolist() == [0, 1]


def test_crop_and_resize_matches_grid_sample_inside_the_map():
    rng = np.random.default_rng(1)
    img = rng.standard_normal((1, 17, 23, 5)).astype(np.float32)
    boxes = np.float32([[0.1, 0.2, 0.8, 0.9], [0.0, 0.0, 1.0, 1.0], [0.3, 0.1, 0.35, 0.6]])
    out = orc.crop_and_resize_tf(img, boxes, np.zeros(3, np.int32), 7, 7)
    t = torch.from_numpy(img).permute(0, 3, 1, 2)
    for k, (y1, x1, y2, x2) in enumerate(boxes):
        ys = torch.linspace(float(y1), float(y2), 7) * 2 - 1
        xs = torch.linspace(float(x1), float(x2), 7) * 2 - 1
        gy, gx = torch.meshgrid(ys, xs, indexing='ij')
        grid = torch.stack([gx, gy], -1)[None]
        ref = torch.nn.functional.grid_sample(t, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
        np.testing.assert_allclose(out[k], ref[0].permute(1, 2, 0).numpy(), rtol=1e-4, atol=1e-5)


def test_known_answers_survey_a8():
    dec = orc.decode_bbox(np.float32([[-84, -40, 99, 55]]), np.zeros((1, 4), np.float32))
    assert dec.tolist() == [[-84.0, -40.0, 100.0, 56.0]]
    assert orc.bboxes_clip_filter(dec, 0, 600, 1000)[0].tolist() == [[0.0, 0.0, 100.0, 56.0]]
    sides = np.float32([300, 150, 500, 60, 20, 900])
    b = np.stack([np.zeros(6, np.float32), np.zeros(6, np.float32), sides, sides], 1)
    assert orc.assign_levels(b)[0].tolist() == [4, 3, 5, 2, 2, 5]
    sq = np.random.default_rng(2).standard_normal((1, 9, 9, 3)).astype(np.float32)
    out = orc.crop_and_resize_tf(sq, np.float32([[0, 0, 1, 1], [0, 0, 1.5, 1]]), np.int32([0, 0]), 9, 9)
    assert np.array_equal(out[0], sq[0]) and (out[1][6:] == 0).all()
    assert syn.anchor_base(16).astype(np.int64).tolist()[0] == [-84, -40, 99, 55]


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 120), st.floats(0.05, 0.95), st.integers(0, 2 ** 31 - 1))
def test_nms_properties(n, thr, seed):
    rng = np.random.default_rng(seed)
    b = syn.random_rois(rng, n, (300, 400))
    s = ((rng.permutation(n) + 1) / (n + 1)).astype(np.float32)
    keep = orc.nms_tf(b, s, n, thr)
    assert (np.diff(s[keep]) < 0).all()                        # selection order = descending score
    assert keep[0] == np.argmax(s)                             # the best box always survives
    k = b[keep]                                                # no kept pair overlaps above the threshold
    area = (k[:, 2] - k[:, 0]) * (k[:, 3] - k[:, 1])
    iw = np.maximum(0, np.minimum(k[:, None, 2], k[None, :, 2]) - np.maximum(k[:, None, 0], k[None, :, 0]))
    ih = np.maximum(0, np.minimum(k[:, None, 3], k[None, :, 3]) - np.maximum(k[:, None, 1], k[None, :, 1]))
    inter = iw * ih
    union = area[:, None] + area[None, :] - inter
    iou = np.where((area[:, None] > 0) & (area[None, :] > 0) & (union > 0), inter / np.where(union > 0, union, 1), 0)
    np.fill_diagonal(iou, 0)
    assert (iou <= np.float32(thr)).all()
    assert np.array_equal(orc.nms_tf(b[keep], s[keep], n, thr), np.arange(len(keep)))   # idempotent


def test_c_twin_agrees_with_numpy_oracle():
    so = os.path.join(ROOT, 'oracle', 'c', 'libboxpath_ref.so')
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle', 'c')])
    lib = ctypes.CDLL(so)
    img = syn.c4_image(2, 1, channels=16)
    n = img['anchors'].shape[0]
    post, P, c = 300, 7, 16
    o_rois = np.zeros((1, post, 4), np.float32); o_idx = np.zeros((1, post), np.int32)
    o_cnt = np.zeros(1, np.int32); o_feat = np.zeros((post, P, P, c), np.float32)
    means, stds = np.zeros(4, np.float32), np.ones(4, np.float32)
    feat = np.ascontiguousarray(img['feat'][None])
    lib.orc_c4_proposal_roi.argtypes = ([ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 4 +
                                        [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.orc_c4_proposal_roi(p(img['anchors']), p(img['deltas']), p(img['scores']), p(feat), 1, n, 38, 63, c, p(means),
                                 p(stds), 600, 1000, 0, post, 0.7, 16.0, P, 1, p(o_rois), p(o_idx), p(o_cnt), p(o_feat))
    assert rc == 0 and o_cnt[0] == post
    rois, idx = orc.region_proposal(img['deltas'], img['anchors'], img['scores'], (600, 1000), post)
    assert np.array_equal(o_idx[0], idx)
    np.testing.assert_allclose(o_rois[0], rois, rtol=1e-5, atol=1e-3)
    assert np.array_equal(o_feat, orc.roi_pool_c4(feat, o_rois[0], 16, P, True))


def test_c_twin_fpn_composite_agrees_with_numpy_oracle():
    """orc_fpn_proposal_roi (the FPN composite of the oracle's C twin) against the numpy oracle: 2 images, C = 8."""
    so = os.path.join(ROOT, 'oracle', 'c', 'libboxpath_ref.so')
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle', 'c')])
    lib = ctypes.CDLL(so)
    hw, b, post, P, c = (600, 1000), 2, 200, 7, 8
    ims = [syn.fpn_image(3, i, hw, channels=c) for i in range(b)]
    anchors = ims[0]['anchors']; n = anchors.shape[0]
    deltas = np.ascontiguousarray(np.stack([im['deltas'] for im in ims]))
    scores = np.ascontiguousarray(np.stack([im['scores'] for im in ims]))
    feats = [np.ascontiguousarray(np.stack([im['feats'][l] for im in ims])) for l in range(4)]
    fptr = (ctypes.c_void_p * 4)(*[f.ctypes.data for f in feats])
    fh = (ctypes.c_int * 4)(*[f.shape[1] for f in feats]); fw = (ctypes.c_int * 4)(*[f.shape[2] for f in feats])
    means, stds = np.zeros(4, np.float32), np.ones(4, np.float32)
    o_rois = np.zeros((b, post, 4), np.float32); o_idx = np.zeros((b, post), np.int32); o_cnt = np.zeros(b, np.int32)
    o_feat = np.zeros((b * post, P, P, c), np.float32); o_ord = np.zeros(b * post, np.int32)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    rc = lib.orc_fpn_proposal_roi(vp(anchors), vp(deltas), vp(scores), fptr, fh, fw, b, n, c, vp(means), vp(stds), hw[0], hw[1],
                                  0, post, ctypes.c_float(0.7), P, vp(o_rois), vp(o_idx), vp(o_cnt), vp(o_feat), vp(o_ord))
    assert rc == 0 and (o_cnt == post).all()
    for i in range(b):
        _, idx = orc.region_proposal(deltas[i], anchors, scores[i], hw, post)
        assert np.array_equal(o_idx[i], idx)
    rois = o_rois.reshape(-1, 4)
    lv, _, order = orc.assign_levels(rois)
    assert np.array_equal(o_ord, order)
    bi = np.repeat(np.arange(b, dtype=np.int32), post)
    pos = 0
    for l in range(4):
        k = order[lv[order] == l + 2]
        if k.size:
            want = orc.roi_pool_fpn(feats[l], rois[k], hw, P, box_ind=bi[k])
            assert np.array_equal(o_feat[pos:pos + k.size], want)
        pos += k.size
    assert pos == b * post


@pytest.mark.skipif(not os.path.isdir('/root/reference/object_detection'), reason='reference sources only exist in the build container')
def test_golden_vectors_regenerate_from_the_reference(tmp_path):
    """Re-run oracle/make_golden.py (reference files on the numpy TF shim) and compare with the committed vectors."""
    env = dict(os.environ)
    src = open(os.path.join(ROOT, 'oracle', 'make_golden.py')).read().replace("OUT = os.path.join(ROOT, 'tests', 'golden')",
                                                                            "OUT = %r" % str(tmp_path))
    script = tmp_path / 'mk.py'
    script.write_text(src.replace("ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))", "ROOT = %r" % ROOT))
    subprocess.check_call(['python', str(script)], env=env, stdout=subprocess.DEVNULL)
    new = dict(np.load(tmp_path / 'reference_on_shim.npz'))
    old = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_on_shim.npz')))
    assert sorted(new) == sorted(old)
    for k in old:
        assert np.array_equal(new[k], old[k]), k
