"""Generate tests/golden/*.npz by executing the REFERENCE's own hot-path files, unmodified, from
/root/reference on the numpy TF shim (oracle/tf_shim).  Runs only in the build container (the GPU box
has no /root/reference); the outputs are committed.  TEST INFRASTRUCTURE.

    python oracle/make_golden.py            # rewrites tests/golden/

What the vectors pin: the reference's Python control flow and op order.  What they do not pin: the
TF kernels themselves (restated in the shim from SURVEY App. B) — "parity unpinned" at that boundary.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf_shim'))
sys.path.insert(0, '/root/reference')

import tensorflow as tf  # noqa: E402  (the shim)
from object_detection.model.anchor_target import AnchorTarget  # noqa: E402
from object_detection.model.proposal_target import ProposalTarget  # noqa: E402
from object_detection.model.region_proposal import RegionProposal  # noqa: E402
from object_detection.model.prediction import post_ops_prediction as ref_post_ops  # noqa: E402
from object_detection.model.roi_pooling import (RoiPoolingCropAndResize, RoiPoolingCropAndResize2,  # noqa: E402
                                                RoiPoolingRoiAlign)
from object_detection.utils import anchor_generator as ref_ag  # noqa: E402
from object_detection.utils import bbox_tf as ref_bbox_tf  # noqa: E402
from object_detection.utils import bbox_transform as ref_bt  # noqa: E402
from object_detection.utils import bbox_np as ref_bbox_np  # noqa: E402

from tf_eager_object_detection_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_proposals(img, training, post_nms_train=2000, post_nms_test=300):
    """RegionProposal.call plus the kept indices (same three calls as region_proposal.py:59-76)."""
    rp = RegionProposal(num_post_nms_train=post_nms_train, num_post_nms_test=post_nms_test)
    rois = rp((img['deltas'], img['anchors'], img['scores'], img['image_shape']), training=training)
    dec = ref_bt.decode_bbox_with_mean_and_std(tf.constant(img['anchors']), tf.constant(img['deltas']),
                                               [0, 0, 0, 0], [1, 1, 1, 1])
    dec, _ = ref_bbox_tf.bboxes_clip_filter(dec, 0, img['image_shape'][0], img['image_shape'][1])
    idx = tf.image.non_max_suppression(tf.to_float(dec), tf.constant(img['scores']),
                                       max_output_size=post_nms_train if training else post_nms_test,
                                       iou_threshold=0.7)
    assert np.array_equal(np.asarray(rois), np.asarray(dec)[np.asarray(idx)])
    return np.asarray(rois), np.asarray(idx), np.asarray(dec)


class _FpnStub:
    """Borrow BaseFPN._assign_levels/_get_roi_features without building the Keras model."""
    _min_level, _max_level = 2, 5
    _level_name_list = ['p2', 'p3', 'p4', 'p5', 'p6']
    _anchor_stride_list = [4, 8, 16, 32, 64]


def load_fpn_methods():
    # base_fpn_model imports the whole model zoo; pull the two methods out of its source instead of
    # importing Keras layers the shim does not provide.
    import ast
    import textwrap
    path = '/root/reference/object_detection/model/fpn/base_fpn_model.py'
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'tf': tf}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ('_assign_levels', '_get_roi_features'):
            code = textwrap.dedent(ast.get_source_segment(src, node))
            exec(compile(code, path, 'exec'), ns)
    _FpnStub._assign_levels = ns['_assign_levels']
    _FpnStub._get_roi_features = ns['_get_roi_features']
    stub = _FpnStub()
    stub._roi_pooling = RoiPoolingCropAndResize2(pool_size=7)
    return stub


def voc_golden():
    import tempfile
    if not hasattr(np, 'bool'):
        np.bool = bool                      # the reference predates numpy 1.24 (detectron_pascal_evaluation_utils.py:150)
    from object_detection.evaluation.detectron_pascal_evaluation_utils import voc_eval as ref_voc_eval
    from tf_eager_object_detection_b200 import evaluation as ev
    from oracle.voc_fixture import synthetic_voc, write_voc_tree
    classes = ev.PASCAL_CLASSES[:5]
    gts, det, cnt = synthetic_voc(np.random.default_rng(syn.seed_for(1, 80)))
    names = ['%06d' % (i + 1) for i in range(len(gts))]
    out = {}
    with tempfile.TemporaryDirectory() as root:
        write_voc_tree(root, names, gts, classes)
        ev.write_voc_results(os.path.join(root, 'det_{:s}.txt'), names, det, cnt, classes)
        for metric07 in (True, False):
            for c in classes[1:]:
                rec, prec, ap = ref_voc_eval(os.path.join(root, 'det_{:s}.txt'), os.path.join(root, 'Annotations', '{:s}.xml'),
                                             os.path.join(root, 'test.txt'), c, os.path.join(root, 'cache'), 0.5, metric07)
                tag = 'voc_%s_%s' % (c, '07' if metric07 else 'area')
                out[tag + '_rec'], out[tag + '_prec'], out[tag + '_ap'] = rec, prec, np.float64(ap)
        out['voc_det_sha'] = np.frombuffer(hashlib.sha256(
            b''.join(open(os.path.join(root, 'det_%s.txt' % c), 'rb').read() for c in classes[1:])).digest(), np.uint8)
    return out


def eval_loop_golden():
    """Runs the reference's own `get_prediction_files` (evaluation/pascal_eval_files_utils.py:19-122), UNMODIFIED, on the
    numpy TF shim: a stub model returns the synthetic roi-head outputs from `im_detect` (with its final `rois / img_scale`,
    base_faster_rcnn_model.py:304), a stub dataset module yields (img, img_scale, raw_h, raw_w).  The golden is the text of
    the 20 VOC result files it writes.  numpy >= 1.25 raises on the reference's `dets == []` (:117) where the numpy of its
    day returned False: the module's `np.array` is wrapped to restore that behaviour."""
    import tempfile
    import types
    from oracle.voc_fixture import eval_loop_inputs
    imgs = eval_loop_inputs()
    fake = types.ModuleType('object_detection.dataset.eval_pascal_tf_dataset')
    names = ['%06d' % (i + 1) for i in range(len(imgs))]
    fake.get_dataset_by_tf_records = lambda *a, **k: ([(None, im['scale'], im['raw_h'], im['raw_w']) for im in imgs], names)
    fake.get_dataset_by_local_file = fake.get_dataset_by_tf_records
    sys.modules['object_detection.dataset.eval_pascal_tf_dataset'] = fake
    import object_detection.evaluation.pascal_eval_files_utils as ref_eval

    class _OldEq(np.ndarray):
        def __eq__(self, other):
            if isinstance(other, list) and len(other) == 0 and self.ndim == 2:
                return False                     # "elementwise comparison failed; returning scalar" of numpy < 1.25
            return np.ndarray.__eq__(self, other)
        __hash__ = None

    class _NpProxy:
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def array(*a, **k):
            return np.array(*a, **k).view(_OldEq)
    ref_eval.np = _NpProxy()
    ref_eval.tqdm = lambda it: it

    class _Model:
        def __init__(self):
            self.i = 0

        def im_detect(self, img, img_scale):
            im = imgs[self.i]
            self.i += 1
            return (tf.constant(im['scores']), tf.constant(im['deltas'].reshape(300, -1)),
                    tf.constant(im['rois']) / tf.to_float(img_scale))
    out = {}
    for tag, max_img in (('eval_voc', 50), ('eval_voc_nocut', 0)):
        with tempfile.TemporaryDirectory() as d:
            ref_eval.get_prediction_files(_Model(), dataset_type='tf', result_file_format=os.path.join(d, '{:s}.txt'),
                                          score_threshold=0.05, iou_threshold=0.3, max_objects_per_class=50,
                                          max_objects_per_image=max_img, min_size=10)
            text = b''.join(open(os.path.join(d, '%s.txt' % c), 'rb').read() for c in ref_eval.class_list[1:])
        out[tag + '_files'] = np.frombuffer(text, np.uint8).copy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    g = {}

    # ---- anchors (inputs to the path; reference generators)
    base = ref_ag.generate_anchor_base(16, [.5, 1, 2], [8, 16, 32])
    c4_anchors = np.asarray(ref_ag.generate_by_anchor_base_tf(base, 16, 38, 63))
    g['anchor_base'] = base.astype(np.float32)
    g['c4_anchors_sha'] = np.frombuffer(bytes.fromhex(sha(c4_anchors)), dtype=np.uint8)
    assert np.array_equal(c4_anchors, syn.c4_anchors(38, 63, 16)), 'synthetic C4 anchors != reference'
    fpn_anchor_list = []
    for b, (h, w), s in zip(syn.FPN_BASE_SIZES, syn.fpn_feature_shapes((600, 1000)), syn.FPN_STRIDES):
        fpn_anchor_list.append(np.asarray(ref_ag.make_anchors(b, [1.], [.5, 1, 2], h, w, s)))
    fpn_anchors = np.concatenate(fpn_anchor_list, 0)
    assert np.array_equal(fpn_anchors, syn.fpn_anchors((600, 1000))), 'synthetic FPN anchors != reference'
    g['fpn_anchors_sha'] = np.frombuffer(bytes.fromhex(sha(fpn_anchors)), dtype=np.uint8)
    g['c4_inside_count'] = np.asarray(ref_bbox_tf.bboxes_range_filter(tf.constant(c4_anchors), 600, 1000).shape[0])

    # ---- a1-a3: C4 proposals (cfg 1 seeds), eval (300) and training (2000)
    img = syn.c4_image(1, 0, channels=8)
    rois_e, idx_e, dec = ref_proposals(img, training=False)
    rois_t, idx_t, _ = ref_proposals(img, training=True)
    g['c4_decoded_clipped_head'] = dec[:2048]
    g['c4_decoded_clipped_sha'] = np.frombuffer(bytes.fromhex(sha(dec)), dtype=np.uint8)
    g['c4_eval_rois'], g['c4_eval_idx'] = rois_e, idx_e
    g['c4_train_rois'], g['c4_train_idx'] = rois_t, idx_t

    # ---- a4: C4 RoI pooling, both flags, on the first 64 eval rois, C=8
    feat = img['feat'][None]
    g['c4_pool_nomax'] = np.asarray(RoiPoolingCropAndResize(7, False)((feat, rois_e[:64], 16)))
    g['c4_pool_max'] = np.asarray(RoiPoolingCropAndResize(7, True)((feat, rois_e[:64], 16)))
    # ---- a8: dormant RoIAlign variant
    g['c4_roialign'] = np.asarray(RoiPoolingRoiAlign(7)((feat, rois_e[:64], 16)))

    # ---- FPN: global proposals over P2-P6 (cfg 3 seeds), level assignment, per-level pooling
    fimg = syn.fpn_image(3, 0, channels=8)
    rois_f, idx_f, _ = ref_proposals(fimg, training=False, post_nms_test=1000)
    g['fpn_eval_rois'], g['fpn_eval_idx'] = rois_f, idx_f
    stub = load_fpn_methods()
    rois_list, order = stub._assign_levels(tf.constant(rois_f))
    g['fpn_level_counts'] = np.asarray([r.shape[0] for r in rois_list], dtype=np.int32)
    g['fpn_level_order'] = np.asarray(order)
    p_list = [f[None] for f in fimg['feats']] + [np.zeros((1, 10, 16, 8), np.float32)]
    g['fpn_roi_features_head'] = np.asarray(stub._get_roi_features(rois_list, p_list, fimg['image_shape']))[:128]
    # standalone rois covering all levels + border extrapolation
    rr = syn.random_rois(np.random.default_rng(syn.seed_for(3, 50)), 256, (600, 1000))
    rl, ro = stub._assign_levels(tf.constant(rr))
    g['rand_level_counts'] = np.asarray([r.shape[0] for r in rl], dtype=np.int32)
    g['rand_level_order'] = np.asarray(ro)
    g['rand_roi_features'] = np.asarray(stub._get_roi_features(rl, p_list, fimg['image_shape']))

    # ---- a9: pairwise IoU (tf version and the reference's importable numpy twin)
    rng = np.random.default_rng(syn.seed_for(4, 0))
    gt, gl = syn.gt_boxes(rng, 100, (600, 1000))
    iou_tf = np.asarray(ref_bbox_tf.pairwise_iou(tf.constant(c4_anchors[:4096]), tf.constant(gt)))
    iou_np = ref_bbox_np.pairwise_iou(c4_anchors[:4096], gt)
    assert np.array_equal(iou_tf, iou_np.astype(np.float32))
    g['iou_anchors4096_gt100'] = iou_tf

    # ---- a10: anchor target with the injected permutation
    perm = rng.permutation(c4_anchors.shape[0])
    inside = np.asarray(ref_bbox_tf.bboxes_range_filter(tf.constant(c4_anchors), 600, 1000))
    with tf.shuffle_hook(lambda idx: idx[np.argsort(perm[inside[idx]], kind='stable')]):
        lab, tg, iw, ow = AnchorTarget()((gt, [600, 1000], c4_anchors))
    g['at_labels'], g['at_targets'], g['at_in_w'], g['at_out_w'] = map(np.asarray, (lab, tg, iw, ow))
    # few-gt case exercises the "no subsampling of fg" branch
    with tf.shuffle_hook(lambda idx: idx[np.argsort(perm[inside[idx]], kind='stable')]):
        lab2, tg2, iw2, ow2 = AnchorTarget()((gt[:3], [600, 1000], c4_anchors))
    g['at3_labels'], g['at3_targets'], g['at3_in_w'], g['at3_out_w'] = map(np.asarray, (lab2, tg2, iw2, ow2))

    # ---- a11: proposal target on the 2000 training proposals
    perm_r = rng.permutation(rois_t.shape[0])
    cycle = {}

    def fake_choice(a, size, replace=True):
        sb = a[np.argsort(perm_r[a], kind='stable')]
        return sb[np.arange(size) % sb.size]
    real_choice = np.random.choice
    np.random.choice = fake_choice
    try:
        for name, kw in (('pt', dict(num_classes=21, pos_iou_threshold=0.5, neg_iou_threshold=0.0,
                                     total_num_samples=128, max_pos_samples=32,
                                     target_stds=[0.1, 0.1, 0.2, 0.2])),
                         ('pt_fpn', dict(num_classes=21, pos_iou_threshold=0.5, neg_iou_threshold=0.0,
                                         total_num_samples=256, max_pos_samples=64,
                                         target_stds=[0.1, 0.1, 0.2, 0.2])),
                         ('pt_pad', dict(num_classes=21, pos_iou_threshold=0.5, neg_iou_threshold=0.1,
                                         total_num_samples=2048, max_pos_samples=512,
                                         target_stds=[0.1, 0.1, 0.2, 0.2]))):
            with tf.shuffle_hook(lambda idx: idx[np.argsort(perm_r[idx], kind='stable')]):
                out = ProposalTarget(**kw)((rois_t, gt, gl))
            for k, v in zip(('rois', 'labels', 'targets', 'in_w', 'out_w'), out):
                g['%s_%s' % (name, k)] = np.asarray(v)
    finally:
        np.random.choice = real_choice
    del cycle

    # ---- f1: post-head detection filtering on synthetic roi-head outputs over the 300 eval rois
    from tf_eager_object_detection_b200.synthetic import roi_head_outputs
    hs, hd = roi_head_outputs(np.random.default_rng(syn.seed_for(1, 77)), rois_e.shape[0], 21)
    pb, pc, ps = ref_post_ops(tf.constant(hs), tf.constant(hd), tf.constant(rois_e), [600, 1000], [0, 0, 0, 0],
                              [0.1, 0.1, 0.2, 0.2], max_num_per_class=50, max_num_per_image=150, nms_iou_threshold=0.3,
                              score_threshold=0.05, extractor_stride=16, num_classes=21)
    g['post_boxes'], g['post_classes'], g['post_scores'] = np.asarray(pb), np.asarray(pc), np.asarray(ps)

    # ---- f3: the reference's losses.py on the RPN / RoI-head shapes it is called with
    from object_detection.model.losses import cls_loss as ref_cls_loss, smooth_l1_loss as ref_sl1
    lrng = np.random.default_rng(syn.seed_for(1, 78))
    n_rpn = g['at_labels'].shape[0]
    rpn_pred = (lrng.normal(0, 1, (n_rpn, 4)) * 0.5).astype(np.float32)
    g['loss_rpn_pred'] = rpn_pred
    g['loss_rpn_reg'] = np.asarray(ref_sl1(tf.constant(rpn_pred), tf.constant(g['at_targets']), tf.constant(g['at_in_w']),
                                           tf.constant(g['at_out_w']), 3.0, dim=[0, 1]))
    rpn_logits = lrng.normal(0, 2, (n_rpn, 2)).astype(np.float32)
    sel = np.nonzero(g['at_labels'] >= 0)[0]                     # base_faster_rcnn_model.py:204-207
    g['loss_rpn_logits'] = rpn_logits
    g['loss_rpn_cls'] = np.asarray(ref_cls_loss(tf.constant(rpn_logits[sel]), tf.constant(g['at_labels'][sel])))
    n_roi = g['pt_labels'].shape[0]
    roi_pred = lrng.normal(0, 1, (n_roi, 84)).astype(np.float32)
    roi_logits = lrng.normal(0, 2, (n_roi, 21)).astype(np.float32)
    g['loss_roi_pred'], g['loss_roi_logits'] = roi_pred, roi_logits
    g['loss_roi_reg'] = np.asarray(ref_sl1(tf.constant(roi_pred), tf.constant(g['pt_targets']), tf.constant(g['pt_in_w']),
                                           tf.constant(g['pt_out_w']), sigma=1.0))
    g['loss_roi_cls'] = np.asarray(ref_cls_loss(tf.constant(roi_logits), tf.constant(g['pt_labels'])))

    # ---- f2: the RPN score layout dances, executed from the reference's own statements
    import ast
    import types as _types
    srng = np.random.default_rng(syn.seed_for(1, 79))
    rpn_score = srng.normal(0, 3, (38 * 63, 18)).astype(np.float32)          # RpnHead output [cells, 2A], :345
    path = '/root/reference/object_detection/model/faster_rcnn/base_faster_rcnn_model.py'
    src = open(path).read()
    stmts = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.Assign) and 149 <= n.lineno <= 152
             and getattr(n.targets[0], 'id', '') == 'scores']
    assert len(stmts) == 4, 'expected the 4 `scores = ...` statements at base_faster_rcnn_model.py:149-152'
    ns = {'tf': tf, 'rpn_score': tf.constant(rpn_score), 'self': _types.SimpleNamespace(_num_anchors=9)}
    for st_ in sorted(stmts, key=lambda n: n.lineno):
        exec(compile(ast.get_source_segment(src, st_), path, 'exec'), ns)
    g['rpn_caffe_logits'], g['rpn_caffe_scores'] = rpn_score, np.asarray(ns['scores'])
    fpn_score = srng.normal(0, 3, (4096, 2)).astype(np.float32)
    g['rpn_pairs_logits'] = fpn_score
    g['rpn_pairs_scores'] = np.asarray(tf.nn.softmax(tf.constant(fpn_score))[:, 1])     # base_fpn_model.py:223

    # ---- f4: the reference's voc_eval on a synthetic VOC tree; detection files written by the package's writer
    g.update(voc_golden())

    # ---- f1, evaluation-loop form: the reference's get_prediction_files run unmodified (VOC result files as text)
    g.update(eval_loop_golden())

    np.savez_compressed(os.path.join(OUT, 'reference_on_shim.npz'), **g)
    sz = os.path.getsize(os.path.join(OUT, 'reference_on_shim.npz'))
    print('wrote %d arrays, %.1f KiB' % (len(g), sz / 1024))
    for k in sorted(g):
        print('  %-28s %-18s %s' % (k, g[k].shape, g[k].dtype))


if __name__ == '__main__':
    main()
