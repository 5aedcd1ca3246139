"""Synthetic PASCAL-VOC tree for the f4 tests (TEST INFRASTRUCTURE ONLY — part of the oracle): ground truth +
noisy detections in the [B,K,6] record format, shared by oracle/make_golden.py (which feeds it to the reference's
voc_eval) and tests/test_evaluation.py."""
import os

import numpy as np

from tf_eager_object_detection_b200 import synthetic as syn


def synthetic_voc(rng, n_images=40, n_classes=5):
    """Ground truth + noisy detections in the [B,K,6] record format; shared by make_golden and the tests."""
    gts, K = [], 24
    det = np.zeros((n_images, K, 6), np.float32)
    cnt = np.zeros(n_images, np.int32)
    for b in range(n_images):
        m = int(rng.integers(0, 6))
        boxes, labels = syn.gt_boxes(rng, m, (375, 500), n_classes)
        difficult = rng.random(m) < 0.15
        gts.append((boxes.astype(np.int64), labels, difficult))
        rows = []
        for k in range(m):                                       # jittered copies (some duplicated), plus clutter
            for _ in range(int(rng.integers(0, 3))):
                rows.append(np.concatenate([boxes[k] + rng.normal(0, 6, 4), [rng.random(), labels[k]]]))
        for _ in range(int(rng.integers(0, 4))):
            cb, cl = syn.gt_boxes(rng, 1, (375, 500), n_classes)
            rows.append(np.concatenate([cb[0], [rng.random() * 0.6, cl[0]]]))
        rows = sorted(rows, key=lambda r: -r[4])[:K]
        cnt[b] = len(rows)
        if rows:
            det[b, :len(rows)] = np.asarray(rows, np.float32)
    return gts, det, cnt


def write_voc_tree(root, names, gts, class_list):
    os.makedirs(os.path.join(root, 'Annotations'), exist_ok=True)
    with open(os.path.join(root, 'test.txt'), 'w') as f:
        f.write(''.join(n + '\n' for n in names))
    for n, (boxes, labels, difficult) in zip(names, gts):
        objs = ''.join('<object><name>%s</name><pose>Unspecified</pose><truncated>0</truncated><difficult>%d</difficult>'
                       '<bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax><ymax>%d</ymax></bndbox></object>'
                       % (class_list[l], d, b[0], b[1], b[2], b[3]) for b, l, d in zip(boxes, labels, difficult))
        with open(os.path.join(root, 'Annotations', n + '.xml'), 'w') as f:
            f.write('<annotation>%s</annotation>' % objs)


def eval_loop_inputs():
    """Three synthetic images for the evaluation-loop goldens: roi-head outputs over 300 rois in network-input pixels,
    resize factor, raw size.  Image 3 has scores quantised to 1 decimal, so exact ties sit on the per-image cut."""
    rng = np.random.default_rng(syn.seed_for(1, 90))
    out = []
    for i, (scale, (raw_h, raw_w)) in enumerate(((1.6, (375, 500)), (1.25, (480, 640)), (2.0, (300, 500)))):
        hs, hd = syn.roi_head_outputs(rng, 300, 21)
        if i == 2:
            hs = np.round(hs, 1).astype(np.float32)
        rois = syn.random_rois(rng, 300, (int(raw_h * scale), int(raw_w * scale)))
        out.append(dict(scores=hs, deltas=hd, rois=rois, scale=np.float32(scale), raw_h=raw_h, raw_w=raw_w))
    return out
