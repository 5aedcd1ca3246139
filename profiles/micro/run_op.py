"""Run one op of the path a few times (for ncu captures).  usage: python profiles/micro/run_op.py {fpn_props|fpn5_props|train_props|vgg_pool|anchor_target}"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import _lib, ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
which = sys.argv[1]
if which == 'fpn_props':
    f = syn.fpn_image(3, 0, with_features=False)
    a, d, s = cu(f['anchors']), cu(f['deltas'])[None], cu(f['scores'])[None]
    fn = lambda: ops.proposals(a, d, s, (600, 1000), 1000)
elif which == 'fpn5_props':
    ims = [syn.fpn_image(5, i, (800, 1333), with_features=False) for i in range(8)]
    a = cu(ims[0]['anchors'])
    d = cu(np.stack([im['deltas'] for im in ims])); s = cu(np.stack([im['scores'] for im in ims]))
    fn = lambda: ops.proposals(a, d, s, (800, 1333), 1000)
elif which == 'train_props':
    im = syn.c4_image(2, 0, with_features=False)
    a, d, s = cu(im['anchors']), cu(im['deltas'])[None], cu(im['scores'])[None]
    fn = lambda: ops.proposals(a, d, s, (600, 1000), 2000, pre_nms_top_k=12000)
elif which == 'vgg_pool':
    im = syn.c4_image(1, 0, with_features=False)
    a, d, s = cu(im['anchors']), cu(im['deltas'])[None], cu(im['scores'])[None]
    rois, _, _ = ops.proposals(a, d, s, (600, 1000), 300)
    feat = torch.randn((1, 38, 63, 512), device=dev)
    fn = lambda: ops.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_MAX2, 7, feat, rois[0])
elif which == 'anchor_target':
    B = 16
    rng = np.random.default_rng(syn.seed_for(4, 1))
    anc = syn.c4_anchors(38, 63)
    gts = np.stack([syn.gt_boxes(rng, 100, (600, 1000))[0] for _ in range(B)])
    perm = np.stack([rng.permutation(anc.shape[0]).astype(np.int32) for _ in range(B)])
    a_, g_, p_ = cu(anc), cu(gts), cu(perm)
    fn = lambda: ops.anchor_target(a_, g_, p_, (600, 1000))
else:
    raise SystemExit('unknown op')
for _ in range(5):
    fn()
torch.cuda.synchronize()
