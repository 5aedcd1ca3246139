"""Device time (CUDA events, median of 20) of every op of the path at the BASELINE.json config sizes.
Run under gpurun: python profiles/micro/op_times.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import _lib, ops, synthetic as syn

dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)


import time


def t(fn, n=30):
    """(device us per call with n calls queued back to back, host us per call to enqueue)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n, (h1 - h0) * 1e6 / n


rows = []
# ---- cfg2 C4
B = 8
imgs = [syn.c4_image(2, i, with_features=False) for i in range(B)]
anchors = cu(imgs[0]['anchors']); deltas = cu(np.stack([i['deltas'] for i in imgs])); scores = cu(np.stack([i['scores'] for i in imgs]))
feat = torch.randn((B, 38, 63, 1024), device=dev)
rows.append(('cfg2 decode_clip  B=8 N=21546', t(lambda: ops.decode_clip(anchors, deltas, image_shape=(600, 1000))), 8 * 21546 * 48))
rows.append(('cfg2 proposals    B=8 6000->300', t(lambda: ops.proposals(anchors, deltas, scores, (600, 1000), 300, pre_nms_top_k=6000)), 8 * (36 * 21546 + 20 * 300)))
rows.append(('cfg2 proposals    B=8 all->300 (reference behaviour)', t(lambda: ops.proposals(anchors, deltas, scores, (600, 1000), 300)), 8 * (36 * 21546 + 20 * 300)))
rows.append(('cfg2 proposals    B=8 12000->2000 (train)', t(lambda: ops.proposals(anchors, deltas, scores, (600, 1000), 2000, pre_nms_top_k=12000)), 8 * (36 * 21546 + 20 * 2000)))
ob, oi, oc = ops.proposals(anchors, deltas, scores, (600, 1000), 300, pre_nms_top_k=6000)
rows.append(('cfg2 roi_pool 7x7x1024 R=2400 (band kernel)', t(lambda: ops.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, ob, roi_counts=oc)), 8 * (4 * 1024 * 38 * 63 + 16 * 300 + 4 * 300 * 49 * 1024)))
os.environ['BX_ROI_DIRECT'] = '1'
rows.append(('cfg2 roi_pool 7x7x1024 R=2400 (direct kernel)', t(lambda: ops.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, ob, roi_counts=oc)), 8 * (4 * 1024 * 38 * 63 + 16 * 300 + 4 * 300 * 49 * 1024)))
del os.environ['BX_ROI_DIRECT']
# ---- cfg1 VGG16
f512 = torch.randn((1, 38, 63, 512), device=dev)
rows.append(('cfg1 roi_pool 14x14+max C=512 R=300 (band)', t(lambda: ops.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_MAX2, 7, f512, ob[0])), 4 * 512 * 38 * 63 + 16 * 300 + 4 * 300 * 49 * 512))
# ---- cfg3 FPN 600x1000, batch 1 image (reference is per image) and batch 16 proposals
fimg = syn.fpn_image(3, 0, with_features=False)
fa = cu(fimg['anchors']); fd = cu(fimg['deltas'])[None]; fs = cu(fimg['scores'])[None]
N3 = fa.shape[0]
rows.append(('cfg3 proposals    B=1 N=150111 all->1000', t(lambda: ops.proposals(fa, fd, fs, (600, 1000), 1000)), 36 * N3 + 20 * 1000))
fd16 = fd.repeat(16, 1, 1); fs16 = fs.repeat(16, 1)
rows.append(('cfg3 proposals    B=16 N=150111 all->1000', t(lambda: ops.proposals(fa, fd16, fs16, (600, 1000), 1000)), 16 * (36 * N3 + 20 * 1000)))
rb, _, _ = ops.proposals(fa, fd, fs, (600, 1000), 1000)
feats = [torch.randn((1, h, w, 256), device=dev) for (h, w) in syn.fpn_feature_shapes((600, 1000))[:4]]
fb = 4 * 256 * sum(h * w for (h, w) in syn.fpn_feature_shapes((600, 1000))[:4])
rows.append(('cfg3 fpn_roi_features R=1000 C=256 (levels + pool)', t(lambda: ops.fpn_roi_features(feats, rb[0], (600, 1000))), fb + 16 * 1000 + 4 * 1000 * 49 * 256))
B3 = 16
feats16 = [torch.randn((B3, h, w, 256), device=dev) for (h, w) in syn.fpn_feature_shapes((600, 1000))[:4]]
rb16 = rb[0].repeat(B3, 1)
bi16 = torch.arange(B3, device=dev, dtype=torch.int32).repeat_interleave(1000)
rows.append(('cfg3 fpn_roi_features B=16 R=16000 C=256 (direct kernel)', t(lambda: ops.fpn_roi_features(feats16, rb16, (600, 1000), box_ind=bi16), n=10), B3 * (fb + 16 * 1000 + 4 * 1000 * 49 * 256)))
del feats16
rows.append(('cfg3 fpn_assign_levels R=1000', t(lambda: ops.fpn_assign_levels(rb[0])), 1000 * 24))
# ---- cfg5 FPN 800x1333
f5 = syn.fpn_image(5, 0, (800, 1333), with_features=False)
a5 = cu(f5['anchors']); d5 = cu(f5['deltas'])[None].repeat(8, 1, 1); s5 = cu(f5['scores'])[None].repeat(8, 1)
rows.append(('cfg5 proposals    B=8 N=267069 all->1000', t(lambda: ops.proposals(a5, d5, s5, (800, 1333), 1000)), 8 * (36 * a5.shape[0] + 20 * 1000)))
# ---- cfg4 targets, batch 16
rng = np.random.default_rng(1)
gts = np.stack([syn.gt_boxes(rng, 100, (600, 1000))[0] for _ in range(16)])
gls = np.stack([syn.gt_boxes(rng, 100, (600, 1000))[1] for _ in range(16)])
perm = np.stack([rng.permutation(21546) for _ in range(16)]).astype(np.int32)
gtc, glc, permc = cu(gts), cu(gls), cu(perm)
rows.append(('cfg4 pairwise_iou 21546x100', t(lambda: ops.pairwise_iou(anchors, gtc[0])), 16 * (21546 + 100) + 4 * 21546 * 100))
rows.append(('cfg4 anchor_target B=16 N=21546 M=100', t(lambda: ops.anchor_target(anchors, gtc, permc, (600, 1000))), 16 * (16 * (21546 + 100) + 52 * 21546)))
tr, _, tc = ops.proposals(anchors, deltas.repeat(2, 1, 1), scores.repeat(2, 1), (600, 1000), 2000)
permr = cu(np.stack([rng.permutation(2000) for _ in range(16)]).astype(np.int32))
rows.append(('cfg4 proposal_target B=16 K=2000 M=100 S=128', t(lambda: ops.proposal_target(tr, gtc, glc, permr, neg_iou_threshold=0.0, stds=(.1, .1, .2, .2), roi_counts=tc)), 16 * (16 * 2100 + 400 + 128 * (20 + 48 * 21))))
# ---- "next" rows f1-f3
hs, hd = syn.roi_head_outputs(np.random.default_rng(7), 16 * 300, 21)
hs_c, hd_c = cu(hs.reshape(16, 300, 21)), cu(hd.reshape(16, 300, 21, 4))
rois16 = ob.repeat(2, 1, 1)
rows.append(('f1 post_ops_prediction B=16 R=300 C=21', t(lambda: ops.post_ops_prediction(hs_c, hd_c, rois16, (600, 1000), stds=(.1, .1, .2, .2))), 16 * 300 * (21 * 20 + 16)))
lg = torch.randn((8, 38 * 63, 18), device=dev)
rows.append(('f2 proposals_rpn  B=8 raw logits 6000->300 (softmax fused)', t(lambda: ops.proposals_rpn(anchors, deltas, lg, _lib.RPN_CAFFE, 9, (600, 1000), 300, pre_nms_top_k=6000)), 8 * (40 * 21546 + 20 * 300)))
rows.append(('f2 generate_anchors FPN 800x1333 (267069 anchors)', t(lambda: __import__('tf_eager_object_detection_b200.anchor_generator', fromlist=['x']).make_fpn_anchors((800, 1333))), 16 * 267069))
go = torch.randn((2400, 7, 7, 1024), device=dev)
rows.append(('f3 roi_pool_grad 7x7x1024 R=2400 -> [8,38,63,1024]', t(lambda: ops.roi_pool_grad(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, ob.reshape(-1, 4), go, roi_counts=oc), n=10), 4 * 2400 * 49 * 1024 + 8 * 4 * 1024 * 38 * 63))   # grad_out read + grad_feat written once (no memset)
gfp = [torch.randn((1, h, w, 256), device=dev) for (h, w) in syn.fpn_feature_shapes((600, 1000))[:1]]
lab = torch.randint(-1, 2, (16 * 21546,), device=dev).float()
lg2 = torch.randn((16 * 21546, 2), device=dev)
rows.append(('f3 cls_loss + grad 344736 x 2 (RPN, labels -1/0/1)', t(lambda: ops.cls_loss(lg2, lab, with_grad=True)), 16 * 21546 * (8 + 4 + 8)))
pr = torch.randn((16 * 21546, 4), device=dev)
rows.append(('f3 smooth_l1_loss + grad 344736 x 4 (RPN)', t(lambda: ops.smooth_l1_loss(pr, pr * 0.5, torch.ones_like(pr), torch.ones_like(pr), 3.0, (0, 1), with_grad=True)), 16 * 21546 * 16 * 5))
peak = 6538.3
print('%-58s %10s %9s %12s %8s' % ('op', 'us/call', 'host us', 'alg. MB', 'of HBM'))
for name, (us, host), b in rows:
    print('%-58s %10.1f %9.1f %12.2f %7.1f%%%s' % (name, us, host, b / 1e6, 100 * b / (us * 1e-6) / 1e9 / peak,
                                                  '  (host-bound)' if host > 0.9 * us else ''))
