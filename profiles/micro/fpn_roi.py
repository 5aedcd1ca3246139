"""Device time of the FPN RoI extractor (level assignment + 14x14 crop + 2x2 max pool, C=256).
usage: [BX_ROI_WHOLE=0|1] python profiles/micro/fpn_roi.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for cfg, hw, B, R in ((3, (600, 1000), 1, 1000), (3, (600, 1000), 16, 512), (3, (600, 1000), 16, 1000),
                      (5, (800, 1333), 8, 1000), (5, (800, 1333), 32, 512)):
    ims = [syn.fpn_image(cfg, i % 4, hw, with_features=False) for i in range(min(B, 4))]
    a = cu(ims[0]['anchors'])
    d = cu(np.stack([im['deltas'] for im in ims])).repeat((B + 3) // 4, 1, 1)[:B]
    s = cu(np.stack([im['scores'] for im in ims])).repeat((B + 3) // 4, 1)[:B]
    rois, _, _ = ops.proposals(a, d, s, hw, R)
    shapes = syn.fpn_feature_shapes(hw)[:4]
    feats = [torch.randn((B, h, w, 256), device=dev) for (h, w) in shapes]
    bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(R)
    rr = rois.reshape(-1, 4).contiguous()
    us = t(lambda: ops.fpn_roi_features(feats, rr, hw, box_ind=bi))
    alg = B * (4 * 256 * sum(h * w for h, w in shapes) + R * (16 + 4 * 49 * 256))
    print('cfg%d B=%d R=%d/img: %.1f us  -> %.2f TB/s algorithmic (%s)' % (cfg, B, R, us, alg / us * 1e-6,
          'BX_ROI_WHOLE=' + os.environ.get('BX_ROI_WHOLE', 'auto')), flush=True)
    del feats
