"""Device time of the FPN-size proposal stage (n > 24576: top-set prefilter + cached-key kernel).
usage: [BX_NO_TOPSET=1] python profiles/micro/fpn_props.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for cfg, hw, B in ((3, (600, 1000), 16), (5, (800, 1333), 8), (5, (800, 1333), 64)):
    ims = [syn.fpn_image(cfg, i % 8, hw, with_features=False) for i in range(min(B, 8))]
    a = cu(ims[0]['anchors'])
    d = cu(np.stack([im['deltas'] for im in ims])).repeat(B // len(ims), 1, 1)
    s = cu(np.stack([im['scores'] for im in ims])).repeat(B // len(ims), 1)
    for post, pre in ((1000, 0), (2000, 12000)):
        us = t(lambda: ops.proposals(a, d, s, hw, post, pre_nms_top_k=pre))
        print('cfg%d N=%d B=%d post=%d pre=%d: %.1f us  (%s)' % (cfg, a.shape[0], B, post, pre, us,
              'single-CTA streaming' if os.environ.get('BX_NO_TOPSET') else 'top-set prefilter'), flush=True)
