"""One proposal-stage configuration a few times (ncu target).  usage: python profiles/micro/props_one.py cfg B post pre"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
cfg, B, post, pre = (int(v) for v in sys.argv[1:5])
hw = (800, 1333) if cfg == 5 else (600, 1000)
if cfg in (3, 5):
    ims = [syn.fpn_image(cfg, i % 8, hw, with_features=False) for i in range(min(B, 8))]
else:
    ims = [syn.c4_image(cfg, i % 8, hw, with_features=False) for i in range(min(B, 8))]
a = cu(ims[0]['anchors'])
d = cu(np.stack([im['deltas'] for im in ims])).repeat((B + 7) // 8, 1, 1)[:B].contiguous()
s = cu(np.stack([im['scores'] for im in ims])).repeat((B + 7) // 8, 1)[:B].contiguous()
for _ in range(4):
    ops.proposals(a, d, s, hw, post, pre_nms_top_k=pre)
torch.cuda.synchronize()
