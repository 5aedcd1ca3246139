"""One FPN RoI extractor configuration a few times (ncu target).  usage: python profiles/micro/fpn_roi_one.py [B] [R]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
hw = (600, 1000)
ims = [syn.fpn_image(3, i, hw, with_features=False) for i in range(4)]
a = cu(ims[0]['anchors'])
d = cu(np.stack([im['deltas'] for im in ims])).repeat((B + 3) // 4, 1, 1)[:B]
s = cu(np.stack([im['scores'] for im in ims])).repeat((B + 3) // 4, 1)[:B]
rois, _, _ = ops.proposals(a, d, s, hw, R)
feats = [torch.randn((B, h, w, 256), device=dev) for (h, w) in syn.fpn_feature_shapes(hw)[:4]]
bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(R)
rr = rois.reshape(-1, 4).contiguous()
for _ in range(4):
    ops.fpn_roi_features(feats, rr, hw, box_ind=bi)
torch.cuda.synchronize()
