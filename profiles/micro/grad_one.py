"""One bx_roi_pool_grad call at the cfg2 shape (8 x 300 rois, 7x7x1024) for ncu.  BX_ROI_GRAD_ATOMIC / BX_ROI_GRAD_VEC select."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import _lib, ops, synthetic as syn
dev = torch.device('cuda', 0)
rng = np.random.default_rng(5)
B = 8
rois = torch.as_tensor(np.stack([syn.random_rois(rng, 300, (600, 1000)) for _ in range(B)]).reshape(-1, 4)).to(dev)
feat = torch.randn((B, 38, 63, 1024), device=dev)
go = torch.randn((B * 300, 7, 7, 1024), device=dev)
counts = torch.full((B,), 300, dtype=torch.int32, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = ops.roi_pool_grad(_lib.ROI_STRIDE_NORM, _lib.POOL_NONE, 7, feat, rois, go, stride=16.0, roi_counts=counts)
torch.cuda.synchronize()
print(float(out.abs().sum()))
