"""VGG16-shape pooled extractor (14x14 crop + 2x2 max, C=512): roi_pool2 (default) vs band kernel (BX_ROI_BAND_POOLED=1)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import _lib, ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
im = syn.c4_image(1, 0, with_features=False)
a, d, s = cu(im['anchors']), cu(im['deltas'])[None], cu(im['scores'])[None]
rois, _, _ = ops.proposals(a, d, s, (600, 1000), 300)
for B in (1, 8):
    f512 = torch.randn((B, 38, 63, 512), device=dev)
    rr = rois[0].repeat(B, 1)
    bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(300)
    out = torch.empty((B * 300, 7, 7, 512), device=dev)
    def fn():
        ops.roi_pool(_lib.ROI_STRIDE_NORM, _lib.POOL_MAX2, 7, f512, rr, box_ind=bi)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print('VGG 14x14+max C=512 B=%d R=%d: %.1f us (%s)' % (B, B * 300, e0.elapsed_time(e1) / 50 * 1e3, 'band' if os.environ.get('BX_ROI_BAND_POOLED') else 'roi_pool2 (default)'), flush=True)
