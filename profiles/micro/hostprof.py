import os, sys, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import ops, synthetic as syn
dev = torch.device('cuda', 0)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(1)
anchors = cu(syn.c4_anchors(38, 63))
gts = cu(np.stack([syn.gt_boxes(rng, 100, (600, 1000))[0] for _ in range(16)]))
gls = cu(np.stack([syn.gt_boxes(rng, 100, (600, 1000))[1] for _ in range(16)]))
perm = cu(np.stack([rng.permutation(21546) for _ in range(16)]).astype(np.int32))
fn = lambda: ops.anchor_target(anchors, gts, perm, (600, 1000))
for _ in range(20): fn()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): fn()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14); print(s.getvalue()[:3500])
