"""A/B of the two backward kernels of the RoI extractors (bx_roi_pool_grad): scatter form with float4 atomics
(BX_ROI_GRAD_ATOMIC=1) against the row-owned, atomic-free form (default), over BX_ROI_GRAD_VEC = channels per lane.
Run under gpurun: python profiles/micro/grad_ab.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_eager_object_detection_b200 import _lib, ops, synthetic as syn

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
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


rng = np.random.default_rng(5)
cases = []
# cfg2 shape: 8 images x 300 rois, 7x7 crop of [38,63,1024]
B = 8
rois = np.stack([syn.random_rois(rng, 300, (600, 1000)) for _ in range(B)])
cases.append(('cfg2 NONE 7x7x1024 R=2400', dict(mode=_lib.ROI_STRIDE_NORM, pool=_lib.POOL_NONE, feat=(B, 38, 63, 1024), rois=rois, stride=16.0, img=(0, 0))))
cases.append(('VGG MAX2 14x14->7x7x512 R=2400', dict(mode=_lib.ROI_STRIDE_NORM, pool=_lib.POOL_MAX2, feat=(B, 38, 63, 512), rois=rois, stride=16.0, img=(0, 0))))
cases.append(('RoIAlign AVG2 7x7x1024 R=2400', dict(mode=_lib.ROI_ALIGN_PAD, pool=_lib.POOL_AVG2, feat=(B, 38, 63, 1024), rois=rois, stride=16.0, img=(0, 0))))
# FPN P2 / P4 of 600x1000, 16 images: rois of the size the level assignment sends there
small = np.stack([syn.random_rois(rng, 400, (600, 1000)) for _ in range(16)])
ctr = (small[..., :2] + small[..., 2:]) / 2
half = rng.uniform(20, 56, size=small.shape[:2] + (1,)).astype(np.float32)
small = np.concatenate([np.maximum(ctr - half, 0), np.minimum(ctr + half, [999, 599])], -1).astype(np.float32)
cases.append(('FPN P2 MAX2 7x7x256 R=6400 [16,150,250]', dict(mode=_lib.ROI_IMAGE_NORM, pool=_lib.POOL_MAX2, feat=(16, 150, 250, 256), rois=small, stride=4.0, img=(600, 1000))))
cases.append(('train B=1 NONE 7x7x1024 R=256', dict(mode=_lib.ROI_STRIDE_NORM, pool=_lib.POOL_NONE, feat=(1, 38, 63, 1024), rois=rois[:1, :256], stride=16.0, img=(0, 0))))

cfgs = [None] + [1, 2, 4]
if len(sys.argv) > 1:
    cfgs = [None] + [int(a) for a in sys.argv[1:]]
print('%-44s %10s  %s' % ('case', 'atomic us', '  '.join('vec %d' % c for c in cfgs[1:])))
for name, c in cases:
    b, fh, fw, ch = c['feat']
    feat = torch.randn(c['feat'], device=dev)
    rr = cu(c['rois'].reshape(-1, 4))
    n_per = c['rois'].shape[1]
    counts = cu(np.full((b,), n_per, np.int32))
    go = torch.randn((b * n_per, 7, 7, ch), device=dev)
    call = lambda: ops.roi_pool_grad(c['mode'], c['pool'], 7, feat, rr, go, stride=c['stride'], image_shape=c['img'], roi_counts=counts)
    os.environ['BX_ROI_GRAD_ATOMIC'] = '1'
    ref = call().clone()
    us_atomic = t(call)
    os.environ['BX_ROI_GRAD_ATOMIC'] = '0'
    line = []
    for cfg in cfgs[1:]:
        os.environ['BX_ROI_GRAD_VEC'] = '%d' % cfg
        a1 = call().clone()
        a2 = call().clone()
        err = float((a1 - ref).abs().max() / ref.abs().max())
        ok = bool(torch.equal(a1, a2)) and err < 1e-5
        line.append('%7.1f%s' % (t(call), '' if ok else '!(%.1e)' % err))
    alg = go.numel() * 4 + feat.numel() * 4
    print('%-44s %10.1f  %s   alg %.0f MB' % (name, us_atomic, '  '.join(line), alg / 1e6))
