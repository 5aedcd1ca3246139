"""Per-CTA timeline of the RoI band kernel (BX_BAND_DEBUG=1): how long CTAs wait for TMA, how long they compute,
how busy each SM is.  Run under gpurun: BX_BAND_DEBUG=1 python profiles/micro/band_timeline.py"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ['BX_BAND_DEBUG'] = '1'
from tf_eager_object_detection_b200 import _lib, ops, synthetic as syn
import bench

w = bench.WORKLOAD
hb = bench.make_batch(w, 0)
dev = torch.device('cuda', 0)
anchors = torch.as_tensor(hb['anchors']).to(dev); deltas = torch.as_tensor(hb['deltas']).to(dev)
scores = torch.as_tensor(hb['scores']).to(dev); feat = torch.as_tensor(hb['feat']).to(dev)
for _ in range(3):
    out = ops.c4_proposal_roi(anchors, deltas, scores, feat, w['image_hw'], 300, pre_nms_top_k=6000)
torch.cuda.synchronize()
lib = _lib.load()
lib.bx_debug_band_dump.restype = ctypes.c_longlong
lib.bx_debug_band_dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
buf = np.zeros((8192, 4), np.uint64); info = np.zeros(4, np.int32)
n = lib.bx_debug_band_dump(list(_lib._handles.values())[0], buf.ctypes.data, 8192, info.ctypes.data)
t = buf[:n].astype(np.int64)
t0 = t[:, 0].min()
start, ready, done, sm = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, t[:, 3]
print('ctas %d  bands %d slices %d threads %d smem %d' % (n, info[0], info[1], info[2], info[3]))
print('kernel span %.1f us' % done.max())
print('wait  (start->band ready): mean %.2f  p50 %.2f  p90 %.2f  max %.2f us' % ((ready - start).mean(), np.median(ready - start), np.percentile(ready - start, 90), (ready - start).max()))
print('work  (band ready->done) : mean %.2f  p50 %.2f  p90 %.2f  max %.2f us' % ((done - ready).mean(), np.median(done - ready), np.percentile(done - ready, 90), (done - ready).max()))
print('sum of CTA lifetimes / (SMs * span) = %.2f' % ((done - start).sum() / (148 * done.max())))
per_sm = {}
for s_, a_, b_ in zip(sm, start, done):
    per_sm.setdefault(int(s_), []).append((a_, b_))
busy = [sum(b - a for a, b in v) for v in per_sm.values()]
last = [max(b for a, b in v) for v in per_sm.values()]
cnt = [len(v) for v in per_sm.values()]
print('SMs used %d; CTAs per SM min/max %d/%d; SM finish time p10 %.1f p50 %.1f max %.1f us' % (len(per_sm), min(cnt), max(cnt), np.percentile(last, 10), np.median(last), max(last)))
band = np.arange(n) // (n // info[0])
for b in range(info[0]):
    m = band == b
    print('  band %d: work mean %.2f us' % (b, (done - ready)[m].mean()))
