"""Mirror of the reference's `object_detection/utils/anchor_generator.py` (same names, argument meaning) — "next" row
f2.  The small per-anchor tables (anchor base, ws/hs) are host numpy exactly as in the reference; the per-cell
enumeration runs on the device (`bx_generate_anchors`)."""
import math

import numpy as np

from . import ops

__all__ = ['generate_anchor_base', 'generate_by_anchor_base_tf', 'make_anchors', 'make_fpn_anchors']


def generate_anchor_base(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """utils/anchor_generator.py:63-134 (py-faster-rcnn table): [len(ratios)*len(scales), 4] float64, ratio-major."""
    out = []
    w0 = h0 = float(base_size)
    ctr = 0.5 * (base_size - 1)
    for r in ratios:
        w = np.round(np.sqrt(w0 * h0 / r))
        h = np.round(w * r)
        for s in scales:
            ws, hs = w * s, h * s
            out.append([ctr - 0.5 * (ws - 1), ctr - 0.5 * (hs - 1), ctr + 0.5 * (ws - 1), ctr + 0.5 * (hs - 1)])
    return np.asarray(out, dtype=np.float64)


_CACHE = {}


def _cached(key, make):
    """The reference regenerates the anchors on every call; they only depend on the arguments, so the device tensor is
    built once per (arguments, device) and handed out again (callers treat anchors as read-only, as the reference does)."""
    import torch
    key = key + (torch.cuda.current_device(),)
    out = _CACHE.get(key)
    if out is None:
        if len(_CACHE) >= 64:
            _CACHE.clear()
        out = _CACHE[key] = make()
    return out


def generate_by_anchor_base_tf(anchor_base, feat_stride, height, width, device=None):
    """utils/anchor_generator.py:46-60: [height*width*A, 4] (x1,y1,x2,y2), cell-major, anchors fastest."""
    base = np.asarray(anchor_base).astype(np.float32)                  # tf.to_float(anchor_base), :57
    key = ('base', base.tobytes(), float(feat_stride), int(height), int(width), str(device))
    return _cached(key, lambda: ops.generate_anchors([(int(height), int(width))], [float(feat_stride)], base[None], device))


def _ratio_tables(base_anchor_size, anchor_scales, anchor_ratios):
    """enum_scales / enum_ratios (:165-178), including the swapped (hs, ws) return at :178."""
    f = np.float32
    size = f(base_anchor_size) * np.asarray(anchor_scales, f)          # [S]
    sq = np.sqrt(np.asarray(anchor_ratios, f))                         # [R]
    ws = (size[None, :] * sq[:, None]).reshape(-1).astype(f)           # what make_anchors calls `ws` (= hs of :177)
    hs = (size[None, :] / sq[:, None]).reshape(-1).astype(f)
    return ws, hs


def _offsets(base_anchor_size, anchor_scales, anchor_ratios):
    ws, hs = _ratio_tables(base_anchor_size, anchor_scales, anchor_ratios)
    hw, hh = np.float32(0.5) * ws, np.float32(0.5) * hs                # `0.5 * box_sizes`, :159-160
    return np.stack([-hw, -hh, hw, hh], axis=1)


def make_anchors(base_anchor_size, anchor_scales, anchor_ratios, featuremap_height, featuremap_width, stride,
                 name='make_anchors', device=None):
    """utils/anchor_generator.py:137-162: one FPN level, centres at (j*stride, i*stride)."""
    key = ('level', float(base_anchor_size), tuple(anchor_scales), tuple(anchor_ratios), int(featuremap_height),
           int(featuremap_width), float(stride), str(device))

    def make():
        off = _offsets(base_anchor_size, anchor_scales, anchor_ratios)
        return ops.generate_anchors([(int(featuremap_height), int(featuremap_width))], [float(stride)], off[None], device)
    return _cached(key, make)


def make_fpn_anchors(image_shape, base_anchor_size_list=(32, 64, 128, 256, 512), anchor_stride_list=(4, 8, 16, 32, 64),
                     anchor_scales=(1.0,), anchor_ratios=(0.5, 1.0, 2.0), device=None):
    """fpn/base_fpn_model.py:163-186 `_get_anchors`: P2..P6 concatenated — one launch for all levels."""
    key = ('fpn', int(image_shape[0]), int(image_shape[1]), tuple(base_anchor_size_list), tuple(anchor_stride_list),
           tuple(anchor_scales), tuple(anchor_ratios), str(device))

    def make():
        shapes = [(math.ceil(image_shape[0] / s), math.ceil(image_shape[1] / s)) for s in anchor_stride_list]
        off = np.stack([_offsets(b, anchor_scales, anchor_ratios) for b in base_anchor_size_list])
        return ops.generate_anchors(shapes, [float(s) for s in anchor_stride_list], off, device)
    return _cached(key, make)
