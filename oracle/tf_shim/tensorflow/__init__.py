"""numpy-backed mini-`tensorflow` (TEST INFRASTRUCTURE ONLY — part of the oracle).

Purpose: let the reference's own hot-path files
(`object_detection/model/{region_proposal,roi_pooling,anchor_target,proposal_target}.py`,
`object_detection/utils/{bbox_tf,bbox_transform,anchor_generator}.py`,
`object_detection/model/fpn/base_fpn_model.py:_assign_levels/_get_roi_features`)
execute UNMODIFIED in the build container, where TensorFlow is not installed, so that
`oracle/make_golden.py` can produce golden vectors that keep the reference's Python
control flow (including its quirks).  The TF *kernels* behind the symbols are restated
here from the TF r1.13 semantics recorded in SURVEY.md Appendix B
(`tf.image.non_max_suppression`, `tf.image.crop_and_resize`, `MaxPooling2D`,
`tf.nn.avg_pool`, `tf.pad(SYMMETRIC)`, `tf.where`, `tf.argmax`, `tf.nn.top_k`).

PARITY UNPINNED at the TF-kernel boundary: the reference ships no tests / golden vectors
and TensorFlow cannot be imported here, so these restatements are cross-witnessed only
(torchvision NMS, `bbox_np.pairwise_iou`, `grid_sample`) — see DESIGN.md.

Never imported by the product package; only `oracle/make_golden.py` and tests put this
directory on `sys.path`.
"""
import contextlib
import types

import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_  # noqa: A001  (tf.bool)
newaxis = None


class EagerTensor(np.ndarray):
    """ndarray with the handful of tf.Tensor methods the reference touches."""

    def numpy(self):
        return np.asarray(self)

    def get_shape(self):
        return _Shape(self.shape)

    def __hash__(self):
        return id(self)


class _Shape(tuple):
    def as_list(self):
        return list(self)


def _t(x, dtype=None):
    """convert_to_tensor: python floats -> float32, python ints -> int32 (TF defaults)."""
    if isinstance(x, EagerTensor) and dtype is None:
        return x
    if dtype is None:
        if isinstance(x, (builtins_float,)):
            dtype = np.float32
        elif isinstance(x, builtins_bool):
            dtype = np.bool_
        elif isinstance(x, builtins_int):
            dtype = np.int32
        elif isinstance(x, (list, tuple)):
            a = np.asarray(x)
            if a.dtype == np.float64:
                dtype = np.float32
            elif a.dtype == np.int64 and not any(isinstance(v, np.ndarray) for v in x):
                dtype = np.int32
    a = np.asarray(x, dtype=dtype)
    return a.view(EagerTensor)


import builtins as _b  # noqa: E402

builtins_float, builtins_int, builtins_bool = _b.float, _b.int, _b.bool


# ---------------------------------------------------------------- basic constructors
def constant(value, dtype=None, shape=None):
    a = _t(value, dtype)
    if shape is not None:
        a = np.broadcast_to(a.reshape(-1) if a.size == int(np.prod(shape)) else a, shape) \
            if a.size != int(np.prod(shape)) else a.reshape(shape)
        a = np.array(a).view(EagerTensor)
    return a


def cast(x, dtype):
    return np.asarray(x).astype(dtype).view(EagerTensor)


def to_float(x):
    return cast(x, np.float32)


def to_int32(x):
    return cast(x, np.int32)


def zeros(shape, dtype=np.float32):
    return np.zeros(_shape_arg(shape), dtype=dtype).view(EagerTensor)


def ones(shape, dtype=np.float32):
    return np.ones(_shape_arg(shape), dtype=dtype).view(EagerTensor)


def _shape_arg(shape):
    if isinstance(shape, (list, tuple)):
        return tuple(builtins_int(np.asarray(s)) for s in shape)
    return tuple(builtins_int(s) for s in np.asarray(shape).reshape(-1))


def zeros_like(x, dtype=None):
    return np.zeros_like(np.asarray(x), dtype=dtype).view(EagerTensor)


def ones_like(x, dtype=None):
    return np.ones_like(np.asarray(x), dtype=dtype).view(EagerTensor)


def range(*args, dtype=None):  # noqa: A001
    args = [builtins_int(np.asarray(a)) if not isinstance(a, builtins_float) else a for a in args]
    a = np.arange(*args)
    if dtype is None:
        dtype = np.int32 if a.dtype.kind in 'iu' else np.float32
    return a.astype(dtype).view(EagerTensor)


def Variable(initial_value, **_):
    return np.array(initial_value, copy=True).view(EagerTensor)


def stop_gradient(x, name=None):
    return _t(x)


# ---------------------------------------------------------------- shape ops
def size(x):
    return np.asarray(np.asarray(x).size, dtype=np.int32).view(EagerTensor)


def shape(x):
    return np.asarray(np.asarray(x).shape, dtype=np.int32).view(EagerTensor)


def reshape(x, shape, name=None):
    return np.reshape(np.asarray(x), [builtins_int(s) for s in np.asarray(shape).reshape(-1)]).view(EagerTensor)


def squeeze(x, axis=None):
    if isinstance(axis, list):
        axis = tuple(axis)
    return np.squeeze(np.asarray(x), axis=axis).view(EagerTensor)


def expand_dims(x, axis):
    return np.expand_dims(np.asarray(x), axis).view(EagerTensor)


def transpose(x, perm=None):
    return np.transpose(np.asarray(x), perm).view(EagerTensor)


def split(x, num, axis=0):
    return [p.view(EagerTensor) for p in np.split(np.asarray(x), num, axis=axis)]


def unstack(x, axis=0):
    x = np.asarray(x)
    return [np.take(x, i, axis=axis).view(EagerTensor) for i in _b.range(x.shape[axis])]


def stack(xs, axis=0):
    return np.stack([np.asarray(_t(v)) for v in xs], axis=axis).view(EagerTensor)


def concat(xs, axis=0, name=None):
    return np.concatenate([np.asarray(_t(v)) for v in xs], axis=axis).view(EagerTensor)


def meshgrid(*xs):
    return [m.view(EagerTensor) for m in np.meshgrid(*[np.asarray(v) for v in xs])]


def pad(x, paddings, mode='CONSTANT'):
    mode = {'CONSTANT': 'constant', 'SYMMETRIC': 'symmetric', 'REFLECT': 'reflect'}[mode]
    return np.pad(np.asarray(x), paddings, mode=mode).view(EagerTensor)


# ---------------------------------------------------------------- gather / scatter / where
def gather(params, indices, axis=0):
    return np.take(np.asarray(params), np.asarray(indices), axis=axis).view(EagerTensor)


def gather_nd(params, indices):
    idx = np.asarray(indices)
    return np.asarray(params)[tuple(idx[..., i] for i in _b.range(idx.shape[-1]))].view(EagerTensor)


def scatter_update(ref, indices, updates):
    ref[np.asarray(indices)] = np.asarray(updates)
    return ref


def scatter_nd_update(ref, indices, updates):
    idx = np.asarray(indices)
    ref[tuple(idx[..., i] for i in _b.range(idx.shape[-1]))] = np.asarray(updates)
    return ref


def where(condition, x=None, y=None):
    c = np.asarray(condition)
    if x is None:
        return np.argwhere(c).astype(np.int64).view(EagerTensor)  # row-major ascending, int64
    return np.where(c, np.asarray(x), np.asarray(y)).view(EagerTensor)


# ---------------------------------------------------------------- math
def _bin(fn):
    def op(a, b, name=None):
        a, b = _t(a), _t(b)
        if a.dtype != b.dtype:  # python scalars follow the tensor operand (TF convert semantics)
            if a.ndim == 0 and b.ndim > 0:
                a = a.astype(b.dtype)
            elif b.ndim == 0:
                b = b.astype(a.dtype)
        return np.asarray(fn(np.asarray(a), np.asarray(b))).view(EagerTensor)
    return op


maximum = _bin(np.maximum)
minimum = _bin(np.minimum)
add = _bin(np.add)
multiply = _bin(np.multiply)
equal = _bin(np.equal)
logical_and = _bin(np.logical_and)


def truediv(a, b):
    return (np.asarray(_t(a)) / np.asarray(_t(b))).view(EagerTensor)


def _un(fn):
    def op(x, name=None):
        return np.asarray(fn(np.asarray(_t(x)))).view(EagerTensor)
    return op


exp = _un(np.exp)
log = _un(np.log)
sqrt = _un(np.sqrt)
floor = _un(np.floor)
abs = _un(np.abs)  # noqa: A001
less = _bin(np.less)


def pow(x, y):  # noqa: A001
    x = np.asarray(x)
    return np.power(x, np.asarray(y, dtype=x.dtype)).view(EagerTensor)


def reduce_mean(x, axis=None):
    x = np.asarray(x)
    axis = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return np.asarray(np.mean(x, axis=axis, dtype=x.dtype)).view(EagerTensor)


def _sparse_softmax_cross_entropy(logits, labels, weights=1.0):
    """tf.losses.sparse_softmax_cross_entropy, TF r1.13 (python/ops/losses/losses_impl.py): per-row
    `logsumexp(x) - x[label]` (kernel: max-shifted, core/kernels/xent_op.h), times `weights`, reduction
    SUM_BY_NONZERO_WEIGHTS = sum / number of rows with a non-zero weight (0 when there are none)."""
    x = np.asarray(logits, np.float32)
    lab = np.asarray(labels).astype(np.int64).reshape(-1)
    w = np.broadcast_to(np.asarray(weights, np.float32), lab.shape)
    z = x - x.max(axis=1, keepdims=True)
    per = np.log(np.exp(z).sum(axis=1, dtype=np.float32)) - z[np.arange(lab.size), lab]
    present = np.float32(np.count_nonzero(w))
    total = np.sum(per.astype(np.float32) * w, dtype=np.float32)
    return np.asarray(total / present if present > 0 else np.float32(0), np.float32).view(EagerTensor)


losses = types.SimpleNamespace(sparse_softmax_cross_entropy=_sparse_softmax_cross_entropy)


def reduce_max(x, axis=None):
    return np.asarray(np.max(np.asarray(x), axis=axis)).view(EagerTensor)


def reduce_sum(x, axis=None):
    x = np.asarray(x)
    axis = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return np.asarray(np.sum(x, axis=axis, dtype=x.dtype)).view(EagerTensor)


def argmax(x, axis=None, output_type=np.int64):
    return np.asarray(np.argmax(np.asarray(x), axis=axis)).astype(output_type).view(EagerTensor)  # first max


# ---------------------------------------------------------------- randomness (permutation-injected)
_shuffle_fn = None


def random_shuffle(x, seed=None):
    """`tf.random_shuffle` is unseeded in the reference (anchor_target.py:74,81; proposal_target.py:68,71).
    The oracle injects the order through `shuffle_hook` (idx sorted by pi[idx])."""
    if _shuffle_fn is None:
        raise RuntimeError('tf_shim.random_shuffle: no shuffle hook installed')
    return _t(_shuffle_fn(np.asarray(x)))


@contextlib.contextmanager
def shuffle_hook(fn):
    global _shuffle_fn
    old, _shuffle_fn = _shuffle_fn, fn
    try:
        yield
    finally:
        _shuffle_fn = old


# ---------------------------------------------------------------- scopes / logging
@contextlib.contextmanager
def name_scope(name):
    yield


variable_scope = name_scope

logging = types.SimpleNamespace(debug=lambda *a, **k: None, info=lambda *a, **k: None,
                                warning=lambda *a, **k: None)


# ---------------------------------------------------------------- tf.nn
def _top_k(x, k=1, sorted=True):  # noqa: A002
    x = np.asarray(x)
    order = np.argsort(-x, kind='stable')[:builtins_int(np.asarray(k))]  # ties -> lower index
    return x[order].view(EagerTensor), order.astype(np.int32).view(EagerTensor)


def _pool2x2(x, ksize, strides, padding, reducer):
    """2x2 / stride 2 pooling, 'SAME'; windows clipped at the border (TF excludes padding)."""
    x = np.asarray(x)
    assert list(ksize) == [1, 2, 2, 1] and list(strides) == [1, 2, 2, 1] and padding.upper() == 'SAME'
    n, h, w, c = x.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = np.empty((n, oh, ow, c), dtype=x.dtype)
    for i in _b.range(oh):
        for j in _b.range(ow):
            out[:, i, j, :] = reducer(x[:, 2 * i:2 * i + 2, 2 * j:2 * j + 2, :].reshape(n, -1, c), axis=1)
    return out.view(EagerTensor)


def _avg_pool(x, ksize, strides, padding, data_format='NHWC', name=None):
    return _pool2x2(x, ksize, strides, padding, lambda v, axis: np.mean(v, axis=axis, dtype=np.float32))


def _softmax(logits, axis=-1):
    """tf.nn.softmax (core/kernels/softmax_op_functor.h): exp(x - max) / sum(exp(x - max)) along the last axis."""
    x = np.asarray(logits, np.float32)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return (e / e.sum(axis=axis, keepdims=True, dtype=np.float32)).view(EagerTensor)


nn = types.SimpleNamespace(top_k=_top_k, avg_pool=_avg_pool, softmax=_softmax)


# ---------------------------------------------------------------- tf.image
def _nms(boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float('-inf'), name=None):
    """`tf.image.non_max_suppression` (CPU NonMaxSuppressionV3) restated — SURVEY App. B.1.
    Greedy in descending score (ties: lower index first), IoU WITHOUT +1 on min/max-normalised
    corners, non-positive-area boxes never suppress / are never suppressed, strict `>`."""
    b = np.asarray(boxes, dtype=np.float32)
    s = np.asarray(scores, dtype=np.float32)
    thr = np.float32(iou_threshold)
    max_out = builtins_int(np.asarray(max_output_size))
    ymin = np.minimum(b[:, 0], b[:, 2]); ymax = np.maximum(b[:, 0], b[:, 2])
    xmin = np.minimum(b[:, 1], b[:, 3]); xmax = np.maximum(b[:, 1], b[:, 3])
    area = (ymax - ymin) * (xmax - xmin)
    order = np.argsort(-s, kind='stable')
    sel = []
    for c in order:
        if len(sel) >= max_out:
            break
        if s[c] <= score_threshold:
            break
        if sel and area[c] > 0:
            k = np.asarray(sel)
            ih = np.maximum(np.float32(0), np.minimum(ymax[c], ymax[k]) - np.maximum(ymin[c], ymin[k]))
            iw = np.maximum(np.float32(0), np.minimum(xmax[c], xmax[k]) - np.maximum(xmin[c], xmin[k]))
            inter = ih * iw
            with np.errstate(divide='ignore', invalid='ignore'):
                iou = inter / (area[c] + area[k] - inter)
            if np.any((area[k] > 0) & (iou > thr)):
                continue
        sel.append(c)
    return np.asarray(sel, dtype=np.int32).view(EagerTensor)


def _crop_and_resize(image, boxes, box_ind, crop_size, method='bilinear', extrapolation_value=0, name=None):
    """`tf.image.crop_and_resize` (bilinear) restated — SURVEY App. B.2 (fp32 op order kept)."""
    img = np.asarray(image, dtype=np.float32)
    bx = np.asarray(boxes, dtype=np.float32)
    bi = np.asarray(box_ind)
    ch, cw = builtins_int(crop_size[0]), builtins_int(crop_size[1])
    _, h, w, c = img.shape
    r = bx.shape[0]
    f = np.float32
    out = np.full((r, ch, cw, c), f(extrapolation_value), dtype=np.float32)
    y1, x1, y2, x2 = bx[:, 0], bx[:, 1], bx[:, 2], bx[:, 3]
    hs = (y2 - y1) * f(h - 1) / f(ch - 1) if ch > 1 else np.zeros(r, np.float32)
    ws = (x2 - x1) * f(w - 1) / f(cw - 1) if cw > 1 else np.zeros(r, np.float32)
    for y in _b.range(ch):
        in_y = y1 * f(h - 1) + f(y) * hs if ch > 1 else f(0.5) * (y1 + y2) * f(h - 1)
        oky = ~((in_y < 0) | (in_y > f(h - 1)))
        top = np.floor(in_y); bot = np.ceil(in_y); ly = (in_y - top).astype(np.float32)
        for x in _b.range(cw):
            in_x = x1 * f(w - 1) + f(x) * ws if cw > 1 else f(0.5) * (x1 + x2) * f(w - 1)
            ok = oky & ~((in_x < 0) | (in_x > f(w - 1)))
            if not ok.any():
                continue
            left = np.floor(in_x); right = np.ceil(in_x); lx = (in_x - left).astype(np.float32)
            k = np.nonzero(ok)[0]
            t_, b_, l_, r_ = (top[k].astype(np.int64), bot[k].astype(np.int64),
                              left[k].astype(np.int64), right[k].astype(np.int64))
            n = bi[k]
            tl = img[n, t_, l_]; tr = img[n, t_, r_]; bl = img[n, b_, l_]; br = img[n, b_, r_]
            tp = tl + (tr - tl) * lx[k, None]
            bt = bl + (br - bl) * lx[k, None]
            out[k, y, x] = tp + (bt - tp) * ly[k, None]
    return out.view(EagerTensor)


image = types.SimpleNamespace(non_max_suppression=_nms, crop_and_resize=_crop_and_resize)


# ---------------------------------------------------------------- tf.keras
class _Model:
    def __init__(self, *a, **k):
        pass

    def __call__(self, inputs, training=None, mask=None):
        # Keras converts ndarray inputs to eager tensors; python lists/ints (image_shape, stride) pass through
        if isinstance(inputs, (list, tuple)):
            inputs = tuple(_t(v) if isinstance(v, np.ndarray) else v for v in inputs)
        elif isinstance(inputs, np.ndarray):
            inputs = _t(inputs)
        return self.call(inputs, training=training, mask=mask)


class _MaxPooling2D:
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', **_):
        assert tuple(np.broadcast_to(pool_size, 2)) == (2, 2) and strides is None and padding == 'same'

    def __call__(self, x):
        return _pool2x2(x, [1, 2, 2, 1], [1, 2, 2, 1], 'SAME', np.max)


class _Concatenate:
    def __init__(self, axis=-1):
        self.axis = axis

    def __call__(self, xs):
        return concat(xs, axis=self.axis)


keras = types.SimpleNamespace(
    Model=_Model,
    layers=types.SimpleNamespace(MaxPooling2D=_MaxPooling2D, Concatenate=_Concatenate),
)
