"""ctypes binding of include/boxpath.h (libboxpath.so).  No fallback: a missing library or a missing CUDA
device raises — the product path never computes on the CPU."""
import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libboxpath.so')

BX_OK, BX_ERR_INVALID, BX_ERR_CUDA, BX_ERR_UNSUPPORTED, BX_ERR_DLPACK = 0, -1, -2, -3, -4
ROI_STRIDE_NORM, ROI_IMAGE_NORM, ROI_ALIGN_PAD = 0, 1, 2
RPN_CAFFE, RPN_PAIRS = 0, 1
POOL_NONE, POOL_MAX2, POOL_AVG2 = 0, 1, 2
CUT_TOP_K, CUT_SCORE_GE = 0, 1

F4 = c_float * 4


class ProposalParams(ctypes.Structure):
    _fields_ = [('means', F4), ('stds', F4), ('image_h', c_int), ('image_w', c_int), ('pre_nms_top_k', c_int),
                ('post_nms', c_int), ('iou_threshold', c_float), ('min_size', c_float)]


class AnchorTargetParams(ctypes.Structure):
    _fields_ = [('pos_iou_threshold', c_float), ('neg_iou_threshold', c_float), ('total_num_samples', c_int),
                ('max_pos_samples', c_int), ('means', F4), ('stds', F4), ('image_h', c_int), ('image_w', c_int)]


class ProposalTargetParams(ctypes.Structure):
    _fields_ = [('num_classes', c_int), ('pos_iou_threshold', c_float), ('neg_iou_threshold', c_float),
                ('total_num_samples', c_int), ('max_pos_samples', c_int), ('means', F4), ('stds', F4)]


class PredictionParams(ctypes.Structure):
    _fields_ = [('means', F4), ('stds', F4), ('image_h', c_int), ('image_w', c_int), ('num_classes', c_int),
                ('max_per_class', c_int), ('max_per_image', c_int), ('nms_iou_threshold', c_float),
                ('score_threshold', c_float), ('min_edge', c_float)]


P = c_void_p  # device pointers travel as integers

# name -> (restype, argtypes); every symbol include/boxpath.h declares
SIGNATURES = {
    'bx_version': (c_int, []),
    'bx_last_error': (c_char_p, []),
    'bx_create': (c_int, [c_int, POINTER(c_void_p)]),
    'bx_destroy': (c_int, [c_void_p]),
    'bx_launch_count': (c_longlong, [c_void_p]),
    'bx_stats': (c_int, [c_void_p, POINTER(c_longlong), c_int]),
    'bx_set_deterministic': (c_int, [c_void_p, c_int]),
    'bx_reserve': (c_int, [c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, c_void_p]),
    'bx_profile_roi': (c_int, [c_void_p, c_int, c_int]),
    'bx_profile_read': (c_int, [c_void_p, POINTER(c_float), c_int, POINTER(c_int)]),
    'bx_dlpack_data': (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(c_int64), c_int, POINTER(c_void_p)]),
    'bx_decode_clip': (c_int, [c_void_p, P, c_int, P, c_int, c_int, F4, F4, c_int, c_int, P, c_void_p]),
    'bx_encode': (c_int, [c_void_p, P, P, c_int, F4, F4, P, c_void_p]),
    'bx_clip_filter': (c_int, [c_void_p, P, c_int, c_float, c_int, c_int, c_float, P, P, P, c_void_p]),
    'bx_range_filter': (c_int, [c_void_p, P, c_int, c_int, c_int, P, P, c_void_p]),
    'bx_nms': (c_int, [c_void_p, P, P, c_int, c_int, c_int, c_float, P, P, c_void_p]),
    'bx_proposals': (c_int, [c_void_p, P, P, P, c_int, c_int, POINTER(ProposalParams), P, P, P, c_void_p]),
    'bx_generate_anchors': (c_int, [c_void_p, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int,
                                    POINTER(c_float), P, c_void_p]),
    'bx_rpn_scores': (c_int, [c_void_p, P, c_int, c_int, c_int, c_int, P, c_void_p]),
    'bx_proposals_rpn': (c_int, [c_void_p, P, P, P, c_int, c_int, c_int, c_int, POINTER(ProposalParams), P, P, P, P,
                                 c_void_p]),
    'bx_crop_and_resize': (c_int, [c_void_p, P, c_int, c_int, c_int, c_int, P, P, c_int, c_int, c_int, c_float, P,
                                   c_void_p]),
    'bx_roi_pool': (c_int, [c_void_p, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P, P, P, c_int, c_float,
                            c_int, c_int, P, c_void_p]),
    'bx_roi_pool_grad': (c_int, [c_void_p, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P, P, P, c_int, c_float,
                                 c_int, c_int, P, P, c_void_p]),
    'bx_allgather_detections': (c_int, [c_void_p, c_void_p, P, P, c_int, c_int, c_int, c_int, P, P, c_void_p]),
    'bx_smooth_l1_loss': (c_int, [c_void_p, P, P, P, P, c_longlong, c_int, c_float, c_int, P, P, c_void_p]),
    'bx_cls_loss': (c_int, [c_void_p, P, P, c_int, c_int, c_float, P, P, P, c_void_p]),
    'bx_fpn_assign_levels': (c_int, [c_void_p, P, c_int, c_int, c_int, P, P, P, c_void_p]),
    'bx_fpn_roi_features': (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), c_int, c_int, c_int,
                                    c_int, P, P, c_int, c_int, c_int, c_int, P, P, P, P, c_void_p]),
    'bx_pairwise_iou': (c_int, [c_void_p, P, c_int, P, c_int, P, c_void_p]),
    'bx_anchor_target': (c_int, [c_void_p, P, c_int, P, P, c_int, c_int, P, POINTER(AnchorTargetParams), P, P, P, P,
                                 P, c_void_p]),
    'bx_proposal_target': (c_int, [c_void_p, P, P, c_int, P, P, P, c_int, c_int, P, POINTER(ProposalTargetParams), P,
                                   P, P, P, P, P, P, c_void_p]),
    'bx_post_ops_prediction': (c_int, [c_void_p, P, P, P, P, c_int, c_int, POINTER(PredictionParams), P, P, c_void_p]),
    'bx_eval_detections': (c_int, [c_void_p, P, P, P, P, P, P, c_int, c_int, POINTER(PredictionParams), c_int, c_int, P, P,
                                   c_void_p]),
    'bx_c4_proposal_roi': (c_int, [c_void_p, P, P, P, P, c_int, c_int, c_int, c_int, c_int, POINTER(ProposalParams),
                                   c_float, c_int, c_int, P, P, P, P, c_void_p]),
    'bx_c4_proposal_roi_host': (c_int, [c_void_p, P, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                        POINTER(ProposalParams), c_float, c_int, c_int, P, P, P, P, c_void_p]),
}

_lib = None
_lock = threading.Lock()
_handles = {}


class BoxpathError(RuntimeError):
    pass


def load():
    """Load libboxpath.so (once).  Raises ImportError if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    'libboxpath.so not found at %s: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                    '(or ./build.sh).  tf_eager_object_detection_b200 has no CPU fallback.' % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError here = header/library mismatch
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def last_error():
    return load().bx_last_error().decode('utf-8', 'replace')


def check(rc):
    if rc == BX_OK:
        return
    msg = last_error()
    if rc == BX_ERR_INVALID:
        raise ValueError(msg)                 # the reference raises ValueError / TF InvalidArgumentError here
    if rc == BX_ERR_DLPACK:
        raise TypeError(msg)
    if rc == BX_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise BoxpathError(msg)


class _ThreadHandles:
    """The bx_handle* objects of one host thread; destroyed (workspaces freed) when the thread ends or at exit."""

    def __init__(self):
        self.by_key = {}

    def close(self):
        lib = _lib
        for key, h in list(self.by_key.items()):
            _handles.pop(key, None)
            if lib is not None:
                try:
                    lib.bx_destroy(h)
                except Exception:
                    pass
        self.by_key.clear()

    def __del__(self):
        self.close()


_tls = threading.local()
_all_sets = []


def handle(device_index, stream=0):
    """bx_handle* for (device, calling thread, stream) — SURVEY §8b: one handle per (device, host thread); keyed by the
    stream as well so that calls issued on different streams never share a workspace.  Handles are destroyed with their
    thread (thread-local owner) and at interpreter exit."""
    key = (int(device_index), threading.get_ident(), int(stream or 0))
    h = _handles.get(key)
    if h is None:
        lib = load()
        out = c_void_p()
        check(lib.bx_create(int(device_index), ctypes.byref(out)))
        owner = getattr(_tls, 'owner', None)
        if owner is None:
            owner = _tls.owner = _ThreadHandles()
            import weakref
            _all_sets.append(weakref.ref(owner))
        h = _handles[key] = owner.by_key[key] = out
    return h


def destroy_handles():
    """Destroy every cached handle (all threads).  Registered with atexit; callable from tests."""
    for ref in list(_all_sets):
        owner = ref()
        if owner is not None:
            owner.close()
    del _all_sets[:]
    _handles.clear()


import atexit  # noqa: E402
atexit.register(destroy_handles)


def launch_count(device_index):
    return int(load().bx_launch_count(handle(device_index)))


def stats(h):
    """bx_stats of a handle as a dict (launches, band_launches, band_fallbacks, workspace / plan / stage bytes)."""
    buf = (c_longlong * 6)()
    check(load().bx_stats(h, buf, 6))
    return dict(zip(('launches', 'band_launches', 'band_fallbacks', 'ws_bytes', 'plan_bytes', 'stage_bytes'),
                    [int(v) for v in buf]))


def total_stats():
    """Sum of stats() over every live handle of the process."""
    tot = {}
    for h in list(_handles.values()):
        for k, v in stats(h).items():
            tot[k] = tot.get(k, 0) + v
    return tot


def f4(values):
    return F4(*[float(v) for v in values])
