"""Tensor hand-off between the host framework and the C ABI.

Rules (DESIGN.md §1 "Boundary"):
  * a CUDA tensor of the expected dtype, C-contiguous and suitably aligned is BORROWED zero-copy — torch tensors by their
    storage pointer after the same checks bx_dlpack_data makes, any other `__dlpack__` producer (a TF >= 2.2 EagerTensor
    through tf.experimental.dlpack in the reference's world), or every tensor under BX_FORCE_DLPACK=1, through a DLPack
    capsule validated by bx_dlpack_data;
  * everything else is normalised HERE, in the Python layer, before the C ABI sees it, and counted in `CONVERSIONS`:
    host data (numpy / lists / CPU tensors) is uploaded ("uploads" — feeding a numpy array to a TF op), other dtypes are
    cast ("casts" — the reference's own tf.to_float / tf.to_int32), non-contiguous views are compacted ("compactions");
    BX_STRICT=1 turns each of them into a TypeError instead, for callers that want to prove their path makes none;
  * the C ABI itself never copies or converts: bx_dlpack_data rejects wrong device / dtype / strides / alignment.

torch is used for device memory (output allocation), streams and nothing else."""
import ctypes
import os
from ctypes import c_char_p, c_int64, c_void_p, py_object

import numpy as np
import torch

from . import _lib

_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = c_void_p
_PyCapsule_GetPointer.argtypes = [py_object, c_char_p]

FLOAT32 = (2, 32)
INT32 = (0, 32)


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.BoxpathError('no CUDA device available: tf_eager_object_detection_b200 runs only on the GPU '
                                '(hand-written sm_100a kernels; no CPU fallback)')


CONVERSIONS = {'uploads': 0, 'casts': 0, 'compactions': 0}


def strict():
    return os.environ.get('BX_STRICT', '') not in ('', '0')


def _converted(kind, what):
    if strict():
        raise TypeError('BX_STRICT: %s (%s) would need a copy before the C ABI' % (what, kind))
    CONVERSIONS[kind] += 1


def to_device(x, dtype, device=None):
    """Normalise one input for the C ABI: torch CUDA tensors of the right dtype pass through untouched (zero-copy);
    host data is uploaded and other dtypes are cast — explicitly, and counted in CONVERSIONS (module docstring)."""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            require_cuda()
            _converted('uploads', 'CPU tensor')
            x = x.cuda(device)
        if x.dtype != dtype:
            _converted('casts', 'dtype %s -> %s' % (x.dtype, dtype))
            x = x.to(dtype)
        return x.detach() if x.requires_grad else x
    require_cuda()
    _converted('uploads', type(x).__name__)
    np_dtype = {torch.float32: np.float32, torch.int32: np.int32}[dtype]
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np_dtype)), device=device or 'cuda')


_TORCH_KIND = {torch.float32: (2, 32), torch.int32: (0, 32)}


def force_dlpack():
    """BX_FORCE_DLPACK=1 routes torch tensors through the DLPack capsule path too (it is always used for tensors of
    other frameworks).  By default a torch.Tensor is validated on its own metadata — the same checks bx_dlpack_data
    makes — and its storage pointer is borrowed directly, which saves ~6 us of host time per tensor."""
    return os.environ.get('BX_FORCE_DLPACK', '') not in ('', '0')


class Borrow:
    """Holds the DLPack capsules of one call alive and resolves them to validated device pointers."""

    def __init__(self, device_index):
        self.device = device_index
        self._caps = []
        self._lib = _lib.load()
        self._dlpack = force_dlpack()

    def ptr(self, t, kind, shape, align=4):
        if t is None:
            return None
        if isinstance(t, torch.Tensor) and not t.is_contiguous():
            _converted('compactions', 'non-contiguous view %s' % (tuple(t.stride()),))
            t = t.contiguous()
            self._caps.append(t)
        if isinstance(t, torch.Tensor) and not self._dlpack:
            # zero-copy borrow of a torch tensor: same rejections as bx_dlpack_data (device, dtype, shape, alignment)
            dv = t.device
            if dv.type != 'cuda':
                raise TypeError('tensor is not on a CUDA device (%s); libboxpath has no CPU path' % dv)
            if dv.index is not None and dv.index != self.device:
                raise TypeError('tensor is on cuda:%d, handle is on cuda:%d' % (dv.index, self.device))
            if _TORCH_KIND.get(t.dtype) != kind:
                raise TypeError('tensor dtype %s != expected (code %d, %d bits)' % (t.dtype, kind[0], kind[1]))
            ts = tuple(t.shape)
            if ts != shape and (len(ts) != len(shape) or any(e >= 0 and e != a for a, e in zip(ts, shape))):
                raise TypeError('tensor shape %s, expected %s' % (ts, tuple(shape)))
            p = t.data_ptr()
            if align > 1 and p % align and t.numel():
                raise TypeError('tensor data is not %d-byte aligned' % align)
            self._caps.append(t)
            return p
        cap = t.__dlpack__()
        self._caps.append(cap)
        dl = _PyCapsule_GetPointer(cap, b'dltensor')
        shp = (c_int64 * len(shape))(*shape)
        out = c_void_p()
        _lib.check(self._lib.bx_dlpack_data(dl, self.device, kind[0], kind[1], len(shape), shp, align, ctypes.byref(out)))
        return out.value if out.value else 0


def device_index_of(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr(device_index):
    """torch's current stream on the device as a cudaStream_t (raw accessor when torch has it: no Stream object)."""
    if _raw_stream is not None:
        return c_void_p(_raw_stream(device_index))
    return c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


_DEVICES = {}


def empty(shape, dtype, device_index):
    dev = _DEVICES.get(device_index)
    if dev is None:
        dev = _DEVICES[device_index] = torch.device('cuda', device_index)
    return torch.empty(shape, dtype=dtype, device=dev)
