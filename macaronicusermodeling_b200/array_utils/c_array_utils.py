"""Drop-in for array_utils/c_array_utils.pyx: same function names, argument meaning, return shapes and error
behaviour, with the arithmetic done by libmlbp.so on the GPU (float64, like the reference).

Contract kept from the reference: 2-D float64 C-contiguous ndarrays in, NEW ndarray out;
``normalize`` returns a new array when the sum is positive, otherwise zero-fills its argument IN PLACE and
returns it (pyx:29-40); the typed functions ``dense_dot`` / ``dense_pointwise_multiply`` raise ValueError on
non-float64 or non-2-D buffers (pyx:90-94).  There is no CPU fallback: without the built library or without a
B200 every call raises.

The batched engine does not go through these per-message helpers (it fuses them into K1..K6); they exist so
that code written against ``au`` keeps working.
"""
import ctypes

import numpy as np
import torch

from .. import _lib

K = 100  # hard-coded top-K of the reference's sparse approximations (pyx:44, 54, 67, 97, 118, 194)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _typed_2d(*arrays):
    for a in arrays:
        if not isinstance(a, np.ndarray) or a.dtype != np.float64:
            raise ValueError("Buffer dtype mismatch, expected 'float64_t'")
        if a.ndim != 2:
            raise ValueError('Buffer has wrong number of dimensions (expected 2, got %d)' % a.ndim)


def pointwise_multiply(m1, m2):
    """pyx:12-16  np.multiply(m1, m2)"""
    lib = _lib.require_device()
    m1 = np.asarray(m1, dtype=np.float64)
    m2 = np.asarray(m2, dtype=np.float64)
    if m1.shape != m2.shape:
        m1, m2 = np.broadcast_arrays(m1, m2)
    a, b = _dev(m1), _dev(m2)
    out = torch.empty_like(a)
    _lib.check(lib.mlbp_pointwise_multiply_f64(_p(a), _p(b), _p(out), a.numel(), _stream()))
    return out.cpu().numpy().reshape(m1.shape)


def normalize(m1):
    """pyx:29-40"""
    lib = _lib.require_device()
    a = _dev(m1)
    out = torch.empty_like(a)
    s = torch.zeros(1, dtype=torch.float64, device='cuda')
    _lib.check(lib.mlbp_normalize_f64(_p(a), _p(out), a.numel(), _p(s), _stream()))
    if float(s.item()) > 0.0:
        return out.cpu().numpy().reshape(np.shape(m1))
    m1.fill(0)
    return m1


def dense_dot(m1, m2):
    """pyx:90-91  m1.dot(m2)"""
    _typed_2d(m1, m2)
    if m1.shape[1] != m2.shape[0]:
        raise ValueError('shapes %s and %s not aligned' % (m1.shape, m2.shape))
    lib = _lib.require_device()
    a, b = _dev(m1), _dev(m2)
    out = torch.empty((m1.shape[0], m2.shape[1]), dtype=torch.float64, device='cuda')
    _lib.check(lib.mlbp_dense_dot_f64(_p(a), _p(b), _p(out), m1.shape[0], m1.shape[1], m2.shape[1], _stream()))
    return out.cpu().numpy()


def dense_pointwise_multiply(m1, m2):
    """pyx:93-94"""
    _typed_2d(m1, m2)
    lib = _lib.require_device()
    a, b = _dev(m1), _dev(m2)
    out = torch.empty_like(a)
    _lib.check(lib.mlbp_dense_pointwise_multiply_f64(_p(a), _p(b), _p(out), a.numel(), _stream()))
    return out.cpu().numpy().reshape(m1.shape)


def induce_s_multiply_threshold(s1, d2):
    """pyx:78-87"""
    raise NotImplementedError("do not use it seems very slow..")


def sd_pointwise_multiply(s1, d2):
    """pyx:149-150"""
    raise NotImplementedError("not implemented pointwise multiply for sparse-dense matrix")
