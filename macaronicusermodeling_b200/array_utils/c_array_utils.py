"""Drop-in for array_utils/c_array_utils.pyx: same function names, argument meaning, return shapes and error
behaviour, with the arithmetic done by libmlbp.so on the GPU (float64, like the reference).

Contract kept from the reference: 2-D float64 C-contiguous ndarrays in, NEW ndarray out;
``normalize`` returns a new array when the sum is positive, otherwise zero-fills its argument IN PLACE and
returns it (pyx:29-40); the typed functions ``dense_dot`` / ``dense_pointwise_multiply`` raise ValueError on
non-float64 or non-2-D buffers (pyx:90-94).  There is no CPU fallback: without the built library or without a
B200 every call raises.

The batched engine does not go through these per-message helpers (it fuses them into K1..K6); they exist so
that code written against ``au`` keeps working.  All 22 functions of the .pyx are here.  Products, dot products and
normalisations run on the device; WHICH entries a top-K variant keeps is decided by the same ``np.argpartition``
call the reference makes (an index selection whose tie-breaking is part of the reference's results), and the
functions that only move data (``clip``, ``make_adapt_phi``, ``set_adaptation*``, ``set_original``) are plain
array assignments like in the reference.
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


def _mul(a, b):
    """elementwise product of two equal-shape float64 arrays on the device"""
    lib = _lib.require_device()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if a.size == 0:
        return a * b
    x, y = _dev(a), _dev(b)
    out = torch.empty_like(x)
    _lib.check(lib.mlbp_pointwise_multiply_f64(_p(x), _p(y), _p(out), x.numel(), _stream()))
    return out.cpu().numpy().reshape(a.shape)


def _dot(a, b):
    """(m, k) . (k, n) on the device"""
    lib = _lib.require_device()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x, y = _dev(a), _dev(b)
    out = torch.empty((a.shape[0], b.shape[1]), dtype=torch.float64, device='cuda')
    _lib.check(lib.mlbp_dense_dot_f64(_p(x), _p(y), _p(out), a.shape[0], a.shape[1], b.shape[1], _stream()))
    return out.cpu().numpy()


def _div_by_sum(a):
    """(a / sum(a), sum(a)) with the sum and the division on the device; sum <= 0 falls back to NumPy's a / s
    (inf / nan like the reference, which divides unconditionally in these helpers)"""
    lib = _lib.require_device()
    a = np.ascontiguousarray(a, dtype=np.float64)
    x = _dev(a)
    out = torch.empty_like(x)
    s = torch.zeros(1, dtype=torch.float64, device='cuda')
    _lib.check(lib.mlbp_normalize_f64(_p(x), _p(out), x.numel(), _p(s), _stream()))
    sv = float(s.item())
    if sv > 0.0:
        return out.cpu().numpy().reshape(a.shape), sv
    with np.errstate(all='ignore'):
        return a / sv, sv


def clip(m1):
    """pyx:18-20 (in place)"""
    m1[m1 < 1.0e-100] = 0.0
    return m1


def sparse_normalize(m1, c_idx, r_idx):
    """pyx:23-26: normalise the (c_idx x r_idx) block in place"""
    ix = np.ix_(c_idx, r_idx)
    m1[ix] = _div_by_sum(m1[ix])[0]
    return m1


def induce_s_pointwise_multiply_clip(d1, d2):
    """pyx:43-50: product restricted to the K largest entries of d1"""
    if __debug__: assert np.shape(d1) == np.shape(d2)
    indices = (-d1).argpartition(K, axis=None)[:K]
    x, y = np.unravel_index(indices, d1.shape)
    result = np.zeros_like(d2)
    result[x, y] = _mul(d1[x, y], d2[x, y])
    return result


def induce_s(m1):
    """pyx:53-63: keep the K largest entries of a column vector"""
    if __debug__: assert np.shape(m1)[1] == 1
    if K > np.size(m1):
        return m1
    indices = (-m1).argpartition(K, axis=None)[:K]
    x, y = np.unravel_index(indices, m1.shape)
    new_m1 = np.zeros_like(m1)
    new_m1[x, y] = m1[x, y]
    return new_m1


def induce_s_mutliply_clip(s1, d2):
    """pyx:66-75 [sic]: d2 . s1 over the K entries of s1 that are largest in magnitude"""
    if __debug__: assert np.shape(d2)[0] < np.shape(d2)[1]
    if __debug__: assert np.shape(s1)[0] == np.shape(d2)[1] and np.shape(s1)[1] == 1
    s1_abs = np.reshape(np.abs(s1), (np.size(s1),))
    max_idx = np.argpartition(s1_abs, -K)[-K:]
    return _dot(d2[:, max_idx], s1[max_idx, :])


def make_sparse_and_dot(m1, m2):
    """pyx:96-105: {(x, y): m1[x, 0] * m2[0, y]} over the top-K x top-K entries"""
    m1_max_idx = np.argpartition(np.reshape(m1, np.size(m1)), -K)[-K:]
    m2_max_idx = np.argpartition(np.reshape(m2, np.size(m2)), -K)[-K:]
    block = _dot(np.asarray(m1)[m1_max_idx, :1], np.asarray(m2)[:1, m2_max_idx])
    d = {}
    for i, x in enumerate(m1_max_idx):
        for j, y in enumerate(m2_max_idx):
            d[x, y] = block[i, j]
    return d


def sparse_pointwise_multiply(sparse_m, c_idx, r_idx, dense_m):
    """pyx:108-114"""
    _typed_2d(sparse_m, dense_m)
    ix = np.ix_(c_idx, r_idx)
    z = np.zeros_like(dense_m)
    z[ix] = _mul(sparse_m[ix], dense_m[ix])
    return z


def sparse_dot(m1, m2):
    """pyx:117-129: outer product of a column and a row vector restricted to their K largest entries ->
    (out (n, n), m1_idx (K,), m2_idx (K,))"""
    _typed_2d(m1, m2)
    assert m1.shape[0] == m2.shape[1]
    assert m1.shape[1] == m2.shape[0] == 1
    n = m1.shape[0]
    out = np.zeros((n, n), dtype=np.float64)
    m1_idx = np.argpartition(-m1, K - 1, axis=0)[:K].ravel()
    m2_idx = np.argpartition(-m2, K - 1)[:, :K].ravel()
    out[np.ix_(m1_idx, m2_idx)] = _dot(m1[m1_idx], m2[:, m2_idx])
    return out, m1_idx, m2_idx


def sparse_multiply_and_normalize(s_m1, m2):
    """pyx:132-142: (dense array, dict) of m2[x, y] * v normalised over the keys of the dict s_m1"""
    keys = list(s_m1.keys())
    xs = np.array([k[0] for k in keys], dtype=np.int64)
    ys = np.array([k[1] for k in keys], dtype=np.int64)
    vals = np.array([s_m1[k] for k in keys], dtype=np.float64)
    m2_z = np.zeros_like(m2)
    m2_d = {}
    if keys:
        prod = _mul(np.asarray(m2)[xs, ys], vals)
        normed = _div_by_sum(prod)[0]
        m2_z[xs, ys] = normed
        for k, v in zip(keys, normed):
            m2_d[k] = v
    return m2_z, m2_d


def _as_dense(x):
    return x.toarray() if hasattr(x, 'toarray') else np.asarray(x)


def sd_matrix_multiply(s1, d2):
    """pyx:145-146  s1.dot(d2) (s1 may be a scipy.sparse matrix)"""
    return _dot(np.atleast_2d(_as_dense(s1)), np.atleast_2d(_as_dense(d2)))


def ss_matix_multiply(s1, s2):
    """pyx:153-154 [sic]"""
    return _dot(np.atleast_2d(_as_dense(s1)), np.atleast_2d(_as_dense(s2)))


def make_adapt_phi(phi, num_adaptations):
    """pyx:157-161"""
    adapt_phi = np.zeros((np.shape(phi)[0], np.shape(phi)[1] * (num_adaptations + 1)))
    adapt_phi[:, list(range(0, np.shape(phi)[1]))] = phi
    return adapt_phi


def set_adaptation(f_size, adapt_phi, active_adaptations):
    """pyx:164-173 (in place)"""
    r_0 = list(range(f_size))
    for i in active_adaptations:
        st = i * f_size
        adapt_phi[:, list(range(st, st + f_size))] = adapt_phi[:, r_0]
    return adapt_phi


def set_adaptation_off(f_size, adapt_phi, active_adaptations):
    """pyx:176-184 (in place)"""
    for i in active_adaptations:
        st = i * f_size
        adapt_phi[:, list(range(st, st + f_size))] = 0
    return adapt_phi


def set_original(phi, adapt_phi):
    """pyx:187-190 (in place)"""
    adapt_phi[:, list(range(0, np.shape(phi)[1]))] = phi
    return adapt_phi


def sparse_vec_mat_dot(vec, mat):
    """pyx:193-205: a row vector returns a 1-D (n,) array, a column vector an (n, 1) array"""
    _typed_2d(vec, mat)
    if vec.shape[0] == 1:
        m_idx = np.argpartition(-vec[0, :], K - 1)[:K]
        return _dot(vec[:1, m_idx], mat[m_idx, :])[0]
    m_idx = np.argpartition(-vec[:, 0], K - 1)[:K]
    return _dot(mat[:, m_idx], vec[m_idx])


def induce_s_multiply_threshold(s1, d2):
    """pyx:78-87"""
    raise NotImplementedError("do not use it seems very slow..")


def sd_pointwise_multiply(s1, d2):
    """pyx:149-150"""
    raise NotImplementedError("not implemented pointwise multiply for sparse-dense matrix")
