"""GPU tier, kernel level: each CUDA kernel through the C ABI against NumPy float64 on the same inputs."""
import ctypes

import numpy as np
import pytest
import torch

from macaronicusermodeling_b200 import _lib, build, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def lib():
    build.build()
    return _lib.require_device()


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def S():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def split_planes(x):
    """float64 [R, V] -> (hi, lo) fp16 device tensors padded to ld = roundup(V, 64)"""
    R, V = x.shape
    ld = (V + 63) // 64 * 64
    x32 = x.astype(np.float32)
    hi = x32.astype(np.float16)
    lo = (x32 - hi.astype(np.float32)).astype(np.float16)
    H = torch.zeros((R, ld), dtype=torch.float16)
    L = torch.zeros((R, ld), dtype=torch.float16)
    H[:, :V] = torch.from_numpy(hi)
    L[:, :V] = torch.from_numpy(lo)
    exact = hi.astype(np.float64) + lo.astype(np.float64)
    return H.cuda(), L.cuda(), exact, ld


def gemm(lib, A, B, M, V, impl, a_row0=0, alpha=1.0):
    Ah, Al, Ax, ld = split_planes(A)
    Bh, Bl, Bx, _ = split_planes(B)
    D = torch.full((M + 3, ld), -7.0, dtype=torch.float32, device='cuda')
    _lib.check(lib.mlbp_factor_to_var_gemm(P(Ah), P(Al), A.shape[0], a_row0, M, P(Bh), P(Bl), V, ld, P(D), 1, ld,
                                           alpha, impl, S()))
    torch.cuda.synchronize()
    ref = alpha * (Ax[a_row0:a_row0 + M] @ Bx.T)
    return D.cpu().numpy(), ref, ld


@pytest.mark.parametrize('M,V', [(128, 256), (5, 200), (300, 1000), (131, 2112), (700, 2500), (257, 4100)])
@pytest.mark.parametrize('impl', [1, 0, 30, 11], ids=['simt', 'tcgen05', 'tcgen05-pair', 'tcgen05-single'])
def test_message_gemm_matches_float64(lib, M, V, impl):
    rng = np.random.default_rng(M * 7 + V)
    A = rng.random((M + 9, V)) * 2.0 ** 14 / V * 2       # message-like magnitudes (2^14 * normalised)
    B = np.exp(rng.normal(size=(V, V)) * 0.7) * 8.0       # table-like magnitudes
    D, ref, ld = gemm(lib, A, B, M, V, impl, a_row0=4, alpha=2.0 ** -17)
    got = D[1:1 + M, :V]
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 3e-6, err
    assert (D[0] == -7.0).all() and (D[1 + M:] == -7.0).all(), 'rows outside [d_row0, d_row0 + M) were written'
    assert _lib.load().mlbp_gemm_barrier_timeout_code() == 0


@pytest.mark.parametrize('M,V', [(131, 2112), (300, 1000), (257, 4100)])
@pytest.mark.parametrize('impl', [1, 0, 30, 11], ids=['simt', 'tcgen05', 'tcgen05-pair', 'tcgen05-single'])
def test_gradient_gemm_a_hi_only(lib, M, V, impl):
    """MLBP_GEMM_A_HI_ONLY (gradient rows): exactly  A_hi . (B_hi + B_lo)'  -- the A_lo plane must not contribute"""
    rng = np.random.default_rng(M + V)
    A = rng.random((M + 9, V)) * 2.0 ** 14 / V * 2
    B = np.exp(rng.normal(size=(V, V)) * 0.7) * 8.0
    Ah, Al, Ax, ld = split_planes(A)
    Bh, Bl, Bx, _ = split_planes(B)
    D = torch.full((M + 3, ld), -7.0, dtype=torch.float32, device='cuda')
    _lib.check(lib.mlbp_factor_to_var_gemm(P(Ah), P(Al), A.shape[0], 4, M, P(Bh), P(Bl), V, ld, P(D), 1, ld,
                                           2.0 ** -17, impl | 256, S()))
    torch.cuda.synchronize()
    ref = 2.0 ** -17 * (Ah.cpu().numpy()[4:4 + M, :V].astype(np.float64) @ Bx.T)
    got = D.cpu().numpy()[1:1 + M, :V]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-6
    full = 2.0 ** -17 * (Ax[4:4 + M] @ Bx.T)
    assert np.abs(got - full).max() / np.abs(full).max() > 1e-6, 'the A_lo term was not dropped'
    assert _lib.load().mlbp_gemm_barrier_timeout_code() == 0


@pytest.mark.parametrize('impl', [1, 0, 30, 11], ids=['simt', 'tcgen05', 'tcgen05-pair', 'tcgen05-single'])
def test_gradient_gemm_one_pass(lib, impl):
    """MLBP_GEMM_A_HI_ONLY | MLBP_GEMM_B_HI_ONLY: exactly  A_hi . B_hi'  (plain fp16 operands, fp32 accumulation)"""
    M, V = 257, 4100
    rng = np.random.default_rng(5)
    A = rng.random((M + 9, V)) * 2.0 ** 14 / V * 2
    B = np.exp(rng.normal(size=(V, V)) * 0.7) * 8.0
    Ah, Al, Ax, ld = split_planes(A)
    Bh, Bl, Bx, _ = split_planes(B)
    D = torch.full((M + 3, ld), -7.0, dtype=torch.float32, device='cuda')
    _lib.check(lib.mlbp_factor_to_var_gemm(P(Ah), P(Al), A.shape[0], 4, M, P(Bh), P(Bl), V, ld, P(D), 1, ld,
                                           2.0 ** -17, impl | 256 | 512, S()))
    torch.cuda.synchronize()
    ref = 2.0 ** -17 * (Ah.cpu().numpy()[4:4 + M, :V].astype(np.float64) @ Bh.cpu().numpy()[:, :V].astype(np.float64).T)
    got = D.cpu().numpy()[1:1 + M, :V]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-6
    assert _lib.load().mlbp_gemm_barrier_timeout_code() == 0


def test_message_gemm_peaked_rows(lib):
    """near-delta messages: one element carries (almost) all mass -- the split must keep the small ones"""
    rng = np.random.default_rng(3)
    M, V = 130, 1536
    A = rng.random((M, V)) * 1e-4
    A[np.arange(M), rng.integers(V, size=M)] = 2.0 ** 14
    B = np.exp(rng.normal(size=(V, V)))
    D, ref, ld = gemm(lib, A, B, M, V, 0)
    got = D[1:1 + M, :V]
    assert (np.abs(got - ref) / np.abs(ref)).max() < 5e-6


def test_pairwise_tables(lib):
    model = synth.make_model(200, 40, seed=1, w1_density=0.4)
    V, ld = 200, 256
    te = np.array([0.7, -0.4, 0.2])
    pmi = torch.zeros((V, ld)); pmi[:, :V] = torch.from_numpy(model['pmi'].astype(np.float32))
    w1 = torch.zeros((V, ld)); w1[:, :V] = torch.from_numpy(model['pmi_w1'].astype(np.float32))
    pmi, w1 = pmi.cuda(), w1.cuda()
    planes = torch.zeros((14, V, ld), dtype=torch.float16, device='cuda')
    cols = torch.zeros((7, V), dtype=torch.float64, device='cuda')
    s = 9
    _lib.check(lib.mlbp_build_pairwise_tables(P(pmi), P(w1), V, ld, te.ctypes.data_as(ctypes.c_void_p), s, P(planes),
                                              V * ld, ld, P(cols), 1, S()))
    torch.cuda.synchronize()
    pl = planes.cpu().numpy().astype(np.float64)
    p32, w32 = model['pmi'].astype(np.float32).astype(np.float64), model['pmi_w1'].astype(np.float32).astype(np.float64)
    T = np.exp(te[0] * p32 + te[2]); T1 = np.exp(te[0] * p32 + te[1] * w32 + te[2])
    want = [T, T.T, T1, T1.T, T * p32, T1 * p32, T1 * w32]
    for i, W in enumerate(want):
        got = (pl[2 * i] + pl[2 * i + 1])[:, :V] * 2.0 ** -s
        assert np.abs(got - W).max() / W.max() < 1e-6, i
        assert (pl[2 * i][:, V:] == 0).all()
    c = cols.cpu().numpy()
    for i, W in enumerate([T, T1, T * p32, T1 * p32, T1 * w32]):
        np.testing.assert_allclose(c[i], W.sum(0), rtol=1e-12)
    np.testing.assert_allclose(c[5], T.sum(1), rtol=1e-12)
    np.testing.assert_allclose(c[6], T1.sum(1), rtol=1e-12)


def test_dense_array_utils(lib):
    z = np.load(__import__('os').path.join(__import__('os').path.dirname(__file__), 'golden', 'au_cases.npz'))
    from macaronicusermodeling_b200.array_utils import c_array_utils as au
    np.testing.assert_allclose(au.pointwise_multiply(z['a'], z['b']), z['pointwise_multiply'], rtol=1e-15)
    np.testing.assert_allclose(au.normalize(z['a'].copy()), z['normalize'], rtol=1e-13)
    zz = np.zeros((5, 1))
    assert (au.normalize(zz) == 0).all()
    np.testing.assert_allclose(au.dense_dot(z['T'], z['a']), z['dense_dot_Tv'], rtol=1e-12)
    np.testing.assert_allclose(au.dense_dot(np.ascontiguousarray(z['a'].T), z['T']), z['dense_dot_vT'], rtol=1e-12)
    np.testing.assert_allclose(au.dense_dot(z['a'], np.ascontiguousarray(z['b'].T)), z['dense_dot_outer'], rtol=1e-15)
    np.testing.assert_allclose(au.dense_pointwise_multiply(z['T'], z['T'].T.copy()), z['dense_pointwise_multiply'], rtol=1e-15)
