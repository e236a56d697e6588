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
@pytest.mark.parametrize('impl', [1, 0, 2, 3], ids=['simt', 'tcgen05', 'tcgen05-pair', 'tcgen05-single'])
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
@pytest.mark.parametrize('impl', [1, 0, 2, 3], ids=['simt', 'tcgen05', 'tcgen05-pair', 'tcgen05-single'])
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


@pytest.mark.parametrize('impl', [1, 0, 2, 3], ids=['simt', 'tcgen05', 'tcgen05-pair', 'tcgen05-single'])
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
    planes = torch.zeros((20, V, ld), dtype=torch.float16, device='cuda')
    cols = torch.zeros((7, V), dtype=torch.float64, device='cuda')
    s = 9
    rpl = torch.zeros((4, V, ld), dtype=torch.float16, device='cuda')
    tbar = np.zeros(1, dtype=np.float32)
    _lib.check(lib.mlbp_build_pairwise_tables(P(pmi), P(w1), V, ld, te.ctypes.data_as(ctypes.c_void_p), s, P(planes),
                                              V * ld, ld, P(cols), 1, P(rpl), tbar.ctypes.data_as(ctypes.c_void_p), S()))
    torch.cuda.synchronize()
    pl = planes.cpu().numpy().astype(np.float64)
    # residual planes R = T - tbar, R1 = T1 - tbar and their transposes (hi halves only: 11 bits of the RESIDUAL), tbar = T(phi = 0)
    np.testing.assert_allclose(tbar[0], np.exp(te[2]) * 2.0 ** s, rtol=1e-6)
    rp = rpl.cpu().numpy().astype(np.float64)
    p32, w32 = model['pmi'].astype(np.float32).astype(np.float64), model['pmi_w1'].astype(np.float32).astype(np.float64)
    T = np.exp(te[0] * p32 + te[2]); T1 = np.exp(te[0] * p32 + te[1] * w32 + te[2])
    want = [T, T.T, T1, T1.T, T * p32, T1 * p32, T1 * w32, (T * p32).T, (T1 * p32).T, (T1 * w32).T]
    for i, W in enumerate(want):
        got = (pl[2 * i] + pl[2 * i + 1])[:, :V] * 2.0 ** -s
        assert np.abs(got - W).max() / W.max() < 1e-6, i
        assert (pl[2 * i][:, V:] == 0).all()
    for i, W in enumerate(want[:4]):
        R = W * 2.0 ** s - float(tbar[0])
        assert np.abs(rp[i][:, :V] - R).max() <= 2.0 ** -10 * np.abs(R).max() + 1e-6, i      # one fp16 ulp (stochastic rounding)
        assert (rp[i][:, V:] == 0).all()
    c = cols.cpu().numpy()
    # K2 evaluates exp in fp32 on a compensated argument (~1 ulp per entry, random) and sums in compensated fp32 / float64:
    # column sums over V = 200 entries are good to ~3e-8 (measured; expf's rounding is not quite unbiased), row sums (plain
    # fp32 across the 64 columns of a tile) to ~1e-7.
    # (Round 1 used a float64 exp per element and asserted 1e-12 here; the consumers of these sums -- the unary gradient and the
    # constant messages -- need 1e-6.)
    for i, W in enumerate([T, T1, T * p32, T1 * p32, T1 * w32]):
        np.testing.assert_allclose(c[i], W.sum(0), rtol=1e-7)
    np.testing.assert_allclose(c[5], T.sum(1), rtol=2e-7)
    np.testing.assert_allclose(c[6], T1.sum(1), rtol=2e-7)


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


def test_gated_gemm_runs_only_when_the_device_flag_matches(lib):
    """mlbp_factor_to_var_gemm_gated: every CTA reads the gate word and returns unless (*gate != 0) == run_if_set"""
    M, V = 300, 2500
    rng = np.random.default_rng(8)
    A = rng.random((M, V)) * 2.0 ** 14 / V * 2
    B = np.exp(rng.normal(size=(V, V)) * 0.5) * 8.0
    Ah, Al, Ax, ld = split_planes(A)
    Bh, Bl, Bx, _ = split_planes(B)
    ref3 = Ax @ Bx.T
    ref2 = Ah.cpu().numpy()[:, :V].astype(np.float64) @ Bx.T
    gate = torch.zeros(4, dtype=torch.int32, device='cuda')
    for flag in (0, 1):
        gate[0] = flag
        D = torch.full((M, ld), -7.0, dtype=torch.float32, device='cuda')
        for impl, run_if_set in ((256, 0), (0, 1)):               # what Engine issues for one slice of message rows
            _lib.check(lib.mlbp_factor_to_var_gemm_gated(P(Ah), P(Al), M, 0, M, P(Bh), P(Bl), V, ld, P(D), 0, ld, 1.0, impl,
                                                         P(gate), run_if_set, 0, 0, 0.0, S()))
        torch.cuda.synchronize()
        got = D.cpu().numpy()[:, :V]
        want, other = (ref3, ref2) if flag else (ref2, ref3)
        assert np.abs(got - want).max() / np.abs(want).max() < 3e-6
        assert np.abs(got - other).max() / np.abs(other).max() > 1e-6, 'the wrong variant ran'
    # nothing runs when neither launch matches
    D = torch.full((M, ld), -7.0, dtype=torch.float32, device='cuda')
    gate[0] = 1
    _lib.check(lib.mlbp_factor_to_var_gemm_gated(P(Ah), P(Al), M, 0, M, P(Bh), P(Bl), V, ld, P(D), 0, ld, 1.0, 256, P(gate), 0, 0, 0, 0.0, S()))
    torch.cuda.synchronize()
    assert (D == -7.0).all()
    assert _lib.load().mlbp_gemm_barrier_timeout_code() == 0


def test_marginals_flags_and_exact_rescore(lib):
    """K5 near-tie detection + K5b: a variable whose D rows were computed WITHOUT the lo half of A has its near-tied top-2
    swapped on purpose; the re-score recomputes the last hop from the full operands and restores the float64 decision."""
    V, ld, n_in = 2304, 2304, 3
    rng = np.random.default_rng(21)
    A = rng.random((n_in, V)) * 2.0 ** 14 / V * 2
    T = np.exp(rng.random((V, V)) * 0.8)
    Ah, Al, Ax, _ = split_planes(A)
    Th, Tl, Tx, _ = split_planes(T * 64.0)
    planes = torch.zeros((8, V, ld), dtype=torch.float16, device='cuda')
    planes[2], planes[3] = Th, Tl                                  # table id 1 (MLBP_TABLE_TT)
    exact = Ax @ Tx.T                                              # the three-pass rows, float64
    two = Ah.cpu().numpy()[:, :V].astype(np.float64) @ Tx.T        # what a two-pass GEMM writes
    U = rng.random(V) + 0.5
    belief = U * exact.prod(axis=0)
    a, b = np.argsort(belief)[-2:][::-1]                           # arg-max a, runner-up b
    # nudge U so that a beats b by 2e-6 on exact rows while the two-pass rows say the opposite (if they do not already)
    U[b] *= belief[a] / belief[b] / (1.0 + 2e-6)
    Uf = U.astype(np.float32)
    bel_exact = Uf.astype(np.float64) * exact.prod(axis=0)
    want = int(np.argmax(bel_exact))
    D = torch.zeros((5 + n_in, ld), dtype=torch.float32, device='cuda')
    D[0] = 1.0
    D[5:, :V] = torch.from_numpy(two.astype(np.float32)).cuda()
    Ud = torch.zeros((1, ld), dtype=torch.float32, device='cuda'); Ud[0, :V] = torch.from_numpy(Uf).cuda()
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device='cuda')
    grp_u, grp_off, in_row = i32([0]), i32([0, n_in]), i32([5, 6, 7])
    label = i32([int(np.argsort(belief)[-30])])                    # a label inside the top-50: its rank is re-scored too
    logp = torch.zeros(1, dtype=torch.float64, device='cuda')
    top1, rank = i32([0]), i32([0])
    aux = torch.zeros((1, 2), dtype=torch.float64, device='cuda')
    cnts, flags, flagged = torch.zeros((1, 2), dtype=torch.int32, device='cuda'), i32([0]), i32([0])
    words = torch.zeros(16, dtype=torch.int32, device='cuda')
    tau, tau_label = 2e-3, 5e-4
    _lib.check(lib.mlbp_marginals(1, P(grp_u), P(grp_off), P(in_row), P(label), P(Ud), P(D), ld, V, P(logp), P(top1), P(rank), None,
                                  50.0, n_in, tau, tau_label, P(aux), P(cnts), P(flags), P(flagged), P(words[1:]), S()))
    torch.cuda.synchronize()
    bel_two = Uf.astype(np.float64) * two.astype(np.float32).astype(np.float64).prod(axis=0)
    assert int(top1[0]) == int(np.argmax(bel_two))
    assert int(flags[0]) & 1 and int(words[1]) == 1 and int(flagged[0]) == 0
    blocks = i32([1, 0, 5, n_in])                                  # one message block: table 1, A row 0, D row 5, n_in rows
    _lib.check(lib.mlbp_rescore_candidates(1, P(flagged), P(words[1:]), P(flags), P(grp_u), P(grp_off), P(in_row), P(label), P(Ud),
                                           P(D), ld, V, P(Ah), P(Al), P(planes), V * ld, P(blocks), 1, n_in, P(aux), P(cnts), tau,
                                           tau_label, 50.0, P(top1), P(rank), P(words[8:]), S()))
    torch.cuda.synchronize()
    assert int(top1[0]) == want, (int(top1[0]), want, a, b)
    lab = int(label[0])
    assert int(rank[0]) == int((bel_exact > bel_exact[lab]).sum())
    assert int(words[8]) == 1 and int(words[9]) == 0 and int(words[10]) == 0


def test_var_to_factor_records_spikes(lib):
    """K3 records the elements that carry more than peak_mult / V of a message's mass (SPIKE word, per-row slots) and keeps the
    largest element it wrote; flat messages record nothing"""
    from macaronicusermodeling_b200.engine import Corpus, Engine
    model = synth.make_model(2304, 64, seed=3, dtype=np.float32)
    sents = synth.make_corpus(model, 3, k=5, g=1, seed=4)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(synth.draw_roots(sents, 3, seed=5))
    for td, expect in (([0.9, -0.5, 0.5, 0.3, 0.4, -0.2], 0), ([0.9, -0.5, 9.0, 9.0, 3.0, -0.2], 1)):
        eng = Engine(model, msg_passes=2)
        eng.set_theta([0.7, 0.4, -0.2], td)
        eng.run(corpus, roots, 3)
        st = eng.pass_stats()
        assert st['msg_two_pass'] and st['spike_flag'] == expect and st['peak_flag'] == 0, st
        assert (st['spiky_rows_last_batch'] > 0) == bool(expect)
        assert (st['max_message_prob'] > 16.0 / 2304) == bool(expect), st


def test_spike_scan_files_the_spikes_of_a_block(lib):
    """K4a: per row the elements above the limit as (column, A_lo value) in ascending columns, their count, the block's list of
    spiky rows; five spikes in one row raise PEAK; rows outside the block and the uninitialised row padding are not read"""
    M, V = 70, 2500
    rng = np.random.default_rng(2)
    A = rng.random((M, V)) * 2.0 ** 14 / V * 2
    want = {5: [17, 900, 2499], 31: [0], 64: [3, 4, 5, 6]}
    for r, cols in want.items():
        for c in cols:
            A[r, c] = 2.0 ** 14 * 0.05 * (1 + 0.01 * c / V)
    A[2, 100] = 2.0 ** 14 * 0.5                                    # row 2 is outside the scanned block
    Ah, Al, Ax, ld = split_planes(A)
    Ah[:, V:] = float('nan')                                       # the padding of A rows is never written: poison it
    words = torch.zeros(8, dtype=torch.int32, device='cuda')
    cnt = torch.full((M,), -1, dtype=torch.int32, device='cuda')
    ent = torch.zeros((M, 4, 2), dtype=torch.int32, device='cuda')
    rows = torch.full((M,), -1, dtype=torch.int32, device='cuda')
    n_list = torch.zeros(1, dtype=torch.int32, device='cuda')
    a0, n = 4, 62
    _lib.check(lib.mlbp_spike_scan(P(Ah), P(Al), ld, V, a0, n, 16.0 / V, P(words), P(cnt), P(ent), P(rows), P(n_list), S()))
    torch.cuda.synchronize()
    c, e, w = cnt.cpu().numpy(), ent.cpu().numpy(), words.cpu().numpy()
    assert (c[:a0] == -1).all() and (c[a0 + n:] == -1).all()
    lo = Al.cpu().numpy().astype(np.float32)
    for r in range(a0, a0 + n):
        cols = want.get(r, [])
        assert c[r] == len(cols), (r, c[r])
        for i, col in enumerate(cols):
            assert e[r, i, 0] == col and e[r, i, 1:2].view(np.float32)[0] == lo[r, col]
    assert sorted(rows.cpu().numpy()[:int(n_list[0])].tolist()) == sorted(want) and int(n_list[0]) == 3
    assert w[0] == 0 and w[3] == 1 and w[4] == 3
    A[40, [1, 2, 3, 4, 5]] = 2.0 ** 14 * 0.05
    Ah2, Al2, _, _ = split_planes(A)
    _lib.check(lib.mlbp_spike_scan(P(Ah2), P(Al2), ld, V, a0, n, 16.0 / V, P(words), P(cnt), P(ent), None, None, S()))
    torch.cuda.synchronize()
    assert int(words[0]) == 1 and int(cnt[40]) == 5


def test_spike_correct_restores_the_dropped_lo_part(lib):
    """two-pass GEMM (A_hi only) + mlbp_spike_correct on the recorded spikes == the full product up to the rounding of the
    un-spiky remainder; rows outside the block and rows without spikes are untouched; a set PEAK word disables it"""
    M, V = 300, 2500
    rng = np.random.default_rng(12)
    A = rng.random((M, V)) * 2.0 ** 14 / V * 0.4                  # un-spiky remainder: 20 % of the mass
    spikes = {}
    for r in (3, 17, 130, 299):
        cols = sorted(rng.choice(V, size=int(rng.integers(1, 5)), replace=False).tolist())
        for c in cols:
            h = np.float64(np.float16(2.0 ** 14 * (0.1 + 0.15 * rng.random())))
            A[r, c] = h * (1.0 + 0.4 * 2.0 ** -11)                  # 0.4 ulp above an fp16 value: a large, known lo part
        spikes[r] = cols
    T = np.exp(rng.normal(size=(V, V)) * 0.5) * 8.0               # B[n, k]; its transpose is the "other orientation" plane pair
    Ah, Al, Ax, ld = split_planes(A)
    Bh, Bl, Bx, _ = split_planes(T)
    Th, Tl, Tx, _ = split_planes(np.ascontiguousarray(T.T))
    words = torch.zeros(8, dtype=torch.int32, device='cuda')
    cnt = torch.zeros(M + 8, dtype=torch.int32)
    ent = torch.zeros((M + 8, 4, 2), dtype=torch.int32)
    rows = torch.zeros(M + 8, dtype=torch.int32)
    a0 = 2                                                         # the block starts at A row 2: rows 0, 1 belong to another block
    x32 = A.astype(np.float32)
    lo_exact = x32 - x32.astype(np.float16).astype(np.float32)
    n_sp = 0
    for r, cols in spikes.items():
        for c in cols:                                             # ascending columns, as mlbp_spike_scan files them
            ent[r, cnt[r], 0] = c
            ent[r, cnt[r], 1] = int(np.float32(lo_exact[r, c]).view(np.int32))
            cnt[r] += 1
        rows[n_sp] = r
        n_sp += 1
    words[4] = n_sp
    n_list = torch.tensor([n_sp], dtype=torch.int32, device='cuda')
    cnt, ent, rows = cnt.cuda(), ent.cuda(), rows.cuda()
    D = torch.full((M + 3, ld), -7.0, dtype=torch.float32, device='cuda')
    n_blk = M - a0
    _lib.check(lib.mlbp_factor_to_var_gemm_gated(P(Ah), P(Al), M, a0, n_blk, P(Bh), P(Bl), V, ld, P(D), 1, ld, 0.5, 256, P(words), 0, 0, 0, 0.0, S()))
    before = D.clone()
    _lib.check(lib.mlbp_spike_correct(P(words), P(cnt), P(ent), P(rows), P(n_list), a0, n_blk, P(Th), P(Tl), V, ld, P(D), 1, ld, 0.5, None, None, 0.0, S()))
    torch.cuda.synchronize()
    got, was = D.cpu().numpy(), before.cpu().numpy()
    full = 0.5 * (Ax[a0:] @ Bx.T)
    two = 0.5 * (Ah.cpu().numpy()[a0:, :V].astype(np.float64) @ Bx.T)
    for r in range(a0, M):
        i = 1 + r - a0
        if r in spikes:
            e_before = np.abs(was[i, :V] - full[r - a0]).max() / np.abs(full[r - a0]).max()
            e_after = np.abs(got[i, :V] - full[r - a0]).max() / np.abs(full[r - a0]).max()
            assert e_after < 5e-6 and e_before > 2e-5, (r, e_before, e_after)
        else:
            np.testing.assert_array_equal(got[i], was[i])
    assert (got[0] == -7.0).all() and (got[1 + n_blk:] == -7.0).all()
    assert (got[:, V:ld][1:1 + n_blk] == 0).all()                  # row padding stays zero
    # PEAK set: the block ran three passes, the correction must not touch it
    words[0] = 1
    D2 = before.clone()
    _lib.check(lib.mlbp_spike_correct(P(words), P(cnt), P(ent), P(rows), P(n_list), a0, n_blk, P(Th), P(Tl), V, ld, P(D2), 1, ld, 0.5, None, None, 0.0, S()))
    torch.cuda.synchronize()
    assert torch.equal(D2, before)
    # ONE-pass rows (A_hi . B_hi): with A_hi passed the correction also restores hi_s * B_lo[:, col_s] at the spikes
    words[0] = 0
    D3 = torch.full((M + 3, ld), -7.0, dtype=torch.float32, device='cuda')
    _lib.check(lib.mlbp_factor_to_var_gemm_gated(P(Ah), P(Al), M, a0, n_blk, P(Bh), P(Bl), V, ld, P(D3), 1, ld, 0.5, 256 | 512, P(words), 0, 0, 0, 0.0, S()))
    was3 = D3.cpu().numpy()
    _lib.check(lib.mlbp_spike_correct(P(words), P(cnt), P(ent), P(rows), P(n_list), a0, n_blk, P(Th), P(Tl), V, ld, P(D3), 1, ld, 0.5, P(Ah), None, 0.0, S()))
    torch.cuda.synchronize()
    got3 = D3.cpu().numpy()
    ah, bh = Ah.cpu().numpy()[:, :V].astype(np.float64), Bh.cpu().numpy()[:, :V].astype(np.float64)
    for r, cols in spikes.items():
        i = 1 + r - a0
        want = was3[i, :V].astype(np.float64)
        for c in cols:                                             # exact spike term minus what the one-pass product held of it
            want += 0.5 * (Ax[r, c] * Bx[:, c] - ah[r, c] * bh[:, c])
        assert np.abs(got3[i, :V] - want).max() / np.abs(want).max() < 2e-6, r
    # RESIDUAL-plane one-pass rows: B_hi := fp16(T - tbar), the GEMM adds the constant alpha * tbar * sum(row); at the spikes the
    # correction brings the element's term to (hi + lo) * (T - tbar) exactly
    tbar = float(np.median(T))
    Rh = torch.zeros((V, ld), dtype=torch.float16); Rh[:, :V] = torch.from_numpy((T - tbar).astype(np.float32).astype(np.float16))
    Rth = torch.zeros((V, ld), dtype=torch.float16); Rth[:, :V] = torch.from_numpy((T.T - tbar).astype(np.float32).astype(np.float16))
    Rh, Rth = Rh.cuda(), Rth.cuda()
    row_sum = 1000.0                                               # any constant: the test rows do not sum to 2^14
    D4 = torch.full((M + 3, ld), -7.0, dtype=torch.float32, device='cuda')
    _lib.check(lib.mlbp_factor_to_var_gemm_gated(P(Ah), P(Al), M, a0, n_blk, P(Rh), P(Bl), V, ld, P(D4), 1, ld, 0.5, 256 | 512, P(words), 0, 0, 0,
                                                 0.5 * tbar * row_sum, S()))
    was4 = D4.cpu().numpy()
    rh = Rh.cpu().numpy()[:, :V].astype(np.float64)
    plain = 0.5 * (ah[a0:] @ rh.T) + 0.5 * tbar * row_sum
    assert np.abs(was4[1:1 + n_blk, :V] - plain).max() / np.abs(plain).max() < 3e-6          # D = alpha A_hi R_hi' + add_const
    _lib.check(lib.mlbp_spike_correct(P(words), P(cnt), P(ent), P(rows), P(n_list), a0, n_blk, P(Th), P(Tl), V, ld, P(D4), 1, ld, 0.5, P(Ah),
                                      P(Rth), tbar, S()))
    torch.cuda.synchronize()
    got4 = D4.cpu().numpy()
    for r, cols in spikes.items():
        i = 1 + r - a0
        want = was4[i, :V].astype(np.float64)
        for c in cols:
            want += 0.5 * (Ax[r, c] * (Bx[:, c] - tbar) - ah[r, c] * rh[:, c])
        assert np.abs(got4[i, :V] - want).max() / np.abs(want).max() < 2e-6, r


def test_gemm_k_ranges_accumulate(lib):
    """a long contraction issued as several launches over K ranges (the engine does it at V = 50 000): ranges that do not start
    at 0 add to D; the result equals one launch up to the fp32 rounding of the partial sums"""
    M, V = 300, 2500
    rng = np.random.default_rng(4)
    A = rng.random((M, V)) * 2.0 ** 14 / V * 2
    B = np.exp(rng.normal(size=(V, V)) * 0.5) * 8.0
    Ah, Al, Ax, ld = split_planes(A)
    Bh, Bl, Bx, _ = split_planes(B)
    ref = Ax @ Bx.T
    D = torch.full((M, ld), -7.0, dtype=torch.float32, device='cuda')
    for k0, k_len in ((0, 896), (896, 896), (1792, 708)):
        _lib.check(lib.mlbp_factor_to_var_gemm_gated(P(Ah), P(Al), M, 0, M, P(Bh), P(Bl), V, ld, P(D), 0, ld, 1.0, 2, None, 0,
                                                     k0, k_len if k0 + k_len < V else 0, 0.0, S()))
    torch.cuda.synchronize()
    got = D.cpu().numpy()[:, :V]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-6
    assert (D.cpu().numpy()[:, V:] == 0).all()
    assert _lib.load().mlbp_gemm_barrier_timeout_code() == 0


@pytest.mark.parametrize('V,K', [(10000, 50), (67, 50), (2049, 1), (513, 513), (5000, 1024)])
def test_topk_rows_matches_numpy(lib, V, K):
    """mlbp_topk_rows == the reference's np.argpartition + np.argsort list (LBP.py:402-411) on rows without exact ties;
    rows WITH exact ties are flagged and listed in (value descending, index ascending) order"""
    rng = np.random.default_rng(V + K)
    n = 37
    ld = (V + 63) // 64 * 64
    X = np.zeros((n, ld), dtype=np.float32)
    X[:, :V] = rng.random((n, V), dtype=np.float32) ** 8                # a peaked, tie-free spread
    X[:, :V] /= X[:, :V].sum(axis=1, keepdims=True)
    X[:, V:] = 7.0                                                      # row padding must never be listed
    X[3, :V] = 1.0 / V                                                  # all equal
    if V > 200:
        X[5, 100:110] = X[5, :V].max() * 2                               # ten equal maxima
        X[7, :V] = 0; X[7, 11] = 1.0                                     # a delta: zeros tie at the boundary (when K > 1)
    x = torch.from_numpy(X).cuda()
    idx = torch.empty((n, K), dtype=torch.int32, device='cuda')
    val = torch.empty((n, K), dtype=torch.float32, device='cuda')
    ties = torch.empty(n, dtype=torch.int32, device='cuda')
    _lib.check(lib.mlbp_topk_rows(P(x), ld, V, n, K, P(idx), P(val), P(ties), S()))
    torch.cuda.synchronize()
    I, Pv, T = idx.cpu().numpy(), val.cpu().numpy(), ties.cpu().numpy()
    for r in range(n):
        a = X[r, :V]
        want = np.lexsort((np.arange(V), -a))[:K]
        np.testing.assert_array_equal(I[r], want, err_msg='row %d' % r)
        np.testing.assert_array_equal(Pv[r], a[want])
        exact_ties = (np.diff(a[want]) == 0).any() or (K < V and np.sort(a)[::-1][K] == a[want][-1])
        assert (T[r] > 0) == bool(exact_ties), (r, T[r])
        if not exact_ties:                                              # the reference's own calls give the same list
            top = np.argpartition(a, -K)[-K:]
            top = top[np.argsort(a[top])][::-1]
            np.testing.assert_array_equal(I[r], top)
    assert lib.mlbp_topk_rows(P(x), ld, V, n, V + 1, P(idx), P(val), P(ties), S()) != 0   # K > V is rejected
