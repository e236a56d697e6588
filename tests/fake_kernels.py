"""TEST DOUBLE for libmlbp.so's device kernels: the same C-ABI calls, emulated with NumPy on host pointers.

Lets the CPU-only test tier run the REAL host logic (engine.py orchestration + the C++ schedule compiler in
csrc/plan.cpp) end to end against the oracle without a GPU.  It mirrors the kernels' storage formats (fp16 hi/lo
split, 2^14 message scale, dropped lo*lo term), not their speed.  Never imported by the package."""
import ctypes

import numpy as np
import torch

A_SCALE = 2.0 ** 14


def _arr(ptr, dtype, count):
    if ptr is None:
        return None
    addr = ptr.value if isinstance(ptr, ctypes.c_void_p) else int(ptr)
    if addr is None:
        return None
    dt = np.dtype(dtype)
    buf = (ctypes.c_char * (count * dt.itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dt, count=count)


def _split(x):
    x32 = x.astype(np.float32)
    hi = x32.astype(np.float16)
    lo = (x32 - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def _cell_hash(a, b):
    """csrc/tables.cu cell_hash"""
    with np.errstate(over='ignore'):
        h = (a.astype(np.uint32) * np.uint32(0x9E3779B1) + b.astype(np.uint32) * np.uint32(0x85EBCA77)).astype(np.uint32)
        h ^= h >> np.uint32(15); h = (h * np.uint32(0x2C1B3C6D)).astype(np.uint32)
        h ^= h >> np.uint32(12); h = (h * np.uint32(0x297A2D39)).astype(np.uint32)
        h ^= h >> np.uint32(15)
    return h


def _split_stochastic(x, a, b, plane):
    """K2's table split: the hi half is rounded stochastically (csrc/tables.cu half_stochastic), lo = RN(x - hi)"""
    x32 = x.astype(np.float32)
    r = (_cell_hash(a, b) >> np.uint32(3 * plane)) & np.uint32(0x1FFF)
    bits = ((x32.view(np.uint32) + r) & np.uint32(0xFFFFE000)).astype(np.uint32)
    hi = bits.view(np.float32).astype(np.float16)          # exact: the low 13 mantissa bits are zero
    lo = (x32 - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


class FakeKernels(object):
    def __init__(self):
        self.device = torch.device('cpu')
        self.calls = []

    def call(self, name, *args):
        self.calls.append(name)
        getattr(self, name)(*args)

    # ---- K2
    def mlbp_build_pairwise_tables(self, pmi, w1, V, ldf, th, scale_exp, planes, ps, ldv, colsums, with_grad, r_planes=None,
                                   h_tbar=None):
        P = _arr(pmi, np.float32, V * ldf).reshape(V, ldf)[:, :V].astype(np.float64)
        W = _arr(w1, np.float32, V * ldf).reshape(V, ldf)[:, :V].astype(np.float64)
        t = _arr(th, np.float64, 3)
        T = np.exp(t[0] * P + t[2]); T1 = np.exp(t[0] * P + t[1] * W + t[2])
        mats = [T, T.T, T1, T1.T, T * P, T1 * P, T1 * W, (T * P).T, (T1 * P).T, (T1 * W).T]
        pl = _arr(planes, np.float16, (20 if with_grad else 8) * ps)
        ai, bi = np.meshgrid(np.arange(V), np.arange(V), indexing='ij')
        base = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 3, 6: 4, 7: 2, 8: 3, 9: 4}      # which of the five tables a plane pair holds
        for i, M in enumerate(mats[: 10 if with_grad else 4]):
            transposed = i in (1, 3, 7, 8, 9)
            src = M.T if transposed else M                                       # the cell (a, b) of the un-transposed table
            hi, lo = _split_stochastic(np.ldexp(src, scale_exp), ai, bi, base[i])
            if transposed:
                hi, lo = hi.T, lo.T
            for j, h in enumerate((hi, lo)):
                v = pl[(2 * i + j) * ps:(2 * i + j) * ps + V * ldv].reshape(V, ldv)
                v[:, :V] = h
        cs = _arr(colsums, np.float64, 7 * V).reshape(7, V)
        cs[0], cs[1], cs[2], cs[3], cs[4] = T.sum(0), T1.sum(0), (T * P).sum(0), (T1 * P).sum(0), (T1 * W).sum(0)
        cs[5], cs[6] = T.sum(1), T1.sum(1)
        tbar = np.float32(np.exp(t[2]) * 2.0 ** scale_exp)
        if h_tbar is not None and getattr(h_tbar, 'value', h_tbar):
            if _arr(h_tbar, np.float32, 1)[0] > 0:
                tbar = np.float32(_arr(h_tbar, np.float32, 1)[0])
            _arr(h_tbar, np.float32, 1)[0] = tbar
        if r_planes is not None and getattr(r_planes, 'value', r_planes):
            # residual planes R = T - tbar, R1 = T1 - tbar (hi halves, stochastic rounding of the magnitude), both orientations
            rp = _arr(r_planes, np.float16, 4 * ps)
            for i, M in enumerate((T, T1)):
                x = (np.ldexp(M, scale_exp).astype(np.float32) - tbar).astype(np.float32)
                hi, _ = _split_stochastic(np.abs(x).astype(np.float64), ai ^ 0x5bd1e995, bi, 0)
                hi = np.where(x < 0, -hi, hi).astype(np.float16)
                rp[(2 * i) * ps:(2 * i) * ps + V * ldv].reshape(V, ldv)[:, :V] = hi
                rp[(2 * i + 1) * ps:(2 * i + 1) * ps + V * ldv].reshape(V, ldv)[:, :V] = hi.T

    def mlbp_build_unary_tables(self, edT, pedT, V, Vd, ldf, th, edstats):
        E = _arr(edT, np.float32, Vd * ldf).reshape(Vd, ldf)[:, :V].astype(np.float64)
        Pd = _arr(pedT, np.float32, Vd * ldf).reshape(Vd, ldf)[:, :V].astype(np.float64)
        t = _arr(th, np.float64, 6)
        psi = np.exp(t[0] * E + t[1] * Pd + t[5])
        st = _arr(edstats, np.float64, Vd * 3).reshape(Vd, 3)
        st[:, 0], st[:, 1], st[:, 2] = psi.sum(1), (psi * E).sum(1), (psi * Pd).sum(1)

    # ---- unary factors
    def _unary_common(self, nv, var_de, sp_off, sp_en, sp_feat, sp_val, giv_off, giv_label, giv_gap1):
        de = _arr(var_de, np.int32, nv)
        so = _arr(sp_off, np.int32, nv + 1); go = _arr(giv_off, np.int32, nv + 1)
        ns, ng = int(so[-1]), int(go[-1])
        return (de, so, _arr(sp_en, np.int32, max(ns, 1)), _arr(sp_feat, np.int32, max(ns, 1)),
                _arr(sp_val, np.float32, max(ns, 1)), go, _arr(giv_label, np.int32, max(ng, 1)),
                _arr(giv_gap1, np.int32, max(ng, 1)))

    def mlbp_unary_stats(self, nv, var_de, var_label, sp_off, sp_en, sp_feat, sp_val, giv_off, giv_label, giv_gap1, pmi,
                         w1, edT, pedT, V, ldf, th, edstats, colsums, inv_sigma, g_unary):
        de, so, se, sf, sv, go, gl, gg = self._unary_common(nv, var_de, sp_off, sp_en, sp_feat, sp_val, giv_off,
                                                            giv_label, giv_gap1)
        lab = _arr(var_label, np.int32, nv)
        Vd = max(int(de.max()) + 1, 1)
        E = _arr(edT, np.float32, Vd * ldf).reshape(Vd, ldf); Pd = _arr(pedT, np.float32, Vd * ldf).reshape(Vd, ldf)
        P = _arr(pmi, np.float32, V * ldf).reshape(V, ldf); W = _arr(w1, np.float32, V * ldf).reshape(V, ldf)
        t = _arr(th, np.float64, 6)
        st = _arr(edstats, np.float64, Vd * 3).reshape(Vd, 3)
        cs = _arr(colsums, np.float64, 5 * V).reshape(5, V)
        isg = _arr(inv_sigma, np.float64, nv); gu = _arr(g_unary, np.float64, nv * 9).reshape(nv, 9)
        for v in range(nv):
            d, y = int(de[v]), int(lab[v])
            if d < 0:
                g = np.zeros(9)
                for j in range(go[v], go[v + 1]):
                    o = int(gl[j])
                    if gg[j]:
                        g[0] += float(P[y, o]) - cs[3, o] / cs[1, o]
                        g[1] += float(W[y, o]) - cs[4, o] / cs[1, o]
                    else:
                        g[0] += float(P[y, o]) - cs[2, o] / cs[0, o]
                isg[v] = 1.0
                gu[v] = g
                continue
            S0, S1, S2 = st[d]
            ent = [(int(se[s]), int(sf[s]), float(sv[s])) for s in range(so[v], so[v + 1])]
            base = lambda e: np.exp(t[0] * float(E[d, e]) + t[1] * float(Pd[d, e]) + t[5])
            delta = lambda e: sum(t[f] * val for (ee, f, val) in ent if ee == e)
            for e in sorted(set(e for e, _, _ in ent)):
                b = base(e); f = b * np.exp(delta(e))
                S0 += f - b; S1 += (f - b) * float(E[d, e]); S2 += (f - b) * float(Pd[d, e])
            g = np.zeros(9)
            g[3] = float(E[d, y]) - S1 / S0
            g[4] = float(Pd[d, y]) - S2 / S0
            for e, f, val in ent:
                g[3 + f] += val * ((1.0 if e == y else 0.0) - base(e) * np.exp(delta(e)) / S0)
            for j in range(go[v], go[v + 1]):
                o = int(gl[j])
                if gg[j]:
                    g[0] += float(P[y, o]) - cs[3, o] / cs[1, o]
                    g[1] += float(W[y, o]) - cs[4, o] / cs[1, o]
                else:
                    g[0] += float(P[y, o]) - cs[2, o] / cs[0, o]
            isg[v] = 1.0 / S0
            gu[v] = g

    def mlbp_unary_products(self, nv, var_de, sp_off, sp_en, sp_feat, sp_val, giv_off, giv_label, giv_gap1, edT, pedT, V,
                            ldf, th, inv_sigma, planes, ps, ldv, scale_exp, colsums, U):
        de, so, se, sf, sv, go, gl, gg = self._unary_common(nv, var_de, sp_off, sp_en, sp_feat, sp_val, giv_off,
                                                            giv_label, giv_gap1)
        Vd = max(int(de.max()) + 1, 1)
        E = _arr(edT, np.float32, Vd * ldf).reshape(Vd, ldf)[:, :V].astype(np.float64)
        Pd = _arr(pedT, np.float32, Vd * ldf).reshape(Vd, ldf)[:, :V].astype(np.float64)
        t = _arr(th, np.float64, 6)
        isg = _arr(inv_sigma, np.float64, nv)
        pl = _arr(planes, np.float16, 8 * ps)
        cs = _arr(colsums, np.float64, 5 * V).reshape(5, V)
        Uo = _arr(U, np.float32, nv * ldv).reshape(nv, ldv)
        for v in range(nv):
            d = int(de[v])
            if d >= 0:
                z = t[0] * E[d] + t[1] * Pd[d] + t[5]
                for s in range(so[v], so[v + 1]):
                    z[int(se[s])] += t[int(sf[s])] * float(sv[s])
                u = np.exp(z) * V * isg[v]
            else:
                u = np.ones(V)
            for j in range(go[v], go[v + 1]):
                o, tp = int(gl[j]), (6 if gg[j] else 2)
                row = (pl[tp * ps + o * ldv: tp * ps + o * ldv + V].astype(np.float64) +
                       pl[(tp + 1) * ps + o * ldv:(tp + 1) * ps + o * ldv + V].astype(np.float64))
                u = u * row * np.ldexp(1.0, -scale_exp) * V / cs[1 if gg[j] else 0, o]
            Uo[v, :V] = u
            Uo[v, V:] = 0

    # ---- messages
    def mlbp_fill_uniform_rows(self, A_hi, A_lo, ldv, V, rows, n_rows, keep):
        r = _arr(rows, np.int32, n_rows)
        hi, lo = _split(np.array([A_SCALE / V]))
        n = (int(r.max()) + 1) * ldv
        H, L = _arr(A_hi, np.float16, n).reshape(-1, ldv), _arr(A_lo, np.float16, n).reshape(-1, ldv)
        km = _arr(keep, np.uint8, V) if keep is not None and keep.value else np.ones(V, dtype=np.uint8)
        H[r, :V], L[r, :V] = np.where(km, hi[0], 0), np.where(km, lo[0], 0)
        H[r, V:], L[r, V:] = 0, 0

    def mlbp_zero_words(self, p, n):
        _arr(p, np.int32, n)[:] = 0

    def mlbp_var_to_factor(self, n_groups, grp_u, grp_off, in_row, dest_off, dest, first_dest, second_dest, U, D, ldv, V, A_hi, A_lo, max_in, range_log2):
        gu = _arr(grp_u, np.int32, n_groups); go = _arr(grp_off, np.int32, n_groups + 1)
        n_in = int(go[-1])
        ir = _arr(in_row, np.int32, n_in); do = _arr(dest_off, np.int32, n_in + 1)
        de = _arr(dest, np.int32, max(int(do[-1]), 1))
        fd, sd = _arr(first_dest, np.int32, n_in), _arr(second_dest, np.int32, n_in)
        for i in range(n_in):     # the plan's per-slot shortcuts must agree with the CSR lists
            assert fd[i] == (de[do[i]] if do[i + 1] > do[i] else -1), (i, fd[i])
            assert sd[i] == (de[do[i] + 1] if do[i + 1] > do[i] + 1 else -1), (i, sd[i])
        big = 1 << 40
        Uo = _arr(U, np.float32, (int(gu.max()) + 1) * ldv).reshape(-1, ldv)
        Dm = _arr(D, np.float32, (max(int(ir.max()), 0) + 1) * ldv).reshape(-1, ldv)
        amax = (int(de[:int(do[-1])].max()) + 1) if int(do[-1]) else 1
        H, L = _arr(A_hi, np.float16, amax * ldv).reshape(-1, ldv), _arr(A_lo, np.float16, amax * ldv).reshape(-1, ldv)
        assert big
        for g in range(n_groups):
            rows = ir[go[g]:go[g + 1]]
            assert len(rows) <= max_in
            ins = [Dm[r, :V].astype(np.float64) if r >= 0 else np.ones(V) for r in rows]
            u = Uo[gu[g], :V].astype(np.float64)
            for j in range(len(rows)):
                i = go[g] + j
                if do[i + 1] == do[i]:
                    continue
                p = u.copy()
                for q, m in enumerate(ins):
                    if q != j:
                        p = p * m
                s = p.sum()
                x = p * (A_SCALE / s) if (s > 0 and np.isfinite(s)) else np.full(V, A_SCALE / V)
                hi, lo = _split(x)
                for tdest in de[do[i]:do[i + 1]]:
                    H[tdest, :V], L[tdest, :V] = hi, lo

    def mlbp_topk_mask_rows(self, A_hi, A_lo, ldv, V, row0, n_rows, K):
        if K >= V:
            return
        H = _arr(A_hi, np.float16, (row0 + n_rows) * ldv).reshape(-1, ldv)
        L = _arr(A_lo, np.float16, (row0 + n_rows) * ldv).reshape(-1, ldv)
        for r in range(row0, row0 + n_rows):
            v = H[r, :V].astype(np.float32) + L[r, :V].astype(np.float32)
            drop = np.ones(V, dtype=bool)
            drop[np.argpartition(-v, K - 1)[:K]] = False
            H[r, :V][drop] = 0
            L[r, :V][drop] = 0

    def mlbp_topk_rows(self, X, ldx, V, n_rows, K, idx, val, n_ties):
        Xm = _arr(X, np.float32, n_rows * ldx).reshape(-1, ldx)
        I = _arr(idx, np.int32, n_rows * K).reshape(-1, K)
        P = _arr(val, np.float32, n_rows * K).reshape(-1, K)
        T = _arr(n_ties, np.int32, n_rows)
        for r in range(n_rows):
            x = np.maximum(Xm[r, :V], 0)
            order = np.lexsort((np.arange(V), -x))[:K]              # descending value, ascending index among equals
            I[r], P[r] = order, x[order]
            if T is not None:
                T[r] = int((np.diff(P[r]) == 0).sum()) + (1 if int((x == P[r, -1]).sum()) > int((P[r] == P[r, -1]).sum()) else 0)

    def mlbp_factor_to_var_gemm(self, A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha,
                                impl, k0=0, k_len=0, add_const=0.0):
        k1 = V if k_len == 0 else min(V, k0 + k_len)                # K range of this launch; k0 > 0 adds to D
        H = _arr(A_hi, np.float16, (a_row0 + n_rows) * ldv).reshape(-1, ldv)[a_row0:, k0:k1].astype(np.float64)
        L = _arr(A_lo, np.float16, (a_row0 + n_rows) * ldv).reshape(-1, ldv)[a_row0:, k0:k1].astype(np.float64)
        Bh = _arr(B_hi, np.float16, V * ldv).reshape(V, ldv)[:, k0:k1].astype(np.float64)
        Bl = _arr(B_lo, np.float16, V * ldv).reshape(V, ldv)[:, k0:k1].astype(np.float64)
        out = H @ Bh.T + (0.0 if impl & 512 else H @ Bl.T) + (0.0 if impl & 256 else L @ Bh.T)   # MLBP_GEMM_{A,B}_HI_ONLY
        Dm = _arr(D, np.float32, (d_row0 + n_rows) * ldd).reshape(-1, ldd)
        if k0 > 0:
            Dm[d_row0:d_row0 + n_rows, :V] += (alpha * out).astype(np.float32)
        else:
            Dm[d_row0:d_row0 + n_rows, :V] = (alpha * out + add_const).astype(np.float32)

    def mlbp_factor_to_var_gemm_gated(self, A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha,
                                      impl, gate, run_if_set, k0=0, k_len=0, add_const=0.0):
        if gate is not None and gate.value and (int(_arr(gate, np.int32, 1)[0]) != 0) != (run_if_set != 0):
            return
        self.mlbp_factor_to_var_gemm(A_hi, A_lo, a_rows_total, a_row0, n_rows, B_hi, B_lo, V, ldv, D, d_row0, ldd, alpha, impl,
                                     k0, k_len, add_const)

    def mlbp_spike_scan(self, A_hi, A_lo, ldv, V, a0, n_rows, spike_prob, words, cnt, entries, block_rows, block_n):
        w = _arr(words, np.int32, 5)
        H = _arr(A_hi, np.float16, (a0 + n_rows) * ldv).reshape(-1, ldv)
        L = _arr(A_lo, np.float16, (a0 + n_rows) * ldv).reshape(-1, ldv)
        c = _arr(cnt, np.int32, a0 + n_rows)
        ent = _arr(entries, np.int32, (a0 + n_rows) * 8).reshape(-1, 4, 2)
        listed = block_rows is not None and block_rows.value
        for r in range(a0, a0 + n_rows):
            cols = np.nonzero(H[r, :V].astype(np.float32) > np.float32(spike_prob * A_SCALE))[0]
            c[r] = min(len(cols), 5)
            if len(cols) == 0:
                continue
            w[3] = 1; w[4] += 1
            if len(cols) > 4:
                w[0] = 1
                continue
            for i, col in enumerate(cols):
                ent[r, i, 0] = col
                ent[r, i, 1:2].view(np.float32)[0] = np.float32(L[r, col])
            w[2:3].view(np.float32)[0] = max(float(w[2:3].view(np.float32)[0]), float(H[r, cols].astype(np.float32).max()))
            if listed:
                n = _arr(block_n, np.int32, 1)
                _arr(block_rows, np.int32, n_rows)[n[0]] = r
                n[0] += 1

    def mlbp_spike_correct(self, words, cnt, entries, rows, n_list, a0, n_rows, Bt_hi, Bt_lo, V, ldv, D, d_row0, ldd, alpha,
                           A_hi_one_pass=None, Rt_hi=None, tbar=0.0):
        w = _arr(words, np.int32, 5)
        if w[0] != 0:
            return
        n = int(_arr(n_list, np.int32, 1)[0])
        rws = _arr(rows, np.int32, max(n, 1))[:n]
        sel = [int(r) for r in rws if a0 <= r < a0 + n_rows]
        if not sel:
            return
        top = max(sel) + 1
        c = _arr(cnt, np.int32, top)
        ent = _arr(entries, np.int32, top * 8).reshape(-1, 4, 2)
        Bh = _arr(Bt_hi, np.float16, V * ldv).reshape(V, ldv)
        Bl = _arr(Bt_lo, np.float16, V * ldv).reshape(V, ldv)
        Dm = _arr(D, np.float32, (d_row0 + n_rows) * ldd).reshape(-1, ldd)
        for r in sel:
            items = sorted((int(ent[r, s, 0]), float(ent[r, s, 1:2].view(np.float32)[0])) for s in range(min(int(c[r]), 4)))
            acc = np.zeros(V, dtype=np.float32)
            for col, lo in items:
                full = Bh[col, :V].astype(np.float32) + Bl[col, :V].astype(np.float32)
                hi = None
                if A_hi_one_pass is not None and getattr(A_hi_one_pass, 'value', A_hi_one_pass):
                    hi = np.float32(_arr(A_hi_one_pass, np.float16, (r + 1) * ldv).reshape(-1, ldv)[r, col])
                if hi is not None and Rt_hi is not None and getattr(Rt_hi, 'value', Rt_hi):   # residual-plane product
                    Rh = _arr(Rt_hi, np.float16, V * ldv).reshape(V, ldv)[col, :V].astype(np.float32)
                    acc = (acc + (hi + np.float32(lo)) * (full - np.float32(tbar)) - hi * Rh).astype(np.float32)
                    continue
                acc = (acc + np.float32(lo) * full).astype(np.float32)
                if hi is not None:
                    acc = (acc + hi * Bl[col, :V].astype(np.float32)).astype(np.float32)
            Dm[d_row0 + r - a0, :V] += np.float32(alpha) * acc

    def mlbp_marginals(self, n_groups, grp_u, grp_off, in_row, label, U, D, ldv, V, logp, top1, rank, beliefs, range_log2,
                       max_in=0, tau=0.0, tau_label=0.0, aux=None, cnts=None, flags=None, flagged=None, n_flagged=None):
        gu = _arr(grp_u, np.int32, n_groups); go = _arr(grp_off, np.int32, n_groups + 1)
        ir = _arr(in_row, np.int32, max(int(go[-1]), 1)); lab = _arr(label, np.int32, n_groups)
        Uo = _arr(U, np.float32, (int(gu.max()) + 1) * ldv).reshape(-1, ldv)
        Dm = _arr(D, np.float32, (max(int(ir.max()), 0) + 1) * ldv).reshape(-1, ldv)
        lp, t1, rk = _arr(logp, np.float64, n_groups), _arr(top1, np.int32, n_groups), _arr(rank, np.int32, n_groups)
        B = _arr(beliefs, np.float32, n_groups * ldv).reshape(-1, ldv) if beliefs is not None and beliefs.value else None
        for g in range(n_groups):
            p = Uo[gu[g], :V].astype(np.float64)
            for r in ir[go[g]:go[g + 1]]:
                if r >= 0:
                    p = p * Dm[r, :V].astype(np.float64)
            s = p.sum()
            b = p / s if s > 0 else np.full(V, 1.0 / V)
            lp[g] = np.log(b[lab[g]]) if b[lab[g]] > 0 else -99.99
            t1[g] = int(np.argmax(b)); rk[g] = int((b > b[lab[g]]).sum())
            if B is not None:
                B[g, :V] = b
            if flags is not None and flags.value:
                srt = np.sort(p)
                pl = p[lab[g]]
                near_hi = int(((p > pl) & (p <= pl * (1 + tau_label))).sum())
                near_lo = int(((p <= pl) & (p >= pl * (1 - tau_label))).sum()) - 1       # the label itself
                f = (1 if (V > 1 and srt[-2] >= srt[-1] * (1 - tau)) else 0) | (2 if (near_hi + near_lo > 0 and rk[g] - near_hi <= 50) else 0)
                _arr(flags, np.int32, n_groups)[g] = f
                _arr(aux, np.float64, 2 * n_groups)[2 * g:2 * g + 2] = (srt[-1], pl)
                _arr(cnts, np.int32, 2 * n_groups)[2 * g:2 * g + 2] = (rk[g], near_hi)
                if f:
                    nf = _arr(n_flagged, np.int32, 1)
                    _arr(flagged, np.int32, n_groups)[nf[0]] = g
                    nf[0] += 1

    def mlbp_rescore_candidates(self, n_vars, flagged, n_flagged, flags, grp_u, grp_off, in_row, label, U, D, ldv, V, A_hi, A_lo,
                                planes, ps, msg_blocks, n_blocks, n_msg_rows, aux, cnts, tau, tau_label, range_log2, top1, rank,
                                counters):
        nf = int(_arr(n_flagged, np.int32, 1)[0])
        fl, fg = _arr(flagged, np.int32, n_vars), _arr(flags, np.int32, n_vars)
        gu = _arr(grp_u, np.int32, n_vars); go = _arr(grp_off, np.int32, n_vars + 1)
        ir = _arr(in_row, np.int32, max(int(go[-1]), 1)); lab = _arr(label, np.int32, n_vars)
        Uo = _arr(U, np.float32, (int(gu.max()) + 1) * ldv).reshape(-1, ldv)
        Dm = _arr(D, np.float32, (max(int(ir.max()), 0) + 1) * ldv).reshape(-1, ldv)
        blk = _arr(msg_blocks, np.int32, 4 * max(n_blocks, 1)).reshape(-1, 4)[:n_blocks]
        ax, cn = _arr(aux, np.float64, 2 * n_vars).reshape(-1, 2), _arr(cnts, np.int32, 2 * n_vars).reshape(-1, 2)
        t1, rk, ct = _arr(top1, np.int32, n_vars), _arr(rank, np.int32, n_vars), _arr(counters, np.int32, 8)
        H = _arr(A_hi, np.float16, (n_msg_rows + 1) * ldv).reshape(-1, ldv)
        L = _arr(A_lo, np.float16, (n_msg_rows + 1) * ldv).reshape(-1, ldv)
        pl = _arr(planes, np.float16, 8 * ps)
        for g in fl[:nf]:
            rows = ir[go[g]:go[g + 1]]
            p = Uo[gu[g], :V].astype(np.float64)
            for r in rows:
                if r >= 0:
                    p = p * Dm[r, :V].astype(np.float64)
            best, plab = ax[g]
            bits = np.zeros(V, dtype=np.int32)
            if fg[g] & 1:
                bits[p >= best * (1 - tau)] |= 1
            if fg[g] & 2:
                m = (p >= plab * (1 - tau_label)) & (p <= plab * (1 + tau_label)); m[lab[g]] = False
                bits[m] |= 2
            cand = [int(lab[g])] + [int(e) for e in np.nonzero(bits)[0] if e != lab[g]]
            if len(cand) > 64:
                ct[1] += 1
                continue
            score = Uo[gu[g], cand].astype(np.float64) / float(Uo[gu[g], cand[0]])
            for r in rows:
                if r < 0:
                    continue
                a_row = r - 5
                t = -1
                for b in blk:
                    if b[1] <= a_row < b[1] + b[3] and r >= 5 and a_row < n_msg_rows:
                        t = int(b[0])
                if t < 0:
                    d = Dm[r, cand].astype(np.float64)
                else:
                    a = H[a_row, :V].astype(np.float64) + L[a_row, :V].astype(np.float64)
                    d = np.array([a @ (pl[2 * t * ps + e * ldv: 2 * t * ps + e * ldv + V].astype(np.float64) +
                                       pl[(2 * t + 1) * ps + e * ldv:(2 * t + 1) * ps + e * ldv + V].astype(np.float64)) for e in cand])
                score = score * d / d[0]
            cb = np.array([bits[e] for e in cand]); cb[0] = bits[lab[g]]
            if fg[g] & 1:
                sel = [i for i in range(len(cand)) if cb[i] & 1]
                bi = min(sel, key=lambda i: (-score[i], cand[i]))
                ct[3] += int(t1[g] != cand[bi]); t1[g] = cand[bi]
            if fg[g] & 2:
                new = int(cn[g, 0] - cn[g, 1] + sum(1 for i in range(1, len(cand)) if (cb[i] & 2) and score[i] > score[0]))
                ct[4] += int(rk[g] != new); rk[g] = new
            ct[0] += 1

    def mlbp_pair_expectations(self, n_factors, c_row, z_row, u0_row, u1_row, u2_row, A_hi, A_lo, D, ldv, V, stats,
                               r_row=None, pair_gap1=None, spike_words=None, spike_cnt=None, spike_entries=None, planes=None,
                               ps=0, alpha=1.0):
        cr, zr, r0, r1, r2 = (_arr(x, np.int32, n_factors) for x in (c_row, z_row, u0_row, u1_row, u2_row))
        H = _arr(A_hi, np.float16, (int(cr.max()) + 1) * ldv).reshape(-1, ldv)
        L = _arr(A_lo, np.float16, (int(cr.max()) + 1) * ldv).reshape(-1, ldv)
        Dm = _arr(D, np.float32, (max(int(r0.max()), int(r1.max()), int(r2.max())) + 1) * ldv).reshape(-1, ldv)
        st = _arr(stats, np.float64, n_factors * 3).reshape(-1, 3)
        for f in range(n_factors):
            c = H[cr[f], :V].astype(np.float64) + L[cr[f], :V].astype(np.float64)
            z = H[zr[f], :V].astype(np.float64) + L[zr[f], :V].astype(np.float64)
            st[f, 0] = z @ Dm[r0[f], :V].astype(np.float64)
            st[f, 1] = c @ Dm[r1[f], :V].astype(np.float64)
            st[f, 2] = c @ Dm[r2[f], :V].astype(np.float64) if r2[f] >= 0 else 0.0
        if spike_words is not None and spike_words.value and _arr(spike_words, np.int32, 1)[0] == 0:
            rr, g1 = _arr(r_row, np.int32, n_factors), _arr(pair_gap1, np.int32, n_factors)
            top = max(int(cr.max()), int(rr.max())) + 1
            cnt = _arr(spike_cnt, np.int32, top)
            ent = _arr(spike_entries, np.int32, top * 8).reshape(-1, 4, 2)
            Hr = _arr(A_hi, np.float16, top * ldv).reshape(-1, ldv)
            Lr = _arr(A_lo, np.float16, top * ldv).reshape(-1, ldv)
            pl = _arr(planes, np.float16, 14 * ps)
            lo_of = lambda table, a, b: float(pl[(2 * table + 1) * ps + a * ldv + b])
            for f in range(n_factors):
                for i in range(min(int(cnt[cr[f]]), 4)):
                    a = int(ent[cr[f], i, 0])
                    ca = float(Hr[cr[f], a]) + float(Lr[cr[f], a])
                    for j in range(min(int(cnt[rr[f]]), 4)):
                        b = int(ent[rr[f], j, 0])
                        w = float(np.float32(alpha)) * ca * float(Hr[rr[f], b])
                        if zr[f] == cr[f]:
                            st[f, 0] += w * lo_of(2 if g1[f] else 0, a, b)
                        st[f, 1] += w * lo_of(5 if g1[f] else 4, a, b)
                        if r2[f] >= 0:
                            st[f, 2] += w * lo_of(6, a, b)

    def mlbp_const_rows(self, colsums, V, ldv, rows):
        cs = _arr(colsums, np.float64, 7 * V).reshape(7, V)
        R = _arr(rows, np.float32, 5 * ldv).reshape(5, ldv)
        R[0] = 1.0
        for t, k in enumerate((5, 0, 6, 1)):
            R[1 + t, :V] = (cs[k] / cs[k].mean()).astype(np.float32)
            R[1 + t, V:] = 0

    def mlbp_batch_reduce(self, n_sent, grad, logp_sent, n_vars, rank, peak_flag, out16):
        o = _arr(out16, np.float64, 16)
        if grad is not None and grad.value and n_sent:
            o[:9] += _arr(grad, np.float64, n_sent * 9).reshape(-1, 9).sum(0)
        if logp_sent is not None and logp_sent.value and n_sent:
            o[9] += _arr(logp_sent, np.float64, n_sent).sum()
        if rank is not None and rank.value and n_vars:
            r = _arr(rank, np.int32, n_vars)
            o[10] += (r == 0).sum(); o[11] += (r < 26).sum(); o[12] += (r < 50).sum(); o[13] += n_vars
        o[14] += n_sent
        if peak_flag is not None and peak_flag.value:
            w = _arr(peak_flag, np.int32, 4)
            o[15] = max(o[15], (1.0 if w[0] else 0.0) + (2.0 if w[3] else 0.0))

    def mlbp_gradient_reduce(self, n_sent, sent_var_off, sent_fac_off, g_unary, pair_stats, v0, v1, var_label, gap1, pmi, w1, ldf,
                             logp_var, grad, logp_sent):
        vo = _arr(sent_var_off, np.int32, n_sent + 1)
        has_f = sent_fac_off is not None and sent_fac_off.value
        fo = _arr(sent_fac_off, np.int32, n_sent + 1) if has_f else np.zeros(n_sent + 1, dtype=np.int32)
        nv, nf = int(vo[-1]), int(fo[-1])
        gu = _arr(g_unary, np.float64, nv * 9).reshape(-1, 9)
        st = _arr(pair_stats, np.float64, max(nf, 1) * 3).reshape(-1, 3)
        if has_f:
            lab = _arr(var_label, np.int32, nv)
            a0, a1 = lab[_arr(v0, np.int32, max(nf, 1))], lab[_arr(v1, np.int32, max(nf, 1))]
            g1 = _arr(gap1, np.int32, max(nf, 1))
        V = ldf   # only cells are read
        lv = _arr(logp_var, np.float64, nv) if logp_var is not None and logp_var.value else None
        G = _arr(grad, np.float64, n_sent * 9).reshape(-1, 9); LS = _arr(logp_sent, np.float64, n_sent)
        for s in range(n_sent):
            g = gu[vo[s]:vo[s + 1]].sum(0)
            for f in range(fo[s], fo[s + 1]):
                cell = int(a0[f]) * ldf + int(a1[f])
                z = st[f, 0]
                g[0] += float(_arr(pmi, np.float32, cell + 1)[cell]) - (st[f, 1] / z if z > 0 else 0.0)
                if g1[f]:
                    g[1] += float(_arr(w1, np.float32, cell + 1)[cell]) - (st[f, 2] / z if z > 0 else 0.0)
                g[2] += 0.0 if z > 0 else 1.0
            G[s] = g
            LS[s] = lv[vo[s]:vo[s + 1]].sum() if lv is not None else 0.0
