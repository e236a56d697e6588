"""CPU ORACLE for the LBP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under macaronicusermodeling_b200/ does (the product path has no CPU fallback).

What it is: a plain NumPy float64 restatement of the reference's algorithm for this path, each function
citing the reference file:line it follows (paths relative to /root/reference).  Two evaluators share one
graph / schedule builder:

* ``run_literal``  -- op-for-op what LBP.py does (per-message NumPy calls, sequential Gauss-Seidel
  updates, dense V x V pairwise beliefs in the gradient, potentials rebuilt per sentence like
  train.py:218-253).  This is the one PINNED against the reference: tests/test_oracle_golden.py compares
  every message, marginal, gradient, log-posterior and precision count with the fixtures that
  tests/golden/make_golden.py produced by running the (py3-patched) reference itself in the build
  container.  Parity status: PINNED (outputs of the reference run here; the reference has no golden
  vectors of its own).
* ``run_fast``     -- same numbers (asserted against run_literal in the tests to 1e-10) but with the
  potentials hoisted out of the per-sentence path and the gradient in closed form
  (E = c'(T o phi_k) r / c' T r), GEMVs of one schedule level batched into one BLAS
  dgemm.  It is the strongest fair CPU variant and is what bench.py times as the
  CPU baseline, with all host threads.

Graph conventions (train.py:133-305): variable id = sentence position of a PREDICTED token; factors are
numbered in creation order: first one en_de unary factor per predicted variable (:255-264), then, for
every position pair i<j (:270-297), a pairwise en_en factor (both predicted) or a unary en_en factor on
the predicted one (observed_dim = label of the given one), gap = |i-j|.
"""
import numpy as np

UNARY, PAIR = 1, 2
T_EN_DE, T_EN_EN = 0, 1


# ----------------------------------------------------------------------------------------- graph
class Factor(object):
    __slots__ = ('id', 'ftype', 'vars', 'gap', 'obs')

    def __init__(self, fid, ftype, vars_, gap, obs):
        self.id, self.ftype, self.vars, self.gap, self.obs = fid, ftype, tuple(vars_), gap, obs

    @property
    def arity(self):
        return len(self.vars)


class Graph(object):
    """Factor list (creation order = id order), facset order per variable (LBP.py:367-369)."""

    def __init__(self, sent):
        self.sent = sent
        self.factors = []
        self.facset = {}
        kind, label, de = sent.kind, sent.label, sent.de
        n = len(kind)
        pred = [p for p in range(n) if kind[p] == 1]
        self.var_ids = pred
        for p in pred:
            self.facset[p] = []
        for p in pred:                                                           # train.py:255-264
            self._add(Factor(len(self.factors), T_EN_DE, [p], 0, int(de[p])))
        for i in range(n):                                                       # train.py:270-297
            for j in range(i + 1, n):
                if kind[i] == 1 and kind[j] == 1:
                    self._add(Factor(len(self.factors), T_EN_EN, [i, j], j - i, None))
                elif kind[i] == 0 and kind[j] == 0:
                    pass
                else:
                    vp, vg = (i, j) if kind[i] == 1 else (j, i)
                    self._add(Factor(len(self.factors), T_EN_EN, [vp], abs(i - j), int(label[vg])))
        self.label = {p: int(label[p]) for p in pred}

    def _add(self, f):
        self.factors.append(f)
        for v in f.vars:
            self.facset[v].append(f.id)

    def has_loops(self, root):
        """LBP.py:174-190 (DFS with an explicit stack, parent-edge excluded)."""
        seen = set()
        stack = [(('X', root), None)]
        while stack:
            n, par = stack.pop()
            if n in seen:
                return True
            seen.add(n)
            if n[0] == 'X':
                stack.extend((('F', f), n) for f in self.facset[n[1]] if ('F', f) != par)
            else:
                stack.extend((('X', v), n) for v in self.factors[n[1]].vars if ('X', v) != par)
        return False

    def schedule(self, root):
        """LBP.py:155-172: FIFO BFS, nodes marked seen when POPPED -> duplicate edges in loopy graphs."""
        sched, seen, queue = [], set(), [('X', root)]
        while queue:
            n = queue.pop(0)
            if n in seen:
                continue
            seen.add(n)
            if n[0] == 'X':
                nb = [('F', f) for f in self.facset[n[1]] if ('F', f) not in seen]
            else:
                nb = [('X', v) for v in self.factors[n[1]].vars if ('X', v) not in seen]
            sched.extend((m, n) for m in nb)
            queue.extend(nb)
        return sched

    def update_sequence(self, roots, sweeps):
        """LBP.py:218-245: the sequence of (src, dst) message updates of treelike_inference.
        roots[0] is the has_loops draw (LBP.py:176), roots[1:] one per sweep (LBP.py:223)."""
        loopy = self.has_loops(roots[0])
        n_it = sweeps if loopy else 1
        seq = []
        for it in range(n_it):
            S = self.schedule(roots[1 + it])
            for frm, to in reversed(S):                                           # leaves -> root
                if to[0] == 'F' and self.factors[to[1]].arity < 2:
                    continue
                seq.append((frm, to))
            for to, frm in S:                                                     # root -> leaves
                if to[0] == 'F' and self.factors[to[1]].arity < 2:
                    continue
                seq.append((frm, to))
        return loopy, seq


# ----------------------------------------------------------------------------------------- potentials
def dense_phi_en_de(model, sent):
    """train.py:176-215 + :605-609: the (V, Vd, 6) en_de feature tensor of ONE sentence."""
    V, Vd = model['ed'].shape
    phi = np.zeros((V, Vd, 6))
    phi[:, :, 0] = model['ed']
    phi[:, :, 1] = model['ped']
    phi[:, :, 5] = 1.0
    for e, d, f, val in sent.sparse:
        phi[int(e), int(d), int(f)] += val
    return phi


def dense_phi_en_en(model):
    """train.py:592-595."""
    ones = np.ones_like(model['pmi'])
    phi_w1 = np.stack([model['pmi'], model['pmi_w1'], ones], axis=2)
    phi = np.stack([model['pmi'], np.zeros_like(model['pmi']), ones], axis=2)
    return phi, phi_w1


TOPK = 100   # hard-coded K of the reference's sparse approximations (c_array_utils.pyx:118, :194)


def _topk_mask(m, k=TOPK):
    """keep the k largest entries of a vector, zero the rest: what au.sparse_vec_mat_dot (pyx:193-205) and au.sparse_dot
    (pyx:117-129) do to a message before contracting it (same NumPy argpartition call, so ties resolve alike)"""
    v = np.asarray(m).reshape(-1)
    idx = np.argpartition(-v, k - 1)[:k]
    out = np.zeros_like(v)
    out[idx] = v[idx]
    return out.reshape(np.shape(m))


def _normalize(m):
    """c_array_utils.pyx:29-40 via Message.renormalize LBP.py:649-657."""
    s = np.sum(m)
    if s > 0:
        return m / s
    return np.full_like(m, 1.0 / m.size)


# ----------------------------------------------------------------------------------------- literal evaluator
def run_literal(model, sent, theta_ee, theta_ed, roots, sweeps=3, reg=0.0, lr=1.0, keep_messages=True,
                approx_inference=False, approx_beliefs=False):
    """One sentence through create_factor_graph / initialize / treelike_inference / get_gradient exactly as
    the reference computes it.  Returns a dict (messages keyed like the reference's graph.messages)."""
    theta_ee = np.asarray(theta_ee, dtype=np.float64).reshape(1, 3)
    theta_ed = np.asarray(theta_ed, dtype=np.float64).reshape(1, 6)
    g = Graph(sent)
    V = model['pmi'].shape[0]
    phi_ee, phi_ee_w1 = dense_phi_en_en(model)
    phi_ed = dense_phi_en_de(model, sent)
    pot_ee = np.exp(phi_ee.dot(theta_ee.T)[:, :, 0])                              # train.py:218,237,252
    pot_ee_w1 = np.exp(phi_ee_w1.dot(theta_ee.T)[:, :, 0])                        # train.py:219,238,253
    pot_ed = np.exp(phi_ed.dot(theta_ed.T)[:, :, 0])                              # train.py:240,249,251

    def table(f):                                                                 # LBP.py:456-467, 695-706
        if f.ftype == T_EN_DE:
            t = pot_ed
        elif f.gap > 1:
            t = pot_ee
        elif f.gap == 1:
            t = pot_ee_w1
        else:
            raise ValueError('only 2 kinds of distances are supported')
        return t[:, f.obs].reshape(V, 1) if f.obs is not None else t

    def phi_of(f):                                                                # LBP.py:469-480
        if f.ftype == T_EN_DE:
            return phi_ed
        return phi_ee if f.gap > 1 else phi_ee_w1

    tables = [table(f) for f in g.factors]
    msgs = {}
    uni = np.full((V, 1), 1.0 / V)
    for f in g.factors:                                                           # LBP.py:201-216
        if f.arity == 1:
            msgs[('F', f.id), ('X', f.vars[0])] = uni.copy()
        else:
            for v in f.vars:
                msgs[('X', v), ('F', f.id)] = uni.copy()
                msgs[('F', f.id), ('X', v)] = uni.copy()

    def var_to_factor(v, fid):                                                    # LBP.py:377-389
        m = uni.copy()
        for of in g.facset[v]:
            if of != fid:
                m = np.nan_to_num(np.multiply(msgs[('F', of), ('X', v)], m))      # LBP.py:717-730
        msgs[('X', v), ('F', fid)] = _normalize(m)

    def factor_to_var(fid, v):                                                    # LBP.py:490-526
        f = g.factors[fid]
        if f.arity == 1:
            msgs[('F', fid), ('X', v)] = _normalize(np.copy(tables[fid]))
            return
        o = f.vars[1] if f.vars[0] == v else f.vars[0]
        o_dim = f.vars.index(o)                                                   # var_id2dim: vars[0]->0, vars[1]->1
        m = msgs[('X', o), ('F', fid)]
        if approx_inference:                                                      # LBP.py:506-507, :515-516
            m = _topk_mask(m)
        if o_dim == 1:
            r = tables[fid].dot(m)                                                # LBP.py:509
        else:
            r = m.T.dot(tables[fid])                                              # LBP.py:518
        msgs[('F', fid), ('X', v)] = _normalize(r.reshape(V, 1))

    loopy, seq = g.update_sequence(roots, sweeps)
    for frm, to in seq:
        if frm[0] == 'X':
            var_to_factor(frm[1], to[1])
        else:
            factor_to_var(frm[1], to[1])

    def marginal(v):                                                              # LBP.py:392-400
        m = uni.copy()
        for fid in g.facset[v]:
            m = np.nan_to_num(np.multiply(msgs[('F', fid), ('X', v)], m))
        return _normalize(m)

    marg = np.stack([marginal(v)[:, 0] for v in g.var_ids])
    logp = 0.0
    for i, v in enumerate(g.var_ids):                                             # LBP.py:247-259
        l = np.log(marg[i, g.label[v]]) if marg[i, g.label[v]] > 0 else -np.inf
        logp += -99.99 if l == -np.inf else l

    g_ee = np.zeros((1, 3))
    g_ed = np.zeros((1, 6))
    for f in g.factors:                                                           # LBP.py:301-320, 592-619
        tb = tables[f.id]
        if f.arity == 1:
            s = np.sum(tb)
            beliefs = tb / s if s > 0 else np.zeros_like(tb)                      # LBP.py:540 (messages ignored)
            obs = np.zeros_like(tb)
            obs[g.label[f.vars[0]], 0] = 1.0                                      # LBP.py:584-589
            cell = obs - beliefs
            grad = np.dot(cell.T, phi_of(f)[:, f.obs, :])                         # LBP.py:600-603
        else:
            c = msgs[('X', f.vars[0]), ('F', f.id)].reshape(V, 1)                 # LBP.py:544-553
            r = msgs[('X', f.vars[1]), ('F', f.id)].reshape(1, V)
            if approx_beliefs:                                                    # LBP.py:554-563: top-K x top-K block only
                c, r = _topk_mask(c), _topk_mask(r)
            b = np.multiply(c.dot(r), tb)                                         # LBP.py:566-568
            s = np.sum(b)
            beliefs = b / s if s > 0 else np.zeros_like(b)
            obs = np.zeros_like(tb)
            obs[g.label[f.vars[0]], g.label[f.vars[1]]] = 1.0
            grad = np.tensordot(obs - beliefs, phi_of(f))                         # LBP.py:610
        grad = grad.reshape(1, -1)
        if f.ftype == T_EN_EN:
            g_ee += grad
        else:
            g_ed += grad
    out = _finish(g, loopy, marg, logp, g_ee, g_ed, theta_ee, theta_ed, reg, lr)
    if keep_messages:
        out['messages'] = {('%s_%d' % a, '%s_%d' % b): m[:, 0] for (a, b), m in msgs.items()}
    return out


def _finish(g, loopy, marg, logp, g_ee, g_ed, theta_ee, theta_ed, reg, lr):
    labels = np.array([g.label[v] for v in g.var_ids])
    p_lab = marg[np.arange(len(labels)), labels]
    # LBP.py:80-106 + :402-411, with the same NumPy calls so that exact ties (theta = 0 -> uniform beliefs)
    # resolve like the reference's argpartition/argsort do; `rank` = position of the label in the top-50 list
    # (V if absent).  Without ties this equals the number of strictly larger beliefs.
    top = min(50, marg.shape[1] - 1)
    rank = np.full(len(labels), marg.shape[1], dtype=np.int64)
    for i in range(len(labels)):
        a = marg[i]
        idx = np.argpartition(a, -top)[-top:]
        idx = idx[np.argsort(a[idx])][::-1]
        hit = np.nonzero(idx == labels[i])[0]
        if len(hit):
            rank[i] = hit[0]
    p0 = int((rank == 0).sum())
    p25 = int((rank < 26).sum())
    p50 = int((rank < 51).sum())
    return {
        'is_loopy': loopy, 'var_ids': np.array(g.var_ids), 'marginals': marg, 'top1': marg.argmax(axis=1),
        'logp': logp, 'g_ee_unreg': g_ee.copy(), 'g_ed_unreg': g_ed.copy(),
        'g_ee_ret': lr * (g_ee - reg * theta_ee),                                 # LBP.py:293-299, 322-327
        'g_ed_ret': lr * (g_ed - reg * theta_ed),
        'precision_counts': np.array([p0, p25, p50, len(labels)]),
        'label_rank': rank,
    }


# ----------------------------------------------------------------------------------------- fast evaluator
class Tables(object):
    """Everything that depends on theta only (hoisted out of train.py:218-253)."""

    def __init__(self, model, theta_ee, theta_ed):
        te = np.asarray(theta_ee, dtype=np.float64).reshape(3)
        td = np.asarray(theta_ed, dtype=np.float64).reshape(6)
        self.te, self.td = te, td
        pmi, w1 = model['pmi'], model['pmi_w1']
        self.T = np.exp(te[0] * pmi + te[2])
        self.T1 = np.exp(te[0] * pmi + te[1] * w1 + te[2])
        self.Tt = np.ascontiguousarray(self.T.T)
        self.T1t = np.ascontiguousarray(self.T1.T)
        self.G = self.T * pmi
        self.G1 = self.T1 * pmi
        self.G1w = self.T1 * w1
        self.edT = np.ascontiguousarray(model['ed'].T)                            # de-major: row d contiguous
        self.pedT = np.ascontiguousarray(model['ped'].T)
        self.model = model


def run_fast(tables, sent, roots, sweeps=3, reg=0.0, lr=1.0, want_grad=True, approx_inference=False,
             approx_beliefs=False):
    """Same results as run_literal; GEMVs of one dependency level batched into one dgemm."""
    model = tables.model
    te, td = tables.te, tables.td
    g = Graph(sent)
    V = model['pmi'].shape[0]
    # unary messages (constant over the sweeps: LBP.py:492-498 recomputes the same value each UP pass)
    unary = {}
    sp = sent.sparse
    for f in g.factors:
        if f.arity != 1:
            continue
        if f.ftype == T_EN_DE:
            z = td[0] * tables.edT[f.obs] + td[1] * tables.pedT[f.obs] + td[5]
            for e, d, k, val in sp:
                if int(d) == f.obs:
                    z[int(e)] += td[int(k)] * val
            t = np.exp(z)
        else:
            t = (tables.T1t if f.gap == 1 else tables.Tt)[f.obs]
        unary[f.id] = t
    uprod = {}
    for v in g.var_ids:
        m = np.full(V, 1.0 / V)
        for fid in g.facset[v]:
            if g.factors[fid].arity == 1:
                m = m * (unary[fid] / unary[fid].sum())
        uprod[v] = m
    uni = np.full(V, 1.0 / V)
    v2f, f2v = {}, {}
    for f in g.factors:
        if f.arity == 2:
            for v in f.vars:
                v2f[v, f.id] = uni
                f2v[f.id, v] = uni
    loopy, seq = g.update_sequence(roots, sweeps)
    # group consecutive pairwise factor->var updates (they never depend on each other inside a run)
    i = 0
    while i < len(seq):
        frm, to = seq[i]
        if frm[0] == 'X':
            v, fid = frm[1], to[1]
            m = uprod[v]
            for of in g.facset[v]:
                if of != fid and g.factors[of].arity == 2:
                    m = m * f2v[of, v]
            s = m.sum()
            v2f[v, fid] = m / s if s > 0 else uni
            i += 1
            continue
        run = []
        while i < len(seq) and seq[i][0][0] == 'F':
            fid, v = seq[i][0][1], seq[i][1][1]
            if g.factors[fid].arity == 2:
                run.append((fid, v))
            i += 1
        groups = {}
        for fid, v in run:
            f = g.factors[fid]
            to_dim0 = (f.vars[0] == v)
            o = f.vars[1] if to_dim0 else f.vars[0]
            key = (f.gap == 1, to_dim0)
            groups.setdefault(key, []).append((fid, v, _topk_mask(v2f[o, fid]) if approx_inference else v2f[o, fid]))
        for (w1, to_dim0), items in groups.items():
            M = np.stack([it[2] for it in items])
            if to_dim0:
                B = tables.T1t if w1 else tables.Tt                               # out[a] = sum_b T[a,b] m[b]
            else:
                B = tables.T1 if w1 else tables.T                                 # out[b] = sum_a m[a] T[a,b]
            D = M.dot(B)
            for (fid, v, _), row in zip(items, D):
                s = row.sum()
                f2v[fid, v] = row / s if s > 0 else uni
    marg = []
    for v in g.var_ids:
        m = uprod[v]
        for fid in g.facset[v]:
            if g.factors[fid].arity == 2:
                m = m * f2v[fid, v]
        s = m.sum()
        marg.append(m / s if s > 0 else uni)
    marg = np.stack(marg)
    logp = 0.0
    for i_, v in enumerate(g.var_ids):
        p = marg[i_, g.label[v]]
        logp += np.log(p) if p > 0 else -99.99
    g_ee = np.zeros((1, 3))
    g_ed = np.zeros((1, 6))
    if want_grad:
        pmi, w1p = model['pmi'], model['pmi_w1']
        pair = [f for f in g.factors if f.arity == 2]
        for gap1 in (False, True):
            fs = [f for f in pair if (f.gap == 1) == gap1]
            if not fs:
                continue
            R = np.stack([v2f[f.vars[1], f.id] for f in fs])
            C = np.stack([v2f[f.vars[0], f.id] for f in fs])
            if approx_beliefs:
                R = np.stack([_topk_mask(x) for x in R])
                C = np.stack([_topk_mask(x) for x in C])
            U0 = R.dot(tables.T1t if gap1 else tables.Tt)
            U1 = R.dot((tables.G1 if gap1 else tables.G).T)
            Z = np.einsum('ij,ij->i', C, U0)
            E1 = np.einsum('ij,ij->i', C, U1) / Z
            for n_, f in enumerate(fs):
                l0, l1 = g.label[f.vars[0]], g.label[f.vars[1]]
                g_ee[0, 0] += pmi[l0, l1] - E1[n_]
            if gap1:
                U2 = R.dot(tables.G1w.T)
                E2 = np.einsum('ij,ij->i', C, U2) / Z
                for n_, f in enumerate(fs):
                    l0, l1 = g.label[f.vars[0]], g.label[f.vars[1]]
                    g_ee[0, 1] += w1p[l0, l1] - E2[n_]
        for f in g.factors:
            if f.arity != 1:
                continue
            t = unary[f.id]
            th = t / t.sum()
            l = g.label[f.vars[0]]
            if f.ftype == T_EN_EN:
                g_ee[0, 0] += pmi[l, f.obs] - th.dot(pmi[:, f.obs])
                if f.gap == 1:
                    g_ee[0, 1] += w1p[l, f.obs] - th.dot(w1p[:, f.obs])
            else:
                g_ed[0, 0] += tables.edT[f.obs, l] - th.dot(tables.edT[f.obs])
                g_ed[0, 1] += tables.pedT[f.obs, l] - th.dot(tables.pedT[f.obs])
                for e, d, k, val in sp:
                    if int(d) == f.obs:
                        g_ed[0, int(k)] += val * ((1.0 if int(e) == l else 0.0) - th[int(e)])
    return _finish(g, loopy, marg, logp, g_ee, g_ed, te.reshape(1, 3), td.reshape(1, 6), reg, lr)


def sgd_trajectory(model, sentences, roots_per_epoch, epochs=2, reg_param=0.2, init_lr=0.1, sweeps=3, fast=True):
    """train.py:617-638: per-sentence SGD, lr = 0.1/(1+0.3 epoch), reg = reg_param/N, theta updated in place
    after every sentence (batch_sgd_accumulate :400-416).  No shuffle: sentence order is an input."""
    te = np.zeros((1, 3))
    td = np.zeros((1, 6))
    N = len(sentences)
    reg = float(reg_param) / float(N)                                             # train.py:158
    traj, logps = [], []
    for epoch in range(epochs):
        lr = init_lr / float(1.0 + epoch * 0.3)                                   # train.py:621
        for si, s in enumerate(sentences):
            if fast:
                r = run_fast(Tables(model, te, td), s, roots_per_epoch[epoch][si], sweeps, reg, lr)
            else:
                r = run_literal(model, s, te, td, roots_per_epoch[epoch][si], sweeps, reg, lr, keep_messages=False)
            te = te + r['g_ee_ret']
            td = td + r['g_ed_ret']
            logps.append(r['logp'])
            traj.append(np.concatenate([te[0], td[0]]))
    return np.array(traj), np.array(logps)


def sgd_trajectory_user_adapt(model, sentences, users_of, roots_per_epoch, users, epochs=2, reg_param=0.2, ua_scale=1.0,
                              init_lr=0.1, sweeps=3):
    """train.py --user_adapt: the per-user theta REPLACES the base theta when the potentials are built (:224-229,
    :242-245); the unregularised gradient then updates BOTH the base theta (regularised with reg) and the user's theta
    (regularised with reg * reg_param_ua_scale), :379-390 and :402-409."""
    te, td = np.zeros((1, 3)), np.zeros((1, 6))
    d2t = {u: (np.zeros((1, 3)), np.zeros((1, 6))) for u in users}
    reg = float(reg_param) / float(len(sentences))
    traj = []
    for epoch in range(epochs):
        lr = init_lr / float(1.0 + epoch * 0.3)
        for si, s in enumerate(sentences):
            u = users_of[si]
            ue, ud = d2t[u]
            r = run_fast(Tables(model, ue, ud), s, roots_per_epoch[epoch][si], sweeps)
            g_ee, g_ed = r['g_ee_unreg'], r['g_ed_unreg']
            ue_new = ue + lr * (g_ee - reg * ua_scale * ue)
            ud_new = ud + lr * (g_ed - reg * ua_scale * ud)
            te = te + lr * (g_ee - reg * te)
            td = td + lr * (g_ed - reg * td)
            d2t[u] = (ue_new, ud_new)
            row = [te[0], td[0]]
            for uu in users:
                row += [d2t[uu][0][0], d2t[uu][1][0]]
            traj.append(np.concatenate(row))
    return np.array(traj)


# ----------------------------------------------------------------------------------------- chunked evaluator (large V)
def run_chunked(model, sents, theta_ee, theta_ed, roots_list, sweeps, block=1024, workers=None):
    """Inference (marginals, top-1, log-posterior, label rank) of several sentences at vocabulary sizes whose V x V float64
    tables do not fit in memory (BASELINE config C5: V = 50 000 -> 20 GB per table).  Same numbers as run_fast / run_literal
    (LBP.py:218-245, :377-400, :490-526, :247-259); the pairwise tables exp(phi . theta) (train.py:218-219, :252-253) are never
    materialised: every schedule level of ALL sentences is contracted in one pass over row blocks of the float32 feature
    planes, T[a0:a1, :] = exp(te0 * pmi[a0:a1, :] + te2) built on the fly in float64 and discarded.

    Sentences advance in lock step through their own update sequences (generators yield the GEMV requests of their next
    run of factor->variable updates).  `model` needs 'pmi', 'pmi_w1' (any float dtype), 'ed', 'ped'."""
    import os as _os
    from concurrent.futures import ThreadPoolExecutor
    try:
        from threadpoolctl import threadpool_limits
    except Exception:                                                             # pragma: no cover
        threadpool_limits = None
    te = np.asarray(theta_ee, dtype=np.float64).reshape(3)
    td = np.asarray(theta_ed, dtype=np.float64).reshape(6)
    pmi, w1p = model['pmi'], model['pmi_w1']
    V = pmi.shape[0]
    uni = np.full(V, 1.0 / V)
    workers = workers or min(32, _os.cpu_count() or 1)

    def sentence(sent, roots):
        g = Graph(sent)
        unary = {}
        for f in g.factors:
            if f.arity != 1:
                continue
            if f.ftype == T_EN_DE:
                z = td[0] * np.asarray(model['ed'][:, f.obs], dtype=np.float64) + \
                    td[1] * np.asarray(model['ped'][:, f.obs], dtype=np.float64) + td[5]
                for e, d, k, val in sent.sparse:
                    if int(d) == f.obs:
                        z[int(e)] += td[int(k)] * val
                t = np.exp(z)
            else:                                                                 # column of pot_en_en(_w1), LBP.py:702-703
                z = te[0] * np.asarray(pmi[:, f.obs], dtype=np.float64) + te[2]
                if f.gap == 1:
                    z = z + te[1] * np.asarray(w1p[:, f.obs], dtype=np.float64)
                t = np.exp(z)
            unary[f.id] = t
        uprod = {}
        for v in g.var_ids:
            m = uni.copy()
            for fid in g.facset[v]:
                if g.factors[fid].arity == 1:
                    m = m * (unary[fid] / unary[fid].sum())
            uprod[v] = m
        v2f, f2v = {}, {}
        for f in g.factors:
            if f.arity == 2:
                for v in f.vars:
                    v2f[v, f.id] = uni
                    f2v[f.id, v] = uni
        loopy, seq = g.update_sequence(roots, sweeps)
        i = 0
        while i < len(seq):
            frm, to = seq[i]
            if frm[0] == 'X':
                v, fid = frm[1], to[1]
                m = uprod[v]
                for of in g.facset[v]:
                    if of != fid and g.factors[of].arity == 2:
                        m = m * f2v[of, v]
                s = m.sum()
                v2f[v, fid] = m / s if s > 0 else uni
                i += 1
                continue
            run = []
            while i < len(seq) and seq[i][0][0] == 'F':
                fid, v = seq[i][0][1], seq[i][1][1]
                if g.factors[fid].arity == 2:
                    run.append((fid, v))
                i += 1
            req = []
            for fid, v in run:
                f = g.factors[fid]
                to_dim0 = (f.vars[0] == v)
                o = f.vars[1] if to_dim0 else f.vars[0]
                req.append(((f.gap == 1, to_dim0), v2f[o, fid]))
            if req:
                res = yield req
                for (fid, v), row in zip(run, res):
                    s = row.sum()
                    f2v[fid, v] = row / s if s > 0 else uni
        marg = []
        for v in g.var_ids:
            m = uprod[v]
            for fid in g.facset[v]:
                if g.factors[fid].arity == 2:
                    m = m * f2v[fid, v]
            s = m.sum()
            marg.append(m / s if s > 0 else uni)
        marg = np.stack(marg)
        logp = 0.0
        for i_, v in enumerate(g.var_ids):
            p = marg[i_, g.label[v]]
            logp += np.log(p) if p > 0 else -99.99
        z13 = np.zeros((1, 3)); z16 = np.zeros((1, 6))
        return _finish(g, loopy, marg, logp, z13, z16, te.reshape(1, 3), td.reshape(1, 6), 0.0, 1.0)

    def contract(requests):
        """requests: list of (key, vector); key = (gap1, to_dim0).  One pass over the row blocks of the planes."""
        by_key = {}
        for n, (key, vec) in enumerate(requests):
            by_key.setdefault(key, []).append(n)
        M = {key: np.stack([requests[n][1] for n in idx]) for key, idx in by_key.items()}
        need_w1 = any(k[0] for k in M)
        out = {key: np.zeros((len(idx), V)) for key, idx in by_key.items()}
        blocks = [(a0, min(V, a0 + block)) for a0 in range(0, V, block)]
        n_w = max(1, min(workers, len(blocks)))
        partial = [{key: np.zeros((len(idx), V)) for key, idx in by_key.items() if not key[1]} for _ in range(n_w)]

        def work(w):
            for bi in range(w, len(blocks), n_w):
                a0, a1 = blocks[bi]
                Tb = np.exp(te[0] * np.asarray(pmi[a0:a1], dtype=np.float64) + te[2])
                T1b = Tb * np.exp(te[1] * np.asarray(w1p[a0:a1], dtype=np.float64)) if need_w1 else None
                for key, m in M.items():
                    tb = T1b if key[0] else Tb
                    if key[1]:                                                    # out[a] = sum_b T[a, b] m[b]
                        out[key][:, a0:a1] = m.dot(tb.T)
                    else:                                                         # out[b] = sum_a m[a] T[a, b]
                        partial[w][key] += m[:, a0:a1].dot(tb)

        from contextlib import nullcontext
        with (threadpool_limits(limits=1, user_api='blas') if threadpool_limits else nullcontext()):
            with ThreadPoolExecutor(n_w) as ex:                                   # NumPy releases the GIL inside exp / dot
                list(ex.map(work, range(n_w)))
        for key in out:
            if not key[1]:
                for w in range(n_w):
                    out[key] += partial[w][key]
        res = [None] * len(requests)
        for key, idx in by_key.items():
            for j, n in enumerate(idx):
                res[n] = out[key][j]
        return res

    gens = [sentence(s, r) for s, r in zip(sents, roots_list)]
    results = [None] * len(gens)
    pending = {}
    for i, gen in enumerate(gens):
        try:
            pending[i] = next(gen)
        except StopIteration as e:
            results[i] = e.value
    while pending:
        order = sorted(pending)
        flat = [rq for i in order for rq in pending[i]]
        res = contract(flat)
        pos = 0
        nxt = {}
        for i in order:
            n = len(pending[i])
            try:
                nxt[i] = gens[i].send(res[pos:pos + n])
            except StopIteration as e:
                results[i] = e.value
            pos += n
        pending = nxt
    return results
