"""CPU tier: the schedule compiler's TEMPLATE path (csrc/plan.cpp: one compiled schedule per distinct (wiring, roots, flags),
relocated into the batch) must emit the same blob, bit for bit, as the literal per-graph path (MLBP_PLAN_TEMPLATES=0) -- on
mixed layouts, every flag combination the engine uses, cold and warm cache."""
import ctypes
import os

import numpy as np
import pytest

from fake_kernels import FakeKernels
from macaronicusermodeling_b200 import _lib, build, synth
from macaronicusermodeling_b200.engine import Corpus, Engine, PLAN_BLOB_WORDS


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build()


def _blob(eng, corpus, roots, sweeps, want_grad, want_marg, fold, reuse_z):
    handle, sizes = eng.compile(corpus, roots, sweeps, want_grad, want_marg, fold=fold, reuse_z=reuse_z)
    try:
        b = np.empty(int(sizes[PLAN_BLOB_WORDS]), dtype=np.int32)
        _lib.check(_lib.load().mlbp_plan_export(handle, ctypes.c_void_p(b.ctypes.data)))
    finally:
        _lib.load().mlbp_plan_destroy(handle)
    return b, [int(x) for x in sizes[:9]], int(sizes[9]), int(sizes[10])


@pytest.mark.parametrize('threads', ['1', '3'])
def test_templated_blob_equals_literal_blob(threads, monkeypatch):
    monkeypatch.setenv('MLBP_PLAN_THREADS', threads)
    model = synth.make_model(64, 16, seed=1)
    eng = Engine(model, kernels=FakeKernels())
    layouts = ['pppp', 'gpgpp', 'ppgpgp', 'pp', 'pgppg', 'gpg', 'ppppppp', 'prpgp', 'ppp', 'gppg', 'p', 'gp']
    sents = [synth.sentence_to_arrays(synth.make_sentence(model, layouts[(i * 7) % len(layouts)], seed=50 + i, n_history=3))
             for i in range(90)]
    sents += synth.make_corpus(model, 40, k=12, g=0, seed=2) + synth.make_corpus(model, 30, k=9, g=3, seed=5)
    corpus = Corpus(sents)
    roots = corpus.roots_from_positions(synth.draw_roots(sents, 3, seed=3))
    seen_hit = False
    for sweeps, wg, wm, fold, rz in [(3, True, True, True, False), (3, True, True, True, True), (3, False, True, True, True),
                                     (3, True, True, False, True), (1, True, False, True, True), (0, False, True, True, True)]:
        r = roots if sweeps == 3 else np.ascontiguousarray(roots[:, :1 + sweeps])
        monkeypatch.setenv('MLBP_PLAN_TEMPLATES', '0')
        b0, s0, h0, m0 = _blob(eng, corpus, r, sweeps, wg, wm, fold, rz)
        assert h0 == 0 and m0 == 0
        monkeypatch.setenv('MLBP_PLAN_TEMPLATES', '1')
        b1, s1, h1, m1 = _blob(eng, corpus, r, sweeps, wg, wm, fold, rz)      # cold for this flag set (partly: repeats inside the batch hit)
        b2, s2, h2, m2 = _blob(eng, corpus, r, sweeps, wg, wm, fold, rz)      # warm: every graph is a hit
        assert h1 + m1 == len(sents) and h2 == len(sents) and m2 == 0
        seen_hit = seen_hit or h1 > 0
        assert s0 == s1 == s2
        np.testing.assert_array_equal(b0, b1)
        np.testing.assert_array_equal(b0, b2)
    assert seen_hit


def test_first_root_is_left_out_of_the_key_for_connected_graphs(monkeypatch):
    """roots[0] only feeds has_loops (LBP.py:176): two draws that differ in it alone share one template, and the blob is the
    literal path's in both cases"""
    model = synth.make_model(64, 16, seed=1)
    eng = Engine(model, kernels=FakeKernels())
    sents = synth.make_corpus(model, 1, k=7, g=1, seed=77)
    corpus = Corpus(sents)
    pos = [int(p) for p in sents[0].predicted]
    monkeypatch.setenv('MLBP_PLAN_CACHE_MB', '64')
    blobs = []
    for first in (pos[0], pos[3]):
        roots = corpus.roots_from_positions([[first, pos[1], pos[2], pos[5]]])
        monkeypatch.setenv('MLBP_PLAN_TEMPLATES', '0')
        lit = _blob(eng, corpus, roots, 3, True, True, True, True)
        monkeypatch.setenv('MLBP_PLAN_TEMPLATES', '1')
        tm = _blob(eng, corpus, roots, 3, True, True, True, True)
        np.testing.assert_array_equal(lit[0], tm[0])
        blobs.append(tm)
    assert blobs[1][2] == 1 and blobs[1][3] == 0          # the second draw was a hit
    np.testing.assert_array_equal(blobs[0][0], blobs[1][0])


def test_invalid_graph_is_reported_by_both_paths(monkeypatch):
    lib = _lib.load()
    var_off = np.array([0, 2], dtype=np.int32)
    pair_off = np.array([0, 1], dtype=np.int32)
    v0 = np.array([0], dtype=np.int32); v1 = np.array([5], dtype=np.int32); gap = np.array([0], dtype=np.int32)   # variable 5 of 2
    roots = np.array([0, 0, 0, 0], dtype=np.int32)
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    for mode in ('0', '1'):
        monkeypatch.setenv('MLBP_PLAN_TEMPLATES', mode)
        out = ctypes.c_void_p()
        rc = lib.mlbp_plan_compile(1, P(var_off), P(pair_off), P(v0), P(v1), P(gap), P(roots), 3, 3, ctypes.byref(out))
        assert rc != 0 and b'invalid graph' in lib.mlbp_last_error()
