"""GPU tier: the drop-in LBP.py / train_compat API through the real kernels, against the reference fixtures."""
import glob
import os

import numpy as np
import pytest

import lbp_api_checks
from macaronicusermodeling_b200 import LBP, build
from macaronicusermodeling_b200 import train_compat as tc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'graph_*.npz')))


@pytest.fixture(autouse=True)
def _real_backend():
    build.build()
    LBP._KERNELS_FACTORY = None
    LBP._ENGINES.clear()
    yield
    LBP._ENGINES.clear()


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_lbp_api_matches_reference_fixture(path):
    lbp_api_checks.check_lbp_api_fixture(path)


def test_per_factor_api_matches_graph_gradient():
    """FactorNode.get_gradient summed over the factors (the reference's own loop, LBP.py:304-319) equals the fused
    engine gradient; cell_gradient == cell_gradient_alt (the reference's commented-out cross-check, LBP.py:595-596)."""
    z = np.load(os.path.join(GOLDEN, 'graph_toy3.npz'), allow_pickle=False)
    fg, spec = lbp_api_checks.graph_from_fixture(z)
    roots = [int(r) for r in z['roots']]
    fg.initialize(roots[0])
    fg.treelike_inference(spec['sweeps'], roots[1:])
    g_ee = np.zeros((1, 3)); g_ed = np.zeros((1, 6))
    for f in fg.factors:
        g = f.get_gradient()
        assert g.shape == (1, 3 if f.factor_type == 'en_en' else 6)
        if f.factor_type == 'en_en':
            g_ee += g
        else:
            g_ed += g
        np.testing.assert_allclose(f.cell_gradient(), f.cell_gradient_alt(), atol=1e-12)
        b = f.get_factor_beliefs()
        assert abs(b.sum() - 1.0) < 1e-9 and (b >= 0).all()
    np.testing.assert_allclose(g_ee, z['g_ee_unreg'], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(g_ed, z['g_ed_unreg'], rtol=1e-4, atol=2e-6)


def test_batch_sgd_drop_in_trajectory():
    """train.py's per-sentence loop written against the drop-in API: batch_sgd + batch_sgd_accumulate, 2 epochs"""
    z = np.load(os.path.join(GOLDEN, 'sgd_trajectory.npz'), allow_pickle=False)
    V, Vd = z['pmi'].shape[0], z['ed'].shape[1]
    en_domain = ['e%d' % i for i in range(V)]
    de_domain = ['d%d' % i for i in range(Vd)]
    en2id = dict((e, i) for i, e in enumerate(en_domain))
    de2id = dict((d, i) for i, d in enumerate(de_domain))
    pw = tc.make_phi_wrapper(z['pmi'], z['pmi_w1'], z['ed'], z['ped'])
    te, td = np.zeros((1, 3)), np.zeros((1, 6))
    sents = [str(s) for s in z['sentences']]
    roots = z['roots'].tolist()
    opts = tc.default_options(session_history=True)
    traj = []
    for epoch in range(2):
        lr = 0.1 / float(1.0 + epoch * 0.3)
        for si, s in enumerate(sents):
            res = tc.batch_sgd(s, tc.F_EN_EN_NAMES, tc.F_EN_DE_NAMES, te, td, pw, lr, en_domain, de2id, en2id, {},
                               options=opts, N=len(sents), de_domain=de_domain, roots=roots[epoch][si])
            tc.batch_sgd_accumulate(res, te, td)
            traj.append(np.concatenate([te[0], td[0]]))
    np.testing.assert_allclose(np.array(traj), z['traj'], rtol=1e-4, atol=2e-7)


def test_oov_label_exits_like_reference():
    with pytest.raises(SystemExit):
        LBP.VariableNode(id=0, var_type=LBP.VAR_TYPE_PREDICTED, domain_type='en', domain=['a', 'b'], supervised_label='zzz')


def test_unsupported_gap_raises_like_reference():
    f = LBP.FactorNode(id=0, factor_type='en_en')
    f.gap = 0
    f.graph = LBP.FactorGraph(tc.F_EN_EN_NAMES, tc.F_EN_DE_NAMES, np.zeros((1, 3)), np.zeros((1, 6)), None, None, None)
    with pytest.raises(BaseException):
        f.get_pot()


def test_user_adapt_drop_in_trajectory():
    lbp_api_checks.check_user_adapt_drop_in()


def test_adapt_trainer_trajectory():
    lbp_api_checks.check_adapt_trainer(lambda m: __import__('macaronicusermodeling_b200.engine', fromlist=['Engine']).Engine(m))


@pytest.mark.parametrize('name', ['toy3', 'tree2', 'revealed', 'k8'])
def test_per_node_update_api(name):
    """VariableNode / FactorNode.update_message_to (eager mode): the caller drives the schedule one message at a time"""
    lbp_api_checks.check_per_node_updates(os.path.join(GOLDEN, 'graph_%s.npz' % name))


def test_explicit_potential_table_graph():
    """PotentialTable(table=...) graphs (run.py / toy style) run through the eager path, pinned on a reference fixture"""
    lbp_api_checks.check_explicit_graph(np.load(os.path.join(GOLDEN, 'graphx_explicit.npz'), allow_pickle=False))


def test_au_all_functions():
    from macaronicusermodeling_b200.array_utils import c_array_utils as au
    lbp_api_checks.check_au_remaining(au)
