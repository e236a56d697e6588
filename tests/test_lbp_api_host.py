"""CPU tier: the drop-in LBP.py / train_compat API (recording, lowering, read-back) on emulated kernels."""
import glob
import os

import pytest

import lbp_api_checks
from fake_kernels import FakeKernels
from macaronicusermodeling_b200 import LBP, build

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'graph_*.npz')))


@pytest.fixture(autouse=True)
def _fake_backend():
    build.build()
    LBP._KERNELS_FACTORY = FakeKernels
    LBP._ENGINES.clear()
    yield
    LBP._KERNELS_FACTORY = None
    LBP._ENGINES.clear()


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[6:-4] for p in CASES])
def test_lbp_api_matches_reference_fixture(path):
    lbp_api_checks.check_lbp_api_fixture(path)


def test_dynamic_features_are_fixed_when_the_graph_is_built():
    lbp_api_checks.check_dynamic_features_are_fixed_at_build_time(os.path.join(GOLDEN, 'graph_peaked.npz'))


def test_explicit_table_graph_initialises_like_the_reference():
    import numpy as np
    lbp_api_checks.check_explicit_graph_structure(np.load(os.path.join(GOLDEN, 'graphx_explicit.npz'), allow_pickle=False))


def test_au_surface_is_complete():
    """every function of c_array_utils.pyx exists under the same name"""
    from macaronicusermodeling_b200.array_utils import c_array_utils as au
    for name in ('pointwise_multiply', 'clip', 'sparse_normalize', 'normalize', 'induce_s_pointwise_multiply_clip', 'induce_s',
                 'induce_s_mutliply_clip', 'induce_s_multiply_threshold', 'dense_dot', 'dense_pointwise_multiply',
                 'make_sparse_and_dot', 'sparse_pointwise_multiply', 'sparse_dot', 'sparse_multiply_and_normalize',
                 'sd_matrix_multiply', 'sd_pointwise_multiply', 'ss_matix_multiply', 'make_adapt_phi', 'set_adaptation',
                 'set_adaptation_off', 'set_original', 'sparse_vec_mat_dot'):
        assert callable(getattr(au, name)), name
    assert au.K == 100


def test_params_roundtrip(tmp_path):
    lbp_api_checks.check_params_roundtrip(tmp_path)


def test_import_surface():
    """train.py:9 imports exactly these names from LBP"""
    from macaronicusermodeling_b200.LBP import (FactorGraph, FactorNode, PhiWrapper, PotentialTable, VariableNode,  # noqa: F401
                                                VAR_TYPE_GIVEN, VAR_TYPE_PREDICTED)
    from macaronicusermodeling_b200.array_utils import c_array_utils as au
    for name in ('pointwise_multiply', 'normalize', 'dense_dot', 'dense_pointwise_multiply'):
        assert callable(getattr(au, name))


def test_user_adapt_drop_in_trajectory():
    lbp_api_checks.check_user_adapt_drop_in()


def test_adapt_trainer_trajectory():
    lbp_api_checks.check_adapt_trainer(lambda m: __import__('macaronicusermodeling_b200.engine', fromlist=['Engine']).Engine(m, kernels=FakeKernels()))
