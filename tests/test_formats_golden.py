"""On-disk formats against files written BY THE REFERENCE (tests/golden/formats.npz, made by make_golden.py --only-formats
from train.save_params / train.read_params, train.py:46-99): the checkpoint writer must produce the same bytes, the reader the
same values.  (The prediction / .dist lines of the same fixture are compared on the GPU in test_gpu_lbp_api.py.)"""
import io
import os

import numpy as np

from macaronicusermodeling_b200 import train_compat as tc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'formats.npz')


def _d2t(z):
    d2t = {}
    for i, d in enumerate(str(x) for x in z['domains']):          # the reference iterates its dict in insertion order
        d2t['en_en', d] = z['d_ee'][i:i + 1]
        d2t['en_de', d] = z['d_ed'][i:i + 1]
    return d2t


def test_save_params_writes_the_reference_bytes(tmp_path):
    z = np.load(GOLDEN, allow_pickle=False)
    p = str(tmp_path / 'ours.params')
    tc.save_params(io.open(p, 'w', encoding='utf8'), z['ee'], z['ed'], list(tc.F_EN_EN_NAMES), list(tc.F_EN_DE_NAMES), _d2t(z))
    assert open(p, 'rb').read() == str(z['params_text']).encode('utf8')


def test_read_params_parses_a_reference_written_file(tmp_path):
    z = np.load(GOLDEN, allow_pickle=False)
    p = str(tmp_path / 'ref.params')
    with open(p, 'wb') as f:
        f.write(str(z['params_text']).encode('utf8'))
    een, eet, edn, edt, d2t = tc.read_params(p)
    assert een == [str(x) for x in z['read_een']] and edn == [str(x) for x in z['read_edn']]
    np.testing.assert_array_equal(eet, z['read_ee'])
    np.testing.assert_array_equal(edt, z['read_ed'])
    for i, d in enumerate(str(x) for x in z['domains']):
        np.testing.assert_array_equal(d2t['en_en', d], z['read_d_ee'][i:i + 1])
        np.testing.assert_array_equal(d2t['en_de', d], z['read_d_ed'][i:i + 1])
    assert len(d2t) == 2 * len(z['domains'])
    # the adapt-aware reader of the CLI front end goes through the same parser
    from macaronicusermodeling_b200 import train_cli
    assert hasattr(train_cli, 'read_params_adapt')
