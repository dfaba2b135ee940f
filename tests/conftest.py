import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA sm_100 device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope='session')
def manifest():
    with open(os.path.join(GOLDEN, 'MANIFEST.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def seeded_state_dict():
    """torch.manual_seed(0); MobilePoserNet() -- bit-identical to the reference's init (test_oracle checks the hashes)."""
    import mobileposer_b200 as mp
    torch.manual_seed(0)
    net = mp.MobilePoserNet()
    return {k: v.clone() for k, v in net.state_dict().items()}


@pytest.fixture(scope='session')
def oracle(seeded_state_dict):
    from oracle.torch_port import OraclePoser
    return OraclePoser(seeded_state_dict)


@pytest.fixture(scope='session')
def oracle64(seeded_state_dict):
    """The same torch CPU kernels in float64: the arbiter between the reference's fp32 result and the CUDA path."""
    from oracle.torch_port import OraclePoser
    return OraclePoser(seeded_state_dict, dtype=torch.float64)


@pytest.fixture(scope='session')
def wc_state_dict(seeded_state_dict, manifest):
    """Well-conditioned seeded weights (synthetic.well_conditioned_state_dict), hash-pinned to what the live reference loaded."""
    import hashlib
    from mobileposer_b200.synthetic import well_conditioned_state_dict
    sd = well_conditioned_state_dict(seeded_state_dict)
    for k, h in manifest['wc_weight_sha256'].items():
        assert hashlib.sha256(sd[k].contiguous().numpy().tobytes()).hexdigest() == h, k
    return sd


@pytest.fixture(scope='session')
def wc_oracle(wc_state_dict):
    from oracle.torch_port import OraclePoser
    return OraclePoser(wc_state_dict)


@pytest.fixture(scope='session')
def wc_oracle64(wc_state_dict):
    from oracle.torch_port import OraclePoser
    return OraclePoser(wc_state_dict, dtype=torch.float64)
