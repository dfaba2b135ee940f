"""CPU suite: the C-ABI library builds, loads, and exports exactly what include/mobileposer_b200.h declares.
No compute call is made here (there is no GPU in this tier)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from mobileposer_b200 import _cabi
from mobileposer_b200.build import LIB_PATH, build

HEADER = os.path.join(ROOT, 'include', 'mobileposer_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mp_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    build()
    assert os.path.exists(LIB_PATH)
    return _cabi.lib()


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'


def test_binding_table_covers_the_header(lib):
    assert sorted(_cabi.SIGNATURES) == declared_symbols()


def test_abi_version_and_struct_sizes(lib):
    assert lib.mp_abi_version() == 1
    # struct mp_rnn_weights: 5 int32 (+4 pad) + 4 pointers + 4 x [2][2] pointers
    assert ctypes.sizeof(_cabi.RnnWeights) == 24 + 4 * 8 + 16 * 8
    assert _cabi.ONLINE_STATE_BYTES == 64


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_no_cpu_fallback_in_the_product_path():
    """The package must not import the oracle, and CPU tensors must be rejected loudly."""
    import torch
    import mobileposer_b200 as mp
    pkg = os.path.join(ROOT, 'mobileposer_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn
            assert 'import_module' not in src and '__import__' not in src, fn
    net = mp.MobilePoserNet().eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net.forward_offline(torch.zeros(1, 4, 60), [4])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net.joints(torch.zeros(1, 4, 60), [4])
