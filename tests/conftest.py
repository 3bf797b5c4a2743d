import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def hostsim():
    """Host-simulator build of the task bodies (tests only; never part of the product)."""
    import __graft_entry__ as g
    g.build_hostsim()
    from zkp_ecdsa_b200.capi import ZkaLib
    # the CPU simulator builds its tables with plain loops: keep them small (13-bit windows also
    # exercise the unaligned digit extraction); the GPU tests run the default 16-bit windows
    old = os.environ.get('ZKA_TOM_W')
    os.environ['ZKA_TOM_W'] = '13'
    os.environ['ZKA_P256_HW'] = '8'
    try:
        return ZkaLib(g.HOSTSIM)
    finally:
        if old is None:
            os.environ.pop('ZKA_TOM_W', None)
        else:
            os.environ['ZKA_TOM_W'] = old
        os.environ.pop('ZKA_P256_HW', None)


@pytest.fixture(scope='session')
def hostsim_war():
    """Host-simulator build with ProofGroup = war256 (-DZKA_PG_WAR256); tests only."""
    import __graft_entry__ as g
    g.build_hostsim(war=True)
    from zkp_ecdsa_b200.capi import ZkaLib
    old = os.environ.get('ZKA_TOM_W')
    os.environ['ZKA_TOM_W'] = '9'
    os.environ['ZKA_P256_HW'] = '8'
    try:
        L = ZkaLib(g.HOSTSIM_WAR)
    finally:
        if old is None:
            os.environ.pop('ZKA_TOM_W', None)
        else:
            os.environ['ZKA_TOM_W'] = old
        os.environ.pop('ZKA_P256_HW', None)
    assert L.group == 'war256' and (L.wp, L.ws) == (65, 32)
    return L


@pytest.fixture(scope='session')
def gpu_engine():
    """The product path: libzkattest.so on cuda:0.  Fails loudly without it."""
    from zkp_ecdsa_b200 import api
    return api.Engine(device=0)


@pytest.fixture(scope='session')
def gpu_engine_war():
    """The war256 build of the product (libzkattest_war256.so) on cuda:0."""
    from zkp_ecdsa_b200 import api
    return api.Engine(device=0, proof_group='war256')
