"""libzkattest.so loads and exports every symbol include/zkattest.h declares (no GPU compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'zkattest.h')).read()
    return sorted(set(re.findall(r'\b(zka_[a-z0-9_]+)\s*\(', hdr)))


def test_header_symbols_match_binding_list():
    from zkp_ecdsa_b200 import capi
    assert _declared() == sorted(capi.SYMBOLS)


def test_shared_library_exports_all_symbols():
    import __graft_entry__ as g
    g.build_lib()
    lib = ctypes.CDLL(g.LIB)
    for s in _declared():
        assert hasattr(lib, s), s
    lib.zka_version.restype = ctypes.c_int
    assert lib.zka_version() >= 1


def test_war256_library_exports_the_same_abi():
    """libzkattest_war256.so (ProofGroup = war256): same entry points, 65-byte points / 32-byte scalars."""
    import __graft_entry__ as g
    g.build_lib_war()
    lib = ctypes.CDLL(g.LIB_WAR)
    for s in _declared():
        assert hasattr(lib, s), s
    name, pb, sb = ctypes.create_string_buffer(32), ctypes.c_int(), ctypes.c_int()
    lib.zka_proof_group(name, 32, ctypes.byref(pb), ctypes.byref(sb))
    assert (name.value, pb.value, sb.value) == (b'war256', 65, 32)
    tom = ctypes.CDLL(g.LIB)
    tom.zka_proof_group(name, 32, ctypes.byref(pb), ctypes.byref(sb))
    assert (name.value, pb.value, sb.value) == (b'tomEdwards256', 67, 33)


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from zkp_ecdsa_b200 import api, capi
    with pytest.raises(capi.ZkaError):
        api.Engine(device=0)
