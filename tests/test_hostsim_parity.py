"""CPU-only parity of the kernel task bodies (host-simulator build) against the oracle.

The simulator is a test aid: it executes the SAME task functors the CUDA kernels run, one
work item at a time, so layout/arithmetic logic is checked in the GPU-less container.  The
GPU tests (tests/test_gpu_parity.py) repeat these checks on the real sm_100a library.
"""
import common


def test_field_ops(hostsim):
    common.check_field_ops(hostsim)


def test_hash80(hostsim):
    common.check_hash(hostsim)


def test_p256_mul(hostsim):
    common.check_p256_mul(hostsim)


def test_params_and_commit(hostsim):
    P, po = common.make_params(hostsim, seed=5)
    common.check_tom_commit(hostsim, P, po)
    hostsim.params_destroy(P)


def test_prove_bit_exact_small_ring(hostsim):
    common.check_prove_parity(hostsim, B=2, N=6, seed=3)


def test_prove_bit_exact_sec_level_16(hostsim):
    # smaller SecLevel keeps the oracle fast while exercising every code path
    common.check_prove_parity(hostsim, B=3, N=17, seed=4, sec_level=16)


def test_prove_few_distinct_keys_wide_key_tables(hostsim):
    common.check_prove_few_keys(hostsim, B=10, N=5, signers=1, seed=41, sec_level=16)     # one key: 6/7-bit windows
    common.check_prove_few_keys(hostsim, B=12, N=5, signers=3, seed=42, sec_level=16, spots=(1, 11))


def test_verify_decisions_match_oracle(hostsim):
    common.check_verify_parity(hostsim, N=6, seed=3, tampers=16)


def test_verify_sample_count_is_a_parameter(hostsim):
    """zka_verify_batch_ex: 5, 33 and all 80 sampled repetitions (80 = four MSM segments), verdicts incl. tampered
    proofs equal the oracle's (C++ port of the reference algorithms) under identical randomness."""
    import __graft_entry__ as g
    from zkp_ecdsa_b200.capi import ZkaLib
    g.build_oracle_cpu()
    cpu = ZkaLib(g.ORACLE_CPU)
    for K in (5, 33, 80):
        common.check_verify_samples(hostsim, K, N=5, seed=25, sec_level=80, tampers=4, oracle=cpu)
    common.check_verify_samples(hostsim, 7, N=4, seed=26, sec_level=20, tampers=2, oracle='python')
