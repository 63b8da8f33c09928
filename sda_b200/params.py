"""Scheme parameters of the BASELINE.json configurations.

The reference hard-codes its parameters in tests (integration-tests/tests/full_loop.rs:57-64:
p=433, omega_secrets=354 of order 8, omega_shares=150 of order 9).  BASELINE.json's 61-bit
configurations name only (k, n); the rest is fixed here and recorded in every bench line:

  * prime p = 2^61 - 1  (p - 1 = 2 . 3^2 . 5^2 . 7 . 11 . 13 . 31 . 41 . 61 . 151 . 331 . 1321)
  * privacy_threshold t = n - k (every share is needed to reconstruct: k + t = n)
  * omega_secrets / omega_shares: smallest-base elements (c^((p-1)/q), c = 2, 3, ...) of two
    different PRIME orders q_s >= k + t + 1 and q_h >= n + 1, so the secret-side points
    {w_s^i} and the share-side points {w_h^j} meet only in 1.  tests/test_params.py re-derives
    them with the oracle's `sdao_find_root_of_order`.
"""
from .crypto import LinearSecretSharingScheme

P61 = (1 << 61) - 1
# a 61-bit prime that is not of Mersenne form: exercises the generic (reciprocal) reduction
P61_GENERIC = 2305843009213693921   # largest prime below 2^61 - 1 - 30; checked in tests/test_params.py

ROOT_ORDER_7 = 69203453413471971
ROOT_ORDER_11 = 54008984094220448
ROOT_ORDER_13 = 844735144842896729
ROOT_ORDER_31 = 484083891529811867   # for shapes beyond the BASELINE ones (k + t + 1 <= 31, n + 1 <= 41)
ROOT_ORDER_41 = 439424789145975530

# the reference's own packed-Shamir test parameters (full_loop.rs:57-64)
REFERENCE_TEST = dict(secret_count=3, share_count=8, privacy_threshold=4, prime_modulus=433,
                      omega_secrets=354, omega_shares=150)


def additive(share_count=3, modulus=P61):
    return LinearSecretSharingScheme.Additive(share_count, modulus)


def config2():
    """additive sharing, dim=1M, 61-bit prime, 1024 participants, 3-way split"""
    return additive(3, P61)


def config3():
    """packed Shamir k=3/n=5 (t=2), 61-bit prime"""
    return LinearSecretSharingScheme.PackedShamir(3, 5, 2, P61, ROOT_ORDER_7, ROOT_ORDER_11)


def config4():
    """packed Shamir k=5/n=9 (t=4), 61-bit prime"""
    return LinearSecretSharingScheme.PackedShamir(5, 9, 4, P61, ROOT_ORDER_11, ROOT_ORDER_13)


def config5():
    """packed Shamir k=3/n=7 (t=4), 61-bit prime (federated-model proxy)"""
    return LinearSecretSharingScheme.PackedShamir(3, 7, 4, P61, ROOT_ORDER_11, ROOT_ORDER_13)


def reference_test():
    return LinearSecretSharingScheme.PackedShamir(**REFERENCE_TEST)
