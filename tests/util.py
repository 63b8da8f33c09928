"""Shared helpers of the test-suite: the same scheme described to the oracle and to the library."""
import hashlib

import numpy as np

import sda_b200
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import LinearSecretSharingScheme as LSS
from sda_b200 import params

P61 = params.P61
PGEN = params.P61_GENERIC


def seed_bytes(tag):
    return hashlib.sha256(str(tag).encode()).digest()


def to_oracle_sharing(O, s):
    c = s.c
    return O.SharingScheme(c.kind, c.share_count, c.secret_count, c.privacy_threshold, c.modulus, c.omega_secrets,
                           c.omega_shares)


def to_oracle_masking(O, s):
    c = s.c
    return O.MaskingScheme(c.kind, c.modulus, c.dimension, c.seed_bitsize)


def canon(O, modulus, a):
    return O.canonical(modulus, np.asarray(a, dtype=np.int64).ravel()).reshape(np.shape(a))


def oracle_generate(O, s, secrets, seed, rounds=20, matrix=False):
    """reference ShareGenerator::generate with OsRng replaced by ChaChaRng::from_seed(seed)"""
    rng = O.rng_from_seed_bytes(seed, rounds)
    return O.share_generate(to_oracle_sharing(O, s), secrets, rng, matrix=matrix)


def rand_secrets(rng, dim, modulus, kind="canonical"):
    if kind == "canonical":
        return rng.integers(0, modulus, size=dim, dtype=np.int64)
    if kind == "signed":      # anything an i64 can hold short of overflow in the reference's adds
        return rng.integers(-(1 << 62), 1 << 62, size=dim, dtype=np.int64)
    raise ValueError(kind)


def packed_scheme(p, k, t, n, O):
    """PackedShamir over any prime p: two roots of distinct prime orders found with the oracle."""
    def prime_orders(lo):
        q = lo
        while True:
            if all(q % d for d in range(2, int(q ** 0.5) + 1)) and (p - 1) % q == 0:
                yield q
            q += 1
            if q > 4096:
                return
    qs = next(prime_orders(k + t + 1))
    qh = next(x for x in prime_orders(n + 1) if x != qs)
    return LSS.PackedShamir(k, n, t, p, O.find_root_of_order(p, qs), O.find_root_of_order(p, qh))
