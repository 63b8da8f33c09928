"""sda_b200/params.py against the oracle's number theory (CPU)."""
from sda_b200 import params
import util


def _order(O, w, p, limit=64):
    x, e = w % p, 1
    while x != 1:
        x = (x * w) % p
        e += 1
        assert e <= limit
    return e


def test_roots_of_unity(oracle):
    p = params.P61
    for q, w in ((7, params.ROOT_ORDER_7), (11, params.ROOT_ORDER_11), (13, params.ROOT_ORDER_13),
                 (31, params.ROOT_ORDER_31), (41, params.ROOT_ORDER_41)):
        assert oracle.find_root_of_order(p, q) == w
        assert _order(oracle, w, p) == q


def test_generic_prime_is_prime():
    p = params.P61_GENERIC
    assert p < (1 << 61) - 1 and p.bit_length() == 61
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):   # deterministic Miller-Rabin below 3.3e24
        d, s = p - 1, 0
        while d % 2 == 0:
            d //= 2
            s += 1
        x = pow(a, d, p)
        ok = x in (1, p - 1)
        for _ in range(s - 1):
            x = x * x % p
            ok |= x == p - 1
        assert ok


def test_config_points_are_disjoint(oracle):
    for s in (params.config3(), params.config4(), params.config5(), params.reference_test()):
        c = s.c
        p = c.modulus
        a = [pow(c.omega_secrets, i, p) for i in range(c.secret_count + c.privacy_threshold + 1)]
        b = [pow(c.omega_shares, j, p) for j in range(1, c.share_count + 1)]
        assert len(set(a)) == len(a) and len(set(b)) == len(b)
        assert not (set(a) & set(b))
        assert c.privacy_threshold + c.secret_count <= c.share_count
