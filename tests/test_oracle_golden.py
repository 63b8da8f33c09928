"""The oracle against everything the reference pins for this path, plus the
properties the reference's tests imply.  CPU only."""
import json
import os

import numpy as np
import pytest

import util
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import LinearSecretSharingScheme as LSS
from sda_b200 import params

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference.json")))
ORC = json.load(open(os.path.join(HERE, "golden", "oracle.json")))


def full_loop(O, sharing, masking, inputs, seeds_tag, missing=()):
    """The reference's participate -> clerk -> reveal flow (participate.rs:53-76, clerk.rs:85-86,
    receive.rs:113-152) with every crypto call going to the oracle."""
    so, mo = util.to_oracle_sharing(O, sharing), util.to_oracle_masking(O, masking)
    dim = len(inputs[0])
    n = sharing.output_size()
    masks, per_clerk = [], [[] for _ in range(n)]
    for pi, secrets in enumerate(inputs):
        rng = O.rng_from_seed_bytes(util.seed_bytes(f"{seeds_tag}/{pi}"))
        mask, masked = O.mask(mo, secrets, rng)
        masks.append(mask)
        shares = O.share_generate(so, masked, rng)
        for c in range(n):
            per_clerk[c].append(shares[c])
    combined = [(c, O.share_combine(sharing.modulus, np.stack(per_clerk[c]))) for c in range(n) if c not in missing]
    masked_out = O.secret_reconstruct(so, dim, [c for c, _ in combined], np.stack([v for _, v in combined]))
    if masking.has_mask():
        mask_sum = O.mask_combine(mo, np.stack(masks))
        out = O.unmask(mo, mask_sum, masked_out)
    else:
        out = masked_out
    return O.positive(sharing.modulus, out)


@pytest.mark.parametrize("masking", ["none", "full", "chacha"])
@pytest.mark.parametrize("sharing", ["additive", "packed_shamir"])
def test_reference_full_loop(oracle, sharing, masking):
    """integration-tests/tests/full_loop.rs:29-67: 2 x [1,2,3,4] mod 433 -> [2,4,6,8]"""
    g = REF["full_loop"]
    s = LSS.Additive(**g["sharing"]["additive"]) if sharing == "additive" else LSS.PackedShamir(**g["sharing"]["packed_shamir"])
    m = {"none": LMS.None_(), "full": LMS.Full(433), "chacha": LMS.ChaCha(**g["masking"]["chacha"])}[masking]
    out = full_loop(oracle, s, m, [g["input"]] * g["participants"], f"fl/{sharing}/{masking}")
    assert out.tolist() == g["expected_positive"]


def test_reference_cli_walkthrough(oracle):
    """README.md:157: `0 2 2 4 4 6 6 8 8 10`  (BASELINE config #1)"""
    g = REF["cli_walkthrough"]
    out = full_loop(oracle, LSS.Additive(**g["sharing"]), LMS.None_(), g["inputs"], "cli")
    assert out.tolist() == g["expected"]


def test_packed_tolerates_missing_clerks(oracle):
    """reconstruction_threshold = t + k = 7 of 8 clerks suffice (protocol/src/crypto.rs:147-153)"""
    g = REF["full_loop"]
    s = LSS.PackedShamir(**g["sharing"]["packed_shamir"])
    for missing in [(0,), (7,), (3,)]:
        out = full_loop(oracle, s, LMS.None_(), [g["input"]] * 2, "miss", missing=missing)
        assert out.tolist() == g["expected_positive"]
    so = util.to_oracle_sharing(oracle, s)
    with pytest.raises(oracle.OracleError, match="Not enough shares to reconstruct"):
        oracle.secret_reconstruct(so, 4, [0, 1, 2, 3, 4, 5], np.zeros((6, 2), dtype=np.int64))


def test_chacha20_keystream_kat(oracle):
    g = REF["chacha20_kat"]
    st = np.zeros(16, dtype=np.uint32)
    st[:4] = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574]
    assert [f"{x:08x}" for x in oracle.chacha_block(st)] == g["block0"]
    st[12] = 1
    assert [f"{x:08x}" for x in oracle.chacha_block(st)[:4]] == g["block1_head"]
    # the rng object walks the same stream: 16 words of block 0 then block 1
    r = oracle.rng_from_seed([])
    words = [oracle.next_u32(r) for _ in range(20)]
    assert [f"{x:08x}" for x in words[:16]] == g["block0"]
    assert [f"{x:08x}" for x in words[16:]] == g["block1_head"]
    # next_u64: first word is the high half
    r = oracle.rng_from_seed([])
    assert oracle.next_u64(r) == (int(g["block0"][0], 16) << 32) | int(g["block0"][1], 16)


def test_tss_kat(oracle):
    g = REF["tss_kat"]
    s = oracle.packed_shamir(3, 8, 4, g["prime"], g["omega_secrets"], 150)
    poly, shares = oracle.tss_share_with_randomness(s, g["secrets"], g["randomness"], want_poly=True)
    assert oracle.canonical(433, poly).tolist() == g["polynomial"]
    assert oracle.canonical(433, shares).tolist() == g["shares_omega150_n8"]
    s26 = oracle.packed_shamir(3, 26, 4, g["prime"], g["omega_secrets"], 17)
    assert oracle.canonical(433, oracle.tss_share_with_randomness(s26, g["secrets"], g["randomness"])).tolist() == \
        g["shares_omega17_n26"]
    # the general interpolate-then-evaluate path is the same map as the FFT path
    for sch, exp in ((s, g["shares_omega150_n8"]), (s26, g["shares_omega17_n26"])):
        gen = oracle.tss_share_with_randomness(sch, g["secrets"], g["randomness"], force_general=True)
        assert oracle.canonical(433, gen).tolist() == exp
    # reconstruct from any 7 of the 8
    for drop in range(8):
        idx = [i for i in range(8) if i != drop]
        rec = oracle.tss_reconstruct(s, idx, [g["shares_omega150_n8"][i] for i in idx])
        assert oracle.canonical(433, rec).tolist() == g["secrets"]


def test_gen_range_model(oracle):
    g = REF["gen_range_model"]
    r = oracle.rng_from_seed(g["seed_words"])
    assert [oracle.gen_range(r, 0, 433) for _ in range(10)] == g["m433"]
    r = oracle.rng_from_seed(g["seed_words"])
    assert [oracle.gen_range(r, 0, params.P61) for _ in range(4)] == g["m2p61m1"]


def test_gen_range_rejection_shifts_stream(oracle):
    """modulus 2^62+1 rejects ~25% of the words; accepted samples are the accepted words in order"""
    m = (1 << 62) + 1
    zone = (1 << 64) - 1 - ((1 << 64) - 1) % m
    r1, r2 = oracle.rng_from_seed([9]), oracle.rng_from_seed([9])
    words = [oracle.next_u64(r1) for _ in range(400)]
    expect = [w % m for w in words if w < zone][:200]
    assert len(expect) == 200 and len([w for w in words if w >= zone]) > 20
    assert [oracle.gen_range(r2, 0, m) for _ in range(200)] == expect


def test_scheme_sizes(oracle):
    """protocol/src/crypto.rs:117-155"""
    a = oracle.additive(3, 433)
    assert (oracle.input_size(a), oracle.output_size(a), oracle.privacy_threshold(a),
            oracle.reconstruction_threshold(a)) == (1, 3, 2, 3)
    p = oracle.packed_shamir(3, 8, 4, 433, 354, 150)
    assert (oracle.input_size(p), oracle.output_size(p), oracle.privacy_threshold(p),
            oracle.reconstruction_threshold(p)) == (3, 8, 4, 7)


def test_combine_semantics(oracle):
    """combiner.rs:19-26 keeps signed representatives: (7, 7, -9) mod 10 -> -5, reordered -> 5"""
    assert oracle.share_combine(10, np.array([[7], [7], [-9]])).tolist() == [-5]
    assert oracle.share_combine(10, np.array([[7], [-9], [7]])).tolist() == [5]
    assert oracle.share_combine(433, np.zeros((0, 0), dtype=np.int64)).tolist() == []


def test_varint_codec(oracle):
    """integer-encoding 1.0 zig-zag LEB128 (sodium.rs:36-41, 83-90)"""
    assert oracle.varint_encode([0]).tolist() == [0]
    assert oracle.varint_encode([-1]).tolist() == [1]
    assert oracle.varint_encode([63]).tolist() == [126]
    assert oracle.varint_encode([432]).tolist() == [0xE0, 0x06]
    assert len(oracle.varint_encode([params.P61 - 1])) == 9
    v = np.array([0, -1, 1, 432, -433, params.P61 - 1, -(1 << 63), (1 << 63) - 1], dtype=np.int64)
    assert oracle.varint_decode(oracle.varint_encode(v)).tolist() == v.tolist()
    with pytest.raises(oracle.OracleError):
        oracle.varint_decode([0x80])


@pytest.mark.parametrize("case", ORC["sharing"], ids=lambda c: c["name"])
def test_frozen_sharing_vectors(oracle, case):
    """oracle.json is what the current oracle produces (guards the fixtures against oracle drift)"""
    kind, n, k, t, m, ws, wh = case["scheme"]
    s = LSS.Additive(n, m) if kind == 0 else LSS.PackedShamir(k, n, t, m, ws, wh)
    shares = util.oracle_generate(oracle, s, np.array(case["secrets"], dtype=np.int64), bytes.fromhex(case["seed"]),
                                  case["rounds"])
    assert util.canon(oracle, m, shares).tolist() == case["shares_canonical"]
    # the precomputed-matrix evaluation is the same map
    if kind == 1:
        fast = util.oracle_generate(oracle, s, np.array(case["secrets"], dtype=np.int64), bytes.fromhex(case["seed"]),
                                    case["rounds"], matrix=True)
        assert fast.tolist() == case["shares_canonical"]


@pytest.mark.parametrize("shape", [(3, 2, 5), (5, 4, 9), (3, 4, 7), (2, 3, 6), (1, 1, 3)])
@pytest.mark.parametrize("p", [params.P61, params.P61_GENERIC, 433, 746497])
def test_roundtrip_and_linearity(oracle, shape, p):
    k, t, n = shape
    if p == 433 and (k + t + 1 > 8 or n + 1 > 8):
        pytest.skip("no such orders mod 433")
    try:
        s = util.packed_scheme(p, k, t, n, oracle)
    except StopIteration:
        pytest.skip("no suitable prime orders")
    so = util.to_oracle_sharing(oracle, s)
    rng = np.random.default_rng(k * 100 + n)
    dim = 7 * k + 1
    a, b = util.rand_secrets(rng, dim, p), util.rand_secrets(rng, dim, p)
    sa = util.canon(oracle, p, util.oracle_generate(oracle, s, a, util.seed_bytes("a")))
    sb = util.canon(oracle, p, util.oracle_generate(oracle, s, b, util.seed_bytes("b")))
    idx = list(range(n))
    assert util.canon(oracle, p, oracle.secret_reconstruct(so, dim, idx, sa)).tolist() == a.tolist()
    # linearity: share(a) + share(b) reconstructs to a + b, from any k + t clerks
    summed = np.stack([oracle.share_combine(p, np.stack([sa[c], sb[c]])) for c in range(n)])
    sub = sorted(rng.permutation(n)[:k + t].tolist())
    rec = oracle.secret_reconstruct(so, dim, sub, summed[sub])
    expect = [(int(x) + int(y)) % p for x, y in zip(a, b)]
    assert util.canon(oracle, p, rec).tolist() == expect
