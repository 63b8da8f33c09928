"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact on canonical residues.
Every test here needs a B200 (`-m gpu`)."""
import json
import os

import numpy as np
import pytest

import util
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import LinearSecretSharingScheme as LSS
from sda_b200 import SdaClientError, params

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference.json")))
ORC = json.load(open(os.path.join(HERE, "golden", "oracle.json")))

P61, PGEN = params.P61, params.P61_GENERIC
M_REJECT = (1 << 62) + 1          # gen_range rejects ~25% of the words: exercises the exact path
M_MAX = (1 << 63) - 1


def scheme_from_fixture(v):
    kind, n, k, t, m, ws, wh = v
    return LSS.Additive(n, m) if kind == 0 else LSS.PackedShamir(k, n, t, m, ws, wh)


# ---- committed fixtures ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", ORC["sharing"], ids=lambda c: c["name"])
def test_share_generate_fixtures(ctx, case):
    s = scheme_from_fixture(case["scheme"])
    ctx.set_rng_rounds(case["rounds"])
    try:
        got = ctx.share_generate(s, np.array(case["secrets"], dtype=np.int64), bytes.fromhex(case["seed"]))
    finally:
        ctx.set_rng_rounds(20)
    assert got.tolist() == case["shares_canonical"]


@pytest.mark.parametrize("case", ORC["masking"], ids=lambda c: c["name"])
def test_mask_fixtures(ctx, case):
    kind, m, dim, bits = case["scheme"]
    ms = LMS.Full(m) if kind == 1 else LMS.ChaCha(m, dim, bits)
    mask, masked = ctx.mask(ms, np.array(case["secrets"], dtype=np.int64), bytes.fromhex(case["seed"]))
    assert mask.tolist() == case["mask"]
    assert masked.tolist() == case["masked_canonical"]


# ---- the reference's own result-pinning tests, through the product ---------------------------------
def full_loop(crypto, sharing, masking, inputs, missing=()):
    """participate.rs:53-76 -> clerk.rs:85-86 -> receive.rs:113-152 on CryptoModule"""
    dim = len(inputs[0])
    n = sharing.output_size()
    masks, per_clerk = [], [[] for _ in range(n)]
    for secrets in inputs:
        mask, masked = crypto.new_secret_masker(masking).mask(secrets)
        masks.append(mask)
        shares = crypto.new_share_generator(sharing).generate(masked)
        for c in range(n):
            per_clerk[c].append(shares[c])
    combiner = crypto.new_share_combiner(sharing)
    indexed = [(c, combiner.combine(per_clerk[c])) for c in range(n) if c not in missing]
    masked_out = crypto.new_secret_reconstructor(sharing, dim).reconstruct(indexed)
    if masking.has_mask():
        mask_sum = crypto.new_mask_combiner(masking).combine(masks)
        return crypto.new_secret_unmasker(masking).unmask((mask_sum, masked_out))
    return masked_out


@pytest.mark.parametrize("masking", ["none", "full", "chacha"])
@pytest.mark.parametrize("sharing", ["additive", "packed_shamir"])
def test_reference_full_loop(crypto, sharing, masking):
    """integration-tests/tests/full_loop.rs:29-67 -> [2,4,6,8]"""
    g = REF["full_loop"]
    s = LSS.Additive(**g["sharing"]["additive"]) if sharing == "additive" else LSS.PackedShamir(**g["sharing"]["packed_shamir"])
    m = {"none": LMS.None_(), "full": LMS.Full(433), "chacha": LMS.ChaCha(**g["masking"]["chacha"])}[masking]
    out = full_loop(crypto, s, m, [g["input"]] * g["participants"])
    assert out.tolist() == g["expected_positive"]


def test_reference_cli_walkthrough(crypto):
    """README.md:157 (BASELINE config #1)"""
    g = REF["cli_walkthrough"]
    out = full_loop(crypto, LSS.Additive(**g["sharing"]), LMS.None_(), g["inputs"])
    assert out.tolist() == g["expected"]


def test_full_loop_missing_clerk(crypto):
    g = REF["full_loop"]
    s = LSS.PackedShamir(**g["sharing"]["packed_shamir"])
    for missing in [(0,), (5,)]:
        assert full_loop(crypto, s, LMS.Full(433), [g["input"]] * 2, missing).tolist() == g["expected_positive"]


# ---- share generation vs the live oracle ----------------------------------------------------------
DIMS = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 127, 128, 129, 1000, 4099]


@pytest.mark.parametrize("modulus", [433, P61, PGEN, 2, 1, M_MAX])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 9, 16, 33])
def test_additive_generate(ctx, oracle, n, modulus):
    s = LSS.Additive(n, modulus)
    rng = np.random.default_rng(n * 1000 + modulus % 997)
    for dim in DIMS:
        for kind in ("canonical", "signed"):
            if modulus > (1 << 62) and (kind == "signed" or n > 2):
                # additive.rs:47 computes `acc - share` in i64 with acc in (-m, m): beyond one
                # subtraction from a canonical secret that wraps (panics in a debug build) once
                # m > 2^62, so the reference -- and its literal oracle -- is undefined there
                continue
            secrets = util.rand_secrets(rng, dim, modulus, kind)
            seed = util.seed_bytes(f"add/{n}/{dim}/{kind}")
            exp = util.canon(oracle, modulus, util.oracle_generate(oracle, s, secrets, seed))
            got = ctx.share_generate(s, secrets, seed)
            assert got.shape == (n, dim)
            assert np.array_equal(got, exp), (n, modulus, dim, kind)
    if n > 1 and modulus in (433, P61, PGEN):
        assert "in-kernel rng" in ctx.last_kernel()      # no share count falls back to draws materialised in HBM


@pytest.mark.parametrize("modulus", [M_REJECT, 3 << 61])
@pytest.mark.parametrize("n", [2, 3, 7])
def test_additive_generate_exact_rejection_path(ctx, oracle, n, modulus):
    """gen_range rejects often for these moduli: the rejected words shift the whole stream"""
    if n > 2 and modulus > (1 << 62):
        pytest.skip("additive.rs:47 wraps i64 for a second subtraction when m > 2^62: reference undefined")
    s = LSS.Additive(n, modulus)
    rng = np.random.default_rng(n)
    for dim in [1, 5, 64, 1000, 5000]:
        secrets = util.rand_secrets(rng, dim, modulus)
        seed = util.seed_bytes(f"rej/{n}/{dim}")
        r = oracle.rng_from_seed_bytes(seed)
        exp = util.canon(oracle, modulus, oracle.share_generate(util.to_oracle_sharing(oracle, s), secrets, r))
        if dim >= 64:
            assert r.rejections > 0
        got = ctx.share_generate(s, secrets, seed)
        assert np.array_equal(got, exp), (n, modulus, dim)


PACKED = [("cfg3", params.config3), ("cfg4", params.config4), ("cfg5", params.config5), ("ref433", params.reference_test)]


@pytest.mark.parametrize("rounds", [20, 12, 8])
@pytest.mark.parametrize("name,mk", PACKED, ids=[p[0] for p in PACKED])
def test_packed_generate_fast_shapes(ctx, oracle, name, mk, rounds):
    s = mk()
    p = s.modulus
    rng = np.random.default_rng(rounds)
    ctx.set_rng_rounds(rounds)
    try:
        for dim in DIMS:
            for kind in ("canonical", "signed"):
                secrets = util.rand_secrets(rng, dim, p, kind) if p > (1 << 40) or kind == "canonical" else \
                    rng.integers(-50 * p, 50 * p, size=dim, dtype=np.int64)
                seed = util.seed_bytes(f"{name}/{dim}/{kind}/{rounds}")
                exp = util.oracle_generate(oracle, s, secrets, seed, rounds, matrix=True)
                got = ctx.share_generate(s, secrets, seed)
                assert got.shape == exp.shape
                assert np.array_equal(got, exp), (name, dim, kind)
        assert "packed_share<" in ctx.last_kernel() and "memory" not in ctx.last_kernel()
    finally:
        ctx.set_rng_rounds(20)


def test_packed_generate_literal_oracle(ctx, oracle):
    """against the LITERAL restatement (per-batch FFT / Newton as tss does), not the matrix form"""
    rng = np.random.default_rng(5)
    for s in (params.reference_test(), params.config3(), params.config4()):
        secrets = util.rand_secrets(rng, 301, s.modulus)
        seed = util.seed_bytes("literal")
        exp = util.canon(oracle, s.modulus, util.oracle_generate(oracle, s, secrets, seed))
        assert np.array_equal(ctx.share_generate(s, secrets, seed), exp)


@pytest.mark.parametrize("shape", [(1, 1, 3), (2, 3, 6), (4, 3, 10), (7, 5, 16), (2, 1, 4), (8, 8, 20), (15, 1, 17), (1, 15, 32),
                                   (3, 3, 7), (6, 7, 14), (5, 2, 9),
                                   (4, 2, 6), (2, 4, 8), (8, 4, 5), (1, 2, 3), (6, 2, 7), (7, 4, 8), (8, 2, 1), (3, 2, 4),
                                   (8, 8, 32), (3, 5, 24), (2, 6, 17), (8, 1, 9), (1, 7, 8), (5, 3, 12), (9, 2, 12), (2, 9, 12)])
@pytest.mark.parametrize("p", [P61, PGEN, 2305843009213693561])
def test_packed_generate_generic_shapes(ctx, oracle, shape, p):
    """any (k, t, n) the reference accepts (packed_shamir.rs:13-27) runs on a tcgen05 kernel -- over 2^61-1 on the
    paired-tile kernel with the share count at run time (k + t up to 16, n up to 32), over other primes on the
    run-time-shaped one -- and no scheme materialises its draws in HBM"""
    k, t, n = shape
    try:
        s = util.packed_scheme(p, k, t, n, oracle)
    except StopIteration:
        pytest.skip("no suitable prime orders in p-1")
    rng = np.random.default_rng(k + 10 * n)
    for dim in [1, k, k + 1, 10 * k - 1, 1000, 256 * k + 1, 3000 * k + 2]:
        for kind in ("canonical", "signed"):
            secrets = util.rand_secrets(rng, dim, p, kind)
            seed = util.seed_bytes(f"gen/{shape}/{dim}/{kind}")
            exp = util.oracle_generate(oracle, s, secrets, seed, matrix=True)
            assert np.array_equal(ctx.share_generate(s, secrets, seed), exp), (shape, p, dim, kind)
    name = ctx.last_kernel()
    assert "tcgen05" in name and ("run-time shape" in name or "at run time" in name), name
    k_, t_, n_ = shape
    assert ("at run time" in name) == (p == P61), name


def test_packed_share_matrix_matches_oracle(ctx, oracle):
    for s in (params.reference_test(), params.config3(), params.config4(), params.config5()):
        c = s.c
        M = ctx.packed_share_matrix(s)
        so = util.to_oracle_sharing(oracle, s)
        w = c.secret_count + c.privacy_threshold
        for i in range(w):
            unit = [0] * w
            unit[i] = 1
            col = oracle.tss_share_with_randomness(so, unit[:c.secret_count], unit[c.secret_count:])
            assert util.canon(oracle, c.modulus, col).tolist() == M[:, i].tolist()


# ---- clerk combine -----------------------------------------------------------------------------------
@pytest.mark.parametrize("modulus", [433, P61, PGEN, M_MAX, 10, 1])
def test_share_combine(ctx, oracle, modulus):
    s = LSS.Additive(3, modulus)
    rng = np.random.default_rng(modulus % 1000)
    for P, L in [(1, 1), (2, 4), (3, 7), (8, 16), (9, 1023), (17, 1024), (64, 1026), (300, 33), (5000, 10),
                 (2, 100001), (100, 4096)]:
        for kind in ("canonical", "signed"):
            if kind == "signed" and modulus > (1 << 62):
                continue                   # the reference's own i64 adds would wrap
            rows = np.stack([util.rand_secrets(rng, L, modulus, kind) for _ in range(P)])
            if modulus > (1 << 62) and P > 1:
                # combiner.rs:24 adds two residues in i64: wraps (panics in debug) once m > 2^62, so the
                # reference is undefined; the kernel is held to the exact integer result instead
                exp = np.array([int(v) % modulus for v in rows.astype(object).sum(axis=0)], dtype=np.int64)
            else:
                exp = util.canon(oracle, modulus, oracle.share_combine(modulus, rows))
            got = ctx.share_combine(s, rows)
            assert np.array_equal(got, exp), (modulus, P, L, kind)
            got_rows = ctx.share_combine(s, [r.copy() for r in rows])
            assert np.array_equal(got_rows, exp)


def test_share_combine_edge_cases(ctx):
    s = LSS.Additive(3, 433)
    assert ctx.share_combine(s, []).tolist() == []                       # combiner.rs:17
    assert ctx.share_combine(s, [[], []]).tolist() == []
    assert ctx.share_combine(s, [[5, -1, 433, 866, -434]]).tolist() == [5, 432, 0, 0, 432]
    with pytest.raises(SdaClientError, match="Wrong dimension"):          # combiner.rs:21
        ctx.share_combine(s, [[1, 2, 3], [1, 2]])
    # extreme magnitudes: no overflow inside the kernel whatever the reference would do
    big = np.full((64, 5), (1 << 63) - 1, dtype=np.int64)
    assert ctx.share_combine(LSS.Additive(3, P61), big).tolist() == [(64 * ((1 << 63) - 1)) % P61] * 5
    small = np.full((64, 5), -(1 << 63), dtype=np.int64)
    assert ctx.share_combine(LSS.Additive(3, PGEN), small).tolist() == [(-64 * (1 << 63)) % PGEN] * 5


# ---- reconstruction ------------------------------------------------------------------------------------
def test_additive_reconstruct(ctx, oracle):
    rng = np.random.default_rng(3)
    for modulus in (433, P61, PGEN):
        s = LSS.Additive(3, modulus)
        for dim in (1, 10, 1001):
            rows = [util.rand_secrets(rng, dim, modulus, "signed") for _ in range(3)]
            exp = util.canon(oracle, modulus, oracle.secret_reconstruct(util.to_oracle_sharing(oracle, s), dim,
                                                                        [0, 1, 2], np.stack(rows)))
            got = ctx.secret_reconstruct(s, dim, list(enumerate(rows)))
            assert np.array_equal(got, exp)
    with pytest.raises(SdaClientError, match="Mismatching dimension"):    # additive.rs:64
        ctx.secret_reconstruct(LSS.Additive(2, 433), 3, [(0, [1, 2, 3]), (1, [1, 2])])


@pytest.mark.parametrize("name,mk", PACKED, ids=[p[0] for p in PACKED])
def test_packed_reconstruct(ctx, oracle, name, mk):
    s = mk()
    c = s.c
    p, k, n, need = c.modulus, c.secret_count, c.share_count, c.secret_count + c.privacy_threshold
    so = util.to_oracle_sharing(oracle, s)
    rng = np.random.default_rng(11)
    for dim in (1, k, k + 1, 100, 3001):
        secrets = util.rand_secrets(rng, dim, p)
        shares = util.oracle_generate(oracle, s, secrets, util.seed_bytes(dim), matrix=True)
        subsets = [list(range(n))]
        if need < n:
            subsets += [sorted(rng.permutation(n)[:need].tolist()), sorted(rng.permutation(n)[:need + (need + 1 < n)].tolist())]
        subsets.append(list(reversed(range(n))))                          # order of indexed_shares is free
        for idx in subsets:
            exp = util.canon(oracle, p, oracle.secret_reconstruct(so, dim, idx, shares[idx]))
            assert exp.tolist() == secrets.tolist()
            got = ctx.secret_reconstruct(s, dim, [(i, shares[i]) for i in idx])
            assert np.array_equal(got, secrets), (name, dim, idx)
            R = ctx.packed_reconstruct_matrix(s, idx)
            assert R.shape == (k, len(idx))
    with pytest.raises(SdaClientError, match="Not enough shares to reconstruct"):   # packed_shamir.rs:75
        ctx.secret_reconstruct(s, k, [(i, [0]) for i in range(need - 1)])
    with pytest.raises(SdaClientError, match="index out of bounds"):                # batched.rs:84 panics
        ctx.secret_reconstruct(s, 5 * k, [(i, [0]) for i in range(n)])


@pytest.mark.parametrize("shape", [(1, 1, 3), (2, 3, 6), (4, 3, 10), (7, 5, 16), (8, 8, 20), (15, 1, 17), (1, 15, 32), (3, 3, 7),
                                   (6, 7, 14), (5, 2, 9), (1, 2, 3), (8, 1, 9), (2, 1, 4), (10, 1, 12)])
def test_packed_reconstruct_generic_shapes(ctx, oracle, shape):
    """reveal for any (k, t, n) over 2^61-1: every clerk-subset size from k + t to n (up to 16 present clerks on the tcgen05
    kernel -- every chunk count and both parities of m' -- more on the CUDA-core one), vectors that end inside a tile and
    inside a batch, against the oracle's Newton interpolation"""
    k, t, n = shape
    try:
        s = util.packed_scheme(P61, k, t, n, oracle)
    except StopIteration:
        pytest.skip("no suitable prime orders in p-1")
    rng = np.random.default_rng(100 * k + n)
    need = k + t
    for dim in (1, k + 1, 128 * k * 3 - 1, 128 * k * 19 + 2):
        secrets = util.rand_secrets(rng, dim, P61)
        shares = util.oracle_generate(oracle, s, secrets, util.seed_bytes(f"rec/{shape}/{dim}"), matrix=True)
        for m in sorted({need, min(need + 1, n), min(16, n), n}):
            idx = sorted(rng.permutation(n)[:m].tolist())
            got = ctx.secret_reconstruct(s, dim, [(i, shares[i]) for i in idx])
            assert np.array_equal(got, secrets), (shape, dim, idx)


# ---- masking ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("modulus", [433, P61, PGEN, M_REJECT, 1])
def test_full_mask_roundtrip_and_parity(ctx, oracle, modulus):
    ms = LMS.Full(modulus)
    mo = util.to_oracle_masking(oracle, ms)
    rng = np.random.default_rng(17)
    for dim in [1, 7, 8, 9, 1000, 4099]:
        secrets = util.rand_secrets(rng, dim, modulus, "signed" if modulus > 433 else "canonical")
        seed = util.seed_bytes(f"full/{dim}")
        emask, emasked = oracle.mask(mo, secrets, oracle.rng_from_seed_bytes(seed))
        mask, masked = ctx.mask(ms, secrets, seed)
        assert np.array_equal(mask, emask) and np.array_equal(masked, util.canon(oracle, modulus, emasked))
        back = ctx.unmask(ms, mask, masked)
        assert np.array_equal(back, util.canon(oracle, modulus, secrets))
    masks = [util.rand_secrets(rng, 33, modulus) for _ in range(5)]
    exp = util.canon(oracle, modulus, oracle.mask_combine(mo, np.stack(masks)))
    assert np.array_equal(ctx.mask_combine(ms, masks), exp)


@pytest.mark.parametrize("bits", [32, 40, 128, 256])
@pytest.mark.parametrize("modulus", [433, P61, PGEN, M_REJECT])
def test_chacha_mask(ctx, oracle, modulus, bits):
    rng = np.random.default_rng(bits)
    for dim in [1, 8, 9, 1000]:
        ms = LMS.ChaCha(modulus, dim, bits)
        mo = util.to_oracle_masking(oracle, ms)
        seeds = []
        for pi in range(3):
            secrets = util.rand_secrets(rng, dim, modulus)
            seed = util.seed_bytes(f"cc/{dim}/{pi}")
            emask, emasked = oracle.mask(mo, secrets, oracle.rng_from_seed_bytes(seed))
            mask, masked = ctx.mask(ms, secrets, seed)
            assert mask.tolist() == emask.tolist() and len(mask) == (bits + 31) // 32
            assert np.array_equal(masked, util.canon(oracle, modulus, emasked))
            seeds.append(mask)
        ecomb = util.canon(oracle, modulus, oracle.mask_combine(mo, np.stack(seeds)))
        comb = ctx.mask_combine(ms, seeds)
        assert np.array_equal(comb, ecomb)
    with pytest.raises(SdaClientError, match="chacha.rs:26"):
        ctx.mask(LMS.ChaCha(433, 5, 128), [1, 2, 3])


def test_chacha_mask_combine_many_seeds(ctx, oracle):
    """seed axis sliced across CTAs + second-stage combine"""
    ms = LMS.ChaCha(P61, 777, 128)
    rng = np.random.default_rng(1)
    seeds = rng.integers(0, 1 << 32, size=(600, 4), dtype=np.int64)
    exp = util.canon(oracle, P61, oracle.mask_combine(util.to_oracle_masking(oracle, ms), seeds))
    assert np.array_equal(ctx.mask_combine(ms, list(seeds)), exp)


def test_none_mask(ctx):
    ms = LMS.None_()
    mask, masked = ctx.mask(ms, [3, 1, 4])
    assert mask.tolist() == [] and masked.tolist() == [3, 1, 4]
    assert ctx.mask_combine(ms, [[], []]).tolist() == []
    assert ctx.unmask(ms, [], [3, 1, 4]).tolist() == [3, 1, 4]
    with pytest.raises(SdaClientError):
        ctx.unmask(ms, [1], [3, 1, 4])                                    # none.rs:30
    with pytest.raises(SdaClientError):
        ctx.unmask(LMS.Full(433), [1, 2], [3, 1, 4])                      # full.rs:58


# ---- scheme validation ---------------------------------------------------------------------------------
def test_scheme_validation(crypto):
    with pytest.raises(SdaClientError):
        crypto.new_share_generator(LSS.Additive(0, 433))                  # additive.rs:42 underflow
    with pytest.raises(SdaClientError, match="gen_range"):
        crypto.new_share_generator(LSS.Additive(3, 0))
    with pytest.raises(SdaClientError, match="collide"):
        crypto.new_share_generator(LSS.PackedShamir(3, 8, 4, 433, 1, 150))
    with pytest.raises(SdaClientError) as e:
        crypto.new_share_generator(LSS.PackedShamir(30, 60, 30, P61, 3, 5))
    assert e.value.code == 4
    # FFT sizes (k + t + 1 = 2^a, n + 1 = 3^b) run tss 0.2's FFTs whatever the roots are: only primitive roots of exactly
    # those orders make that the interpolation; 17 has order 27 mod 433, 151 has order 16 (354 = order 8, 150 = order 9)
    with pytest.raises(SdaClientError, match="primitive 9-th root"):
        crypto.new_share_generator(LSS.PackedShamir(3, 8, 4, 433, 354, 17))
    with pytest.raises(SdaClientError, match="primitive 8-th root"):
        crypto.new_share_generator(LSS.PackedShamir(3, 8, 4, 433, 151, 150))
    crypto.new_share_generator(LSS.PackedShamir(3, 8, 4, 433, 354, 150))


@pytest.mark.parametrize("name,mk", PACKED + [("additive3", lambda: LSS.Additive(3, P61))], ids=[p[0] for p in PACKED] + ["additive3"])
@pytest.mark.parametrize("mask_kind", ["full", "chacha", "none"])
def test_mask_share_generate_host(ctx, oracle, name, mk, mask_kind):
    """sda_mask_share_generate on host vectors == mask then share_generate through the separate trait calls (and the
    oracle): every masking scheme, packed (fused kernel for the 2^61-1 shapes, two steps for p = 433) and additive sharing"""
    ss = mk()
    p = ss.modulus
    rng = np.random.default_rng(len(name))
    for dim in (1, 7, 1000, 4099):
        ms = {"full": LMS.Full(p), "chacha": LMS.ChaCha(p, dim, 64), "none": LMS.None_()}[mask_kind]
        secrets = util.rand_secrets(rng, dim, p, "signed" if p > 433 else "canonical")
        mseed, sseed = util.seed_bytes(f"hm/{name}/{dim}"), util.seed_bytes(f"hs/{name}/{dim}")
        mask, shares = ctx.mask_share_generate(ms, ss, secrets, mseed, sseed)
        if mask_kind == "none":
            emask, emasked = np.zeros(0, dtype=np.int64), secrets
        else:
            emask, emasked = ctx.mask(ms, secrets, mseed)
            omask, omasked = oracle.mask(util.to_oracle_masking(oracle, ms), secrets, oracle.rng_from_seed_bytes(mseed))
            assert np.array_equal(emask, np.asarray(omask, dtype=np.int64))
        assert np.array_equal(mask, emask), (name, mask_kind, dim)
        assert np.array_equal(shares, ctx.share_generate(ss, emasked, sseed)), (name, mask_kind, dim)


def test_mask_share_generate_errors(ctx):
    """the fused entry point keeps the reference's failure modes: chacha.rs:26 (scheme dimension != number of secrets),
    an invalid sharing scheme, an invalid masking modulus"""
    ss = params.config3()
    secrets = np.arange(10, dtype=np.int64)
    with pytest.raises(SdaClientError, match="chacha.rs:26"):
        ctx.mask_share_generate(LMS.ChaCha(P61, 11, 128), ss, secrets, b"\1" * 32, b"\2" * 32)
    with pytest.raises(SdaClientError, match="low >= high"):
        ctx.mask_share_generate(LMS.Full(0), ss, secrets, b"\1" * 32, b"\2" * 32)
    bad = LSS.PackedShamir(3, 5, 2, P61, 1, ss.c.omega_shares)                       # secret points collide
    with pytest.raises(SdaClientError, match="omega_secrets has order"):
        ctx.mask_share_generate(LMS.Full(P61), bad, secrets, b"\1" * 32, b"\2" * 32)
    # an empty vector is not an error: no mask, no shares
    mask, shares = ctx.mask_share_generate(LMS.Full(P61), ss, np.zeros(0, dtype=np.int64), b"\1" * 32, b"\2" * 32)
    assert len(mask) == 0 and shares.shape == (5, 0)
