"""BASELINE config #5 in miniature, end to end on one GPU: float updates -> fixed point -> ChaCha mask ->
packed Shamir k=3/t=4 with 8 clerks (one more than config #5's 7, so that a clerk may go missing) -> per-clerk
sums (materialised AND fused) -> reveal from 7 of the 8 -> unmask -> mean as floats.  Every stage against the oracle; the two clerk-sum paths against each other."""
import numpy as np
import pytest

import sda_b200
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import params

import util

pytestmark = pytest.mark.gpu
P61 = params.P61
FRAC = 24


@pytest.fixture(scope="module")
def ctx():
    c = sda_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def test_fixed_point_codec_matches_oracle(ctx, oracle):
    import torch as t
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.standard_normal(10_001).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, -1.0, 2.0 ** -25, 3 * 2.0 ** -25, -(2.0 ** -25), 1e4, -1e4, 2.0 ** -24],
                                 dtype=np.float32)])
    for m in (P61, params.P61_GENERIC, (1 << 40) + 15):
        d_q = t.empty(len(x), dtype=t.int64, device="cuda")
        ctx.fixed_encode_dev(m, FRAC, t.from_numpy(x).cuda(), len(x), d_q)
        ctx.synchronize()
        q = d_q.cpu().numpy()
        assert np.array_equal(q, oracle.fixed_encode(x, FRAC, m))
        for div in (1, 7, 8192):
            d_y = t.empty(len(x), dtype=t.float32, device="cuda")
            ctx.fixed_decode_dev(m, FRAC, div, d_q, len(x), d_y)
            ctx.synchronize()
            assert np.array_equal(d_y.cpu().numpy(), oracle.fixed_decode(q, FRAC, m, div))
    back = oracle.fixed_decode(oracle.fixed_encode(x, FRAC, P61), FRAC, P61, 1)
    assert np.max(np.abs(back - x)) <= 2.0 ** -25 * 1.0001 + 1e4 * 2.0 ** -24   # one rounding step (+ float ulp at 1e4)


def test_federated_round_trip(ctx, oracle):
    import torch as t
    s = sda_b200.LinearSecretSharingScheme.PackedShamir(3, 9, 4, P61, params.ROOT_ORDER_11, params.ROOT_ORDER_13)   # n + 1 = 10: not an FFT size
    n, k = s.output_size(), s.input_size()
    P, dim = 37, 5000
    B = s.batches(dim)
    rng = np.random.default_rng(12)
    updates = rng.standard_normal((P, dim)).astype(np.float32)
    ms = LMS.ChaCha(P61, dim, 128)
    mo = util.to_oracle_masking(oracle, ms)

    # participants: encode, mask (mask seeds go to the recipient), share
    d_x = t.from_numpy(updates).cuda()
    d_q = t.empty((P, dim), dtype=t.int64, device="cuda")
    ctx.fixed_encode_dev(P61, FRAC, d_x, P * dim, d_q)
    ctx.synchronize()
    q = d_q.cpu().numpy()
    assert np.array_equal(q, oracle.fixed_encode(updates.ravel(), FRAC, P61).reshape(P, dim))
    masked = np.empty_like(q)
    seeds_for_recipient = []
    for pi in range(P):
        mask, masked[pi] = ctx.mask(ms, q[pi], util.seed_bytes(f"fed/mask/{pi}"))
        emask, emasked = oracle.mask(mo, q[pi], oracle.rng_from_seed_bytes(util.seed_bytes(f"fed/mask/{pi}")))
        assert mask.tolist() == emask.tolist() and np.array_equal(masked[pi], util.canon(oracle, P61, emasked))
        seeds_for_recipient.append(mask)
    share_seeds = b"".join(util.seed_bytes(f"fed/share/{pi}") for pi in range(P))
    d_masked = t.from_numpy(masked).cuda()
    d_shares = t.empty((P, n, B), dtype=t.int64, device="cuda")
    ctx.share_generate_dev(s, d_masked, dim, P, dim, share_seeds, d_shares)

    # clerks: per-clerk sums on the materialised shares, and the fused kernel that never materialises them
    d_sums = t.empty((n, B), dtype=t.int64, device="cuda")
    for c in range(n):
        ctx.share_combine_dev(s, d_shares[:, c, :], n * B, P, B, d_sums[c])
    d_fused = t.empty((n, B), dtype=t.int64, device="cuda")
    ctx.share_generate_combine_dev(s, d_masked, dim, P, dim, share_seeds, d_fused)
    ctx.synchronize()
    assert t.equal(d_sums, d_fused)

    # recipient: clerks 2 and 8 never answered; reveal from the other seven, combine the mask seeds, unmask, decode the mean
    idx = [0, 1, 3, 4, 5, 6, 7]
    got = ctx.secret_reconstruct(s, dim, [(i, d_sums[i].cpu().numpy()) for i in idx])
    assert np.array_equal(got, masked.astype(object).sum(axis=0) % P61)
    total_mask = ctx.mask_combine(ms, seeds_for_recipient)
    summed = ctx.unmask(ms, total_mask, got)
    assert np.array_equal(summed, q.astype(object).sum(axis=0) % P61)
    d_mean = t.empty(dim, dtype=t.float32, device="cuda")
    ctx.fixed_decode_dev(P61, FRAC, P, t.from_numpy(summed).cuda(), dim, d_mean)
    ctx.synchronize()
    mean = d_mean.cpu().numpy()
    assert np.array_equal(mean, oracle.fixed_decode(summed, FRAC, P61, P))
    assert np.max(np.abs(mean - updates.astype(np.float64).mean(axis=0))) < 2.0 ** -24


@pytest.mark.parametrize("modulus", [P61, params.P61_GENERIC, 433])
def test_fused_encode_mask_equals_encode_then_mask(ctx, oracle, modulus):
    """sda_fixed_encode_mask_dev == sda_fixed_encode_dev + sda_mask_dev, and both == the oracle, for every mask scheme"""
    import torch as t
    rng = np.random.default_rng(31)
    for dim in (1, 7, 8, 9, 1000, 40003):
        x = (rng.standard_normal(dim) * 3).astype(np.float32)
        q = oracle.fixed_encode(x, FRAC, modulus)
        d_x = t.from_numpy(x).cuda()
        for ms in (LMS.None_(), LMS.Full(modulus), LMS.ChaCha(modulus, dim, 128), LMS.ChaCha(modulus, dim, 40)):
            seed = util.seed_bytes(f"fusedmask/{dim}/{modulus}")
            nmask = ms.mask_len(dim)
            d_mask = t.zeros(max(nmask, 1), dtype=t.int64, device="cuda")
            d_masked = t.empty(dim, dtype=t.int64, device="cuda")
            ctx.fixed_encode_mask_dev(ms, modulus, FRAC, d_x[1:] if False else d_x, dim, seed, d_mask, d_masked)
            ctx.synchronize()
            emask, emasked = oracle.mask(util.to_oracle_masking(oracle, ms), q, oracle.rng_from_seed_bytes(seed))
            assert np.array_equal(d_masked.cpu().numpy(), util.canon(oracle, modulus, emasked)), (dim, modulus, ms)
            assert d_mask.cpu().numpy()[:nmask].tolist() == util.canon(oracle, modulus, emask).tolist() if nmask and ms.c.kind == 1 \
                else d_mask.cpu().numpy()[:nmask].tolist() == list(emask)
    # extreme inputs (around |q| = 2^60, 2^61, 2^63, infinities, NaN, denormals, -0.0): the fused pass encodes exactly as
    # sda_fixed_encode_dev does (the fast float path and its general branch)
    if modulus == P61:
        xe = np.array([2.0 ** 43, -2.0 ** 43, 2.0 ** 44, -2.0 ** 44, 2.0 ** 44 - 2.0 ** 21, 2.0 ** 45, -2.0 ** 45, 2.0 ** 46, 2.0 ** 47,
                       -2.0 ** 47, 1e30, -1e30, np.inf, -np.inf, np.nan, 1e-45, -1e-45, -0.0, 0.0, 3.4e38, -3.4e38, 0.5 * 2.0 ** -16,
                       1.5 * 2.0 ** -16, 2.5 * 2.0 ** -16, -0.5 * 2.0 ** -16], dtype=np.float32)
        d_xe = t.from_numpy(xe).cuda()
        d_q = t.empty(len(xe), dtype=t.int64, device="cuda")
        ctx.fixed_encode_dev(P61, FRAC, d_xe, len(xe), d_q)
        for ms in (LMS.None_(), LMS.Full(P61)):
            seed = util.seed_bytes("fusedmask/extreme")
            d_m1, d_o1 = t.zeros(len(xe), dtype=t.int64, device="cuda"), t.empty(len(xe), dtype=t.int64, device="cuda")
            d_m2, d_o2 = t.zeros(len(xe), dtype=t.int64, device="cuda"), t.empty(len(xe), dtype=t.int64, device="cuda")
            ctx.fixed_encode_mask_dev(ms, P61, FRAC, d_xe, len(xe), seed, d_m1, d_o1)
            ctx.mask_dev(ms, d_q, len(xe), seed, d_m2, d_o2)
            ctx.synchronize()
            assert t.equal(d_o1, d_o2) and t.equal(d_m1, d_m2), (ms.c.kind, d_o1.cpu().numpy(), d_o2.cpu().numpy())
    # an unaligned float vector takes the scalar loads
    x = rng.standard_normal(1001).astype(np.float32)
    d_x = t.from_numpy(x).cuda()
    d_masked = t.empty(1000, dtype=t.int64, device="cuda")
    d_mask = t.zeros(1000, dtype=t.int64, device="cuda")
    ms = LMS.Full(P61)
    seed = util.seed_bytes("fusedmask/unaligned")
    ctx.fixed_encode_mask_dev(ms, P61, FRAC, d_x[1:], 1000, seed, d_mask, d_masked)
    ctx.synchronize()
    _, emasked = oracle.mask(util.to_oracle_masking(oracle, ms), oracle.fixed_encode(x[1:], FRAC, P61), oracle.rng_from_seed_bytes(seed))
    assert np.array_equal(d_masked.cpu().numpy(), util.canon(oracle, P61, emasked))
    with pytest.raises(sda_b200.SdaClientError, match="differs"):
        ctx.fixed_encode_mask_dev(LMS.Full(433), P61, FRAC, d_x, 10, seed, d_mask, d_masked)
    # several participants per call, strided rows: the same as one call per participant
    P, dim, ld = 5, 777, 800
    xs = rng.standard_normal((P, ld)).astype(np.float32)
    seeds = b"".join(util.seed_bytes(f"fusedmask/batch/{pi}") for pi in range(P))
    for ms in (LMS.Full(P61), LMS.ChaCha(P61, dim, 128)):
        nm = ms.mask_len(dim)
        d_masks = t.zeros((P, nm), dtype=t.int64, device="cuda")
        d_out = t.zeros((P, ld), dtype=t.int64, device="cuda")
        ctx.fixed_encode_mask_dev(ms, P61, FRAC, t.from_numpy(xs).cuda(), dim, seeds, d_masks, d_out, P=P, x_ld=ld, masked_ld=ld)
        ctx.synchronize()
        for pi in range(P):
            emask, emasked = oracle.mask(util.to_oracle_masking(oracle, ms), oracle.fixed_encode(xs[pi, :dim], FRAC, P61),
                                         oracle.rng_from_seed_bytes(seeds[32 * pi:32 * pi + 32]))
            assert np.array_equal(d_out[pi, :dim].cpu().numpy(), util.canon(oracle, P61, emasked))
            assert d_masks[pi].cpu().numpy().tolist() == [int(v) % P61 for v in emask]
