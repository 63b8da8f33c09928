"""The host-buffer entry points with PINNED buffers take a sliced path (copies of neighbouring slices overlap the
kernel, csrc/api.cu `share_generate_sliced` / `combine_host`): same results as the plain path and as the oracle,
bit for bit, for vector lengths around the slice boundaries."""
import numpy as np
import pytest

import util
from sda_b200 import LinearSecretSharingScheme as LSS
from sda_b200 import params

pytestmark = pytest.mark.gpu

P61 = params.P61


def pinned_copy(ctx, a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    h = ctx.pinned_empty(a.size).reshape(a.shape)
    h[...] = a
    return h


@pytest.mark.parametrize("dim", [524_288, 600_001, 8 * 512 * 3 * 41, 1_572_865])
def test_share_generate_pinned_matches_pageable_and_oracle(ctx, oracle, dim):
    s = params.config3()                       # packed Shamir k=3 / n=5 / t=2 over 2^61 - 1 (tensor-core shape)
    rng = np.random.default_rng(dim)
    sec = util.rand_secrets(rng, dim, P61)
    sec[:7] = [-1, -(1 << 62), P61, P61 - 1, 0, 1, (1 << 62) - 12345]   # negative / non-canonical inputs
    seed = util.seed_bytes(("hostpipe", dim))
    plain = ctx.share_generate(s, sec, seed)                           # pageable buffers: one slice
    h_sec = pinned_copy(ctx, sec)
    h_out = ctx.pinned_empty(plain.size).reshape(plain.shape)
    h_out[...] = -7
    got = ctx.share_generate(s, h_sec, seed, out=h_out)
    assert got is h_out
    assert np.array_equal(got, plain)
    want = util.canon(oracle, P61, util.oracle_generate(oracle, s, sec, seed))
    assert np.array_equal(np.asarray(got), want)


def test_share_generate_pinned_other_round_counts(ctx):
    s = params.config3()
    dim = 700_003
    sec = util.rand_secrets(np.random.default_rng(3), dim, P61)
    seed = util.seed_bytes("hostpipe-rounds")
    h_sec = pinned_copy(ctx, sec)
    try:
        for r in (8, 12):
            ctx.set_rng_rounds(r)
            plain = ctx.share_generate(s, sec, seed)
            h_out = ctx.pinned_empty(plain.size).reshape(plain.shape)
            assert np.array_equal(ctx.share_generate(s, h_sec, seed, out=h_out), plain)
    finally:
        ctx.set_rng_rounds(20)


@pytest.mark.parametrize("modulus", [P61, params.P61_GENERIC, 433])
@pytest.mark.parametrize("P,L", [(1, 524_288), (5, 600_001), (3, 8 * 1024 * 70 + 1)])
def test_share_combine_pinned_rows_and_matrix(ctx, oracle, modulus, P, L):
    s = LSS.Additive(3, modulus)
    rng = np.random.default_rng(P * 1000 + L % 1000)
    m = rng.integers(-(1 << 62), 1 << 62, size=(P, L), dtype=np.int64)
    want = util.canon(oracle, modulus, oracle.share_combine(modulus, m))
    # `Vec<Vec<Share>>` as P separately pinned rows
    rows = [pinned_copy(ctx, m[p]) for p in range(P)]
    out = ctx.pinned_empty(L)
    out[...] = -7
    got = ctx.share_combine(s, rows, out=out)
    assert np.array_equal(np.asarray(got), want)
    # one pinned [P][L] matrix
    h_m = pinned_copy(ctx, m)
    out2 = ctx.pinned_empty(L)
    got2 = ctx.share_combine(s, h_m, out=out2)
    assert np.array_equal(np.asarray(got2), want)
    # pageable inputs, pinned output and the other way round stay on the plain path
    assert np.array_equal(ctx.share_combine(s, m), want)
    assert np.array_equal(np.asarray(ctx.share_combine(s, [m[p] for p in range(P)], out=ctx.pinned_empty(L))), want)


def test_share_combine_pinned_more_rows_than_one_staging_tile(ctx, oracle):
    """the staging buffer holds 1 GB of rows: 300 rows of 4 MB take two row tiles, each walked in column slices,
    the second accumulating onto the first (matrix and row-pointer forms)"""
    P, L = 300, 524_288
    s = LSS.Additive(3, P61)
    rng = np.random.default_rng(300)
    h_m = ctx.pinned_empty(P * L).reshape(P, L)
    h_m[...] = rng.integers(0, P61, size=(P, L), dtype=np.int64)
    h_m[7, ::3] = -5
    want = util.canon(oracle, P61, oracle.share_combine(P61, np.asarray(h_m)))
    out = ctx.pinned_empty(L)
    out[...] = -1
    assert np.array_equal(np.asarray(ctx.share_combine(s, h_m, out=out)), want)
    out[...] = -1
    assert np.array_equal(np.asarray(ctx.share_combine(s, [h_m[p] for p in range(P)], out=out)), want)
