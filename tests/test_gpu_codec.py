"""Share wire codec on the GPU (sda_varint_*) against the oracle's integer-encoding 1.0 restatement:
client/src/crypto/encryption/sodium.rs:35-41 (encode loop) and :83-90 (decode loop)."""
import numpy as np
import pytest

import sda_b200
import util
from sda_b200 import SdaClientError, params

pytestmark = pytest.mark.gpu


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


def values(rng, n, kind):
    if kind == "shares61":                       # canonical 61-bit shares: 9 bytes each
        return rng.integers(0, params.P61, size=n, dtype=np.int64)
    if kind == "small":                          # mod-433 shares: 1-2 bytes
        return rng.integers(0, 433, size=n, dtype=np.int64)
    if kind == "signed":                         # the reference's signed representatives
        return rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
    mags = rng.integers(0, 64, size=n)           # every byte length 1..10, both signs
    v = (rng.integers(0, 1 << 62, size=n, dtype=np.int64) >> (62 - np.minimum(mags, 62))).astype(np.int64)
    v[mags == 63] = np.iinfo(np.int64).min
    v[mags == 62] = np.iinfo(np.int64).max
    return np.where(rng.integers(0, 2, size=n) == 1, -v, v).astype(np.int64)


@pytest.mark.parametrize("kind", ["shares61", "small", "signed", "mixed"])
def test_varint_encode_decode_match_oracle(ctx, oracle, kind):
    rng = np.random.default_rng(5)
    for n in [0, 1, 2, 7, 255, 256, 2047, 2048, 2049, 4097, 100_003]:
        v = values(rng, n, kind)
        exp = oracle.varint_encode(v)
        got = ctx.varint_encode(v)
        assert got.dtype == np.uint8 and np.array_equal(got, exp), (kind, n)
        back = ctx.varint_decode(got)
        assert np.array_equal(back, v), (kind, n)
        assert np.array_equal(back, oracle.varint_decode(exp))


def test_varint_known_answers(ctx):
    assert ctx.varint_encode([0]).tolist() == [0]
    assert ctx.varint_encode([-1]).tolist() == [1]
    assert ctx.varint_encode([63]).tolist() == [126]
    assert ctx.varint_encode([432]).tolist() == [0xE0, 0x06]
    assert len(ctx.varint_encode([params.P61 - 1])) == 9
    assert len(ctx.varint_encode([np.iinfo(np.int64).min])) == 10
    assert ctx.varint_decode(bytes([0xE0, 0x06, 0x00, 0x01])).tolist() == [432, 0, -1]


def test_varint_decode_unaligned_and_offsets(ctx, oracle, torch_cuda):
    """device buffers at every byte alignment, streams crossing the 4096-byte chunks"""
    t = torch_cuda
    rng = np.random.default_rng(9)
    v = values(rng, 5000, "mixed")
    enc = oracle.varint_encode(v)
    for shift in (0, 1, 3, 8, 15):
        d_buf = t.zeros(len(enc) + 32, dtype=t.uint8, device="cuda")
        d_buf[shift:shift + len(enc)] = t.from_numpy(enc).cuda()
        d_out = t.empty(len(v), dtype=t.int64, device="cuda")
        cnt = ctx.varint_decode_dev(d_buf[shift:], len(enc), d_out, len(v))
        assert cnt == len(v) and np.array_equal(d_out.cpu().numpy(), v)
        d_enc = t.zeros(10 * len(v) + 32, dtype=t.uint8, device="cuda")
        ln = ctx.varint_encode_dev(t.from_numpy(v).cuda(), len(v), d_enc[shift:])
        assert ln == len(enc) and np.array_equal(d_enc[shift:shift + ln].cpu().numpy(), enc)
        assert int(d_enc[:shift].sum()) == 0 and int(d_enc[shift + ln:].sum()) == 0   # nothing written outside


def test_varint_decode_rejects_malformed_streams(ctx):
    with pytest.raises(SdaClientError, match="ends inside a value"):
        ctx.varint_decode(bytes([0x01, 0x80]))
    with pytest.raises(SdaClientError, match="longer than 10 bytes"):
        ctx.varint_decode(bytes([0x80] * 10 + [0x01]))
    with pytest.raises(SdaClientError, match="capacity"):
        ctx.varint_decode(bytes([1, 2, 3]), cap=2)
    assert ctx.varint_decode(bytes([0xFF] * 9 + [0x01])).tolist() == [np.iinfo(np.int64).min]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    return torch


# ---- server snapshot transpose: server/src/snapshot.rs:11-27, stores.rs:86-101 ------------------
@pytest.mark.parametrize("P,n,maxlen", [(1, 1, 5), (3, 2, 0), (7, 3, 40), (64, 9, 3000), (1000, 5, 17), (2, 8, 100_000)])
def test_snapshot_transpose_matches_the_reference_regrouping(ctx, oracle, P, n, maxlen):
    """variable-length blobs (empty ones included), every alignment: clerk-major bytes and offsets against the oracle"""
    import torch as t
    rng = np.random.default_rng(P * 131 + n)
    blobs = [[rng.integers(0, 256, size=int(rng.integers(0, maxlen + 1)), dtype=np.uint8).tobytes() for _ in range(n)]
             for _ in range(P)]
    flat = b"".join(b for part in blobs for b in part)
    offsets = np.cumsum([0] + [len(b) for part in blobs for b in part]).astype(np.uint64)
    d_in = t.from_numpy(np.frombuffer(flat + b"\0", dtype=np.uint8).copy()).cuda()
    d_out = t.zeros(len(flat) + 1, dtype=t.uint8, device="cuda")
    out_off = ctx.snapshot_transpose_dev(d_in, offsets, P, n, d_out)
    ctx.synchronize()
    got = d_out.cpu().numpy().tobytes()
    expect = oracle.snapshot_transpose(blobs)
    assert got[:len(flat)] == b"".join(b for job in expect for b in job)
    for c in range(n):
        for p in range(P):
            assert got[int(out_off[c * P + p]):int(out_off[c * P + p + 1])] == blobs[p][c]
    assert int(out_off[-1]) == len(flat)


def test_snapshot_transpose_of_varint_coded_shares_feeds_the_clerk(ctx, oracle):
    """participant shares -> varint wire coding per clerk -> snapshot transpose -> the clerk decodes its job and sums"""
    import torch as t
    from sda_b200 import params
    s = params.config3()
    P, dim = 5, 3001
    n, B = s.output_size(), s.batches(dim)
    rng = np.random.default_rng(3)
    shares = [ctx.share_generate(s, rng.integers(0, s.modulus, size=dim, dtype=np.int64), util.seed_bytes(f"snap/{p}"))
              for p in range(P)]
    blobs = [[ctx.varint_encode(shares[p][c]) for c in range(n)] for p in range(P)]
    flat = b"".join(b for part in blobs for b in part)
    offsets = np.cumsum([0] + [len(b) for part in blobs for b in part]).astype(np.uint64)
    d_out = t.zeros(len(flat), dtype=t.uint8, device="cuda")
    out_off = ctx.snapshot_transpose_dev(t.from_numpy(np.frombuffer(flat, dtype=np.uint8).copy()).cuda(), offsets, P, n, d_out)
    job = d_out.cpu().numpy().tobytes()
    for c in range(n):
        rows = [ctx.varint_decode(job[int(out_off[c * P + p]):int(out_off[c * P + p + 1])]) for p in range(P)]
        assert all(len(r) == B for r in rows)
        assert np.array_equal(ctx.share_combine(s, np.stack(rows)),
                              oracle.canonical(s.modulus, oracle.share_combine(s.modulus, np.stack([shares[p][c] for p in range(P)]))))
