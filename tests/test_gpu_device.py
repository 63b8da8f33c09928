"""Device-pointer entry points and BASELINE-size properties (`-m gpu`).  torch is only the
device allocator here; every computation under test is a libsda_b200 kernel."""
import numpy as np
import pytest

import util
from sda_b200 import LinearMaskingScheme as LMS
from sda_b200 import LinearSecretSharingScheme as LSS
from sda_b200 import params

pytestmark = pytest.mark.gpu

P61 = params.P61


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    torch.cuda.init()
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def test_synth_fill_matches_oracle(ctx, oracle, torch_cuda):
    t = torch_cuda
    for m in (433, P61, params.P61_GENERIC):
        for start, count in [(0, 1), (0, 8), (3, 5), (5, 100), (1 << 40, 4097), (7, 1)]:
            out = t.empty(count, dtype=t.int64, device="cuda")
            ctx.synth_fill_dev(3, m, start, count, out)
            ctx.synchronize()
            assert np.array_equal(host(out), oracle.synth_fill(3, m, start, count)), (m, start, count)


@pytest.mark.parametrize("path", [1, 2, 3, 4], ids=["cuda_cores", "tensor_cores", "tensor_cores_v1", "tensor_cores_any_shape"])
@pytest.mark.parametrize("mk", [params.config3, params.config4, params.config5], ids=["cfg3", "cfg4", "cfg5"])
def test_packed_kernels_agree_with_oracle(ctx, oracle, torch_cuda, mk, path):
    """every Mersenne-61 share-gen kernel a BASELINE shape can be sent to (IMAD.WIDE limbs; tcgen05 byte-limb GEMM: paired
    tiles, first generation, share count at run time) against the oracle: several participants, dims that end inside a
    tile, negative and out-of-range secrets"""
    t = torch_cuda
    s = mk()
    k = s.input_size()
    rng = np.random.default_rng(7)
    ctx.set_packed_path(path)
    try:
        for P, dim in [(1, 1), (2, 128 * k), (3, 512 * k + 1), (2, 3 * 512 * k - 2), (4, 40000)]:
            n, B = s.output_size(), s.batches(dim)
            secrets = rng.integers(0, s.modulus, size=(P, dim), dtype=np.int64)
            secrets[0, ::7] = rng.integers(-(1 << 63), 1 << 63, size=secrets[0, ::7].shape, dtype=np.int64)
            secrets[-1, ::5] = np.array([0, s.modulus - 1, s.modulus, -1, (1 << 63) - 1, -(1 << 63)] * dim,
                                        dtype=np.int64)[:secrets[-1, ::5].size]
            seeds = b"".join(util.seed_bytes(f"paths/{P}/{dim}/{pi}") for pi in range(P))
            d_out = t.empty((P, n, B), dtype=t.int64, device="cuda")
            ctx.share_generate_dev(s, dev(t, secrets), dim, P, dim, seeds, d_out)
            ctx.synchronize()
            assert ("tcgen05" in ctx.last_kernel()) == (path >= 2)
            assert ("at run time" in ctx.last_kernel()) == (path == 4) and ("paired" in ctx.last_kernel()) == (path in (2, 4))
            got = host(d_out)
            for pi in range(P):
                exp = util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32], matrix=True)
                assert np.array_equal(got[pi], util.canon(oracle, s.modulus, exp)), (path, P, dim, pi)
    finally:
        ctx.set_packed_path(0)


@pytest.mark.parametrize("shape", [(4, 2, 6), (2, 4, 8), (8, 2, 3), (1, 4, 5), (7, 4, 7), (3, 2, 6),
                                   (3, 3, 7), (5, 1, 12), (2, 6, 9), (7, 5, 16), (8, 8, 32), (1, 7, 3), (4, 3, 20), (1, 1, 2),
                                   (15, 1, 17), (1, 15, 18), (9, 7, 30), (11, 3, 15), (2, 10, 13), (5, 11, 19), (3, 12, 16), (2, 14, 17)])
@pytest.mark.parametrize("offset,ld_pad", [(0, 0), (1, 1)])
def test_paired_tile_kernel_with_run_time_share_count(ctx, oracle, torch_cuda, shape, offset, ld_pad):
    """packed_tc2n.cu: (k, t) templated (k + t <= 16, odd t included), n <= 32 at run time in groups of 8, over 2^61-1:
    several participants, aligned (bulk copy, 16-byte stores) and unaligned sources, vectors ending inside a pass,
    negative secrets -- every share against the oracle"""
    t = torch_cuda
    k, tt, n = shape
    try:
        s = util.packed_scheme(P61, k, tt, n, oracle)
    except StopIteration:
        pytest.skip("no suitable prime orders in p-1")
    rng = np.random.default_rng(k * 10 + tt + n)
    for P, dim in [(1, 2), (3, 5 * 512 * k + 2 * k + 1), (2, 1024 * k)]:
        ld = dim + (dim & 1) + ld_pad
        B = s.batches(dim)
        secrets = rng.integers(0, P61, size=(P, dim), dtype=np.int64)
        secrets[-1, ::7] = rng.integers(-(1 << 63), 1 << 63, size=secrets[-1, ::7].shape, dtype=np.int64)
        flat = np.zeros(offset + P * ld, dtype=np.int64)
        for pi in range(P):
            flat[offset + pi * ld: offset + pi * ld + dim] = secrets[pi]
        d_in = dev(t, flat)[offset:]
        seeds = b"".join(util.seed_bytes(f"tc2n/{shape}/{P}/{dim}/{pi}") for pi in range(P))
        d_out = t.empty((P, n, B), dtype=t.int64, device="cuda")
        ctx.share_generate_dev(s, d_in, ld, P, dim, seeds, d_out)
        ctx.synchronize()
        assert "at run time" in ctx.last_kernel() and "tcgen05" in ctx.last_kernel(), ctx.last_kernel()
        got = host(d_out)
        for pi in range(P):
            exp = util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32], matrix=True)
            assert np.array_equal(got[pi], util.canon(oracle, P61, exp)), (shape, P, dim, pi)


@pytest.mark.parametrize("shape", [(2, 3, 6), (7, 5, 16), (1, 1, 2), (4, 1, 9), (9, 7, 32), (12, 2, 20), (2, 10, 13)])
@pytest.mark.parametrize("p,rounds", [(P61, 12), (P61, 8), (params.P61_GENERIC, 20)])
def test_runtime_shaped_kernel_many_participants(ctx, oracle, torch_cuda, shape, p, rounds):
    """packed_tcg.cu (other primes; 2^61-1 with 8 / 12 rounds) on device buffers: several participants, vectors spanning
    many passes and ending inside one, strided rows, negative secrets -- every share against the oracle"""
    t = torch_cuda
    k, tt, n = shape
    try:
        s = util.packed_scheme(p, k, tt, n, oracle)
    except StopIteration:
        pytest.skip("no suitable prime orders in p-1")
    rng = np.random.default_rng(k * 100 + n)
    ctx.set_rng_rounds(rounds)
    try:
        _runtime_shaped_cases(ctx, oracle, t, s, shape, p, rounds, rng)
    finally:
        ctx.set_rng_rounds(20)


def _runtime_shaped_cases(ctx, oracle, t, s, shape, p, rounds, rng):
    k, tt, n = shape
    for P, dim, ld in [(1, 1, 1), (3, 256 * k * 3 + 1, 256 * k * 3 + 4), (2, 40000, 40000), (5, 256 * k, 256 * k + 2)]:
        B = s.batches(dim)
        secrets = np.zeros((P, ld), dtype=np.int64)
        secrets[:, :dim] = rng.integers(0, p, size=(P, dim), dtype=np.int64)
        secrets[0, :dim:9] = rng.integers(-(1 << 63), 1 << 63, size=secrets[0, :dim:9].shape, dtype=np.int64)
        seeds = b"".join(util.seed_bytes(f"tcg/{shape}/{P}/{dim}/{pi}") for pi in range(P))
        d_out = t.empty((P, n, B), dtype=t.int64, device="cuda")
        ctx.share_generate_dev(s, dev(t, secrets), ld, P, dim, seeds, d_out)
        ctx.synchronize()
        assert "run-time shape" in ctx.last_kernel() and "tcgen05" in ctx.last_kernel()
        got = host(d_out)
        for pi in range(P):
            exp = util.oracle_generate(oracle, s, secrets[pi, :dim], seeds[32 * pi:32 * pi + 32], rounds, matrix=True)
            assert np.array_equal(got[pi], util.canon(oracle, p, exp)), (shape, P, dim, pi)


@pytest.mark.parametrize("mk", [params.config3, params.config4, params.config5], ids=["cfg3", "cfg4", "cfg5"])
@pytest.mark.parametrize("offset,ld_pad", [(0, 0), (1, 0), (0, 1), (1, 1), (2, 2)])
def test_packed_tc_secret_sources_of_every_alignment(ctx, oracle, torch_cuda, mk, offset, ld_pad):
    """the tensor-core kernel takes a pass's secrets by bulk copy when the source is 16-byte aligned and every
    participant starts at an even element, and by per-thread loads otherwise: same shares either way"""
    t = torch_cuda
    s = mk()
    k = s.input_size()
    P, dim = 3, 5 * 512 * k + 2 * k + 1          # interior passes and a ragged last one
    ld = dim + (dim & 1) + ld_pad                # even (+ pad): odd strides break the alignment of participants 1, 2
    rng = np.random.default_rng(offset * 10 + ld_pad)
    secrets = rng.integers(0, s.modulus, size=(P, dim), dtype=np.int64)
    secrets[1, ::11] = rng.integers(-(1 << 63), 1 << 63, size=secrets[1, ::11].shape, dtype=np.int64)
    flat = np.zeros(offset + P * ld, dtype=np.int64)
    for pi in range(P):
        flat[offset + pi * ld: offset + pi * ld + dim] = secrets[pi]
    d_flat = dev(t, flat)
    d_in = d_flat[offset:]                       # data pointer 8 * offset bytes past a 256-byte aligned allocation
    assert (d_in.data_ptr() % 16 == 0) == (offset % 2 == 0)
    n, B = s.output_size(), s.batches(dim)
    seeds = b"".join(util.seed_bytes(f"align/{offset}/{ld_pad}/{pi}") for pi in range(P))
    d_out = t.empty((P, n, B), dtype=t.int64, device="cuda")
    ctx.share_generate_dev(s, d_in, ld, P, dim, seeds, d_out)
    ctx.synchronize()
    assert "tcgen05" in ctx.last_kernel()
    got = host(d_out)
    for pi in range(P):
        exp = util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32], matrix=True)
        assert np.array_equal(got[pi], util.canon(oracle, s.modulus, exp)), (offset, ld_pad, pi)


@pytest.mark.parametrize("mk", [params.config2, params.config3, params.config4, params.config5,
                                lambda: LSS.Additive(3, 433), params.reference_test])
def test_share_generate_dev_multi_participant(ctx, oracle, torch_cuda, mk):
    """P participants in one launch, strided secrets, one seed each"""
    t = torch_cuda
    s = mk()
    rng = np.random.default_rng(2)
    for P, dim, ld in [(1, 10, 10), (3, 1001, 1004), (5, 4096, 4096), (2, 12289, 12292)]:
        n, B = s.output_size(), s.batches(dim)
        secrets = np.zeros((P, ld), dtype=np.int64)
        secrets[:, :dim] = rng.integers(0, s.modulus, size=(P, dim), dtype=np.int64)
        seeds = b"".join(util.seed_bytes(f"dev/{P}/{pi}") for pi in range(P))
        d_in = dev(t, secrets)
        d_out = t.empty((P, n, B), dtype=t.int64, device="cuda")
        ctx.share_generate_dev(s, d_in, ld, P, dim, seeds, d_out)
        ctx.synchronize()
        got = host(d_out)
        for pi in range(P):
            exp = util.oracle_generate(oracle, s, secrets[pi, :dim], seeds[32 * pi:32 * pi + 32], matrix=True)
            assert np.array_equal(got[pi], util.canon(oracle, s.modulus, exp)), (P, dim, pi)


def test_share_combine_dev_accumulate(ctx, oracle, torch_cuda):
    """streaming tiles into a running sum == one combine over all rows (clerk.rs:71-72 FIXME)"""
    t = torch_cuda
    s = params.config4()
    rng = np.random.default_rng(4)
    P, L, ld = 300, 5000, 5004
    rows = np.zeros((P, ld), dtype=np.int64)
    rows[:, :L] = rng.integers(0, P61, size=(P, L), dtype=np.int64)
    exp = oracle.share_combine(P61, rows[:, :L].copy())
    d_rows = dev(t, rows)
    acc = t.zeros(L, dtype=t.int64, device="cuda")
    for p0 in range(0, P, 64):
        pc = min(64, P - p0)
        ctx.share_combine_dev(s, d_rows[p0:], ld, pc, L, acc, d_acc_in=acc if p0 else None)
    ctx.synchronize()
    assert np.array_equal(host(acc), exp)
    one = t.empty(L, dtype=t.int64, device="cuda")
    ctx.share_combine_dev(s, d_rows, ld, P, L, one)
    ctx.synchronize()
    assert np.array_equal(host(one), exp)


def test_share_generate_combine_dev(ctx, oracle, torch_cuda):
    """fused participant->clerk path == generate then per-clerk combine"""
    t = torch_cuda
    for s in (params.config3(), params.config2()):
        P, dim = 7, 3000
        n, B = s.output_size(), s.batches(dim)
        rng = np.random.default_rng(8)
        secrets = rng.integers(0, P61, size=(P, dim), dtype=np.int64)
        seeds = b"".join(util.seed_bytes(f"fuse/{pi}") for pi in range(P))
        out = t.empty((n, B), dtype=t.int64, device="cuda")
        ctx.share_generate_combine_dev(s, dev(t, secrets), dim, P, dim, seeds, out)
        ctx.synchronize()
        shares = np.stack([util.canon(oracle, P61, util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32],
                                                                      matrix=True)) for pi in range(P)])
        exp = np.stack([oracle.share_combine(P61, np.ascontiguousarray(shares[:, c, :])) for c in range(n)])
        assert np.array_equal(host(out), exp)


# a 61-bit prime not of Mersenne form; 2^31-1; the largest prime = 1 mod 2002 below 2^62 (the oracle's i64 sums limit
# it to p < 2^62; the widest byte limbs it can check); and 1.3 * 2^61, for which gen_range rejects 2.5 % of the words,
# so the kernel must notice and hand the call over to the exact path
@pytest.mark.parametrize("p", [params.P61_GENERIC, (1 << 31) - 1, 4611686018427297811, 2997595911977817151],
                         ids=["61bit", "mersenne31", "62bit", "61bit_rejecting"])
# (3, 4, 8) is left to the reference's own parameters (tests/test_gpu_parity.py, p = 433): with k+t+1 = 2^3 and n+1 = 3^2
# tss takes its FFT path, which is only defined for roots of exactly those orders
@pytest.mark.parametrize("shape", [(3, 2, 5), (5, 4, 9), (3, 4, 7)], ids=["k3t2n5", "k5t4n9", "k3t4n7"])
def test_packed_tensor_core_kernel_any_prime(ctx, oracle, torch_cuda, shape, p):
    """the byte-limb GEMM kernel over primes that are not 2^61-1: generic draw reduction and compose"""
    t = torch_cuda
    k, tt, n = shape
    try:
        s = util.packed_scheme(p, k, tt, n, oracle)
    except StopIteration:
        pytest.skip("no suitable prime orders in p - 1")
    rng = np.random.default_rng(31)
    for P, dim in [(1, k), (2, 512 * k + 2), (3, 20000)]:
        B = s.batches(dim)
        secrets = rng.integers(0, p, size=(P, dim), dtype=np.int64)
        secrets[0, ::5] = rng.integers(-(1 << 63), 1 << 63, size=secrets[0, ::5].shape, dtype=np.int64)
        seeds = b"".join(util.seed_bytes(f"anyp/{p}/{shape}/{P}/{pi}") for pi in range(P))
        d_out = t.empty((P, n, B), dtype=t.int64, device="cuda")
        ctx.share_generate_dev(s, dev(t, secrets), dim, P, dim, seeds, d_out)
        ctx.synchronize()
        got = host(d_out)
        for pi in range(P):
            exp = util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32], matrix=True)
            assert np.array_equal(got[pi], util.canon(oracle, p, exp)), (p, shape, P, dim, pi)


@pytest.mark.parametrize("mk", [params.config3, params.config4, params.config5], ids=["cfg3", "cfg4", "cfg5"])
def test_fused_share_generate_combine_tmem_accumulation(ctx, oracle, torch_cuda, mk):
    """the fused tensor-core kernel sums the participants inside TMEM (drained every 256): participant counts
    around the pass (4) and drain (256) boundaries, a running sum passed in, ranges ending inside a tile,
    negative secrets"""
    t = torch_cuda
    s = mk()
    k, n, m = s.input_size(), s.output_size(), s.modulus
    rng = np.random.default_rng(21)
    for P, dim in [(1, 5), (3, 128 * k), (4, 129 * k + 1), (7, 1000), (258, 300 * k), (515, 130)]:
        B = s.batches(dim)
        secrets = rng.integers(0, m, size=(P, dim), dtype=np.int64)
        secrets[0, ::3] = rng.integers(-(1 << 63), 1 << 63, size=secrets[0, ::3].shape, dtype=np.int64)
        seeds = b"".join(util.seed_bytes(f"fused/{P}/{dim}/{pi}") for pi in range(P))
        acc_in = rng.integers(0, m, size=(n, B), dtype=np.int64)
        out = t.empty((n, B), dtype=t.int64, device="cuda")
        ctx.share_generate_combine_dev(s, dev(t, secrets), dim, P, dim, seeds, out, d_acc_in=dev(t, acc_in))
        ctx.synchronize()
        assert "TMEM-accumulated" in ctx.last_kernel()
        total = acc_in.astype(object)
        for pi in range(P):
            sh = util.canon(oracle, m, util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32], matrix=True))
            total = total + sh.astype(object)
        assert np.array_equal(host(out), (total % m).astype(np.int64)), (P, dim)


def test_mod_reduce_and_unmask_dev(ctx, oracle, torch_cuda):
    t = torch_cuda
    rng = np.random.default_rng(6)
    v = rng.integers(-(1 << 63), (1 << 63) - 1, size=10001, dtype=np.int64)
    for m in (433, P61, params.P61_GENERIC, (1 << 63) - 1):
        out = t.empty(len(v), dtype=t.int64, device="cuda")
        ctx.mod_reduce_dev(m, dev(t, v), len(v), out)
        ctx.synchronize()
        assert np.array_equal(host(out), util.canon(oracle, m, v))
    ms = LMS.Full(P61)
    a, b = rng.integers(0, P61, size=999, dtype=np.int64), rng.integers(0, P61, size=999, dtype=np.int64)
    out = t.empty(999, dtype=t.int64, device="cuda")
    ctx.unmask_dev(ms, dev(t, a), dev(t, b), 999, out)
    ctx.synchronize()
    assert np.array_equal(host(out), util.canon(oracle, P61, oracle.unmask(util.to_oracle_masking(oracle, ms), a, b)))


# ---- BASELINE sizes: parity on the full vector + size-independent properties -------------------------
def test_config3_full_size(ctx, oracle, torch_cuda):
    """packed Shamir k=3/n=5, dim = 10M, p = 2^61-1: bit-exact against the oracle on the whole
    vector, then share -> clerk-sum -> reconstruct == sum of secrets for 3 participants"""
    t = torch_cuda
    s = params.config3()
    dim, P = 10_000_000, 3
    n, B = 5, s.batches(dim)
    d_sec = t.empty((P, dim), dtype=t.int64, device="cuda")
    for pi in range(P):
        ctx.synth_fill_dev(3, P61, pi * dim, dim, d_sec[pi])
    seeds = b"".join(util.seed_bytes(f"cfg3/{pi}") for pi in range(P))
    d_sh = t.empty((P, n, B), dtype=t.int64, device="cuda")
    ctx.share_generate_dev(s, d_sec, dim, P, dim, seeds, d_sh)
    ctx.synchronize()
    sec0 = oracle.synth_fill(3, P61, 0, dim)
    assert np.array_equal(host(d_sec[0]), sec0)
    exp0 = util.oracle_generate(oracle, s, sec0, seeds[:32], matrix=True)
    assert np.array_equal(host(d_sh[0]), exp0)
    # clerk sums, then reveal from all 5 clerks
    d_sum = t.empty((n, B), dtype=t.int64, device="cuda")
    for c in range(n):
        ctx.share_combine_dev(s, d_sh[:, c, :], n * B, P, B, d_sum[c])
    d_rec = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.secret_reconstruct_dev(s, dim, list(range(n)), d_sum, B, n, B, d_rec)
    d_tot = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.share_combine_dev(s, d_sec, dim, P, dim, d_tot)
    ctx.synchronize()
    assert t.equal(d_rec, d_tot)
    assert int(d_sh.min()) >= 0 and int(d_sh.max()) < P61


@pytest.mark.parametrize("mk", [params.config3, params.config4, params.config5], ids=["cfg3", "cfg4", "cfg5"])
@pytest.mark.parametrize("out_offset", [0, 1])
def test_reveal_many_tiles_any_alignment(ctx, oracle, torch_cuda, mk, out_offset):
    """reveal_tc.cu on device buffers: more tiles than resident CTAs (every CTA loops), a vector that ends inside a tile and
    inside a batch, an output that is / is not 16-byte aligned (bulk stores / per-thread stores), shares that are negative
    representatives, a subset of the clerks -- against the oracle's reconstruction"""
    t = torch_cuda
    s = mk()
    c = s.c
    p, k, n, need = c.modulus, c.secret_count, c.share_count, c.secret_count + c.privacy_threshold
    so = util.to_oracle_sharing(oracle, s)
    rng = np.random.default_rng(k + out_offset)
    for dim in (128 * k * 2500 + 128 * k - 1, 128 * k * 3, 128 * k * 2 + k + 1):
        B = s.batches(dim)
        idx = sorted(rng.permutation(n)[:need + (need < n)].tolist())
        shares = rng.integers(0, p, size=(len(idx), B), dtype=np.int64)
        exp = util.canon(oracle, p, oracle.secret_reconstruct(so, dim, idx, shares))
        shares[:, ::5] -= p                                   # another representative of the same residue
        d_out = t.full((dim + 2,), -1, dtype=t.int64, device="cuda")
        ctx.secret_reconstruct_dev(s, dim, idx, dev(t, shares), B, len(idx), B, d_out[out_offset:])
        ctx.synchronize()
        got = host(d_out)
        assert np.array_equal(got[out_offset:out_offset + dim], exp), (dim, idx)
        assert (got[:out_offset] == -1).all() and (got[out_offset + dim:] == -1).all()      # nothing beyond the vector


def test_config2_full_size(ctx, oracle, torch_cuda):
    """additive 3-way, dim = 1M, p = 2^61-1, 16 participants: split parity on one participant,
    then sum of the three clerk sums == sum of secrets"""
    t = torch_cuda
    s = params.config2()
    dim, P, n = 1_000_000, 16, 3
    d_sec = t.empty((P, dim), dtype=t.int64, device="cuda")
    ctx.synth_fill_dev(2, P61, 0, P * dim, d_sec)
    seeds = b"".join(util.seed_bytes(f"cfg2/{pi}") for pi in range(P))
    d_sh = t.empty((P, n, dim), dtype=t.int64, device="cuda")
    ctx.share_generate_dev(s, d_sec, dim, P, dim, seeds, d_sh)
    ctx.synchronize()
    sec5 = oracle.synth_fill(2, P61, 5 * dim, dim)
    exp = util.canon(oracle, P61, util.oracle_generate(oracle, s, sec5, seeds[5 * 32:6 * 32]))
    assert np.array_equal(host(d_sh[5]), exp)
    d_sum = t.empty((n, dim), dtype=t.int64, device="cuda")
    for c in range(n):
        ctx.share_combine_dev(s, d_sh[:, c, :], n * dim, P, dim, d_sum[c])
    d_rec = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.secret_reconstruct_dev(s, dim, [0, 1, 2], d_sum, dim, n, dim, d_rec)
    d_tot = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.share_combine_dev(s, d_sec, dim, P, dim, d_tot)
    ctx.synchronize()
    assert t.equal(d_rec, d_tot)


def test_config4_clerk_sum_slice(ctx, oracle, torch_cuda):
    """k=5/n=9 clerk job slice [2048][2M] (config #4 shape, 32.8 GB): checksum-of-checksums --
    the combine of column sums equals the column sums of row-block combines, and sampled
    columns match the oracle"""
    t = torch_cuda
    s = params.config4()
    P, L = 2048, 2_000_000
    d_rows = t.empty((P, L), dtype=t.int64, device="cuda")
    ctx.synth_fill_dev(4, P61, 0, P * L, d_rows)
    out = t.empty(L, dtype=t.int64, device="cuda")
    ctx.share_combine_dev(s, d_rows, L, P, L, out)
    parts = t.empty((4, L), dtype=t.int64, device="cuda")
    for q in range(4):
        ctx.share_combine_dev(s, d_rows[q * 512:], L, 512, L, parts[q])
    out2 = t.empty(L, dtype=t.int64, device="cuda")
    ctx.share_combine_dev(s, parts, L, 4, L, out2)
    ctx.synchronize()
    assert t.equal(out, out2)
    cols = [0, 1, 2, 3, 1023, 1024, 999_999, L - 1]
    sample = host(d_rows[:, cols])
    exp = oracle.share_combine(P61, np.ascontiguousarray(sample))
    assert host(out[cols]).tolist() == exp.tolist()
    del d_rows
    t.cuda.empty_cache()


def test_fused_in_place_redo_does_not_double_count(oracle, torch_cuda, monkeypatch):
    """sda_share_generate_combine_dev called in place (d_acc_in == d_out) must survive a set gen_range rejection
    flag: the redo on the materialising path starts from the untouched running sum.  SDA_B200_DEBUG_FORCE_REJECT=1 (read
    at context creation) makes the library treat the fused kernel's flag as set."""
    import sda_b200
    t = torch_cuda
    s = params.config3()
    n, dim, P = s.output_size(), 3000, 6
    B = s.batches(dim)
    rng = np.random.default_rng(77)
    secrets = rng.integers(0, s.modulus, size=(2 * P, dim), dtype=np.int64)
    seeds = b"".join(util.seed_bytes(f"redo/{pi}") for pi in range(2 * P))
    plain = sda_b200.Context(0)
    monkeypatch.setenv("SDA_B200_DEBUG_FORCE_REJECT", "1")
    forced = sda_b200.Context(0)
    monkeypatch.delenv("SDA_B200_DEBUG_FORCE_REJECT")
    results = []
    for c in (plain, forced):
        d_acc = t.empty((n, B), dtype=t.int64, device="cuda")
        c.share_generate_combine_dev(s, dev(t, secrets[:P]), dim, P, dim, seeds[:32 * P], d_acc)
        c.share_generate_combine_dev(s, dev(t, secrets[P:]), dim, P, dim, seeds[32 * P:], d_acc, d_acc_in=d_acc)   # in place
        c.synchronize()
        results.append(host(d_acc))
    assert "draws" not in plain.last_kernel()
    # expected: per-clerk sums of the oracle's shares of all 2P participants
    exp = np.zeros((n, B), dtype=object)
    for pi in range(2 * P):
        exp += util.canon(oracle, s.modulus, util.oracle_generate(oracle, s, secrets[pi], seeds[32 * pi:32 * pi + 32], matrix=True)).astype(object)
    exp = (exp % s.modulus).astype(np.int64)
    assert np.array_equal(results[0], exp)
    assert np.array_equal(results[1], exp)
    plain.close()
    forced.close()


@pytest.mark.parametrize("sharing", ["cfg3", "cfg4", "cfg5", "k4t2n6", "additive3", "cfg3_generic_prime"])
@pytest.mark.parametrize("mask_kind", ["full", "chacha", "none"])
@pytest.mark.parametrize("offset,ld_pad", [(0, 0), (1, 1)])
def test_mask_share_generate_matches_mask_then_generate(ctx, oracle, torch_cuda, sharing, mask_kind, offset, ld_pad):
    """sda_mask_share_generate_dev == SecretMasker::mask then ShareGenerator::generate per participant (oracle): the fused
    kernel (masks added in shared memory; BASELINE shapes over 2^61-1) and the unfused path behind the same entry point
    (another shape, additive sharing, another prime), Full / ChaCha / no masking, aligned and unaligned sources, vectors
    that end inside a pass and inside a batch, negative secrets"""
    t = torch_cuda
    if sharing == "k4t2n6":
        ss = util.packed_scheme(P61, 4, 2, 6, oracle)
    elif sharing == "additive3":
        ss = LSS.Additive(3, P61)
    elif sharing == "cfg3_generic_prime":
        ss = util.packed_scheme(params.P61_GENERIC, 3, 2, 5, oracle)
    else:
        ss = {"cfg3": params.config3, "cfg4": params.config4, "cfg5": params.config5}[sharing]()
    p = ss.modulus
    k = ss.input_size()
    packed = sharing != "additive3"
    rng = np.random.default_rng(len(sharing) + len(mask_kind))
    for P, dim in [(1, 1), (3, 512 * k * 3 + 2 * k + 1), (2, 1024 * k), (2, 7)]:
        ms = {"full": LMS.Full(p), "chacha": LMS.ChaCha(p, dim, 128), "none": LMS.None_()}[mask_kind]
        ld = dim + (dim & 1) + ld_pad
        secrets = rng.integers(0, p, size=(P, dim), dtype=np.int64)
        secrets[-1, ::3] = rng.integers(-(1 << 62), 1 << 62, size=secrets[-1, ::3].shape, dtype=np.int64)
        flat = np.zeros(offset + P * ld, dtype=np.int64)
        for pi in range(P):
            flat[offset + pi * ld: offset + pi * ld + dim] = secrets[pi]
        mseeds = b"".join(util.seed_bytes(f"mk/{sharing}/{P}/{dim}/{pi}") for pi in range(P))
        sseeds = b"".join(util.seed_bytes(f"sh/{sharing}/{P}/{dim}/{pi}") for pi in range(P))
        mask_len = {"full": dim, "chacha": 4, "none": 0}[mask_kind]
        n_out = ss.output_size() * (ss.batches(dim) if packed else dim)
        d_masks = t.full((P, max(mask_len, 1)), -1, dtype=t.int64, device="cuda")
        d_shares = t.empty((P, n_out), dtype=t.int64, device="cuda")
        ctx.mask_share_generate_dev(ms, ss, dev(t, flat)[offset:], ld, P, dim, mseeds, sseeds, d_masks, d_shares)
        ctx.synchronize()
        fused = mask_kind != "none" and sharing in ("cfg3", "cfg4", "cfg5")
        assert ("mask+packed_share" in ctx.last_kernel()) == fused, ctx.last_kernel()
        got_m, got_s = host(d_masks), host(d_shares)
        for pi in range(P):
            if mask_kind == "none":
                exp_s = util.oracle_generate(oracle, ss, secrets[pi], sseeds[32 * pi:32 * pi + 32], matrix=packed)
                exp_m = np.zeros(0, dtype=np.int64)
            else:
                mo = util.to_oracle_masking(oracle, ms)
                emask, emasked = oracle.mask(mo, secrets[pi], oracle.rng_from_seed_bytes(mseeds[32 * pi:32 * pi + 32]))
                exp_m = np.asarray(emask, dtype=np.int64)
                exp_s = util.oracle_generate(oracle, ss, np.asarray(emasked, dtype=np.int64), sseeds[32 * pi:32 * pi + 32],
                                             matrix=packed)
            assert np.array_equal(got_s[pi], util.canon(oracle, p, np.asarray(exp_s)).reshape(-1)), (sharing, mask_kind, P, dim, pi)
            assert np.array_equal(got_m[pi, :mask_len], exp_m), (sharing, mask_kind, P, dim, pi)


def test_mask_share_generate_redo_after_rejection(oracle, torch_cuda, monkeypatch):
    """the fused mask -> share kernel with its rejection flag forced: the call is redone on the unfused path and returns the
    same masks and shares"""
    import sda_b200
    t = torch_cuda
    ss, dim, P = params.config3(), 5000, 3
    ms = LMS.Full(ss.modulus)
    rng = np.random.default_rng(5)
    secrets = rng.integers(0, ss.modulus, size=(P, dim), dtype=np.int64)
    mseeds = b"".join(util.seed_bytes(f"rm/{pi}") for pi in range(P))
    sseeds = b"".join(util.seed_bytes(f"rs/{pi}") for pi in range(P))
    plain = sda_b200.Context(0)
    monkeypatch.setenv("SDA_B200_DEBUG_FORCE_REJECT", "1")
    forced = sda_b200.Context(0)
    monkeypatch.delenv("SDA_B200_DEBUG_FORCE_REJECT")
    res = []
    for c in (plain, forced):
        d_masks = t.empty((P, dim), dtype=t.int64, device="cuda")
        d_shares = t.empty((P, ss.output_size(), ss.batches(dim)), dtype=t.int64, device="cuda")
        c.mask_share_generate_dev(ms, ss, dev(t, secrets), dim, P, dim, mseeds, sseeds, d_masks, d_shares)
        c.synchronize()
        res.append((host(d_masks), host(d_shares), c.last_kernel()))
    assert "mask+packed_share" in res[0][2] and "mask+packed_share" not in res[1][2]
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    mo = util.to_oracle_masking(oracle, ms)
    emask, emasked = oracle.mask(mo, secrets[1], oracle.rng_from_seed_bytes(mseeds[32:64]))
    assert np.array_equal(res[0][0][1], np.asarray(emask, dtype=np.int64))
    exp = util.oracle_generate(oracle, ss, np.asarray(emasked, dtype=np.int64), sseeds[32:64], matrix=True)
    assert np.array_equal(res[0][1][1], util.canon(oracle, ss.modulus, exp))
    plain.close()
    forced.close()


@pytest.mark.parametrize("mk", [params.config3, params.config5], ids=["cfg3", "cfg5"])
@pytest.mark.parametrize("mask_kind", ["full", "chacha"])
def test_masked_pipeline_reveals_the_sum(ctx, oracle, torch_cuda, mk, mask_kind):
    """participant -> clerks -> recipient at a size the oracle does not finish in seconds, through the fused kernels:
    mask + share in one kernel for 12 participants x 1M secrets, per-clerk sums, reveal from a clerk subset, mask
    combine, unmask == the column sums of the secrets; the same clerk sums from the share-gen -> clerk-sum kernel fed
    with the masked secrets (the two fused paths agree with each other and with the unfused arithmetic)"""
    t = torch_cuda
    ss = mk()
    p, n, k, need = ss.modulus, ss.output_size(), ss.input_size(), ss.c.secret_count + ss.c.privacy_threshold
    P, dim = 12, 1_000_003
    B = ss.batches(dim)
    ms = LMS.Full(p) if mask_kind == "full" else LMS.ChaCha(p, dim, 128)
    mask_len = dim if mask_kind == "full" else 4
    d_sec = t.empty((P, dim), dtype=t.int64, device="cuda")
    ctx.synth_fill_dev(21, p, 0, P * dim, d_sec)
    mseeds = b"".join(util.seed_bytes(f"pl/m/{pi}") for pi in range(P))
    sseeds = b"".join(util.seed_bytes(f"pl/s/{pi}") for pi in range(P))
    d_masks = t.empty((P, mask_len), dtype=t.int64, device="cuda")
    d_shares = t.empty((P, n, B), dtype=t.int64, device="cuda")
    ctx.mask_share_generate_dev(ms, ss, d_sec, dim, P, dim, mseeds, sseeds, d_masks, d_shares)
    assert "mask+packed_share" in ctx.last_kernel()
    d_sum = t.empty((n, B), dtype=t.int64, device="cuda")
    for c in range(n):
        ctx.share_combine_dev(ss, d_shares[:, c, :], n * B, P, B, d_sum[c])
    # the unfused mask, then the fused share-gen -> clerk-sum kernel on the masked secrets: the same clerk sums
    d_masked = t.empty((P, dim), dtype=t.int64, device="cuda")
    d_m2 = t.empty((P, mask_len), dtype=t.int64, device="cuda")
    for pi in range(P):
        ctx.mask_dev(ms, d_sec[pi], dim, mseeds[32 * pi:32 * pi + 32], d_m2[pi], d_masked[pi])
    assert t.equal(d_m2, d_masks)
    d_sum2 = t.empty((n, B), dtype=t.int64, device="cuda")
    ctx.share_generate_combine_dev(ss, d_masked, dim, P, dim, sseeds, d_sum2)
    assert "paired tiles, TMEM-accumulated" in ctx.last_kernel()
    ctx.synchronize()
    assert t.equal(d_sum, d_sum2)
    # recipient: reveal from a subset of the clerks, combine the masks, unmask
    idx = sorted(np.random.default_rng(3).permutation(n)[:need + (need < n)].tolist())
    d_sub = d_sum[idx].contiguous()
    d_rec = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.secret_reconstruct_dev(ss, dim, idx, d_sub, B, len(idx), B, d_rec)
    d_mask_total = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.mask_combine_dev(ms, d_masks, P, mask_len, d_mask_total)
    d_out = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.unmask_dev(ms, d_mask_total, d_rec, dim, d_out)
    d_tot = t.empty(dim, dtype=t.int64, device="cuda")
    ctx.share_combine_dev(LSS.Additive(2, p), d_sec, dim, P, dim, d_tot)
    ctx.synchronize()
    assert t.equal(d_out, d_tot)


def test_deferred_rejection_checks(oracle, torch_cuda, monkeypatch):
    """sda_ctx_set_deferred_checks: *_dev calls that draw randomness return without a stream synchronisation and produce the same
    results; synchronize() reports SDA_OK -- or, with the rejection flag forced, SDA_ERR_REJECTED for the queued calls; the
    host entry points of the same context still check at once"""
    import sda_b200
    from sda_b200 import _lib
    t = torch_cuda
    ss, dim, P = params.config3(), 30_000, 4
    ms = LMS.Full(ss.modulus)
    rng = np.random.default_rng(9)
    secrets = rng.integers(0, ss.modulus, size=(P, dim), dtype=np.int64)
    seeds = b"".join(util.seed_bytes(f"df/{pi}") for pi in range(P))
    n, B = ss.output_size(), ss.batches(dim)

    def run(c):
        d_sh = t.empty((P, n, B), dtype=t.int64, device="cuda")
        d_mask, d_masked = t.empty(dim, dtype=t.int64, device="cuda"), t.empty(dim, dtype=t.int64, device="cuda")
        d_acc = t.empty((n, B), dtype=t.int64, device="cuda")
        d_sh3 = t.empty((P, 3, dim), dtype=t.int64, device="cuda")
        for _ in range(3):                                     # several calls queued back to back
            c.share_generate_dev(ss, dev(t, secrets), dim, P, dim, seeds, d_sh)
            c.mask_dev(ms, dev(t, secrets[0]), dim, seeds[:32], d_mask, d_masked)
            c.share_generate_combine_dev(ss, dev(t, secrets), dim, P, dim, seeds, d_acc)
            c.share_generate_dev(LSS.Additive(3, ss.modulus), dev(t, secrets), dim, P, dim, seeds, d_sh3)
        return d_sh, d_mask, d_masked, d_acc, d_sh3

    plain = sda_b200.Context(0)
    ref = run(plain)
    plain.synchronize()
    deferred = sda_b200.Context(0)
    deferred.set_deferred_checks(True)
    got = run(deferred)
    deferred.synchronize()                                     # SDA_OK: nothing was rejected
    for a, b in zip(ref, got):
        assert t.equal(a, b)
    exp = util.oracle_generate(oracle, ss, secrets[2], seeds[64:96], matrix=True)
    assert np.array_equal(host(got[0][2]), util.canon(oracle, ss.modulus, exp))
    # a host entry point of a deferred context checks at once and is exact
    assert np.array_equal(deferred.share_generate(ss, secrets[1], seeds[32:64]), host(ref[0][1]))
    deferred.set_deferred_checks(False)
    # forced flag: the deferred calls report it at the next synchronisation, not before
    monkeypatch.setenv("SDA_B200_DEBUG_FORCE_REJECT", "1")
    forced = sda_b200.Context(0)
    monkeypatch.delenv("SDA_B200_DEBUG_FORCE_REJECT")
    forced.set_deferred_checks(True)
    d_sh = t.empty((P, n, B), dtype=t.int64, device="cuda")
    d_masks = t.empty((P, dim), dtype=t.int64, device="cuda")
    forced.mask_share_generate_dev(ms, ss, dev(t, secrets), dim, P, dim, seeds, seeds, d_masks, d_sh)    # returns SDA_OK
    with pytest.raises(sda_b200.SdaClientError, match="deferred call #0") as ei:
        forced.synchronize()
    assert ei.value.code == _lib.SDA_ERR_REJECTED
    forced.synchronize()                                       # reported once; nothing pending any more
    for c in (plain, deferred, forced):
        c.close()
