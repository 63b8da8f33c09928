"""Integer models of the two Mersenne-61 share-generation kernels' inner arithmetic.

csrc/packed_m61.cu (centred 31-bit limbs, two 64-bit accumulators, renormalisation every 4 terms,
offset-cancelling row constant, compare-free fold) and csrc/packed_tc.cu (byte-limb GEMM + compose)
are exact only because every intermediate stays inside 64 (or 32) bits.  These models restate the
kernels' word-level operations with Python integers, assert every bound the kernels rely on, and
compare against plain `sum(c * v) % p` on random and worst-case operands.  No GPU involved.
"""
import random

P = (1 << 61) - 1
M64 = (1 << 64) - 1
OX1, OA2, OX2 = 1 << 63, -(1 << 60), 6 << 60
DELTA = (1 << 30) + (1 << 60)


def u64(x):
    assert 0 <= x <= M64, hex(x)
    return x


def s64(x):
    assert -(1 << 63) <= x < (1 << 63), hex(x)
    return x


def u32(x):
    assert 0 <= x < (1 << 32), hex(x)
    return x


# ---- packed_m61.cu ------------------------------------------------------------------------------
def centre_matrix(c):
    c %= P
    if c > P // 2:
        c -= P
    m1 = (c + (1 << 30)) >> 31
    m0 = c - (m1 << 31)
    assert -(1 << 30) <= m0 < (1 << 30) and abs(m1) <= (1 << 29)
    return m0, m1, 2 * m1


def chunks(w):
    out, i, first = [], 0, True
    while i < w:
        n = 5 if first else 4
        out.append((i, min(w, i + n)))
        i += n
        first = False
    return out


def row_const(crow_centred):
    nb = len(chunks(len(crow_centred))) - 1
    s = sum(c * DELTA for c in crow_centred) + 1 - nb * OA2 - (OX1 + nb * OX2) * (1 << 31)
    return (s % P) * pow(2, 30, P) % P


def dot_m61(limbs, kp, xs):
    a, x = 0, OX1 + kp
    for ci, (lo, hi) in enumerate(chunks(len(xs))):
        if ci:
            a = s64((a & P) + (a >> 61) + OA2)
            x = u64(((x & P) | OX2) + (x >> 61))
        for i in range(lo, hi):
            m0, m1, d1 = limbs[i]
            x0, x1 = xs[i]
            assert -(1 << 31) <= x0 < (1 << 31) and -(1 << 31) <= x1 < (1 << 31)
            a = s64(a + m0 * x0 + d1 * x1)
            x = u64(x + m0 * x1 + m1 * x0)
    a_lo, a_hi = a & 0xffffffff, (a >> 32) & 0xffffffff
    t = u64((a_lo | ((a_hi & 0x1fffffff) << 32)) + (x & 0xffffffff) * (1 << 31))
    t = u64(t + (x >> 32) * 4)
    t = u64(t + (a >> 61))                       # arithmetic shift: signed top three bits
    assert t >= 1
    s = u64(t + (t >> 61))
    return ((t + (s >> 61) - 1) & M64) & P


def limbs_secret(v):
    assert 0 <= v < (1 << 61)
    lo, hi = v & 0xffffffff, v >> 32
    return (lo & 0x7fffffff) - (1 << 30), (((hi << 1) | (lo >> 31)) & 0xffffffff) - (1 << 29)


def limbs_draw(v):
    w0, w1 = v >> 32, v & 0xffffffff
    l1 = ((w0 << 1) | (w1 >> 31)) & 0x3fffffff
    return (w1 & 0x7fffffff) + 2 * (w0 >> 29) - (1 << 30), l1 - (1 << 29), l1 == 0x3fffffff


EXTREME_C = [1, P - 1, (1 << 60) - 1, 1 << 60, (1 << 60) + 1, P // 2, P // 2 + 1, 1 << 30, (1 << 31) - 1, P - (1 << 30)]
EXTREME_V = [0, 1, P - 1, P, (1 << 61) - 1, (1 << 31) - 1, 1 << 31, 1 << 60]


def run_m61(w, t, trials, extreme, rng):
    k = w - t
    for _ in range(trials):
        crow = [rng.choice(EXTREME_C + [rng.randrange(P)]) if extreme else rng.randrange(P) for _ in range(w)]
        limbs = [centre_matrix(c) for c in crow]
        kp = row_const([c if c <= P // 2 else c - P for c in crow])
        xs, vals = [], []
        for _ in range(k):
            v = rng.choice(EXTREME_V) if extreme and rng.random() < .7 else rng.randrange(1 << 61)
            xs.append(limbs_secret(v))
            vals.append(v)
        for _ in range(t):
            while True:
                v = rng.choice([0, M64 - 17, (7 << 61) + (1 << 61) - 20, rng.randrange(1 << 64),
                                (rng.randrange(8) << 61) | ((1 << 61) - 1 - rng.randrange(1 << 31))]) \
                    if extreme else rng.randrange(1 << 64)
                x0, x1, unusual = limbs_draw(v)
                if not unusual:                   # the kernel raises `flag` for these and the host redoes the call
                    break
            assert (v & P) + 2 * (v >> 61) == v % (P - 1)
            xs.append((x0, x1))
            vals.append(v % (P - 1))
        assert dot_m61(limbs, kp, xs) == sum(c * v for c, v in zip(crow, vals)) % P


def test_packed_m61_inner_arithmetic_is_exact():
    rng = random.Random(1)
    for w, t in [(5, 2), (7, 4), (9, 4), (16, 8), (1, 0), (2, 1), (6, 3), (10, 5), (13, 6)]:
        run_m61(w, t, 3000, False, rng)
        run_m61(w, t, 3000, True, rng)


def test_packed_m61_sign_aligned_worst_case():
    for w in (5, 9, 16):
        for crow in ([(1 << 60) - 1] * w, [P - ((1 << 60) - 1)] * w, [(-(1 << 30)) % P] * w):
            limbs = [centre_matrix(c) for c in crow]
            kp = row_const([c if c <= P // 2 else c - P for c in crow])
            for v in (0, (1 << 31) - 1, (1 << 60), (1 << 61) - 1):
                assert dot_m61(limbs, kp, [limbs_secret(v)] * w) == sum(c * v for c in crow) % P


# ---- packed_tc.cu -------------------------------------------------------------------------------
def compose(d):
    """limb sums d[s] < 2^23 -> canonical sum_s d[s] 2^{8s} mod p, word for word as the kernel"""
    assert all(0 <= x < (1 << 23) for x in d)
    e = [u32(d[2 * i] + (d[2 * i + 1] << 8)) for i in range(4)]
    lo64, hi64 = e[0] + e[1] * 65536, e[2] + e[3] * 65536
    l_lo, l_hi, h_lo, h_hi = lo64 & 0xffffffff, lo64 >> 32, hi64 & 0xffffffff, hi64 >> 32
    small = u32((h_lo >> 29) + h_hi * 8 + 1)
    t = u64((l_lo | (u32(l_hi + (h_lo & 0x1fffffff)) << 32)) + small)
    assert 1 <= t < (1 << 62)
    qm1 = (t >> 61) - 1
    return ((t + qm1) & M64) & P


def tc_share(crow, xs):
    """one row of the byte-limb GEMM: A = bytes of x, B = bytes of (c 2^{8c'} mod p)"""
    d = [0] * 8
    for c, x in zip(crow, xs):
        assert 0 <= x <= M64
        for byte in range(8):
            cst = c * pow(2, 8 * byte, P) % P
            xb = (x >> (8 * byte)) & 0xff
            for s in range(8):
                d[s] += xb * ((cst >> (8 * s)) & 0xff)
    return compose(d)


def test_packed_tc_byte_limb_gemm_is_exact():
    rng = random.Random(2)
    for w in (5, 7, 9, 12):                      # 12 values x 8 bytes = 96 bytes of K: the widest tile
        for trial in range(300):
            crow = [rng.choice(EXTREME_C + [rng.randrange(P)]) for _ in range(w)]
            xs = [rng.choice([0, M64, P, P - 1, (1 << 61) + 13, rng.randrange(1 << 64)]) for _ in range(w)]
            assert tc_share(crow, xs) == sum(c * x for c, x in zip(crow, xs)) % P
    assert tc_share([P - 1] * 12, [M64] * 12) == (P - 1) * M64 * 12 % P      # every limb sum at its maximum


def test_draw_reduction_matches_gen_range_outside_the_flagged_band():
    rng = random.Random(3)
    for _ in range(20000):
        v = rng.choice([rng.randrange(1 << 64), (rng.randrange(8) << 61) | ((1 << 61) - 33 - rng.randrange(1 << 20))])
        hi, w1 = (v >> 32) & 0x1fffffff, v & 0xffffffff
        bad = hi == 0x1fffffff and w1 >= 0xffffffe0
        if not bad:
            assert v < ((1 << 64) - 16)                                     # accepted by Range::new(0, p - 1)
            assert (v & P) + 2 * (v >> 61) == v % (P - 1)


# ---- packed_tc.cu: the persistent CTAs' walk over (participant, pass) units ------------------------
def walk_units(grid, participants, unit_begin, units_per_p):
    """packed_share_tc_kernel's unit bookkeeping: CTA x starts at unit x and strides by the grid; a launch
    covers passes unit_begin .. unit_begin + units_per_p - 1 of every participant (the host entry point's slices)."""
    unit_end = unit_begin + units_per_p
    units_total = units_per_p * participants
    seen = []
    for cta in range(min(grid, units_total)):
        p, u = cta // units_per_p, unit_begin + cta % units_per_p
        unit = cta
        while unit < units_total:
            seen.append((p, u))
            pn, un = p, u + grid
            while un >= unit_end:
                un -= units_per_p
                pn += 1
            p, u = pn, un
            unit += grid
    return seen


def test_packed_tc_unit_walk_covers_every_slice_exactly_once():
    rng = random.Random(5)
    for _ in range(300):
        grid = rng.choice([1, 2, 3, 7, 148, 592])
        participants = rng.randint(1, 9)
        units_per_p = rng.randint(1, 40)
        unit_begin = rng.choice([0, 1, 5, 813])
        seen = walk_units(grid, participants, unit_begin, units_per_p)
        want = [(p, u) for p in range(participants) for u in range(unit_begin, unit_begin + units_per_p)]
        assert sorted(seen) == want


def test_host_slices_tile_the_vector():
    """share_generate_sliced (csrc/api.cu): 8 slices, each a multiple of the kernel's pass size, cover [0, B)."""
    for dim in (524_288, 600_001, 10_000_000, 25_000_000, 3 * 512 * 8 * 5):
        k, unit, slices = 3, 512, 8
        B = (dim + k - 1) // k
        per = ((B + slices - 1) // slices + unit - 1) // unit * unit
        covered, b0 = 0, 0
        while b0 < B:
            nb = min(per, B - b0)
            assert b0 % unit == 0
            covered += nb
            b0 += per
        assert covered == B and (B + per - 1) // per <= slices


# ---- packed_tc2.cu: uneven limb plan, one wide term, paired tiles -----------------------------------
LOW29 = (1 << 29) - 1


def limb_plan2(kt):
    """Shape2::W5 / limb_plan(): widths (8,8,8,8,8,W5,8,13-W5), W5 the widest with e2 = d4 + 256 d5 < 2^29"""
    w5 = next(w for w in range(8, 4, -1) if 8 * kt * 255 * (255 + ((1 << w) - 1) * 256) < (1 << 29))
    return [8, 8, 8, 8, 8, w5, 8, 13 - w5], [0, 8, 16, 24, 32, 40, 40 + w5, 48 + w5]


def compose2(d, w, kt):
    """compose2<W5> of packed_tc2.cu, word for word; every intermediate asserted to fit its register"""
    w5 = w[5]
    for s in range(8):
        assert 0 <= d[s] <= 8 * kt * 255 * ((1 << w[s]) - 1)
    e = [u32(d[2 * i] + (d[2 * i + 1] << 8)) for i in range(4)]
    x = u64((e[0] | (e[2] << 32)) + e[1] * 65536)                       # mad.wide.u32 e1, 65536, {e0, e2}
    sh3 = 8 + w5
    m3 = ((e[3] << sh3) & 0xffffffff) & (LOW29 & ~((1 << sh3) - 1))
    assert m3 == (e[3] & ((1 << (21 - w5)) - 1)) << sh3
    s3 = u32((e[3] >> (21 - w5)) + 1)
    t = u64(x + (s3 | (m3 << 32)))
    assert 1 <= t < (1 << 62)
    qm1 = (t >> 61) - 1
    r = (t + qm1) & M64
    return r & ((LOW29 << 32) | 0xffffffff)


def tc2_share(crow, xs):
    kt = len(crow)
    w, pos = limb_plan2(kt)
    assert pos[7] + w[7] == 61
    d = [0] * 8
    for c, x in zip(crow, xs):
        assert 0 <= x <= M64
        for byte in range(8):
            cst = c * pow(2, 8 * byte, P) % P
            xb = (x >> (8 * byte)) & 0xff
            for s in range(8):
                d[s] += xb * ((cst >> pos[s]) & ((1 << w[s]) - 1))
    return compose2(d, w, kt)


def test_packed_tc2_limb_plan_and_compose_are_exact():
    rng = random.Random(7)
    for kt in range(2, 17):
        for trial in range(200):
            crow = [rng.choice(EXTREME_C + [rng.randrange(P)]) for _ in range(kt)]
            xs = [rng.choice([0, M64, P, P - 1, (1 << 61) + 13, 1 << 63, rng.randrange(1 << 64)]) for _ in range(kt)]
            assert tc2_share(crow, xs) == sum(c * x for c, x in zip(crow, xs)) % P
        w, _ = limb_plan2(kt)
        compose2([8 * kt * 255 * ((1 << w[s]) - 1) for s in range(8)], w, kt)      # every limb sum at its maximum
        assert tc2_share([P - 1] * kt, [M64] * kt) == (P - 1) * M64 * kt % P


def test_packed_tc2_draw_operand_is_congruent_to_gen_range():
    """reduce_draw2: X = v + (v >> 61) as a u64 is congruent mod p to gen_range's v mod (p - 1) outside the flagged band"""
    rng = random.Random(8)
    for _ in range(20000):
        v = rng.choice([rng.randrange(1 << 64), (rng.randrange(8) << 61) | ((1 << 61) - 33 - rng.randrange(1 << 20)),
                        (7 << 61) | rng.randrange(1 << 61)])
        hi, w1 = (v >> 32) & LOW29, v & 0xffffffff
        if hi == LOW29 and w1 >= 0xffffffe0:
            continue                                                       # the kernel raises `flag`; the host redoes the call
        x = v + (v >> 61)
        assert x <= M64 and v < (1 << 64) - 16
        assert x % P == (v % (P - 1)) % P


def walk_units2(grid, participants, unit_begin, units_per_p):
    """packed_share_tc2_kernel's 32-bit unit bookkeeping: one conditional subtraction per step"""
    unit_end, units_total = unit_begin + units_per_p, units_per_p * participants
    step_p, step_u = grid // units_per_p, grid % units_per_p
    seen = []
    for cta in range(min(grid, units_total)):
        p, u, unit = cta // units_per_p, unit_begin + cta % units_per_p, cta
        while unit < units_total:
            seen.append((p, u))
            p, u = p + step_p, u + step_u
            if u >= unit_end:
                u -= units_per_p
                p += 1
            unit += grid
    return seen


def test_packed_tc2_unit_walk_covers_every_slice_exactly_once():
    rng = random.Random(9)
    for _ in range(300):
        grid = rng.choice([1, 2, 3, 7, 148, 592])
        participants, units_per_p, unit_begin = rng.randint(1, 9), rng.randint(1, 40), rng.choice([0, 1, 5, 813])
        want = [(p, u) for p in range(participants) for u in range(unit_begin, unit_begin + units_per_p)]
        assert sorted(walk_units2(grid, participants, unit_begin, units_per_p)) == want


def test_packed_tc2_pair_layout_maps_every_draw_and_secret_once():
    """stage_draws2 / the staging loop: (tile, row, chunk) of every draw chunk and secret word of a pass"""
    for k, t in ((3, 2), (5, 4), (3, 4), (4, 2), (2, 8)):
        dc, sc = t // 2, (k + 1) // 2
        g = 4 // __import__("math").gcd(t, 4)
        pairs, nb = g, t // __import__("math").gcd(t, 4)
        assert pairs * 256 * t == 8 * 128 * nb                            # the pass's draws are whole keystream blocks
        seen = {}
        for slot in range(128 * nb):
            gc0 = slot * 4
            beta0, c0 = gc0 // dc, gc0 % dc
            for cb in range(4):
                gc = gc0 + cb
                beta, c = gc // dc, gc % dc                               # what the chunk is: batch beta, chunk c
                tile, row = (beta >> 8) * 2 + (beta & 1), (beta & 255) >> 1
                # what the kernel computes from chunk 0 of the block plus a compile-time delta
                tile0, row0 = (beta0 >> 8) * 2 + (beta0 & 1), (beta0 & 255) >> 1
                if dc == 1:
                    got = (tile0 + (cb & 1), row0 + (cb >> 1), c0)
                elif dc == 2:
                    got = (tile0 + (cb >> 1), row0, c0 + (cb & 1))
                else:
                    got = (tile0, row0, c0 + cb)
                assert got == (tile, row, c)
                assert (row0 & 7) + (cb >> 1 if dc == 1 else 0) < 8       # the delta stays inside the 8-row group
                seen[(tile, row, c)] = seen.get((tile, row, c), 0) + 1
        assert len(seen) == pairs * 2 * 128 * dc and set(seen.values()) == {1}
        # secrets: thread r of pair q reads words 0..k-1 at element (q*256 + 2r)*k; E row r gets words 0..sc-1, O row r words k-sc..k-1
        for r in range(128):
            base = 2 * r * k
            e_elems = [base + 2 * c + v for c in range(sc) for v in range(2)]
            o_elems = [base + 2 * (k - sc + c) + v for c in range(sc) for v in range(2)]
            for slot_i, el in enumerate(e_elems):
                idx = slot_i                                             # image E: secret index = 2c + v
                assert (el == base + idx) and (idx < k or el >= base + k)
            for slot_i, el in enumerate(o_elems):
                idx = slot_i - (k & 1)                                    # image O: half a chunk late for odd k
                assert el == base + k + idx


# ---- packed_tc2f.cu: limb sums accumulated over participants, composed once per drain ------------------------------
def compose_wide2(d, w5):
    """compose_wide2<W5> of packed_tc2f.cu, word for word"""
    lo = d[0] + (d[1] << 8) + (d[2] << 16) + (d[3] << 24)
    hi = d[4] + (d[5] << 8) + (d[6] << (8 + w5)) + (d[7] << (16 + w5))
    assert lo < (1 << 56) and hi < (1 << 56)
    v = lo + ((hi & LOW29) << 32) + (hi >> 29)
    assert v <= M64
    v = (v & P) + (v >> 61)
    return v - P if v >= P else v


def test_fused_accumulation_fits_s32_and_composes_exactly():
    """MAX_ACCUM2 = 256 participants of limb sums fit the s32 accumulators for every k + t <= 16, and the wide compose of
    the accumulated sums equals the sum of the participants' shares mod p"""
    rng = random.Random(11)
    for kt in range(2, 17):
        w, pos = limb_plan2(kt)
        for s in range(8):
            assert 256 * 8 * kt * 255 * ((1 << w[s]) - 1) < (1 << 31)
        crow = [rng.randrange(P) for _ in range(kt)]
        acc, total = [0] * 8, 0
        for _ in range(256):
            xs = [rng.choice([M64, rng.randrange(1 << 64)]) for _ in range(kt)]
            total += sum(c * x for c, x in zip(crow, xs))
            for c, x in zip(crow, xs):
                for byte in range(8):
                    cst = c * pow(2, 8 * byte, P) % P
                    xb = (x >> (8 * byte)) & 0xff
                    for s in range(8):
                        acc[s] += xb * ((cst >> pos[s]) & ((1 << w[s]) - 1))
        assert max(acc) < (1 << 31)
        assert compose_wide2(acc, w[5]) == total % P
        assert compose_wide2([256 * 8 * kt * 255 * ((1 << w[s]) - 1) for s in range(8)], w[5]) < P     # every limb at its maximum


# ---- packed_tc2m.cu / sharegen.cu: masks over 2^61 - 1 without a reduction -------------------------------------------
def test_masked_operand_and_lazy_mask_sum_are_congruent():
    """add_masks2: secret + ((v & p) + (v >> 61)) is a u64 congruent to the masked secret, for any i64 secret (negative ones
    canonicalised first) and any accepted word v; chacha_mask_combine_kernel: the same terms summed 7 at a time before a fold
    stay below 2^64 and reduce to the sum of gen_range's values.  Words in the flagged band are the host's business."""
    rng = random.Random(12)
    acc, exact, since = 0, 0, 0
    for _ in range(20000):
        v = rng.choice([rng.randrange(1 << 64), (rng.randrange(8) << 61) | ((1 << 61) - 9 - rng.randrange(1 << 16)), (1 << 64) - 9 - rng.randrange(100)])
        hi, w1 = (v >> 32) & LOW29, v & 0xffffffff
        if hi == LOW29 and w1 >= 0xfffffff8:
            continue                                        # low 61 bits within 8 of 2^61: the kernels raise `flag`
        assert v < (1 << 64) - 8                            # accepted by gen_range(0, p): zone = 2^64 - 8
        sd = (v & P) + (v >> 61)
        assert sd < P and sd == v % P                       # gen_range's value, no compare needed
        secret = rng.choice([rng.randrange(P), -rng.randrange(1, 1 << 63), (1 << 63) - 1, P, 0])
        x = secret if secret >= 0 else (secret % P)         # canon_negative
        operand = x + sd
        assert operand <= M64 and operand % P == (secret + v % P) % P
        # the mask kernel's fast path: secret <= p, one conditional subtraction
        if 0 <= secret <= P:
            t = secret + sd
            assert (t - P if t >= P else t) == (secret + sd) % P
        acc += sd
        exact += v % P
        since += 1
        assert acc <= M64
        if since == 7:
            acc, since = (acc & P) + (acc >> 61), 0
    a = (acc & P) + (acc >> 61)
    a = (a & P) + (a >> 61)
    assert (a - P if a >= P else a) == exact % P


def test_suspect_pairs_never_miss_a_flagged_word():
    """suspect_of_block with SDA_TC2_SUSPECT_PAIRS: (w0_a | w0_b) & LOW29 == LOW29 is necessary for either word to have its
    bits 0..28 all ones, so the running maximum over pairs reaches LOW29 whenever a single-word test would"""
    rng = random.Random(13)
    for _ in range(20000):
        ws = [rng.choice([rng.randrange(1 << 32), LOW29 | (rng.randrange(8) << 29), 0xffffffff]) for _ in range(8)]
        single = max(w & LOW29 for w in ws)
        pairs = max((ws[d] | ws[d + 1]) & LOW29 for d in range(0, 8, 2))
        assert pairs >= single and (single == LOW29) <= (pairs == LOW29)


def test_float_encode_fast_path_matches_the_double_definition():
    """mask_kernel<FLOAT_IN> over 2^61 - 1: rint(float(x) * float(2^f)) computed in float equals rint(double(x) * 2^f) (a
    power-of-two scale is exact in float short of overflow, where both saturate), and q + (p & (q >> 63)) is the canonical
    residue whenever |q| < 2^60"""
    import numpy as np
    rng = np.random.default_rng(14)
    xs = np.concatenate([rng.standard_normal(20000).astype(np.float32) * np.float32(1000.0),
                         np.array([0.0, -0.0, 1e-45, -1e-45, 2.0 ** 43, -2.0 ** 43, 2.0 ** 44 - 2.0 ** 21, 1.5 * 2.0 ** -16, 2.5 * 2.0 ** -16,
                                   -0.5 * 2.0 ** -16, 65504.0, 1e20, -1e20], dtype=np.float32)])
    for frac in (0, 16, 24, 40):
        with np.errstate(over="ignore"):
            qf = np.rint(xs * np.float32(2.0 ** frac))                       # float product, as the kernel forms it
        qd = np.rint(xs.astype(np.float64) * 2.0 ** frac)
        finite = np.isfinite(qf)
        assert np.array_equal(qf[finite].astype(np.float64), qd[finite])     # the product is exact in float
        assert (np.abs(qd[~finite]) >= 2.0 ** 63).all()                      # float overflow only where the double saturates too
        small = np.abs(qd) < 2.0 ** 60
        q = qd[small].astype(np.int64)
        fast = (q.astype(object) + np.where(q < 0, P, 0).astype(object))
        assert all(int(f) == int(v) % P for f, v in zip(fast[:2000], q[:2000]))
