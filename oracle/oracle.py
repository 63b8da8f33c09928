"""ctypes loader for the CPU oracle (oracle/libsda_oracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the sda_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsda_oracle.so")

ADDITIVE, PACKED_SHAMIR = 0, 1
MASK_NONE, MASK_FULL, MASK_CHACHA = 0, 1, 2


class SharingScheme(C.Structure):
    _fields_ = [("kind", C.c_int32), ("share_count", C.c_uint64), ("secret_count", C.c_uint64),
                ("privacy_threshold", C.c_uint64), ("modulus", C.c_int64),
                ("omega_secrets", C.c_int64), ("omega_shares", C.c_int64)]


class MaskingScheme(C.Structure):
    _fields_ = [("kind", C.c_int32), ("modulus", C.c_int64), ("dimension", C.c_uint64),
                ("seed_bitsize", C.c_uint64)]


class Rng(C.Structure):
    _fields_ = [("kind", C.c_int), ("rounds", C.c_int), ("state", C.c_uint32 * 16),
                ("buf", C.c_uint32 * 16), ("idx", C.c_int), ("draws", C.c_uint64),
                ("rejections", C.c_uint64)]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "sda_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.sdao_last_error.restype = C.c_char_p
        L.sdao_rng_next_u32.restype = C.c_uint32
        L.sdao_rng_next_u64.restype = C.c_uint64
        L.sdao_gen_range.restype = C.c_int64
        L.sdao_gen_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.sdao_mod_pow.restype = C.c_int64
        L.sdao_mod_pow.argtypes = [C.c_int64, C.c_uint64, C.c_int64]
        L.sdao_mod_inverse.restype = C.c_int64
        L.sdao_mod_inverse.argtypes = [C.c_int64, C.c_int64]
        L.sdao_find_root_of_order.restype = C.c_int64
        L.sdao_find_root_of_order.argtypes = [C.c_int64, C.c_uint64]
        for f in ("sdao_input_size", "sdao_output_size", "sdao_privacy_threshold",
                  "sdao_reconstruction_threshold", "sdao_varint_encode", "sdao_varint_decode"):
            getattr(L, f).restype = C.c_size_t
        _lib = L
    return _lib


class OracleError(Exception):
    pass


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _check(rc):
    if rc:
        raise OracleError(lib().sdao_last_error().decode())


def additive(share_count, modulus):
    return SharingScheme(ADDITIVE, share_count, 0, 0, modulus, 0, 0)


def packed_shamir(secret_count, share_count, privacy_threshold, prime_modulus, omega_secrets, omega_shares):
    return SharingScheme(PACKED_SHAMIR, share_count, secret_count, privacy_threshold, prime_modulus,
                         omega_secrets, omega_shares)


def rng_from_seed(seed_words, rounds=20):
    r = Rng()
    s = np.ascontiguousarray(np.asarray(seed_words, dtype=np.uint32))
    lib().sdao_rng_from_seed_rounds(C.byref(r), _p(s), C.c_size_t(len(s)), C.c_int(rounds))
    return r


def rng_from_seed_bytes(seed32, rounds=20):
    """32 bytes of caller entropy -> 8 little-endian key words (the C-ABI's rng_seed)."""
    return rng_from_seed(np.frombuffer(bytes(seed32), dtype="<u4"), rounds)


def rng_os():
    r = Rng()
    lib().sdao_rng_os(C.byref(r))
    return r


def chacha_block(state16, rounds=20):
    st = np.ascontiguousarray(np.asarray(state16, dtype=np.uint32))
    out = np.empty(16, dtype=np.uint32)
    lib().sdao_chacha_block(_p(st), C.c_int(rounds), _p(out))
    return out


def next_u32(r):
    return lib().sdao_rng_next_u32(C.byref(r))


def next_u64(r):
    return lib().sdao_rng_next_u64(C.byref(r))


def gen_range(r, low, high):
    return lib().sdao_gen_range(C.byref(r), low, high)


def input_size(s):
    return lib().sdao_input_size(C.byref(s))


def output_size(s):
    return lib().sdao_output_size(C.byref(s))


def privacy_threshold(s):
    return lib().sdao_privacy_threshold(C.byref(s))


def reconstruction_threshold(s):
    return lib().sdao_reconstruction_threshold(C.byref(s))


def tss_share_with_randomness(s, secrets, randomness, force_general=False, want_poly=False):
    sec, rnd = _i64(secrets), _i64(randomness)
    out = np.empty(s.share_count, dtype=np.int64)
    poly = np.zeros(s.secret_count + s.privacy_threshold + 1, dtype=np.int64)
    _check(lib().sdao_tss_share_with_randomness(C.byref(s), _p(sec), _p(rnd), C.c_int(int(force_general)),
                                                _p(poly), _p(out)))
    return (poly, out) if want_poly else out


def tss_reconstruct(s, indices, shares):
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
    sh = _i64(shares)
    out = np.empty(s.secret_count, dtype=np.int64)
    _check(lib().sdao_tss_reconstruct(C.byref(s), _p(idx), _p(sh), C.c_size_t(len(idx)), _p(out)))
    return out


def share_generate(s, secrets, rng, matrix=False):
    sec = _i64(secrets)
    k, n = input_size(s), output_size(s)
    B = (len(sec) + k - 1) // k
    out = np.empty((n, B), dtype=np.int64)
    f = lib().sdao_share_generate_matrix if matrix else lib().sdao_share_generate
    _check(f(C.byref(s), _p(sec), C.c_size_t(len(sec)), C.byref(rng), _p(out)))
    return out


def share_combine(modulus, shares, ld=None):
    sh = _i64(shares)
    if sh.ndim == 1:
        sh = sh.reshape(0, 0) if sh.size == 0 else sh.reshape(1, -1)
    P, L = sh.shape
    out = np.empty(L, dtype=np.int64)
    _check(lib().sdao_share_combine(C.c_int64(modulus), _p(sh), C.c_size_t(P), C.c_size_t(L),
                                    C.c_size_t(L if ld is None else ld), _p(out)))
    return out


def secret_reconstruct(s, dimension, indices, shares):
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
    sh = _i64(shares)
    m = len(idx)
    Bc = sh.shape[1] if sh.ndim == 2 else 0
    out = np.empty(max(dimension, Bc) + 8, dtype=np.int64)
    n = C.c_size_t(0)
    _check(lib().sdao_secret_reconstruct(C.byref(s), C.c_size_t(dimension), _p(idx), _p(sh), C.c_size_t(m),
                                         C.c_size_t(Bc), _p(out), C.byref(n)))
    return out[:n.value].copy()


def mask(ms, secrets, rng):
    sec = _i64(secrets)
    dim = len(sec)
    mk = np.empty(max(dim, 64), dtype=np.int64)
    masked = np.empty(dim, dtype=np.int64)
    n = C.c_size_t(0)
    _check(lib().sdao_mask(C.byref(ms), _p(sec), C.c_size_t(dim), C.byref(rng), _p(mk), C.byref(n), _p(masked)))
    return mk[:n.value].copy(), masked


def mask_combine(ms, masks):
    mk = _i64(masks)
    if mk.ndim == 1:
        mk = mk.reshape(len(masks), -1) if len(masks) else mk.reshape(0, 0)
    P, ml = mk.shape
    out = np.empty(max(ml, int(ms.dimension), 1), dtype=np.int64)
    n = C.c_size_t(0)
    _check(lib().sdao_mask_combine(C.byref(ms), _p(mk), C.c_size_t(P), C.c_size_t(ml), _p(out), C.byref(n)))
    return out[:n.value].copy()


def unmask(ms, mask_, masked):
    mk, md = _i64(mask_), _i64(masked)
    out = np.empty(len(md), dtype=np.int64)
    _check(lib().sdao_unmask(C.byref(ms), _p(mk), C.c_size_t(len(mk)), _p(md), C.c_size_t(len(md)), _p(out)))
    return out


def positive(modulus, values):
    v = _i64(values).copy()
    lib().sdao_positive(C.c_int64(modulus), _p(v), C.c_size_t(v.size))
    return v


def canonical(modulus, values):
    v = _i64(values).copy()
    lib().sdao_canonical(C.c_int64(modulus), _p(v), C.c_size_t(v.size))
    return v


def varint_encode(values):
    v = _i64(values)
    out = np.empty(10 * len(v) + 1, dtype=np.uint8)
    n = lib().sdao_varint_encode(_p(v), C.c_size_t(len(v)), _p(out))
    return out[:n].copy()


def varint_decode(buf, max_out=None):
    b = np.ascontiguousarray(np.asarray(buf, dtype=np.uint8))
    max_out = len(b) if max_out is None else max_out
    out = np.empty(max(max_out, 1), dtype=np.int64)
    n = lib().sdao_varint_decode(_p(b), C.c_size_t(len(b)), _p(out), C.c_size_t(max_out))
    if n == C.c_size_t(-1).value:
        raise OracleError("truncated varint buffer")
    return out[:n].copy()


def fixed_encode(x, frac_bits, modulus):
    v = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    out = np.empty(len(v), dtype=np.int64)
    lib().sdao_fixed_encode(_p(v), C.c_size_t(len(v)), C.c_int(frac_bits), C.c_int64(modulus), _p(out))
    return out


def fixed_decode(values, frac_bits, modulus, divisor=1):
    v = np.ascontiguousarray(np.asarray(values, dtype=np.int64))
    out = np.empty(len(v), dtype=np.float32)
    lib().sdao_fixed_decode(_p(v), C.c_size_t(len(v)), C.c_int(frac_bits), C.c_int64(modulus), C.c_uint64(divisor), _p(out))
    return out


def synth_fill(stream, modulus, start, count):
    out = np.empty(count, dtype=np.int64)
    lib().sdao_synth_fill(C.c_uint32(stream), C.c_int64(modulus), C.c_uint64(start), C.c_size_t(count), _p(out))
    return out


def find_root_of_order(p, q):
    return lib().sdao_find_root_of_order(p, q)


def snapshot_transpose(blobs):
    """server/src/stores.rs:86-101 `iter_snapshot_clerk_jobs_data`: blobs[p][c] (bytes) of P participations x n clerks ->
    one list per clerk holding that clerk's blob of every participation, in participation order (`shares[ix].push(share.1)`)."""
    n = len(blobs[0]) if blobs else 0
    shares = [[] for _ in range(n)]
    for participation in blobs:
        for ix, share in enumerate(participation):
            shares[ix].append(share)
    return shares
