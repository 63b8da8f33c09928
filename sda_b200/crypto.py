"""Host-side mirror of `sda_client::crypto` over the C ABI (include/sda_b200.h).

Same names, argument meaning and error behaviour as the reference's trait surface
(client/src/crypto/sharing/mod.rs:10-33, client/src/crypto/masking/mod.rs:9-31) so that tests
read like the reference's own (`integration-tests/tests/full_loop.rs`):

    crypto = CryptoModule()
    shares = crypto.new_share_generator(scheme).generate(secrets)          # Vec<Vec<Share>>
    summed = crypto.new_share_combiner(scheme).combine(rows)               # Vec<Share>
    output = crypto.new_secret_reconstructor(scheme, dim).reconstruct(indexed_shares)

Every method is a thin call into libsda_b200.so (hand-written sm_100a kernels).  There is no
Python/numpy implementation of any of them here; without the library or a GPU they raise.
Where the reference returns `Err(msg)` or panics, `SdaClientError` carries the same message.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import sda_masking_scheme, sda_sharing_scheme


class SdaClientError(Exception):
    """`SdaClientResult::Err` (client/src/errors.rs) or a reference panic; `.code` is the C error class."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code
        self.message = message


# ---- protocol/src/crypto.rs:79-155 -------------------------------------------------------------
class LinearSecretSharingScheme:
    """`sda_protocol::LinearSecretSharingScheme` (protocol/src/crypto.rs:79-114)."""

    def __init__(self, c):
        self.c = c

    @classmethod
    def Additive(cls, share_count, modulus):
        return cls(sda_sharing_scheme(_lib.SHARING_ADDITIVE, share_count, 0, 0, modulus, 0, 0))

    @classmethod
    def PackedShamir(cls, secret_count, share_count, privacy_threshold, prime_modulus, omega_secrets, omega_shares):
        return cls(sda_sharing_scheme(_lib.SHARING_PACKED_SHAMIR, share_count, secret_count, privacy_threshold,
                                      prime_modulus, omega_secrets, omega_shares))

    @property
    def modulus(self):
        return self.c.modulus

    def is_additive(self):
        return self.c.kind == _lib.SHARING_ADDITIVE

    def input_size(self):
        return _lib.load().sda_input_size(C.byref(self.c))

    def output_size(self):
        return _lib.load().sda_output_size(C.byref(self.c))

    def privacy_threshold(self):
        return _lib.load().sda_privacy_threshold(C.byref(self.c))

    def reconstruction_threshold(self):
        return _lib.load().sda_reconstruction_threshold(C.byref(self.c))

    def batches(self, dim):
        return _lib.load().sda_share_batches(C.byref(self.c), dim)

    def __repr__(self):
        c = self.c
        if self.is_additive():
            return f"Additive{{share_count: {c.share_count}, modulus: {c.modulus}}}"
        return (f"PackedShamir{{secret_count: {c.secret_count}, share_count: {c.share_count}, "
                f"privacy_threshold: {c.privacy_threshold}, prime_modulus: {c.modulus}, "
                f"omega_secrets: {c.omega_secrets}, omega_shares: {c.omega_shares}}}")


class LinearMaskingScheme:
    """`sda_protocol::LinearMaskingScheme` (protocol/src/crypto.rs:43-64)."""

    def __init__(self, c):
        self.c = c

    @classmethod
    def None_(cls):
        return cls(sda_masking_scheme(_lib.MASK_NONE, 0, 0, 0))

    @classmethod
    def Full(cls, modulus):
        return cls(sda_masking_scheme(_lib.MASK_FULL, modulus, 0, 0))

    @classmethod
    def ChaCha(cls, modulus, dimension, seed_bitsize):
        return cls(sda_masking_scheme(_lib.MASK_CHACHA, modulus, dimension, seed_bitsize))

    def has_mask(self):   # protocol/src/crypto.rs:66-75
        return self.c.kind != _lib.MASK_NONE

    def mask_len(self, dim):
        return _lib.load().sda_mask_len(C.byref(self.c), dim)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _out(out, shape):
    """a caller-provided output array the C side may write `shape` int64 elements into, or a fresh one"""
    count = int(np.prod(shape))
    if out is None:
        return np.empty(shape, dtype=np.int64)
    if not isinstance(out, np.ndarray) or out.dtype != np.int64 or not out.flags.c_contiguous or not out.flags.writeable \
            or out.size < count:
        raise ValueError(f"out must be a writeable C-contiguous int64 array of at least {count} elements")
    return out


def _seed(rng_seed):
    """32 bytes of entropy; the reference draws from OsRng at this point (additive.rs:17, full.rs:16)."""
    s = os.urandom(32) if rng_seed is None else bytes(rng_seed)
    if len(s) != 32:
        raise ValueError("rng_seed must be 32 bytes")
    return (C.c_uint8 * 32).from_buffer_copy(s)


def _dev_ptr(t):
    """device address of a torch CUDA tensor (or a raw int / None)."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


class Context:
    """Owns one `sda_ctx` (device, stream, scratch).  Not thread-safe; one per thread."""

    def __init__(self, device=0, rng_rounds=20):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.sda_ctx_create(device, C.byref(h))
        if rc:
            raise SdaClientError(rc, self._lib.sda_last_error(None).decode())
        self._h = h
        self.device = device
        if rng_rounds != 20:
            self.set_rng_rounds(rng_rounds)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sda_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc:
            raise SdaClientError(rc, self._lib.sda_last_error(self._h).decode())

    # -- configuration ---------------------------------------------------------------------------
    def set_rng_rounds(self, rounds):
        self.check(self._lib.sda_ctx_set_rng_rounds(self._h, rounds))

    def set_packed_path(self, path):
        """0 auto (tensor cores), 1 CUDA cores, 2 tensor cores: which kernel shares over 2^61-1"""
        self.check(self._lib.sda_ctx_set_packed_path(self._h, int(path)))

    def rng_rounds(self):
        return self._lib.sda_ctx_get_rng_rounds(self._h)

    def set_stream(self, cuda_stream):
        """cuda_stream: raw cudaStream_t as int (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        self.check(self._lib.sda_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def stream(self):
        return self._lib.sda_ctx_get_stream(self._h)

    def synchronize(self):
        self.check(self._lib.sda_ctx_synchronize(self._h))

    def set_deferred_checks(self, on):
        """*_dev calls that draw randomness stop synchronising; `synchronize()` reports a rejected keystream word (raises
        SdaClientError with code SDA_ERR_REJECTED) for the calls queued since"""
        self.check(self._lib.sda_ctx_set_deferred_checks(self._h, 1 if on else 0))

    def launch_count(self):
        return self._lib.sda_ctx_launch_count(self._h)

    def last_kernel(self):
        return self._lib.sda_ctx_last_kernel(self._h).decode()

    def pinned_empty(self, n, dtype=np.int64):
        """numpy array over cudaMallocHost memory (kept alive by the array's base object)."""
        nbytes = int(n) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self._lib.sda_host_alloc(self._h, max(nbytes, 1), C.byref(p)))
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n)).view(_PinnedArray)
        arr._owner = _PinnedOwner(self, p)
        return arr

    # -- scheme helpers --------------------------------------------------------------------------
    def validate(self, scheme):
        self.check(self._lib.sda_sharing_scheme_validate(self._h, C.byref(scheme.c)))

    def packed_share_matrix(self, scheme):
        n, w = scheme.c.share_count, scheme.c.secret_count + scheme.c.privacy_threshold
        out = np.empty((n, w), dtype=np.int64)
        self.check(self._lib.sda_packed_share_matrix(self._h, C.byref(scheme.c), _ptr(out)))
        return out

    def packed_reconstruct_matrix(self, scheme, indices):
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        out = np.empty((scheme.c.secret_count, len(idx)), dtype=np.int64)
        self.check(self._lib.sda_packed_reconstruct_matrix(self._h, C.byref(scheme.c), _ptr(idx), len(idx), _ptr(out)))
        return out

    # -- host-pointer entry points ---------------------------------------------------------------
    def share_generate(self, scheme, secrets, rng_seed=None, out=None):
        sec = _i64(secrets)
        n, B = scheme.output_size(), scheme.batches(len(sec))
        out = _out(out, (n, B))
        self.check(self._lib.sda_share_generate(self._h, C.byref(scheme.c), _ptr(sec), len(sec), _seed(rng_seed),
                                                _ptr(out)))
        return out

    def mask_share_generate(self, masking, sharing, secrets, mask_rng_seed=None, share_rng_seed=None):
        """participate.rs:53-54 then :75-76 in one call: (mask, shares[output_size][B]); the masked secrets stay on the
        device.  Same results as `mask` followed by `share_generate`."""
        sec = _i64(secrets)
        dim = len(sec)
        n, B = sharing.output_size(), sharing.batches(dim)
        ml = masking.mask_len(dim)
        mk = np.empty(max(ml, 1), dtype=np.int64)
        out = np.empty((n, B), dtype=np.int64)
        self.check(self._lib.sda_mask_share_generate(self._h, C.byref(masking.c), C.byref(sharing.c), _ptr(sec), dim,
                                                     _seed(mask_rng_seed), _seed(share_rng_seed), _ptr(mk), _ptr(out)))
        return mk[:ml], out

    def share_combine(self, scheme, shares, out=None):
        """shares: 2-D array [P][L] (contiguous fast path) or a list of 1-D rows (`Vec<Vec<Share>>`)."""
        if isinstance(shares, np.ndarray) and shares.ndim == 2 and shares.dtype == np.int64:
            sh = np.ascontiguousarray(shares)
            P, L = sh.shape
            out = _out(out, (L if P else 0,))
            self.check(self._lib.sda_share_combine(self._h, C.byref(scheme.c), _ptr(sh), P, L, _ptr(out)))
            return out
        rows = [_i64(r) for r in shares]
        P = len(rows)
        ptrs = (C.c_void_p * max(P, 1))(*[r.ctypes.data for r in rows])
        lens = (C.c_size_t * max(P, 1))(*[len(r) for r in rows])
        L = len(rows[0]) if P else 0
        out = _out(out, (L,))
        n = C.c_size_t(0)
        self.check(self._lib.sda_share_combine_rows(self._h, C.byref(scheme.c), ptrs, lens, P, _ptr(out), C.byref(n)))
        return out[:n.value]

    def secret_reconstruct(self, scheme, dimension, indexed_shares):
        """indexed_shares: list of (clerk index, share vector) -- `&Vec<(usize, Vec<Share>)>`."""
        idx = np.ascontiguousarray(np.asarray([i for i, _ in indexed_shares], dtype=np.uint64))
        rows = [_i64(r) for _, r in indexed_shares]
        m = len(rows)
        ptrs = (C.c_void_p * max(m, 1))(*[r.ctypes.data for r in rows])
        lens = (C.c_size_t * max(m, 1))(*[len(r) for r in rows])
        cap = max(dimension, len(rows[0]) if m else 0, 1)
        out = np.empty(cap, dtype=np.int64)
        n = C.c_size_t(0)
        self.check(self._lib.sda_secret_reconstruct_rows(self._h, C.byref(scheme.c), dimension, _ptr(idx), ptrs, lens, m,
                                                         _ptr(out), C.byref(n)))
        return out[:n.value]

    def mask(self, scheme, secrets, rng_seed=None):
        sec = _i64(secrets)
        dim = len(sec)
        mk = np.empty(max(scheme.mask_len(dim), 1), dtype=np.int64)
        masked = np.empty(dim, dtype=np.int64)
        n = C.c_size_t(0)
        self.check(self._lib.sda_mask(self._h, C.byref(scheme.c), _ptr(sec), dim, _seed(rng_seed), _ptr(mk), C.byref(n),
                                      _ptr(masked)))
        return mk[:n.value], masked

    def mask_combine(self, scheme, masks):
        rows = [_i64(r) for r in masks]
        P = len(rows)
        ml = len(rows[0]) if P else 0
        for r in rows:
            if len(r) != ml:   # full.rs:43 assert_eq!
                raise SdaClientError(_lib.SDA_ERR_INVALID, "assertion failed: `(left == right)` (full.rs:43)")
        mat = np.ascontiguousarray(np.stack(rows)) if P else np.empty((0, 0), dtype=np.int64)
        cap = max(ml, int(scheme.c.dimension), 1)
        out = np.empty(cap, dtype=np.int64)
        n = C.c_size_t(0)
        self.check(self._lib.sda_mask_combine(self._h, C.byref(scheme.c), _ptr(mat), P, ml, _ptr(out), C.byref(n)))
        return out[:n.value]

    def unmask(self, scheme, mask, masked):
        mk, md = _i64(mask), _i64(masked)
        out = np.empty(len(md), dtype=np.int64)
        self.check(self._lib.sda_unmask(self._h, C.byref(scheme.c), _ptr(mk), len(mk), _ptr(md), len(md), _ptr(out)))
        return out

    # -- device-pointer entry points (torch CUDA tensors or raw addresses) ------------------------
    def share_generate_dev(self, scheme, d_secrets, secrets_ld, P, dim, seeds, d_shares_out):
        seeds = bytes(seeds)
        if len(seeds) != 32 * P:
            raise ValueError("seeds must be P x 32 bytes")
        buf = (C.c_uint8 * max(len(seeds), 1)).from_buffer_copy(seeds or b"\0")
        self.check(self._lib.sda_share_generate_dev(self._h, C.byref(scheme.c), _dev_ptr(d_secrets), secrets_ld, P, dim,
                                                    buf, _dev_ptr(d_shares_out)))

    def mask_share_generate_dev(self, masking, sharing, d_secrets, secrets_ld, P, dim, mask_rng_seeds, share_rng_seeds,
                                d_masks_out, d_shares_out):
        """participate.rs:53-54 then :75-76 for P participants in one call: mask, then share the masked secrets (which are
        never written to memory where the fused kernel applies).  Seeds are P x 32 bytes each."""
        ms, sh = bytes(mask_rng_seeds), bytes(share_rng_seeds)
        if len(ms) != 32 * P or len(sh) != 32 * P:
            raise ValueError("mask_rng_seeds and share_rng_seeds must be P x 32 bytes")
        mbuf = (C.c_uint8 * max(len(ms), 1)).from_buffer_copy(ms or b"\0")
        sbuf = (C.c_uint8 * max(len(sh), 1)).from_buffer_copy(sh or b"\0")
        self.check(self._lib.sda_mask_share_generate_dev(self._h, C.byref(masking.c), C.byref(sharing.c), _dev_ptr(d_secrets),
                                                         secrets_ld, P, dim, mbuf, sbuf, _dev_ptr(d_masks_out),
                                                         _dev_ptr(d_shares_out)))

    def share_combine_dev(self, scheme, d_shares, ld, P, L, d_out, d_acc_in=None):
        self.check(self._lib.sda_share_combine_dev(self._h, C.byref(scheme.c), _dev_ptr(d_shares), ld, P, L,
                                                   _dev_ptr(d_acc_in), _dev_ptr(d_out)))

    def share_generate_combine_dev(self, scheme, d_secrets, secrets_ld, P, dim, seeds, d_out, d_acc_in=None):
        seeds = bytes(seeds)
        if len(seeds) != 32 * P:
            raise ValueError("seeds must be P x 32 bytes")
        buf = (C.c_uint8 * max(len(seeds), 1)).from_buffer_copy(seeds or b"\0")
        self.check(self._lib.sda_share_generate_combine_dev(self._h, C.byref(scheme.c), _dev_ptr(d_secrets), secrets_ld,
                                                            P, dim, buf, _dev_ptr(d_acc_in), _dev_ptr(d_out)))

    # -- multi-GPU clerk sum: the one exchange of the path, NCCL inside the library ------------
    @staticmethod
    def nccl_unique_id():
        """128 bytes rank 0 hands to the other ranks (any transport) before comm_init_rank"""
        lib = _lib.load()
        buf = (C.c_uint8 * 128)()
        rc = lib.sda_nccl_unique_id(buf)
        if rc:
            raise SdaClientError(rc, lib.sda_last_error(None).decode())
        return bytes(buf)

    def comm_init_rank(self, unique_id, nranks, rank):
        """collective: every rank calls it with the same id"""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self.check(self._lib.sda_ctx_comm_init_rank(self._h, buf, nranks, rank))

    def comm_rank(self):
        return self._lib.sda_ctx_comm_rank(self._h)

    def comm_size(self):
        return self._lib.sda_ctx_comm_size(self._h)

    def partial_sums_reduce_dev(self, modulus, d_partials, count, root=0):
        """in place: this rank's canonical column sums -> (on `root`) the sums over all ranks mod `modulus`"""
        self.check(self._lib.sda_partial_sums_reduce_dev(self._h, modulus, _dev_ptr(d_partials), count, root))

    def share_combine_ranks_dev(self, scheme, d_shares, ld, P_local, L, d_out, d_acc_in=None, root=0):
        self.check(self._lib.sda_share_combine_ranks_dev(self._h, C.byref(scheme.c), _dev_ptr(d_shares), ld, P_local, L,
                                                         _dev_ptr(d_acc_in), _dev_ptr(d_out), root))

    def mod_reduce_dev(self, modulus, d_in, n, d_out, unsigned=False):
        f = self._lib.sda_mod_reduce_u64_dev if unsigned else self._lib.sda_mod_reduce_dev
        self.check(f(self._h, modulus, _dev_ptr(d_in), n, _dev_ptr(d_out)))

    def secret_reconstruct_dev(self, scheme, dimension, indices, d_shares, ld, m, B, d_out):
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
        self.check(self._lib.sda_secret_reconstruct_dev(self._h, C.byref(scheme.c), dimension, _ptr(idx),
                                                        _dev_ptr(d_shares), ld, m, B, _dev_ptr(d_out)))

    def mask_dev(self, scheme, d_secrets, dim, rng_seed, d_mask_out, d_masked_out):
        self.check(self._lib.sda_mask_dev(self._h, C.byref(scheme.c), _dev_ptr(d_secrets), dim, _seed(rng_seed),
                                          _dev_ptr(d_mask_out), _dev_ptr(d_masked_out)))

    def mask_combine_dev(self, scheme, d_masks, P, mask_len, d_out):
        self.check(self._lib.sda_mask_combine_dev(self._h, C.byref(scheme.c), _dev_ptr(d_masks), P, mask_len,
                                                  _dev_ptr(d_out)))

    def unmask_dev(self, scheme, d_mask, d_masked, dim, d_out):
        self.check(self._lib.sda_unmask_dev(self._h, C.byref(scheme.c), _dev_ptr(d_mask), _dev_ptr(d_masked), dim,
                                            _dev_ptr(d_out)))

    # -- share wire codec (encryption/sodium.rs:35-41, 83-90) --------------------------------------
    def varint_encode(self, shares):
        """`for share in shares { share.encode_var(..) }`: zig-zag LEB128, concatenated.  Returns uint8 array."""
        v = _i64(shares)
        out = np.empty(max(10 * len(v), 1), dtype=np.uint8)
        n = C.c_size_t(0)
        self.check(self._lib.sda_varint_encode(self._h, _ptr(v), len(v), _ptr(out), C.byref(n)))
        return out[:n.value]

    def varint_decode(self, buf, cap=None):
        """`while !reader.is_empty() { Share::decode_var(reader) }`.  Returns int64 array."""
        b = np.ascontiguousarray(np.frombuffer(bytes(buf), dtype=np.uint8)) if not isinstance(buf, np.ndarray) \
            else np.ascontiguousarray(buf, dtype=np.uint8)
        cap = len(b) if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.int64)
        n = C.c_size_t(0)
        self.check(self._lib.sda_varint_decode(self._h, _ptr(b), len(b), _ptr(out), cap, C.byref(n)))
        return out[:n.value]

    def varint_encode_dev(self, d_shares, n, d_out):
        ln = C.c_size_t(0)
        self.check(self._lib.sda_varint_encode_dev(self._h, _dev_ptr(d_shares), n, _dev_ptr(d_out), C.byref(ln)))
        return ln.value

    def varint_decode_dev(self, d_buf, length, d_out, cap):
        cnt = C.c_size_t(0)
        self.check(self._lib.sda_varint_decode_dev(self._h, _dev_ptr(d_buf), length, _dev_ptr(d_out), cap, C.byref(cnt)))
        return cnt.value

    # -- fixed-point codec of real-valued vectors (not in the reference; BASELINE config #5) ---------
    def fixed_encode_dev(self, modulus, frac_bits, d_x, n, d_out):
        self.check(self._lib.sda_fixed_encode_dev(self._h, modulus, frac_bits, _dev_ptr(d_x), n, _dev_ptr(d_out)))

    def snapshot_transpose_dev(self, d_blobs, offsets, P, n, d_out):
        """server/src/snapshot.rs:11-27: blob (p, c) -> clerk-major; returns the clerk-major offsets (numpy u64, P n + 1)"""
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.uint64))
        if off.size != P * n + 1:
            raise ValueError("offsets must have P * n + 1 entries")
        out_off = np.zeros(P * n + 1, dtype=np.uint64)
        self.check(self._lib.sda_snapshot_transpose_dev(self._h, _dev_ptr(d_blobs), _ptr(off), P, n, _dev_ptr(d_out), _ptr(out_off)))
        return out_off

    def fixed_encode_mask_dev(self, scheme, modulus, frac_bits, d_x, dim, rng_seed, d_mask_out, d_masked_out, P=1, x_ld=None,
                              masked_ld=None):
        """fixed_encode_dev + mask_dev in one pass for P participants (float32 rows in, masked residues out); rng_seed is
        P x 32 bytes"""
        seeds = bytes(rng_seed)
        if len(seeds) != 32 * P:
            raise ValueError("rng_seed must be P x 32 bytes")
        buf = (C.c_uint8 * len(seeds)).from_buffer_copy(seeds)
        self.check(self._lib.sda_fixed_encode_mask_dev(self._h, C.byref(scheme.c), modulus, frac_bits, _dev_ptr(d_x),
                                                       dim if x_ld is None else x_ld, P, dim, buf, _dev_ptr(d_mask_out),
                                                       _dev_ptr(d_masked_out), dim if masked_ld is None else masked_ld))

    def fixed_decode_dev(self, modulus, frac_bits, divisor, d_in, n, d_out):
        self.check(self._lib.sda_fixed_decode_dev(self._h, modulus, frac_bits, divisor, _dev_ptr(d_in), n, _dev_ptr(d_out)))

    def synth_fill_dev(self, stream_id, modulus, start, count, d_out):
        self.check(self._lib.sda_synth_fill_dev(self._h, stream_id, modulus, start, count, _dev_ptr(d_out)))


class _PinnedOwner:
    def __init__(self, ctx, ptr):
        self.lib, self.h, self.ptr = ctx._lib, ctx._h, ptr
        self.ctx = ctx   # keep the context alive

    def __del__(self):
        try:
            if self.ptr and self.ctx._h:
                self.lib.sda_host_free(self.ctx._h, self.ptr)
        except Exception:
            pass


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


# ---- the trait objects (client/src/crypto/{sharing,masking}/mod.rs) ------------------------------
class ShareGenerator:
    """`ShareGenerator::generate(&mut self, secrets: &[Secret]) -> SdaClientResult<Vec<Vec<Share>>>`"""

    def __init__(self, ctx, scheme):
        self.ctx, self.scheme = ctx, scheme

    def generate(self, secrets, rng_seed=None):
        return self.ctx.share_generate(self.scheme, secrets, rng_seed)


class ShareCombiner:
    """`ShareCombiner::combine(&self, shares: &Vec<Vec<Share>>) -> SdaClientResult<Vec<Share>>`"""

    def __init__(self, ctx, scheme):
        self.ctx, self.scheme = ctx, scheme

    def combine(self, shares):
        return self.ctx.share_combine(self.scheme, shares)


class SecretReconstructor:
    """`SecretReconstructor::reconstruct(&self, &Vec<(usize, Vec<Share>)>) -> SdaClientResult<Vec<Secret>>`"""

    def __init__(self, ctx, scheme, dimension):
        self.ctx, self.scheme, self.dimension = ctx, scheme, dimension

    def reconstruct(self, indexed_shares):
        return self.ctx.secret_reconstruct(self.scheme, self.dimension, indexed_shares)


class SecretMasker:
    """`SecretMasker::mask(&mut self, secrets: &[Secret]) -> (Vec<Mask>, Vec<MaskedSecret>)`"""

    def __init__(self, ctx, scheme):
        self.ctx, self.scheme = ctx, scheme

    def mask(self, secrets, rng_seed=None):
        return self.ctx.mask(self.scheme, secrets, rng_seed)


class MaskCombiner:
    """`MaskCombiner::combine(&self, masks: &Vec<Vec<Mask>>) -> Vec<Mask>`"""

    def __init__(self, ctx, scheme):
        self.ctx, self.scheme = ctx, scheme

    def combine(self, masks):
        return self.ctx.mask_combine(self.scheme, masks)


class SecretUnmasker:
    """`SecretUnmasker::unmask(&self, values: &(Vec<Mask>, Vec<MaskedSecret>)) -> Vec<Secret>`"""

    def __init__(self, ctx, scheme):
        self.ctx, self.scheme = ctx, scheme

    def unmask(self, values):
        mask, masked = values
        return self.ctx.unmask(self.scheme, mask, masked)


class CryptoModule:
    """`sda_client::crypto::CryptoModule` restricted to the sharing / masking constructions
    (client/src/crypto/mod.rs:58-66; sharing/mod.rs:35-96; masking/mod.rs:33-94)."""

    def __init__(self, device=0, rng_rounds=20):
        self.ctx = Context(device, rng_rounds)

    def new_share_generator(self, scheme):
        self.ctx.validate(scheme)
        return ShareGenerator(self.ctx, scheme)

    def new_share_combiner(self, scheme):
        return ShareCombiner(self.ctx, scheme)

    def new_secret_reconstructor(self, scheme, dimension):
        return SecretReconstructor(self.ctx, scheme, dimension)

    def new_secret_masker(self, scheme):
        return SecretMasker(self.ctx, scheme)

    def new_mask_combiner(self, scheme):
        return MaskCombiner(self.ctx, scheme)

    def new_secret_unmasker(self, scheme):
        return SecretUnmasker(self.ctx, scheme)
