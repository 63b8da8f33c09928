//! Bindings of `libsda_b200.so` (C ABI: `include/sda_b200.h`, ABI version 1).
//!
//! * `ffi` -- the `extern "C"` declarations, one per entry point the sda-client shim uses
//!   (`bindings/rust/client-patch/b200.rs`), field-for-field mirrors of `sda_sharing_scheme` /
//!   `sda_masking_scheme` (`protocol/src/crypto.rs:79-114`, `:43-64`).
//! * `Context` -- an owned `sda_ctx` (one per thread: a context is not thread-safe) with `Result`-returning
//!   wrappers; the error string is the reference's own `Err(..)` / panic text.
//! * `PinnedVec` -- a `Vec<i64>`-like owner of pinned host memory (`sda_host_alloc`): with pinned input and
//!   output buffers the host entry points overlap the copy in, the kernel and the copy out.
//!
//! SOURCE ONLY: the image this repository is developed in has no Rust toolchain, so this crate has not been
//! compiled there.  The same ABI is exercised end to end by `tests/` through the ctypes mirror and by
//! `tests/c_abi_smoke.c` from plain C.

use std::ffi::CStr;
use std::ops::{Deref, DerefMut};
use std::os::raw::{c_char, c_int, c_void};
use std::ptr;

pub const SDA_OK: c_int = 0;
pub const SDA_ERR_INVALID: c_int = 1;
pub const SDA_ERR_CUDA: c_int = 2;
pub const SDA_ERR_NCCL: c_int = 3;
pub const SDA_ERR_UNSUPPORTED: c_int = 4;
pub const SDA_ERR_REJECTED: c_int = 5;      // deferred checks: a *_dev call queued earlier must be redone

pub const SDA_SHARING_ADDITIVE: i32 = 0;
pub const SDA_SHARING_PACKED_SHAMIR: i32 = 1;
pub const SDA_MASK_NONE: i32 = 0;
pub const SDA_MASK_FULL: i32 = 1;
pub const SDA_MASK_CHACHA: i32 = 2;

/// `sda_sharing_scheme` (protocol/src/crypto.rs:79-114).  Additive uses `share_count` and `modulus` only.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct SharingScheme {
    pub kind: i32,
    pub share_count: u64,
    pub secret_count: u64,
    pub privacy_threshold: u64,
    pub modulus: i64,
    pub omega_secrets: i64,
    pub omega_shares: i64,
}

/// `sda_masking_scheme` (protocol/src/crypto.rs:43-64).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MaskingScheme {
    pub kind: i32,
    pub modulus: i64,
    pub dimension: u64,
    pub seed_bitsize: u64,
}

impl SharingScheme {
    pub fn additive(share_count: usize, modulus: i64) -> SharingScheme {
        SharingScheme { kind: SDA_SHARING_ADDITIVE, share_count: share_count as u64, secret_count: 0,
                        privacy_threshold: 0, modulus: modulus, omega_secrets: 0, omega_shares: 0 }
    }
    pub fn packed_shamir(secret_count: usize, share_count: usize, privacy_threshold: usize, prime_modulus: i64,
                         omega_secrets: i64, omega_shares: i64) -> SharingScheme {
        SharingScheme { kind: SDA_SHARING_PACKED_SHAMIR, share_count: share_count as u64,
                        secret_count: secret_count as u64, privacy_threshold: privacy_threshold as u64,
                        modulus: prime_modulus, omega_secrets: omega_secrets, omega_shares: omega_shares }
    }
    pub fn input_size(&self) -> usize { unsafe { ffi::sda_input_size(self) } }
    pub fn output_size(&self) -> usize { unsafe { ffi::sda_output_size(self) } }
    pub fn privacy_threshold(&self) -> usize { unsafe { ffi::sda_privacy_threshold(self) } }
    pub fn reconstruction_threshold(&self) -> usize { unsafe { ffi::sda_reconstruction_threshold(self) } }
    /// ceil(dim / input_size): length of each clerk's share vector (batched.rs:23)
    pub fn batches(&self, dim: usize) -> usize { unsafe { ffi::sda_share_batches(self, dim) } }
}

impl MaskingScheme {
    pub fn none() -> MaskingScheme { MaskingScheme { kind: SDA_MASK_NONE, modulus: 0, dimension: 0, seed_bitsize: 0 } }
    pub fn full(modulus: i64) -> MaskingScheme {
        MaskingScheme { kind: SDA_MASK_FULL, modulus: modulus, dimension: 0, seed_bitsize: 0 }
    }
    pub fn chacha(modulus: i64, dimension: usize, seed_bitsize: usize) -> MaskingScheme {
        MaskingScheme { kind: SDA_MASK_CHACHA, modulus: modulus, dimension: dimension as u64, seed_bitsize: seed_bitsize as u64 }
    }
    /// length of the mask `SecretMasker::mask` returns: 0 | dim | ceil(seed_bitsize / 32)
    pub fn mask_len(&self, dim: usize) -> usize { unsafe { ffi::sda_mask_len(self, dim) } }
}

pub mod ffi {
    use super::{MaskingScheme, SharingScheme};
    use std::os::raw::{c_char, c_int, c_void};

    pub enum SdaCtx {}

    extern "C" {
        pub fn sda_abi_version() -> c_int;
        pub fn sda_ctx_create(device: c_int, out: *mut *mut SdaCtx) -> c_int;
        pub fn sda_ctx_create_multi(devices: *const c_int, ndev: c_int, out: *mut *mut SdaCtx) -> c_int;
        pub fn sda_ctx_multi_count(ctx: *const SdaCtx) -> c_int;
        pub fn sda_ctx_destroy(ctx: *mut SdaCtx);
        pub fn sda_last_error(ctx: *const SdaCtx) -> *const c_char;
        pub fn sda_ctx_set_rng_rounds(ctx: *mut SdaCtx, rounds: c_int) -> c_int;
        pub fn sda_ctx_synchronize(ctx: *mut SdaCtx) -> c_int;
        pub fn sda_ctx_set_deferred_checks(ctx: *mut SdaCtx, on: c_int) -> c_int;
        pub fn sda_host_alloc(ctx: *mut SdaCtx, bytes: usize, out: *mut *mut c_void) -> c_int;
        pub fn sda_host_free(ctx: *mut SdaCtx, ptr: *mut c_void) -> c_int;

        pub fn sda_input_size(s: *const SharingScheme) -> usize;
        pub fn sda_output_size(s: *const SharingScheme) -> usize;
        pub fn sda_privacy_threshold(s: *const SharingScheme) -> usize;
        pub fn sda_reconstruction_threshold(s: *const SharingScheme) -> usize;
        pub fn sda_share_batches(s: *const SharingScheme, dim: usize) -> usize;
        pub fn sda_mask_len(s: *const MaskingScheme, dim: usize) -> usize;
        pub fn sda_sharing_scheme_validate(ctx: *mut SdaCtx, s: *const SharingScheme) -> c_int;

        pub fn sda_share_generate(ctx: *mut SdaCtx, s: *const SharingScheme, secrets: *const i64, dim: usize,
                                  rng_seed: *const u8, shares_out: *mut i64) -> c_int;
        pub fn sda_share_combine_rows(ctx: *mut SdaCtx, s: *const SharingScheme, rows: *const *const i64,
                                      row_lens: *const usize, p: usize, out: *mut i64, out_len: *mut usize) -> c_int;
        pub fn sda_share_combine_rows_multi(ctx: *mut SdaCtx, s: *const SharingScheme, rows: *const *const i64,
                                            row_lens: *const usize, p: usize, out: *mut i64, out_len: *mut usize) -> c_int;
        pub fn sda_secret_reconstruct_rows(ctx: *mut SdaCtx, s: *const SharingScheme, dimension: usize,
                                           indices: *const u64, rows: *const *const i64, row_lens: *const usize,
                                           m: usize, secrets_out: *mut i64, out_len: *mut usize) -> c_int;
        pub fn sda_mask(ctx: *mut SdaCtx, s: *const MaskingScheme, secrets: *const i64, dim: usize, rng_seed: *const u8,
                        mask_out: *mut i64, mask_len: *mut usize, masked_out: *mut i64) -> c_int;
        pub fn sda_mask_share_generate(ctx: *mut SdaCtx, ms: *const MaskingScheme, ss: *const SharingScheme, secrets: *const i64,
                                       dim: usize, mask_rng_seed: *const u8, share_rng_seed: *const u8, mask_out: *mut i64,
                                       shares_out: *mut i64) -> c_int;
        pub fn sda_mask_combine(ctx: *mut SdaCtx, s: *const MaskingScheme, masks: *const i64, p: usize, mask_len: usize,
                                out: *mut i64, out_len: *mut usize) -> c_int;
        pub fn sda_unmask(ctx: *mut SdaCtx, s: *const MaskingScheme, mask: *const i64, mask_len: usize,
                          masked: *const i64, dim: usize, out: *mut i64) -> c_int;

        pub fn sda_varint_max_bytes(n: usize) -> usize;
        pub fn sda_varint_encode(ctx: *mut SdaCtx, shares: *const i64, n: usize, out: *mut u8, out_len: *mut usize) -> c_int;
        pub fn sda_varint_decode(ctx: *mut SdaCtx, buf: *const u8, len: usize, shares_out: *mut i64, cap: usize,
                                 n: *mut usize) -> c_int;
    }
}

/// Error class + message of a failed call.  For class `SDA_ERR_INVALID` the message is the reference's own
/// `Err(..)` string or panic message ("Wrong dimension", "Not enough shares to reconstruct", ...).
#[derive(Debug, Clone)]
pub struct Error {
    pub code: c_int,
    pub message: String,
}

impl ::std::fmt::Display for Error {
    fn fmt(&self, f: &mut ::std::fmt::Formatter) -> ::std::fmt::Result { write!(f, "{}", self.message) }
}
impl ::std::error::Error for Error {
    fn description(&self) -> &str { &self.message }
}

pub type Result<T> = ::std::result::Result<T, Error>;

fn last_error(ctx: *const ffi::SdaCtx) -> String {
    unsafe {
        let p: *const c_char = ffi::sda_last_error(ctx);
        if p.is_null() { String::new() } else { CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}

/// An owned `sda_ctx`: device, streams, scratch, and -- for `Context::multi` -- the member contexts of the other GPUs
/// and the NCCL communicator.  NOT `Sync`; create one per thread (they are cheap).
pub struct Context {
    raw: *mut ffi::SdaCtx,
}

impl Drop for Context {
    fn drop(&mut self) { unsafe { ffi::sda_ctx_destroy(self.raw) } }
}

impl Context {
    /// `sda_ctx_create(device)`.  Fails with class `SDA_ERR_CUDA` when there is no sm_100 device: there is no CPU fallback.
    pub fn new(device: i32) -> Result<Context> {
        let mut raw: *mut ffi::SdaCtx = ptr::null_mut();
        let rc = unsafe { ffi::sda_ctx_create(device as c_int, &mut raw) };
        if rc != SDA_OK { return Err(Error { code: rc, message: last_error(ptr::null()) }) }
        Ok(Context { raw: raw })
    }

    /// One process, several GPUs: `sda_ctx_create_multi`.  Single-GPU calls run on `devices[0]`; `share_combine_rows`
    /// shards the participants over all devices and sums the partial results with one NCCL reduce inside the library.
    pub fn multi(devices: &[i32]) -> Result<Context> {
        let devs: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
        let mut raw: *mut ffi::SdaCtx = ptr::null_mut();
        let rc = unsafe { ffi::sda_ctx_create_multi(devs.as_ptr(), devs.len() as c_int, &mut raw) };
        if rc != SDA_OK { return Err(Error { code: rc, message: last_error(ptr::null()) }) }
        Ok(Context { raw: raw })
    }

    pub fn as_ptr(&self) -> *mut ffi::SdaCtx { self.raw }
    pub fn device_count(&self) -> usize { unsafe { ffi::sda_ctx_multi_count(self.raw) as usize } }

    fn check(&self, rc: c_int) -> Result<()> {
        if rc == SDA_OK { Ok(()) } else { Err(Error { code: rc, message: last_error(self.raw) }) }
    }

    pub fn validate(&self, s: &SharingScheme) -> Result<()> {
        self.check(unsafe { ffi::sda_sharing_scheme_validate(self.raw, s) })
    }

    /// `ShareGenerator::generate` (sharing/mod.rs:14-17): one `Vec<Share>` per clerk, each `ceil(dim / input_size)` long.
    /// `rng_seed`: 32 bytes of fresh entropy (the reference draws from `OsRng` at this point).
    pub fn share_generate(&self, s: &SharingScheme, secrets: &[i64], rng_seed: &[u8; 32]) -> Result<Vec<Vec<i64>>> {
        let n = s.output_size();
        let l = s.batches(secrets.len());
        let mut flat = vec![0i64; n * l];
        try!(self.check(unsafe {
            ffi::sda_share_generate(self.raw, s, secrets.as_ptr(), secrets.len(), rng_seed.as_ptr(), flat.as_mut_ptr())
        }));
        Ok((0..n).map(|r| flat[r * l..(r + 1) * l].to_vec()).collect())
    }

    /// `SecretMasker::mask` followed by `ShareGenerator::generate` on its output (participate.rs:53-54, :75-76) in one
    /// call: `(mask, shares per clerk)`.  Same results as the two calls; the masked secrets stay on the device.
    pub fn mask_share_generate(&self, ms: &MaskingScheme, ss: &SharingScheme, secrets: &[i64], mask_rng_seed: &[u8; 32],
                               share_rng_seed: &[u8; 32]) -> Result<(Vec<i64>, Vec<Vec<i64>>)> {
        let n = ss.output_size();
        let l = ss.batches(secrets.len());
        let mut mask = vec![0i64; ms.mask_len(secrets.len())];
        let mut flat = vec![0i64; n * l];
        try!(self.check(unsafe {
            ffi::sda_mask_share_generate(self.raw, ms, ss, secrets.as_ptr(), secrets.len(), mask_rng_seed.as_ptr(),
                                         share_rng_seed.as_ptr(), mask.as_mut_ptr(), flat.as_mut_ptr())
        }));
        Ok((mask, (0..n).map(|r| flat[r * l..(r + 1) * l].to_vec()).collect()))
    }

    /// Same into caller-provided (ideally pinned) storage `[n][l]`, row r = clerk r: no per-row allocation.
    pub fn share_generate_into(&self, s: &SharingScheme, secrets: &[i64], rng_seed: &[u8; 32], out: &mut [i64]) -> Result<()> {
        if out.len() < s.output_size() * s.batches(secrets.len()) {
            return Err(Error { code: SDA_ERR_INVALID, message: "output buffer too small".to_string() })
        }
        self.check(unsafe {
            ffi::sda_share_generate(self.raw, s, secrets.as_ptr(), secrets.len(), rng_seed.as_ptr(), out.as_mut_ptr())
        })
    }

    /// `ShareCombiner::combine` on a `&Vec<Vec<Share>>` as it lies in memory (combiner.rs:15-29).  On a multi-GPU
    /// context the rows are sharded over the devices.
    pub fn share_combine_rows(&self, s: &SharingScheme, shares: &[Vec<i64>]) -> Result<Vec<i64>> {
        let rows: Vec<*const i64> = shares.iter().map(|r| r.as_ptr()).collect();
        let lens: Vec<usize> = shares.iter().map(|r| r.len()).collect();
        let mut out = vec![0i64; lens.first().cloned().unwrap_or(0)];
        let mut out_len = 0usize;
        let rc = unsafe {
            if self.device_count() > 1 {
                ffi::sda_share_combine_rows_multi(self.raw, s, rows.as_ptr(), lens.as_ptr(), rows.len(), out.as_mut_ptr(), &mut out_len)
            } else {
                ffi::sda_share_combine_rows(self.raw, s, rows.as_ptr(), lens.as_ptr(), rows.len(), out.as_mut_ptr(), &mut out_len)
            }
        };
        try!(self.check(rc));
        out.truncate(out_len);
        Ok(out)
    }

    /// `SecretReconstructor::reconstruct` (additive.rs:55-73, batched.rs:68-97 + packed_shamir.rs:73-77).
    pub fn secret_reconstruct_rows(&self, s: &SharingScheme, dimension: usize, indexed_shares: &[(usize, Vec<i64>)]) -> Result<Vec<i64>> {
        let idx: Vec<u64> = indexed_shares.iter().map(|&(i, _)| i as u64).collect();
        let rows: Vec<*const i64> = indexed_shares.iter().map(|&(_, ref v)| v.as_ptr()).collect();
        let lens: Vec<usize> = indexed_shares.iter().map(|&(_, ref v)| v.len()).collect();
        // additive: the output is as long as the first share vector (additive.rs:56-58); packed: `dimension`
        let cap = ::std::cmp::max(dimension, lens.first().cloned().unwrap_or(0));
        let mut out = vec![0i64; cap];
        let mut out_len = 0usize;
        try!(self.check(unsafe {
            ffi::sda_secret_reconstruct_rows(self.raw, s, dimension, idx.as_ptr(), rows.as_ptr(), lens.as_ptr(), rows.len(),
                                             out.as_mut_ptr(), &mut out_len)
        }));
        out.truncate(out_len);
        Ok(out)
    }

    /// `SecretMasker::mask` -> (mask, masked) (none.rs:13-19, full.rs:21-35, chacha.rs:24-54).
    pub fn mask(&self, s: &MaskingScheme, secrets: &[i64], rng_seed: &[u8; 32]) -> Result<(Vec<i64>, Vec<i64>)> {
        let mut mask = vec![0i64; s.mask_len(secrets.len())];
        let mut masked = vec![0i64; secrets.len()];
        let mut mask_len = 0usize;
        try!(self.check(unsafe {
            ffi::sda_mask(self.raw, s, secrets.as_ptr(), secrets.len(), rng_seed.as_ptr(), mask.as_mut_ptr(), &mut mask_len,
                          masked.as_mut_ptr())
        }));
        mask.truncate(mask_len);
        Ok((mask, masked))
    }

    /// `MaskCombiner::combine` (none.rs:21-26, full.rs:37-52, chacha.rs:56-77).  `dim`: the vector dimension (Full: the
    /// mask length; ChaCha: `scheme.dimension`).
    pub fn mask_combine(&self, s: &MaskingScheme, masks: &[Vec<i64>], dim: usize) -> Result<Vec<i64>> {
        let mask_len = masks.first().map(|m| m.len()).unwrap_or(0);
        let mut flat: Vec<i64> = Vec::with_capacity(masks.len() * mask_len);
        for m in masks {
            if m.len() != mask_len {
                return Err(Error { code: SDA_ERR_INVALID, message: "assertion failed: `(left == right)` (mask lengths differ)".to_string() })
            }
            flat.extend_from_slice(m);
        }
        let mut out = vec![0i64; ::std::cmp::max(dim, mask_len)];
        let mut out_len = 0usize;
        try!(self.check(unsafe {
            ffi::sda_mask_combine(self.raw, s, flat.as_ptr(), masks.len(), mask_len, out.as_mut_ptr(), &mut out_len)
        }));
        out.truncate(out_len);
        Ok(out)
    }

    /// `SecretUnmasker::unmask` (none.rs:28-33, full.rs:54-66, chacha.rs:79-92).
    pub fn unmask(&self, s: &MaskingScheme, mask: &[i64], masked: &[i64]) -> Result<Vec<i64>> {
        let mut out = vec![0i64; masked.len()];
        try!(self.check(unsafe {
            ffi::sda_unmask(self.raw, s, mask.as_ptr(), mask.len(), masked.as_ptr(), masked.len(), out.as_mut_ptr())
        }));
        Ok(out)
    }

    /// the encoding loop of `ShareEncryptor::encrypt` (encryption/sodium.rs:35-41): zig-zag LEB128, concatenated
    pub fn varint_encode(&self, shares: &[i64]) -> Result<Vec<u8>> {
        let mut out = vec![0u8; unsafe { ffi::sda_varint_max_bytes(shares.len()) }];
        let mut len = 0usize;
        try!(self.check(unsafe { ffi::sda_varint_encode(self.raw, shares.as_ptr(), shares.len(), out.as_mut_ptr(), &mut len) }));
        out.truncate(len);
        Ok(out)
    }

    /// the decoding loop of `ShareDecryptor::decrypt` (encryption/sodium.rs:83-90)
    pub fn varint_decode(&self, buf: &[u8]) -> Result<Vec<i64>> {
        let mut out = vec![0i64; buf.len()];      // a value takes at least one byte
        let mut n = 0usize;
        try!(self.check(unsafe { ffi::sda_varint_decode(self.raw, buf.as_ptr(), buf.len(), out.as_mut_ptr(), out.len(), &mut n) }));
        out.truncate(n);
        Ok(out)
    }
}

/// `Vec<i64>`-like owner of pinned host memory.  Host entry points given pinned input AND output walk vectors of
/// 4 MB and more in slices so that the copy in, the kernel and the copy out overlap (`include/sda_b200.h`).
pub struct PinnedVec<'a> {
    ptr: *mut i64,
    len: usize,
    ctx: &'a Context,
}

impl<'a> PinnedVec<'a> {
    pub fn zeroed(ctx: &'a Context, len: usize) -> Result<PinnedVec<'a>> {
        let mut p: *mut c_void = ptr::null_mut();
        let bytes = ::std::cmp::max(len, 1) * 8;
        try!(ctx.check(unsafe { ffi::sda_host_alloc(ctx.raw, bytes, &mut p) }));
        unsafe { ptr::write_bytes(p as *mut u8, 0, bytes) };
        Ok(PinnedVec { ptr: p as *mut i64, len: len, ctx: ctx })
    }
    pub fn from_slice(ctx: &'a Context, src: &[i64]) -> Result<PinnedVec<'a>> {
        let mut v = try!(PinnedVec::zeroed(ctx, src.len()));
        v.copy_from_slice(src);
        Ok(v)
    }
}
impl<'a> Deref for PinnedVec<'a> {
    type Target = [i64];
    fn deref(&self) -> &[i64] { unsafe { ::std::slice::from_raw_parts(self.ptr, self.len) } }
}
impl<'a> DerefMut for PinnedVec<'a> {
    fn deref_mut(&mut self) -> &mut [i64] { unsafe { ::std::slice::from_raw_parts_mut(self.ptr, self.len) } }
}
impl<'a> Drop for PinnedVec<'a> {
    fn drop(&mut self) { unsafe { ffi::sda_host_free(self.ctx.raw, self.ptr as *mut c_void); } }
}

#[cfg(test)]
mod tests {
    use super::*;

    // integration-tests/tests/full_loop.rs:11-27,113,148: two participants [1,2,3,4] -> [2,4,6,8] (needs a B200)
    fn full_loop(sharing: SharingScheme, masking: MaskingScheme) {
        let ctx = Context::new(0).expect("no CUDA device: libsda_b200 has no CPU fallback");
        let secrets = [1i64, 2, 3, 4];
        let n = sharing.output_size();
        let mut per_clerk: Vec<Vec<Vec<i64>>> = vec![Vec::new(); n];
        let mut masks = Vec::new();
        for p in 0..2u8 {
            let (mask, masked) = ctx.mask(&masking, &secrets, &[p + 1; 32]).unwrap();
            masks.push(mask);
            let shares = ctx.share_generate(&sharing, &masked, &[p + 101; 32]).unwrap();
            for (c, row) in shares.into_iter().enumerate() { per_clerk[c].push(row); }
        }
        let combined: Vec<(usize, Vec<i64>)> =
            per_clerk.iter().enumerate().map(|(c, rows)| (c, ctx.share_combine_rows(&sharing, rows).unwrap())).collect();
        let masked_sum = ctx.secret_reconstruct_rows(&sharing, secrets.len(), &combined).unwrap();
        let mask_sum = ctx.mask_combine(&masking, &masks, secrets.len()).unwrap();
        assert_eq!(ctx.unmask(&masking, &mask_sum, &masked_sum).unwrap(), vec![2, 4, 6, 8]);
    }

    #[test] fn additive() { full_loop(SharingScheme::additive(3, 433), MaskingScheme::none()); }
    #[test] fn additive_full_mask() { full_loop(SharingScheme::additive(3, 433), MaskingScheme::full(433)); }
    #[test] fn additive_chacha_mask() { full_loop(SharingScheme::additive(3, 433), MaskingScheme::chacha(433, 4, 128)); }
    #[test] fn packed_shamir() { full_loop(SharingScheme::packed_shamir(3, 8, 4, 433, 354, 150), MaskingScheme::none()); }
}
