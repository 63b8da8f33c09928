//! `client/src/crypto/b200.rs` -- sda-client's sharing / masking / reconstruction traits on `libsda_b200.so`.
//!
//! Drop this file into `client/src/crypto/`, add `mod b200;` to `client/src/crypto/mod.rs`, the dependency
//! `sda-b200-sys = { path = "<repo>/bindings/rust/sda-b200-sys" }` to `client/Cargo.toml` and `extern crate sda_b200_sys;`
//! to `client/src/lib.rs`, and replace the bodies of the six `impl ...Construction<...> for CryptoModule` blocks
//! (`crypto/sharing/mod.rs:35-96`, `crypto/masking/mod.rs:33-94`) by the one-liners at the bottom of this file
//! (`README.md` next to it has the diff).  Callers (`participate.rs:53-54,75-76`, `clerk.rs:85-86`,
//! `receive.rs:113-116,140-144,149-152`) do not change.
//!
//! Every construction of `sharing/mod.rs` and `masking/mod.rs` and every trait method is here:
//!
//!   new_share_generator       -> Generator      : ShareGenerator::generate          -> sda_share_generate
//!   new_share_combiner        -> Combiner       : ShareCombiner::combine            -> sda_share_combine_rows[_multi]
//!   new_secret_reconstructor  -> Reconstructor  : SecretReconstructor::reconstruct  -> sda_secret_reconstruct_rows
//!   new_secret_masker         -> Masker         : SecretMasker::mask                -> sda_mask
//!   new_mask_combiner         -> Masker         : MaskCombiner::combine             -> sda_mask_combine
//!   new_secret_unmasker       -> Masker         : SecretUnmasker::unmask            -> sda_unmask
//!
//! Semantics that differ from the CPU implementation, both invisible to the callers above:
//!   * outputs are canonical residues in [0, m) -- what `RecipientOutput::positive()` (receive.rs:13-21) yields;
//!   * `OsRng` is read once per call (32 bytes) instead of once per draw; the library expands the seed with the
//!     rand-0.3 `ChaChaRng` keystream and draws with `gen_range` in the reference's order.
//!
//! SOURCE ONLY: not compiled in the image this repository is developed in (no Rust toolchain there).

use std::rc::Rc;

use rand::{OsRng, Rng};
use sda_b200_sys as b200;
use sda_protocol::{LinearMaskingScheme, LinearSecretSharingScheme};

use super::{Mask, MaskedSecret, Secret, Share};
use super::masking::{MaskCombiner, SecretMasker, SecretUnmasker};
use super::sharing::{SecretReconstructor, ShareCombiner, ShareGenerator};
use errors::SdaClientResult;

/// Devices the clerk-side sum may use: `SDA_B200_DEVICES=0,1,2,3` (default: device 0 only).
fn devices() -> Vec<i32> {
    ::std::env::var("SDA_B200_DEVICES").ok()
        .map(|s| s.split(',').filter_map(|d| d.trim().parse().ok()).collect::<Vec<i32>>())
        .filter(|v| !v.is_empty())
        .unwrap_or_else(|| vec![0])
}

thread_local! {
    // one context per thread (a context is not thread-safe; the crypto traits carry no Send/Sync bounds)
    static CONTEXT: Rc<b200::Context> = {
        let devs = devices();
        let ctx = if devs.len() > 1 { b200::Context::multi(&devs) } else { b200::Context::new(devs[0]) };
        Rc::new(ctx.expect("libsda_b200: no usable CUDA device (there is no CPU fallback)"))
    };
}

fn context() -> Rc<b200::Context> { CONTEXT.with(|c| c.clone()) }

/// 32 bytes from the OS, where the reference opens `OsRng` (additive.rs:17, full.rs:16, chacha.rs:29, tss `share`)
fn seed() -> [u8; 32] {
    let mut s = [0u8; 32];
    OsRng::new().expect("Unable to get randomness source").fill_bytes(&mut s);
    s
}

pub fn sharing_scheme(scheme: &LinearSecretSharingScheme) -> b200::SharingScheme {
    match *scheme {
        LinearSecretSharingScheme::Additive { share_count, modulus } => b200::SharingScheme::additive(share_count, modulus),
        LinearSecretSharingScheme::PackedShamir { secret_count, share_count, privacy_threshold, prime_modulus,
                                                  omega_secrets, omega_shares } =>
            b200::SharingScheme::packed_shamir(secret_count, share_count, privacy_threshold, prime_modulus,
                                               omega_secrets, omega_shares),
    }
}

pub fn masking_scheme(scheme: &LinearMaskingScheme) -> b200::MaskingScheme {
    match *scheme {
        LinearMaskingScheme::None => b200::MaskingScheme::none(),
        LinearMaskingScheme::Full { modulus } => b200::MaskingScheme::full(modulus),
        LinearMaskingScheme::ChaCha { modulus, dimension, seed_bitsize } => b200::MaskingScheme::chacha(modulus, dimension, seed_bitsize),
    }
}

// ---- sharing (crypto/sharing/mod.rs:10-33) -----------------------------------------------------------------------

pub struct Generator { ctx: Rc<b200::Context>, scheme: b200::SharingScheme }

impl Generator {
    pub fn new(scheme: &LinearSecretSharingScheme) -> SdaClientResult<Generator> {
        let g = Generator { ctx: context(), scheme: sharing_scheme(scheme) };
        g.ctx.validate(&g.scheme).map_err(|e| e.message)?;
        Ok(g)
    }
}

impl ShareGenerator for Generator {
    /// batched.rs:18-53 over additive.rs:32-51 / packed_shamir.rs:40-43: row r is the `Vec<Share>` for clerk r
    fn generate(&mut self, secrets: &[Secret]) -> SdaClientResult<Vec<Vec<Share>>> {
        Ok(self.ctx.share_generate(&self.scheme, secrets, &seed()).map_err(|e| e.message)?)
    }
}

pub struct Combiner { ctx: Rc<b200::Context>, scheme: b200::SharingScheme }

impl Combiner {
    pub fn new(scheme: &LinearSecretSharingScheme) -> SdaClientResult<Combiner> {
        Ok(Combiner { ctx: context(), scheme: sharing_scheme(scheme) })
    }
}

impl ShareCombiner for Combiner {
    /// combiner.rs:15-29; "Wrong dimension" comes back as the error text.  With SDA_B200_DEVICES naming several GPUs the
    /// participants' rows are sharded over them and summed with one NCCL reduce inside the library.
    fn combine(&self, shares: &Vec<Vec<Share>>) -> SdaClientResult<Vec<Share>> {
        Ok(self.ctx.share_combine_rows(&self.scheme, shares).map_err(|e| e.message)?)
    }
}

pub struct Reconstructor { ctx: Rc<b200::Context>, scheme: b200::SharingScheme, dimension: usize }

impl Reconstructor {
    pub fn new(scheme: &LinearSecretSharingScheme, dimension: usize) -> SdaClientResult<Reconstructor> {
        Ok(Reconstructor { ctx: context(), scheme: sharing_scheme(scheme), dimension: dimension })
    }
}

impl SecretReconstructor for Reconstructor {
    /// additive.rs:55-73 ("Mismatching dimension") / batched.rs:68-97 + packed_shamir.rs:73-77
    /// ("Not enough shares to reconstruct"); any subset of >= t + k clerks works, the output is truncated to `dimension`
    fn reconstruct(&self, indexed_shares: &Vec<(usize, Vec<Share>)>) -> SdaClientResult<Vec<Secret>> {
        Ok(self.ctx.secret_reconstruct_rows(&self.scheme, self.dimension, indexed_shares).map_err(|e| e.message)?)
    }
}

// ---- masking (crypto/masking/mod.rs:9-31): infallible by signature, the reference asserts -> so does the shim ------

pub struct Masker { ctx: Rc<b200::Context>, scheme: b200::MaskingScheme, dimension: usize }

impl Masker {
    pub fn new(scheme: &LinearMaskingScheme) -> SdaClientResult<Masker> {
        let dimension = match *scheme { LinearMaskingScheme::ChaCha { dimension, .. } => dimension, _ => 0 };
        Ok(Masker { ctx: context(), scheme: masking_scheme(scheme), dimension: dimension })
    }
}

impl SecretMasker for Masker {
    /// none.rs:13-19 / full.rs:21-35 / chacha.rs:24-54 (the ChaCha mask is its seed words, zero-extended to i64)
    fn mask(&mut self, secrets: &[Secret]) -> (Vec<Mask>, Vec<MaskedSecret>) {
        match self.ctx.mask(&self.scheme, secrets, &seed()) {
            Ok(pair) => pair,
            Err(e) => panic!("{}", e.message),      // chacha.rs:26 assert_eq!, gen_range's assert!(low < high)
        }
    }
}

impl MaskCombiner for Masker {
    /// none.rs:21-26 / full.rs:37-52 / chacha.rs:56-77 (re-expands every participant's seed on the GPU)
    fn combine(&self, masks: &Vec<Vec<Mask>>) -> Vec<Mask> {
        let dim = if self.dimension > 0 { self.dimension } else { masks.first().map(|m| m.len()).unwrap_or(0) };
        match self.ctx.mask_combine(&self.scheme, masks, dim) {
            Ok(sum) => sum,
            Err(e) => panic!("{}", e.message),      // none.rs:23, full.rs:43 assertions
        }
    }
}

impl SecretUnmasker for Masker {
    /// none.rs:28-33 / full.rs:54-66 / chacha.rs:79-92
    fn unmask(&self, values: &(Vec<Mask>, Vec<MaskedSecret>)) -> Vec<Secret> {
        match self.ctx.unmask(&self.scheme, &values.0, &values.1) {
            Ok(secrets) => secrets,
            Err(e) => panic!("{}", e.message),      // none.rs:30, full.rs:58, chacha.rs:83 assertions
        }
    }
}

// ---- optional: the participant's two steps in one call --------------------------------------------------------------
/// `participate.rs:53-54` (`secret_masker.mask(&secrets)`) followed by `:75-76` (`share_generator.generate(&masked)`):
/// returns `(mask, shares per clerk)` with the same values as the two trait calls, but uploads the secrets once and --
/// for the 2^61-1 shapes -- never writes the masked secrets anywhere.  `new_participation` can call this instead of
/// the two lines when it wants the fast path; the traits above keep working unchanged.
pub fn mask_and_share(masking: &LinearMaskingScheme, sharing: &LinearSecretSharingScheme, secrets: &[Secret])
                      -> SdaClientResult<(Vec<Mask>, Vec<Vec<Share>>)> {
    let ctx = context();
    let (ms, ss) = (masking_scheme(masking), sharing_scheme(sharing));
    ctx.validate(&ss).map_err(|e| e.message)?;
    Ok(ctx.mask_share_generate(&ms, &ss, secrets, &seed(), &seed()).map_err(|e| e.message)?)
}

// ---- the six constructions: the new bodies of crypto/sharing/mod.rs:35-96 and crypto/masking/mod.rs:33-94 ---------
//
// impl ShareGeneratorConstruction<LinearSecretSharingScheme> for CryptoModule {
//     fn new_share_generator(&self, scheme: &LinearSecretSharingScheme) -> SdaClientResult<Box<ShareGenerator>> {
//         Ok(Box::new(b200::Generator::new(scheme)?))
//     }
// }
// impl ShareCombinerConstruction<LinearSecretSharingScheme> for CryptoModule {
//     fn new_share_combiner(&self, scheme: &LinearSecretSharingScheme) -> SdaClientResult<Box<ShareCombiner>> {
//         Ok(Box::new(b200::Combiner::new(scheme)?))
//     }
// }
// impl SecretReconstructorConstruction<LinearSecretSharingScheme> for CryptoModule {
//     fn new_secret_reconstructor(&self, scheme: &LinearSecretSharingScheme, dimension: usize) -> SdaClientResult<Box<SecretReconstructor>> {
//         Ok(Box::new(b200::Reconstructor::new(scheme, dimension)?))
//     }
// }
// impl SecretMaskerConstruction<LinearMaskingScheme> for CryptoModule {
//     fn new_secret_masker(&self, scheme: &LinearMaskingScheme) -> SdaClientResult<Box<SecretMasker>> {
//         Ok(Box::new(b200::Masker::new(scheme)?))
//     }
// }
// impl MaskCombinerConstruction<LinearMaskingScheme> for CryptoModule {
//     fn new_mask_combiner(&self, scheme: &LinearMaskingScheme) -> SdaClientResult<Box<MaskCombiner>> {
//         Ok(Box::new(b200::Masker::new(scheme)?))
//     }
// }
// impl SecretUnmaskerConstruction<LinearMaskingScheme> for CryptoModule {
//     fn new_secret_unmasker(&self, scheme: &LinearMaskingScheme) -> SdaClientResult<Box<SecretUnmasker>> {
//         Ok(Box::new(b200::Masker::new(scheme)?))
//     }
// }
