/*
 * sda_oracle.h -- CPU ORACLE for the SDA sharing / masking / clerk-sum / reveal hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the checker
 * (or as the timed CPU baseline).  libsda_b200.so never links, loads or calls it.
 *
 * It restates, in plain C, the algorithm of the reference (snipsco/sda @ 8cf97f2):
 *   - in-tree code literally (signed i64, truncating %), citing client/src/crypto/...:line
 *   - the external crates the path depends on, which are NOT in /root/reference:
 *       threshold-secret-sharing 0.2 (client/Cargo.toml:15)  -> sdao_tss_*
 *       rand 0.3                      (client/Cargo.toml:18) -> sdao_chacha_*, sdao_gen_range
 *       integer-encoding 1.0          (client/Cargo.toml:17) -> sdao_varint_*
 *     restated from their published algorithms; pinned by the known-answer vectors in
 *     tests/golden/ (ChaCha20 keystream KAT, tss polynomial/share KATs) and by the
 *     reference's own result-pinning tests (integration-tests/tests/full_loop.rs:148,
 *     README.md:157).
 *
 * Parity status: in-tree functions = pinned by the reference's goldens; individual share
 * values of tss and stream values of rand are pinned by KATs of the published algorithms
 * (the crates cannot be run here: no Rust toolchain, crates not vendored).
 */
#ifndef SDA_ORACLE_H
#define SDA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- rand 0.3 model ------------------------------------------------------------------ */

/* kind 0: ChaChaRng::from_seed (deterministic; parity mode);  kind 1: OsRng (getrandom(2)
 * per draw; faithful timing mode). */
typedef struct {
    int      kind;
    int      rounds;      /* 20 in rand 0.3; other values only for the spec extension */
    uint32_t state[16];   /* constants | key | 128-bit block counter */
    uint32_t buf[16];
    int      idx;         /* next word in buf; 16 = exhausted */
    uint64_t draws;       /* number of u64 drawn, incl. rejected (diagnostics) */
    uint64_t rejections;
} sdao_rng;

void     sdao_rng_from_seed(sdao_rng *r, const uint32_t *seed_words, size_t n_words);
void     sdao_rng_from_seed_rounds(sdao_rng *r, const uint32_t *seed_words, size_t n_words, int rounds);
void     sdao_rng_os(sdao_rng *r);
uint32_t sdao_rng_next_u32(sdao_rng *r);
uint64_t sdao_rng_next_u64(sdao_rng *r);
int64_t  sdao_gen_range(sdao_rng *r, int64_t low, int64_t high);
/* raw block function, for the keystream KAT */
void     sdao_chacha_block(const uint32_t state[16], int rounds, uint32_t out[16]);

/* ---- scheme descriptors (protocol/src/crypto.rs:43-64, 79-114) -------------------------- */

enum { SDAO_SHARING_ADDITIVE = 0, SDAO_SHARING_PACKED_SHAMIR = 1 };
enum { SDAO_MASK_NONE = 0, SDAO_MASK_FULL = 1, SDAO_MASK_CHACHA = 2 };

typedef struct {
    int32_t  kind;
    uint64_t share_count, secret_count, privacy_threshold;
    int64_t  modulus, omega_secrets, omega_shares;
} sdao_sharing_scheme;

typedef struct {
    int32_t  kind;
    int64_t  modulus;
    uint64_t dimension, seed_bitsize;
} sdao_masking_scheme;

size_t sdao_input_size(const sdao_sharing_scheme *s);
size_t sdao_output_size(const sdao_sharing_scheme *s);
size_t sdao_privacy_threshold(const sdao_sharing_scheme *s);
size_t sdao_reconstruction_threshold(const sdao_sharing_scheme *s);

/* ---- tss 0.2 model ----------------------------------------------------------------------- */

int64_t sdao_mod_pow(int64_t x, uint64_t e, int64_t p);
int64_t sdao_mod_inverse(int64_t k, int64_t p);
/* radix-2 / radix-3 transforms; n must be a power of 2 / 3.  in/out length n. */
void    sdao_fft2(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out);
void    sdao_fft2_inverse(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out);
void    sdao_fft3(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out);
void    sdao_fft3_inverse(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out);

/* PackedSecretSharing::share with explicit randomness (tss test helper shape):
 * values = [0] ++ secrets ++ randomness -> polynomial -> shares.  Uses the FFT path when
 * (k+t+1) is a power of two and (n+1) a power of three, else the general
 * interpolate-then-evaluate path (spec extension).  force_general != 0 forces the latter. */
int     sdao_tss_share_with_randomness(const sdao_sharing_scheme *s, const int64_t *secrets,
                                       const int64_t *randomness, int force_general,
                                       int64_t *poly_out /* k+t+1 or NULL */,
                                       int64_t *shares_out /* n */);
/* PackedSecretSharing::reconstruct */
int     sdao_tss_reconstruct(const sdao_sharing_scheme *s, const uint64_t *indices,
                             const int64_t *shares, size_t m, int64_t *secrets_out /* k */);

/* ---- the trait surface (client/src/crypto/sharing, masking) --------------------------- */

/* error codes: 0 ok, 1 = the reference returns Err / panics (message via sdao_last_error) */
const char *sdao_last_error(void);

/* ShareGenerator::generate  (batched.rs:18-53 + additive.rs:32-51 | packed_shamir.rs:40-43).
 * shares_out is [output_size][ceil(dim/input_size)], clerk-major. */
int sdao_share_generate(const sdao_sharing_scheme *s, const int64_t *secrets, size_t dim,
                        sdao_rng *rng, int64_t *shares_out);
/* same map, evaluated through a precomputed n x (k+t) matrix built with the oracle's own
 * Newton machinery (share of unit vectors).  Used as the *fair* CPU baseline: the literal
 * path recomputes the interpolation per batch. */
int sdao_share_generate_matrix(const sdao_sharing_scheme *s, const int64_t *secrets, size_t dim,
                               sdao_rng *rng, int64_t *shares_out);

/* ShareCombiner::combine (combiner.rs:15-29); rows[p] + L elements each, row stride ld. */
int sdao_share_combine(int64_t modulus, const int64_t *shares, size_t P, size_t L, size_t ld,
                       int64_t *out);

/* SecretReconstructor::reconstruct (additive.rs:55-73 | batched.rs:68-97 + packed_shamir.rs:73-77)
 * shares is [m][B] (row stride B). */
int sdao_secret_reconstruct(const sdao_sharing_scheme *s, size_t dimension,
                            const uint64_t *indices, const int64_t *shares, size_t m, size_t B,
                            int64_t *secrets_out, size_t *out_len);

/* SecretMasker::mask (none.rs:13-19 | full.rs:21-35 | chacha.rs:24-54) */
int sdao_mask(const sdao_masking_scheme *s, const int64_t *secrets, size_t dim, sdao_rng *rng,
              int64_t *mask_out, size_t *mask_len, int64_t *masked_out);
/* MaskCombiner::combine (none.rs:21-26 | full.rs:37-52 | chacha.rs:56-77); masks [P][mask_len] */
int sdao_mask_combine(const sdao_masking_scheme *s, const int64_t *masks, size_t P,
                      size_t mask_len, int64_t *out, size_t *out_len);
/* SecretUnmasker::unmask (none.rs:28-33 | full.rs:54-66 | chacha.rs:79-92) */
int sdao_unmask(const sdao_masking_scheme *s, const int64_t *mask, size_t mask_len,
                const int64_t *masked, size_t dim, int64_t *out);

/* RecipientOutput::positive (client/src/receive.rs:13-21) */
void sdao_positive(int64_t modulus, int64_t *values, size_t n);
/* full canonical residue ((x % m) + m) % m -- the parity comparison map */
void sdao_canonical(int64_t modulus, int64_t *values, size_t n);

/* ---- integer-encoding 1.0 model (sodium.rs:34-41, 83-90) ------------------------------- */
size_t sdao_varint_encode(const int64_t *values, size_t n, uint8_t *out /* >= 10 n */);
/* returns number of values decoded, or (size_t)-1 on a truncated buffer */
size_t sdao_varint_decode(const uint8_t *buf, size_t len, int64_t *out, size_t max_out);

/* ---- synthetic benchmark inputs (shared definition with the CUDA side) ------ */
/* value(e) = next_u64 of ChaCha20(key = "sda-b200-synthetic-v1" zero padded, key word 7 =
 * stream) at draw index e, reduced  % modulus.  Fills out[0..count) for e = start.. */
/* fixed-point codec of real-valued vectors (not in the reference) */
void sdao_fixed_encode(const float *x, size_t n, int frac_bits, int64_t p, int64_t *out);
void sdao_fixed_decode(const int64_t *in, size_t n, int frac_bits, int64_t p, uint64_t divisor, float *out);

void sdao_synth_fill(uint32_t stream, int64_t modulus, uint64_t start, size_t count, int64_t *out);

/* ---- parameter search helpers (orders / generators in Z_p^*) --------------------------- */
/* smallest c >= 2 whose power c^((p-1)/q) has order exactly q (q | p-1, q given with its
 * prime factorisation implicit: verified by checking w^q == 1 and w^(q/f) != 1 for every
 * prime f | q).  returns 0 if q does not divide p-1. */
int64_t sdao_find_root_of_order(int64_t p, uint64_t q);

#ifdef __cplusplus
}
#endif
#endif
