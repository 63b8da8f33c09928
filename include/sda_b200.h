/*
 * sda_b200.h -- C ABI of libsda_b200.so: the B200-native replacement for the data-parallel
 * hot path of snipsco/sda (participant-side masking + secret sharing, clerk-side share
 * summation, recipient-side reconstruction + unmasking).
 *
 * The reference (Rust, CPU) has no FFI seam; the seam is its trait surface.  Each entry point
 * below is what a Rust shim implementing that trait on `CryptoModule` binds (bindings/rust,
 * INTEGRATION.md).  Citations are paths relative to the reference tree (snipsco/sda @ 8cf97f2).
 *
 *   trait / type                                   reference                                entry point
 *   ---------------------------------------------  ---------------------------------------  --------------------------
 *   LinearSecretSharingScheme                      protocol/src/crypto.rs:79-114            sda_sharing_scheme
 *   LinearMaskingScheme                            protocol/src/crypto.rs:43-64             sda_masking_scheme
 *   input_size/output_size/..._threshold           protocol/src/crypto.rs:117-155           sda_input_size ...
 *   ShareGenerator::generate                       client/src/crypto/sharing/mod.rs:14-17   sda_share_generate
 *     (batched.rs:18-53, additive.rs:32-51, packed_shamir.rs:40-43)
 *   ShareCombiner::combine                         client/src/crypto/sharing/mod.rs:23-25   sda_share_combine[_rows]
 *     (combiner.rs:15-29)
 *   SecretReconstructor::reconstruct               client/src/crypto/sharing/mod.rs:31-33   sda_secret_reconstruct[_rows]
 *     (additive.rs:55-73, batched.rs:68-97, packed_shamir.rs:73-77)
 *   SecretMasker::mask                             client/src/crypto/masking/mod.rs:13-15   sda_mask
 *     (none.rs:13-19, full.rs:21-35, chacha.rs:24-54)
 *   MaskCombiner::combine                          client/src/crypto/masking/mod.rs:21-23   sda_mask_combine
 *     (none.rs:21-26, full.rs:37-52, chacha.rs:56-77)
 *   SecretUnmasker::unmask                         client/src/crypto/masking/mod.rs:29-31   sda_unmask
 *     (none.rs:28-33, full.rs:54-66, chacha.rs:79-92)
 *   RecipientOutput::positive                      client/src/receive.rs:13-21              (outputs are already canonical)
 *   ShareEncryptor/Decryptor varint loops          client/src/crypto/encryption/sodium.rs   sda_varint_encode / _decode
 *     (:35-41, :83-90; the sealed box itself stays with libsodium)
 *
 * Conventions
 *   - Element type is int64_t (`Secret/Mask/MaskedSecret/Share = i64`, client/src/crypto/mod.rs:33-36).
 *   - Inputs may be ANY i64 (negative, >= modulus).  Outputs are always canonical residues in
 *     [0, modulus) -- i.e. what the reference yields after RecipientOutput::positive(); the
 *     reference's raw outputs are signed representatives of the same classes.
 *   - Caller allocates every output; sizes follow from the scheme (helpers below).  Nothing is
 *     allocated across the boundary, no callbacks, no global mutable state outside sda_ctx.
 *   - Randomness is injected: `rng_seed` is 32 bytes of caller entropy (the Rust shim fills it
 *     from OsRng).  The result is exactly the reference algorithm run with its `OsRng` replaced
 *     by rand-0.3 `ChaChaRng::from_seed(seed as 8 LE u32 words)` (ChaCha with the context's
 *     round count, default 20) and draws taken by `Rng::gen_range` in the reference's order.
 *   - Return value: SDA_OK, or an error class; sda_last_error() gives the message, which for
 *     class SDA_ERR_INVALID is the reference's own Err string / panic message.
 *   - A context is not thread-safe; use one per thread (they are cheap) or lock externally.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SDA_ERR_CUDA.
 */
#ifndef SDA_B200_H
#define SDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDA_B200_ABI_VERSION 1

enum {
    SDA_OK = 0,
    SDA_ERR_INVALID = 1,     /* the reference returns Err(..) or panics on this input */
    SDA_ERR_CUDA = 2,        /* CUDA runtime / driver failure (incl. no device) */
    SDA_ERR_NCCL = 3,        /* NCCL failure (incl. libnccl.so.2 not loadable) in a multi-GPU entry point */
    SDA_ERR_UNSUPPORTED = 4, /* parameters outside what the kernels implement */
    SDA_ERR_REJECTED = 5     /* deferred checks only: gen_range rejected a word in an earlier *_dev call (see below) */
};

enum { SDA_SHARING_ADDITIVE = 0, SDA_SHARING_PACKED_SHAMIR = 1 };
enum { SDA_MASK_NONE = 0, SDA_MASK_FULL = 1, SDA_MASK_CHACHA = 2 };

/* protocol/src/crypto.rs:79-114.  Additive uses share_count + modulus only. */
typedef struct {
    int32_t  kind;
    uint64_t share_count, secret_count, privacy_threshold;
    int64_t  modulus;        /* `modulus` | `prime_modulus` */
    int64_t  omega_secrets, omega_shares;
} sda_sharing_scheme;

/* protocol/src/crypto.rs:43-64 */
typedef struct {
    int32_t  kind;
    int64_t  modulus;
    uint64_t dimension, seed_bitsize;   /* ChaCha only */
} sda_masking_scheme;

typedef struct sda_ctx sda_ctx;

/* ---- context -------------------------------------------------------------------------- */
int         sda_abi_version(void);
int         sda_ctx_create(int device, sda_ctx **out);
void        sda_ctx_destroy(sda_ctx *ctx);
/* message of the last failing call on this context (ctx == NULL: of the last failing
 * sda_ctx_create on this thread) */
const char *sda_last_error(const sda_ctx *ctx);
/* ChaCha rounds of the injected sharing / Full-mask randomness: 8, 12 or 20 (default 20).
 * The ChaCha *mask scheme* (chacha.rs) is wire format and always uses 20.
 * SECURITY: this randomness is what hides the secrets.  Fewer than 20 rounds trades margin for speed (12 is what
 * rand >= 0.8 `StdRng` uses, 8 has no margin to speak of); leave the default unless the deployment has decided
 * otherwise.  Every call needs its OWN 32-byte rng_seed from a CSPRNG: the same seed given to sda_mask and to
 * sda_share_generate yields the same keystream for both (no domain separation), i.e. correlated mask and shares. */
int         sda_ctx_set_rng_rounds(sda_ctx *ctx, int rounds);
int         sda_ctx_get_rng_rounds(const sda_ctx *ctx);
/* Which kernels evaluate the packed-Shamir maps over 2^61-1 (share generation for the instantiated
 * shapes, reconstruction for k, m' <= 16): the tcgen05 (tensor-core, byte-limb GEMM) ones or the
 * CUDA-core ones.  All are exact and produce identical results; AUTO picks the tensor-core kernels.
 * TENSOR_CORES_V1 keeps share generation on the first-generation tensor-core kernel (one batch per thread
 * and tile) instead of the paired-tile one, and TENSOR_CORES_ANY_SHAPE sends even the instantiated shapes to the
 * kernels that serve every other (k, t, n); both for side-by-side measurements. */
enum { SDA_PACKED_PATH_AUTO = 0, SDA_PACKED_PATH_CUDA_CORES = 1, SDA_PACKED_PATH_TENSOR_CORES = 2,
       SDA_PACKED_PATH_TENSOR_CORES_V1 = 3, SDA_PACKED_PATH_TENSOR_CORES_ANY_SHAPE = 4 };
int         sda_ctx_set_packed_path(sda_ctx *ctx, int path);
/* stream (cudaStream_t) the *_dev entry points launch on; default: a context-owned stream */
int         sda_ctx_set_stream(sda_ctx *ctx, void *cuda_stream);
void       *sda_ctx_get_stream(const sda_ctx *ctx);
int         sda_ctx_synchronize(sda_ctx *ctx);
/* Deferred rejection checks.  A *_dev entry point that draws randomness normally ends with one stream synchronisation:
 * it reads the flag that says whether rand-0.3's gen_range would have rejected a keystream word (probability ~2^-57 per
 * draw), and redoes the call on the exact path if so.  With deferred checks ON those entry points only QUEUE their work
 * and the copy of their flag word and return at once (calls overlap, small vectors do not pay a round trip each);
 * sda_ctx_synchronize() then waits for the stream and looks at every queued flag: SDA_OK, or SDA_ERR_REJECTED naming
 * the first affected call, whose outputs the caller must recompute with deferred checks off.  Host-buffer entry points,
 * the varint codec and an in-place sda_share_generate_combine_dev are never deferred.  Switching the mode checks what is
 * pending and returns that status. */
int         sda_ctx_set_deferred_checks(sda_ctx *ctx, int on);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t    sda_ctx_launch_count(const sda_ctx *ctx);
/* kernel variant the last sharing call dispatched to ("packed<3,2,5>/mersenne61", ...) */
const char *sda_ctx_last_kernel(const sda_ctx *ctx);
/* pinned host memory for callers that want DMA-speed host entry points (optional).  With pinned input AND output
 * buffers sda_share_generate (tensor-core packed shapes) and sda_share_combine[_rows] / additive reconstruct / Full mask
 * combine walk vectors of 4 MB and more in 8 slices on three streams, so the copy in, the kernel and the copy out of
 * neighbouring slices overlap (PCIe is full duplex); results are identical, pageable buffers take the plain path. */
int         sda_host_alloc(sda_ctx *ctx, size_t bytes, void **out);
int         sda_host_free(sda_ctx *ctx, void *ptr);

/* ---- derived scheme properties (protocol/src/crypto.rs:117-155) ------------------------ */
size_t sda_input_size(const sda_sharing_scheme *s);
size_t sda_output_size(const sda_sharing_scheme *s);
size_t sda_privacy_threshold(const sda_sharing_scheme *s);
size_t sda_reconstruction_threshold(const sda_sharing_scheme *s);
/* ceil(dim / input_size): length of each clerk's share vector (batched.rs:23) */
size_t sda_share_batches(const sda_sharing_scheme *s, size_t dim);
/* length of the mask SecretMasker::mask returns: 0 | dim | ceil(seed_bitsize/32) */
size_t sda_mask_len(const sda_masking_scheme *s, size_t dim);
/* 0 if the scheme is usable, else an error class (message via sda_last_error(ctx)) */
int    sda_sharing_scheme_validate(sda_ctx *ctx, const sda_sharing_scheme *s);
/* the n x (k+t) share matrix M (row-major, canonical) with shares = M.[secrets;randomness]:
 * what packed_shamir.rs:42 (tss share) evaluates per batch.  Diagnostic / test hook. */
int    sda_packed_share_matrix(sda_ctx *ctx, const sda_sharing_scheme *s, int64_t *out);
/* the k x m reconstruction matrix R for a clerk index subset (packed_shamir.rs:76) */
int    sda_packed_reconstruct_matrix(sda_ctx *ctx, const sda_sharing_scheme *s,
                                     const uint64_t *indices, size_t m, int64_t *out);

/* ---- host-pointer entry points (the drop-in trait methods) ----------------------------- */

/* ShareGenerator::generate.  shares_out is [output_size][ceil(dim/input_size)], clerk-major
 * (row r is the Vec<Share> addressed to clerk r). */
int sda_share_generate(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *secrets, size_t dim,
                       const uint8_t rng_seed[32], int64_t *shares_out);

/* ShareCombiner::combine on a contiguous participant-major matrix shares[P][L]. */
int sda_share_combine(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *shares, size_t P, size_t L,
                      int64_t *out /* [L] */);
/* Same on `&Vec<Vec<Share>>` as it lies in Rust memory: P row pointers + lengths.  Rows whose
 * length differs from row 0's fail with "Wrong dimension" (combiner.rs:21).  *out_len = row 0's
 * length (0 when P == 0). */
int sda_share_combine_rows(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *const *rows,
                           const size_t *row_lens, size_t P, int64_t *out, size_t *out_len);

/* SecretReconstructor::reconstruct on shares[m][B] (row s = clerk indices[s]'s combined vector).
 * Additive: column sum, indices ignored, *out_len = B (additive.rs:55-73).
 * Packed:   *out_len = dimension; needs m >= reconstruction_threshold
 *           ("Not enough shares to reconstruct") and B >= ceil(dimension/k). */
int sda_secret_reconstruct(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                           const int64_t *shares, size_t m, size_t B, int64_t *secrets_out, size_t *out_len);
int sda_secret_reconstruct_rows(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension,
                                const uint64_t *indices, const int64_t *const *rows, const size_t *row_lens,
                                size_t m, int64_t *secrets_out, size_t *out_len);

/* SecretMasker::mask -> (mask, masked).  mask_out needs sda_mask_len() elements. */
int sda_mask(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *secrets, size_t dim,
             const uint8_t rng_seed[32], int64_t *mask_out, size_t *mask_len, int64_t *masked_out);
/* MaskCombiner::combine on masks[P][mask_len]; out needs dim (Full: mask_len; ChaCha:
 * scheme.dimension) elements. */
int sda_mask_combine(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *masks, size_t P, size_t mask_len,
                     int64_t *out, size_t *out_len);
/* The participant's two steps in one call (client/src/participate.rs:53-54 SecretMasker::mask, then :75-76
 * ShareGenerator::generate on the masked secrets): mask_out[sda_mask_len(ms, dim)], shares_out[output_size][B].  Same
 * results as sda_mask followed by sda_share_generate; the masked secrets stay on the device (never cross the link, and
 * where sda_mask_share_generate_dev's fused kernel applies never reach its memory either). */
int sda_mask_share_generate(sda_ctx *ctx, const sda_masking_scheme *ms, const sda_sharing_scheme *ss, const int64_t *secrets,
                            size_t dim, const uint8_t mask_rng_seed[32], const uint8_t share_rng_seed[32], int64_t *mask_out,
                            int64_t *shares_out);
/* SecretUnmasker::unmask on (mask, masked). */
int sda_unmask(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *mask, size_t mask_len,
               const int64_t *masked, size_t dim, int64_t *out);

/* ---- device-pointer entry points (what the benchmark times) ---------------------------- */
/* All pointers are device memory of the context's device; launches go to the context's
 * stream.  Calls that draw randomness synchronise that stream once at the end to read the
 * gen_range rejection flag (and redo the call on the exact path if it is set) -- unless the context
 * runs with deferred checks (sda_ctx_set_deferred_checks). */

/* ShareGenerator::generate for P participants at once: secrets[P][dim] (row stride
 * secrets_ld elements), seeds[P][32] (HOST memory), shares_out[P][output_size][B]. */
int sda_share_generate_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_secrets, size_t secrets_ld,
                           size_t P, size_t dim, const uint8_t *seeds, int64_t *d_shares_out);

/* The participant's two steps in one call for P participants (client/src/participate.rs:53-54: SecretMasker::mask, then
 * :75-76: ShareGenerator::generate on the masked secrets): secrets[P][dim] (row stride secrets_ld), mask_rng_seeds[P][32]
 * and share_rng_seeds[P][32] (HOST memory), masks_out[P][sda_mask_len(ms, dim)] (Full: the masks; ChaCha: the seed words;
 * None: unused), shares_out[P][output_size][B].  Same results as sda_mask_dev followed by sda_share_generate_dev per
 * participant.  Where both schemes are over 2^61-1 and the sharing scheme has an instantiated shape the masked secrets are
 * never written to memory: the masks are drawn and added while the secrets are staged as operand rows of the share GEMM. */
int sda_mask_share_generate_dev(sda_ctx *ctx, const sda_masking_scheme *ms, const sda_sharing_scheme *ss,
                                const int64_t *d_secrets, size_t secrets_ld, size_t P, size_t dim,
                                const uint8_t *mask_rng_seeds, const uint8_t *share_rng_seeds, int64_t *d_masks_out,
                                int64_t *d_shares_out);

/* ShareCombiner::combine: out[i] = sum_p shares[p*ld + i] mod m.  If d_acc_in != NULL it is
 * added as one more row (streaming tiles into a running sum: clerk.rs:71-72 FIXME). */
int sda_share_combine_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_shares, size_t ld, size_t P,
                          size_t L, const int64_t *d_acc_in, int64_t *d_out);

/* Fused participant->clerk path for one box: shares of P participants are
 * generated and summed per clerk without materialising [P][n][B]:
 * out[n][B] (+= d_acc_in[n][B] if given) = sum_p generate(secrets[p]) .
 * d_acc_in is NULL, equal to d_out (in place), or disjoint from it. */
int sda_share_generate_combine_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_secrets,
                                   size_t secrets_ld, size_t P, size_t dim, const uint8_t *seeds,
                                   const int64_t *d_acc_in, int64_t *d_out);

/* x -> x mod m over a vector (final pass after an NCCL sum of canonical partial sums). */
int sda_mod_reduce_dev(sda_ctx *ctx, int64_t modulus, const int64_t *d_in, size_t n, int64_t *d_out);
/* same with the input read as u64: the pass after an ncclSum of G canonical partial sums,
 * exact while G * modulus <= 2^64 (8 GPUs at a 61-bit modulus). */
int sda_mod_reduce_u64_dev(sda_ctx *ctx, int64_t modulus, const uint64_t *d_in, size_t n, int64_t *d_out);

/* d_shares[m][ld], the first B columns of each row are that clerk's vector */
int sda_secret_reconstruct_dev(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                               const int64_t *d_shares, size_t ld, size_t m, size_t B, int64_t *d_secrets_out);

int sda_mask_dev(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_secrets, size_t dim,
                 const uint8_t rng_seed[32], int64_t *d_mask_out, int64_t *d_masked_out);
int sda_mask_combine_dev(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_masks, size_t P,
                         size_t mask_len, int64_t *d_out);
int sda_unmask_dev(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_mask, const int64_t *d_masked,
                   size_t dim, int64_t *d_out);

/* ---- multi-GPU clerk sum --------------------------------------------------------------------- */
/* The one step of the path with an exchange (ShareCombiner::combine, combiner.rs:15-29, called at clerk.rs:85-86):
 * the participants' rows are sharded over the GPUs of a box, every GPU sums its rows, and the canonical partial sums
 * are added over NVLink by ONE NCCL collective inside the library (ncclReduce of u64 + one mod pass on the root while
 * ranks * modulus <= 2^64, e.g. 8 GPUs at a 61-bit modulus; otherwise ncclAllGather + the modular combine kernel).
 * NCCL is bound at run time (dlopen of libnccl.so.2, the process's copy if it has one); hosts that never call these
 * entry points do not need it.  Two forms:
 *
 * (1) one process per GPU (what bench.py --gpus N and an MPI-style host use): rank 0 calls sda_nccl_unique_id, the
 *     host application hands the 128 bytes to the other ranks, every rank calls sda_ctx_comm_init_rank on its own
 *     context (collective), then sda_share_combine_ranks_dev / sda_partial_sums_reduce_dev (collective, in the same
 *     order on every rank, launched on each context's stream). */
#define SDA_NCCL_UNIQUE_ID_BYTES 128
int sda_nccl_unique_id(uint8_t id_out[SDA_NCCL_UNIQUE_ID_BYTES]);
int sda_ctx_comm_init_rank(sda_ctx *ctx, const uint8_t id[SDA_NCCL_UNIQUE_ID_BYTES], int nranks, int rank);
int sda_ctx_comm_rank(const sda_ctx *ctx);
int sda_ctx_comm_size(const sda_ctx *ctx);   /* 1 without a communicator */
/* d_partials[count]: this rank's canonical column sums, in place; afterwards the root's buffer holds the sums over
 * all ranks mod `modulus` (canonical), the other ranks' buffers are unspecified.  No-op for a single rank. */
int sda_partial_sums_reduce_dev(sda_ctx *ctx, int64_t modulus, int64_t *d_partials, size_t count, int root);
/* sda_share_combine_dev on this rank's rows (P_local may be 0), then the exchange: root's d_out = the clerk sum
 * over every rank's rows. */
int sda_share_combine_ranks_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_shares, size_t ld,
                                size_t P_local, size_t L, const int64_t *d_acc_in, int64_t *d_out, int root);
/* (2) one process, several GPUs (what a Rust clerk binds: no launcher, no torch): the returned context is member 0
 *     (device devices[0]) and owns one member context per further device plus the communicator; every single-GPU
 *     entry point keeps working on it (on device devices[0]).  sda_ctx_destroy releases all of it. */
int sda_ctx_create_multi(const int *devices, int ndev, sda_ctx **out);
int      sda_ctx_multi_count(const sda_ctx *ctx);          /* 1 for a plain context */
sda_ctx *sda_ctx_multi_member(sda_ctx *ctx, int i);        /* member i's context (its device, its stream) */
/* rows resident on the devices: d_shares[i] = member i's [P_per_device[i]][ld] block (device i memory),
 * d_partials[i] = L elements of scratch on device i (i >= 1; entry 0 unused), d_out = L elements on device 0. */
int sda_share_combine_multi_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *const *d_shares, size_t ld,
                                const size_t *P_per_device, size_t L, int64_t *const *d_partials, int64_t *d_out);
/* ShareCombiner::combine on `&Vec<Vec<Share>>` host rows (sda_share_combine_rows' contract): contiguous blocks of
 * participants go to the devices over their own PCIe links (one host thread per device), then the exchange. */
int sda_share_combine_rows_multi(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *const *rows,
                                 const size_t *row_lens, size_t P, int64_t *out, size_t *out_len);

/* ---- share wire codec: the step either side of the path ------------------------------------ */
/* ShareEncryptor::encrypt's encoding loop (client/src/crypto/encryption/sodium.rs:35-41) and
 * ShareDecryptor::decrypt's decoding loop (sodium.rs:83-90): `integer-encoding 1.0` varints, i.e.
 * zig-zag then unsigned LEB128, values concatenated without a count.  The sealed box around the
 * bytes stays with libsodium.  `out` of encode needs sda_varint_max_bytes(n) = 10 n bytes.
 * decode fails with SDA_ERR_INVALID on a stream that ends inside a value, holds a value longer than
 * 10 bytes (the reference's decode_var would read past a u64 there), or holds more than `cap` values. */
size_t sda_varint_max_bytes(size_t n);
int sda_varint_encode(sda_ctx *ctx, const int64_t *shares, size_t n, uint8_t *out, size_t *out_len);
int sda_varint_decode(sda_ctx *ctx, const uint8_t *buf, size_t len, int64_t *shares_out, size_t cap, size_t *n);
/* device buffers; the length / count comes back through a host pointer (one stream synchronisation) */
int sda_varint_encode_dev(sda_ctx *ctx, const int64_t *d_shares, size_t n, uint8_t *d_out, size_t *out_len);
int sda_varint_decode_dev(sda_ctx *ctx, const uint8_t *d_buf, size_t len, int64_t *d_shares_out, size_t cap, size_t *n);

/* ---- server snapshot transpose: the layout producer of the clerk's matrix --------------------------- */
/* server/src/snapshot.rs:11-27 -> stores.rs:86-101 (iter_snapshot_clerk_jobs_data): every participation holds one
 * encrypted share vector per clerk; a snapshot regroups them into one job per clerk.  Here the blobs are byte strings
 * concatenated in device memory: blob (p, c) of d_blobs is bytes [offsets[p n + c], offsets[p n + c + 1]) (P n + 1
 * offsets, HOST memory, participation-major); d_out receives clerk 0's P blobs, then clerk 1's, ... and
 * out_offsets[c P + p] (P n + 1 entries, HOST, filled by the call) says where blob (p, c) went.  d_out needs
 * offsets[P n] - offsets[0] bytes.  Byte-exact data movement; which bytes they are (sealed boxes, varints, raw i64
 * rows) is the caller's business.  With equal-length raw rows the clerk's rows need no transpose at all:
 * sda_share_combine_dev takes the strided view shares[:, c, :] directly. */
int sda_snapshot_transpose_dev(sda_ctx *ctx, const uint8_t *d_blobs, const uint64_t *offsets, size_t P, size_t n,
                               uint8_t *d_out, uint64_t *out_offsets);

/* ---- fixed-point codec for real-valued vectors (model updates) -------------------------------- */
/* Not in the reference, whose API takes Vec<i64> (client/src/participate.rs:10,25): the adjacent step
 * BASELINE config #5 needs.  encode: rint(x * 2^frac_bits) (ties to even) as a residue in [0, m);
 * decode: centred lift in (-m/2, m/2], divided by 2^frac_bits and by `divisor` (e.g. the participant
 * count, for a mean) in double, rounded to float.  frac_bits in [0, 52].  Device buffers. */
int sda_fixed_encode_dev(sda_ctx *ctx, int64_t modulus, int frac_bits, const float *d_x, size_t n, int64_t *d_out);
int sda_fixed_decode_dev(sda_ctx *ctx, int64_t modulus, int frac_bits, uint64_t divisor, const int64_t *d_in, size_t n,
                         float *d_out);
/* sda_fixed_encode_dev followed by sda_mask_dev (participate.rs:53-54 on a real-valued update) in ONE pass over the
 * vector, for P participants per call: d_x[P][x_ld] float32 in, d_masked_out[P][masked_ld] residues out, seeds[P][32]
 * (HOST), d_mask_out[P][mask_len] (Full: the masks; ChaCha: the seed words; may be NULL).  Reads 4 B, writes 8 B per
 * element, the fixed-point vector never exists in memory; the P kernels are queued back to back with one read-back at the
 * end.  Same results as the two calls per participant; `modulus` is the encoding modulus and must equal the masking
 * scheme's (None has none of its own). */
int sda_fixed_encode_mask_dev(sda_ctx *ctx, const sda_masking_scheme *s, int64_t modulus, int frac_bits, const float *d_x,
                              size_t x_ld, size_t P, size_t dim, const uint8_t *seeds, int64_t *d_mask_out,
                              int64_t *d_masked_out, size_t masked_ld);

/* Synthetic benchmark / test inputs, generated on the device: out[i] = value(start + i) where
 * value(e) = u64 draw e of ChaCha20(key "sda-b200-synthetic-v1", key word 7 = stream) mod m.
 * Same definition as the oracle's sdao_synth_fill. */
int sda_synth_fill_dev(sda_ctx *ctx, uint32_t stream, int64_t modulus, uint64_t start, size_t count,
                       int64_t *d_out);

#ifdef __cplusplus
}
#endif
#endif /* SDA_B200_H */
