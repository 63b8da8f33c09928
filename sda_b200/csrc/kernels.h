// kernels.h -- internal launch interface between the C-ABI host layer (api.cu) and the
// sm_100a kernels.  Every launcher enqueues on `stream`, never synchronises, bumps *nlaunch
// once per kernel launched and returns the launch status.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "chacha.cuh"
#include "field.cuh"

namespace sda {

constexpr int MAX_W = 16;   // k + t   (columns of the share matrix)
constexpr int MAX_N = 32;   // share_count (rows of the share matrix / columns of R)
constexpr int MAX_K = 16;   // secret_count (rows of R)

// n x w share matrix (or k x m' reconstruction matrix), canonical entries, row-major
struct Matrix {
    uint64_t e[MAX_N * MAX_W];
    int rows, cols;
};

struct LaunchCtx {
    cudaStream_t stream;
    uint64_t *nlaunch;
    int sm_count;
    const char **kernel_name;   // variant label of the last sharing launch
};

// ---- K3: column-wise modular sums -------------------------------------------------------
// out[i] = (sum_p rows[p*ld + i] + acc_in[i]) mod m, canonical.  `scratch` (>= scratch_elems
// i64) holds per-slice partial sums when the participant axis is split across CTAs.
cudaError_t launch_combine(const LaunchCtx &lc, const FieldParams &f, const int64_t *rows, size_t ld, size_t P,
                           size_t L, const int64_t *acc_in, int64_t *out, int64_t *scratch, size_t scratch_elems);
size_t combine_scratch_elems(int sm_count, size_t P, size_t L);

// out[i] = in[i] mod m (canonical)
cudaError_t launch_mod_reduce(const LaunchCtx &lc, const FieldParams &f, const int64_t *in, size_t n, int64_t *out,
                              bool input_unsigned = false);
// out[i] = (a[i] - b[i]) mod m  (unmask, full.rs:60-63)
cudaError_t launch_submod(const LaunchCtx &lc, const FieldParams &f, const int64_t *a, const int64_t *b, size_t n,
                          int64_t *out);

// ---- K1: additive split / masks ---------------------------------------------------------
// flag[0] |= 1 when some draw was rejected by gen_range (caller then uses the exact path).
// shares_out[P][n][dim]; keys[P] device array.  draws != nullptr: read pre-drawn canonical
// samples draws[p][dim*(n-1)] instead of generating them (exact path).
cudaError_t launch_additive_split(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds,
                                  int n, const int64_t *secrets, size_t ld, size_t P, size_t dim,
                                  const ChaChaKey *keys, const uint64_t *draws, int64_t *shares_out,
                                  unsigned *flag, uint32_t *d_key_scratch = nullptr);
// d_key_scratch (optional): packed_share_tc2_key_scratch_bytes(P) bytes for the keys' first-round constants (chacha_pre.cuh)
// mask_out (may be null) [dim], masked_out [dim]: Full mask (full.rs:24-31) and the
// participant side of the ChaCha mask (chacha.rs:36-45)
cudaError_t launch_mask(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds,
                        const int64_t *secrets, size_t dim, const ChaChaKey &key, const uint64_t *draws,
                        int64_t *mask_out, int64_t *masked_out, unsigned *flag, const float *fx = nullptr, int frac_bits = 0);
// fx != nullptr: the secrets are the fixed-point encodings of fx[dim] (sda_fixed_encode_dev's definition), computed in
// the same pass; `secrets` is ignored
// ChaCha mask re-expansion (chacha.rs:60-73): out[i] = sum_p draw_p(i) mod m over P keys
cudaError_t launch_chacha_mask_combine(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr,
                                       const ChaChaKey *keys, size_t P, size_t dim, int64_t *out,
                                       int64_t *scratch, size_t scratch_elems, unsigned *flag);
size_t chacha_mask_combine_scratch_elems(int sm_count, size_t P, size_t dim);

// exact gen_range stream: out[0..count) = first `count` accepted samples of the key's stream.
// Needs scratch of draw_exact_scratch_elems(count) u64.  Device-side only, no host sync:
// *status (device) gets 1 if the provisioned stream window was too short (caller retries
// with a larger `window`).
cudaError_t launch_draw_exact(const LaunchCtx &lc, const DrawParams &dr, int rounds, const ChaChaKey &key,
                              size_t count, size_t window, uint64_t *out, uint64_t *scratch, unsigned *status);
size_t draw_exact_scratch_elems(size_t window);

// ---- K2: packed-Shamir share generation ---------------------------------------------------
// shares_out[P][n][B], B = ceil(dim/k).  Mtx is n x (k+t); d_mat is its device copy (row-major
// u64), needed only with draws != nullptr.
cudaError_t launch_packed_share(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int k,
                                int t, int n, const Matrix &mtx, const int64_t *secrets, size_t ld, size_t P,
                                size_t dim, const ChaChaKey *keys, const uint64_t *draws, const uint64_t *d_mat,
                                int64_t *shares_out, unsigned *flag);
// the Mersenne-61 instantiation of the above (packed_m61.cu); shapes of packed_share_has_fast_path
cudaError_t launch_packed_share_m61(const LaunchCtx &lc, int rounds, int k, int t, int n, const Matrix &mtx,
                                    const int64_t *secrets, size_t ld, size_t P, size_t dim, const ChaChaKey *keys,
                                    int64_t *shares_out, unsigned *flag);
// the tensor-core instantiation (packed_tc.cu): D = bytes(x) . bytes(M 2^8c)^T with tcgen05.mma.kind::i8.
// The constant operand is built on the host (image_bytes > 0 iff the shape is instantiated) and
// passed as a 16-byte aligned device copy.
size_t packed_share_tc_image_bytes(int k, int t, int n);
void packed_share_tc_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img);
// any prime below 2^63 (f = the field, dr = gen_range over [0, p - 1)); 2^61 - 1 takes the shift-and-add path.
// Generates batches first_batch .. first_batch + n_batches - 1 (clipped to the vector) of every participant;
// first_batch is a multiple of packed_share_tc_slice_batches(k, t, n); pointers address whole vectors.
size_t packed_share_tc_slice_batches(int k, int t, int n);
cudaError_t launch_packed_share_tc(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int k, int t,
                                   int n, const int64_t *secrets, size_t ld, size_t P, size_t dim, size_t first_batch,
                                   size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *shares_out,
                                   unsigned *flag);
// the paired-tile generation of the same kernel (packed_tc2.cu): 2^61 - 1 only, same slicing contract; two operand
// images (image_bytes covers both).  `supported` also bounds the vector so that a participant's n share rows stay
// within 32-bit byte offsets.
bool packed_share_tc2_supported(int k, int t, int n, size_t dim);
size_t packed_share_tc2_image_bytes(int k, int t, int n);
void packed_share_tc2_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img);
size_t packed_share_tc2_slice_batches(int k, int t, int n);
// d_key_scratch: packed_share_tc2_key_scratch_bytes(P) bytes of device memory, 16-byte aligned (per-participant
// constants of the keystream's first round, written by a small kernel ahead of the main one)
size_t packed_share_tc2_key_scratch_bytes(size_t P);
cudaError_t launch_packed_share_tc2(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets, size_t ld,
                                    size_t P, size_t dim, size_t first_batch, size_t n_batches, const ChaChaKey *keys,
                                    uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *shares_out, unsigned *flag);
// share generation fused with the clerk sums on the paired-tile machinery (packed_tc2f.cu): 2^61 - 1, the instantiated
// shapes; operand images as packed_share_tc2_build_image, d_key_scratch as launch_packed_share_tc2
bool packed_share_combine_tc2_supported(int k, int t, int n, size_t dim);
cudaError_t launch_packed_share_combine_tc2(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets,
                                            size_t ld, size_t P, size_t dim, const ChaChaKey *keys, uint32_t *d_key_scratch,
                                            const uint8_t *d_b_image, const int64_t *acc_in, int64_t *out, unsigned *flag);
// mask -> share generation in one kernel (packed_tc2m.cu): both schemes over 2^61 - 1, 20 rounds, the instantiated shapes.
// mask_keys[P]: the key of every participant's mask stream; mask_out: [P][dim] (Full scheme) or nullptr;
// d_key_scratch: twice packed_share_tc2_key_scratch_bytes(P).  Whole vectors only.
bool packed_share_tc2_masked_supported(int k, int t, int n, size_t dim, int rounds);
cudaError_t launch_packed_share_tc2_masked(const LaunchCtx &lc, int k, int t, int n, const int64_t *secrets, size_t ld, size_t P,
                                           size_t dim, const ChaChaKey *share_keys, const ChaChaKey *mask_keys,
                                           uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *mask_out,
                                           int64_t *shares_out, unsigned *flag);
// the same kernel with the share count n <= 32 as a run-time value, per (k, t) with k <= 8, t <= 8 (packed_tc2n.cu);
// 2^61 - 1 and 20 rounds only
bool packed_share_tc2n_supported(int k, int t, int n, size_t dim, int rounds);
size_t packed_share_tc2n_image_bytes(int k, int t, int n);
void packed_share_tc2n_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img);
size_t packed_share_tc2n_slice_batches(int k, int t);
cudaError_t launch_packed_share_tc2n(const LaunchCtx &lc, int k, int t, int n, const int64_t *secrets, size_t ld, size_t P,
                                     size_t dim, size_t first_batch, size_t n_batches, const ChaChaKey *keys,
                                     uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *shares_out, unsigned *flag);
// the run-time-shaped kernel (packed_tcg.cu): any k + t <= 16, n <= 32, t >= 1, any prime; same slicing contract
bool packed_share_tcg_supported(int k, int t, int n);
size_t packed_share_tcg_image_bytes(int k, int t, int n);
void packed_share_tcg_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img);
size_t packed_share_tcg_slice_batches(int k, int t, int n);
cudaError_t launch_packed_share_tcg(const LaunchCtx &lc, const FieldParams &f, const DrawParams &dr, int rounds, int k, int t,
                                    int n, const int64_t *secrets, size_t ld, size_t P, size_t dim, size_t first_batch,
                                    size_t n_batches, const ChaChaKey *keys, const uint8_t *d_b_image, int64_t *shares_out,
                                    unsigned *flag);
// fused: out[n][B] = acc_in[n][B] + sum over the P participants of their shares, accumulated in TMEM
cudaError_t launch_packed_share_combine_tc(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets,
                                           size_t ld, size_t P, size_t dim, const ChaChaKey *keys, const uint8_t *d_b_image,
                                           const int64_t *acc_in, int64_t *out, unsigned *flag,
                                           uint32_t *d_key_scratch = nullptr);
// true when launch_packed_share / launch_additive_split have an in-kernel-rng instantiation
bool packed_share_has_fast_path(int k, int t, int n);
bool additive_split_has_fast_path(int n);

// ---- reveal: secrets = R . shares ------------------------------------------------------------
// shares[m][ld] -> secrets_out[dimension]; R is k x m.
cudaError_t launch_packed_reconstruct(const LaunchCtx &lc, const FieldParams &f, int k, int m, const Matrix &R,
                                      const int64_t *shares, size_t ld, size_t dimension, int64_t *secrets_out);

// the same map over 2^61 - 1 on the tensor cores (reveal_tc.cu), any k <= 16, m <= 16; the constant
// operand is built on the host and passed as a 16-byte aligned device copy
bool reveal_tc_supported(int k, int m);
size_t reveal_tc_image_bytes(int k, int m);
void reveal_tc_build_image(int k, int m, const Matrix &R, uint8_t *img);
cudaError_t launch_reveal_tc(const LaunchCtx &lc, int k, int m, const int64_t *shares, size_t ld, size_t dimension,
                             const uint8_t *d_b_image, int64_t *secrets_out);

// ---- share wire codec (codec.cu): zig-zag LEB128 varints, encryption/sodium.rs:35-41,83-90 ----------
// `scratch` holds *_scratch_elems u64; the encoded length / value count is left in scratch[elems - 1]
size_t varint_encode_scratch_elems(size_t n);
size_t varint_decode_scratch_elems(size_t len);
cudaError_t launch_varint_encode(const LaunchCtx &lc, const int64_t *in, size_t n, uint8_t *out, uint64_t *scratch);
cudaError_t launch_varint_decode(const LaunchCtx &lc, const uint8_t *buf, size_t len, int64_t *out, size_t cap,
                                 uint64_t *scratch, unsigned *status);

// ---- fixed-point codec of real-valued vectors (misc.cu; definition in include/sda_b200.h) ------------
cudaError_t launch_fixed_encode(const LaunchCtx &lc, const FieldParams &f, int frac_bits, const float *x, size_t n,
                                int64_t *out);
cudaError_t launch_fixed_decode(const LaunchCtx &lc, const FieldParams &f, int frac_bits, uint64_t divisor,
                                const int64_t *in, size_t n, float *out);

// ---- server snapshot transpose (snapshot.cu; server/src/snapshot.rs:11-27, stores.rs:86-101) ---------------------
// blob (p, c) of `in` (bytes [in_off[p n + c], in_off[p n + c + 1])) -> `out` at out_off[c P + p]; offsets are device arrays
cudaError_t launch_snapshot_transpose(const LaunchCtx &lc, const uint8_t *in, const uint64_t *d_in_off, size_t P, size_t n,
                                      uint8_t *out, const uint64_t *d_out_off);

// ---- synthetic inputs ---------------------------------------------------------------------
cudaError_t launch_synth_fill(const LaunchCtx &lc, const FieldParams &f, uint32_t stream_id, uint64_t start,
                              size_t count, int64_t *out);

}  // namespace sda
