// packed_tc2n.cu -- the paired-tile share-generation kernel (packed_tc2.cuh) with the SHARE COUNT as a run-time value:
// instantiated for every (k, t) with k + t <= 16 (120 kernels), p = 2^61 - 1, ChaCha20.  Operand images and accumulators
// are sized for 8 shares (two 64-column accumulators, the TMEM footprint of the fully templated shapes); a scheme with
// more shares, up to 32, runs its shares through the accumulators in groups of 8, one pair of operand images per group.
// A scheme here runs at nearly the speed of a fully templated shape (the fold takes four shares per TMEM load instead of n;
// odd t stores its draws 8 bytes at a time); other primes and 8 / 12 rounds take the run-time-shaped kernel of
// packed_tcg.cu.
//
// This file compiles once per value of t (-DSDA_TC2N_T=t, Makefile) so that the kernels build in parallel, and once
// without it for the dispatch below.
#include "packed_tc2.cuh"

namespace sda {

namespace {
constexpr int NCAP = 8;      // shares per group: 8 n <= 64 columns per accumulator
constexpr int NMAX = NCAP * SDA_TC2_MAX_GROUPS;
constexpr int KTMAX = 16;     // k + t: the limb plan's bound (w5_for)
}

#define SDA_TC2N_K(X, T) X(1, T) X(2, T) X(3, T) X(4, T) X(5, T) X(6, T) X(7, T) X(8, T) X(9, T) X(10, T) X(11, T) X(12, T) \
    X(13, T) X(14, T) X(15, T)

#ifdef SDA_TC2N_T

#define SDA_CAT_(a, b) a##b
#define SDA_CAT(a, b) SDA_CAT_(a, b)

// the kernels of one t: k = 1 .. 16 - t
template <int K, int T>
cudaError_t launch_kt(const LaunchCtx &lc, int n, const int64_t *secrets, size_t ld, size_t P, size_t dim, size_t first_batch,
                      size_t n_batches, const ChaChaKey *keys, uint32_t *d_key_scratch, const uint8_t *d_b_image,
                      int64_t *shares_out, unsigned *flag) {
    if constexpr (K + T <= KTMAX)
        return launch2<K, T, NCAP, 20, true>(lc, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch, d_b_image,
                                             shares_out, flag, n);
    else
        return cudaErrorInvalidValue;
}
template <int K, int T>
void group_image_kt(int n_group, const Matrix &rows, uint64_t p, uint8_t *img) {
    if constexpr (K + T <= KTMAX) build_b_image2<K, T, NCAP>(rows, p, img, n_group);
}
template <int K, int T>
constexpr size_t group_bytes_kt() {
    if constexpr (K + T <= KTMAX) return 2 * Shape2<K, T, NCAP>::B_IMG;
    else return 0;
}
template <int K, int T>
constexpr size_t pass_kt() {
    if constexpr (K + T <= KTMAX) return (size_t)Shape2<K, T, NCAP>::PASS;
    else return 0;
}

cudaError_t SDA_CAT(launch_packed_share_tc2n_t, SDA_TC2N_T)(const LaunchCtx &lc, int k, int n, const int64_t *secrets, size_t ld,
                                                            size_t P, size_t dim, size_t first_batch, size_t n_batches,
                                                            const ChaChaKey *keys, uint32_t *d_key_scratch,
                                                            const uint8_t *d_b_image, int64_t *shares_out, unsigned *flag) {
#define X(K, T)                                                                                                           \
    if (k == K)                                                                                                           \
        return launch_kt<K, T>(lc, n, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch, d_b_image, shares_out, flag);
    SDA_TC2N_K(X, SDA_TC2N_T)
#undef X
    return cudaErrorInvalidValue;
}

void SDA_CAT(packed_share_tc2n_group_image_t, SDA_TC2N_T)(int k, int n_group, const Matrix &rows, uint64_t p, uint8_t *img) {
#define X(K, T) if (k == K) return group_image_kt<K, T>(n_group, rows, p, img);
    SDA_TC2N_K(X, SDA_TC2N_T)
#undef X
}

size_t SDA_CAT(packed_share_tc2n_group_bytes_t, SDA_TC2N_T)(int k) {
#define X(K, T) if (k == K) return group_bytes_kt<K, T>();
    SDA_TC2N_K(X, SDA_TC2N_T)
#undef X
    return 0;
}

size_t SDA_CAT(packed_share_tc2n_pass_t, SDA_TC2N_T)(int k) {
#define X(K, T) if (k == K) return pass_kt<K, T>();
    SDA_TC2N_K(X, SDA_TC2N_T)
#undef X
    return 0;
}

#else

#define SDA_TC2N_TS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)
#define X(T)                                                                                                              \
    cudaError_t launch_packed_share_tc2n_t##T(const LaunchCtx &, int, int, const int64_t *, size_t, size_t, size_t, size_t, \
                                              size_t, const ChaChaKey *, uint32_t *, const uint8_t *, int64_t *, unsigned *); \
    void packed_share_tc2n_group_image_t##T(int, int, const Matrix &, uint64_t, uint8_t *);                                \
    size_t packed_share_tc2n_group_bytes_t##T(int);                                                                        \
    size_t packed_share_tc2n_pass_t##T(int);
SDA_TC2N_TS(X)
#undef X

bool packed_share_tc2n_supported(int k, int t, int n, size_t dim, int rounds) {
    if (k < 1 || t < 1 || k + t > KTMAX || n < 1 || n > NMAX || rounds != 20) return false;
    const size_t B = (dim + (size_t)k - 1) / (size_t)k;
    return ((B * (size_t)t + 7) / 8 >> 32) == 0;         // a participant's keystream stays below 2^32 blocks
}

static size_t group_bytes(int k, int t) {
#define X(T) if (t == T) return packed_share_tc2n_group_bytes_t##T(k);
    SDA_TC2N_TS(X)
#undef X
    return 0;
}

size_t packed_share_tc2n_image_bytes(int k, int t, int n) { return (size_t)((n + NCAP - 1) / NCAP) * group_bytes(k, t); }

// group g holds the operand images (E, O) of shares 8 g .. 8 g + 7
void packed_share_tc2n_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img) {
    const size_t gb = group_bytes(k, t);
    for (int g = 0; g * NCAP < n; g++) {
        const int ng = std::min(NCAP, n - g * NCAP);
        Matrix rows;
        rows.rows = ng;
        rows.cols = mtx.cols;
        memcpy(rows.e, mtx.e + (size_t)g * NCAP * (k + t), sizeof(uint64_t) * (size_t)ng * (k + t));
#define X(T) if (t == T) packed_share_tc2n_group_image_t##T(k, ng, rows, p, img + (size_t)g * gb);
        SDA_TC2N_TS(X)
#undef X
    }
}

size_t packed_share_tc2n_slice_batches(int k, int t) {
#define X(T) if (t == T) return packed_share_tc2n_pass_t##T(k);
    SDA_TC2N_TS(X)
#undef X
    return 0;
}

cudaError_t launch_packed_share_tc2n(const LaunchCtx &lc, int k, int t, int n, const int64_t *secrets, size_t ld, size_t P,
                                     size_t dim, size_t first_batch, size_t n_batches, const ChaChaKey *keys,
                                     uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *shares_out, unsigned *flag) {
    if (k < 1 || t < 1 || k + t > KTMAX || n < 1 || n > NMAX) return cudaErrorInvalidValue;
    *lc.kernel_name = "packed_share<k,t templated, n<=32 at run time>/mersenne61 tcgen05.mma.kind::i8, paired tiles";
#define X(T)                                                                                                        \
    if (t == T)                                                                                                     \
        return launch_packed_share_tc2n_t##T(lc, k, n, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch, \
                                             d_b_image, shares_out, flag);
    SDA_TC2N_TS(X)
#undef X
    return cudaErrorInvalidValue;
}

#endif

}  // namespace sda
