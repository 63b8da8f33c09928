// packed_tc2n.cu -- the paired-tile share-generation kernel (packed_tc2.cuh) with the SHARE COUNT as a run-time value:
// instantiated per (k, t) for k = 1..8 and t = 2, 4 with operand images and accumulators sized for up to 8 shares
// (two 64-column accumulators, four CTAs per SM like the fully templated shapes), p = 2^61 - 1, ChaCha20.  A scheme
// whose (k, t) is here and whose n <= 8 runs at nearly the speed of a fully templated shape (the fold loses its unrolled
// 32-column TMEM loads, nothing else); everything else -- odd t, k > 8, n > 8, other primes, 8 / 12 rounds -- takes the
// run-time-shaped kernel of packed_tcg.cu.
#include "packed_tc2.cuh"

namespace sda {

namespace {
constexpr int NCAP = 8;      // share-count capacity: 8 n <= 64 columns per accumulator
}

#define SDA_TC2N_KT(X) X(1, 2) X(2, 2) X(3, 2) X(4, 2) X(5, 2) X(6, 2) X(7, 2) X(8, 2) \
                       X(1, 4) X(2, 4) X(3, 4) X(4, 4) X(5, 4) X(6, 4) X(7, 4) X(8, 4)

bool packed_share_tc2n_supported(int k, int t, int n, size_t dim, int rounds) {
    if (n < 1 || n > NCAP || rounds != 20) return false;
    const size_t B = (dim + (size_t)k - 1) / (size_t)k;
    if ((B * (size_t)t + 7) / 8 >> 32) return false;     // a participant's keystream stays below 2^32 blocks
#define X(K, T) if (k == K && t == T) return true;
    SDA_TC2N_KT(X)
#undef X
    return false;
}

size_t packed_share_tc2n_image_bytes(int k, int t) {
#define X(K, T) if (k == K && t == T) return 2 * Shape2<K, T, NCAP>::B_IMG;
    SDA_TC2N_KT(X)
#undef X
    return 0;
}

void packed_share_tc2n_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img) {
#define X(K, T) if (k == K && t == T) return build_b_image2<K, T, NCAP>(mtx, p, img, n);
    SDA_TC2N_KT(X)
#undef X
}

size_t packed_share_tc2n_slice_batches(int k, int t) {
#define X(K, T) if (k == K && t == T) return (size_t)Shape2<K, T, NCAP>::PASS;
    SDA_TC2N_KT(X)
#undef X
    return 0;
}

cudaError_t launch_packed_share_tc2n(const LaunchCtx &lc, int k, int t, int n, const int64_t *secrets, size_t ld, size_t P,
                                     size_t dim, size_t first_batch, size_t n_batches, const ChaChaKey *keys,
                                     uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *shares_out, unsigned *flag) {
    if (n < 1 || n > NCAP) return cudaErrorInvalidValue;
#define X(K, T)                                                                                                       \
    if (k == K && t == T) {                                                                                           \
        *lc.kernel_name = "packed_share<" #K "," #T ",n<=8 at run time>/mersenne61 tcgen05.mma.kind::i8, paired tiles"; \
        return launch2<K, T, NCAP, 20, true>(lc, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch,    \
                                             d_b_image, shares_out, flag, n);                                         \
    }
    SDA_TC2N_KT(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sda
