// packed_tc2.cu -- the paired-tile share-generation kernel (packed_tc2.cuh) instantiated for the shapes BASELINE names
#include "packed_tc2.cuh"

namespace sda {

#define SDA_TC2_SHAPES(X) X(3, 2, 5) X(5, 4, 9) X(3, 4, 7) X(3, 4, 8)

// shapes and sizes this instantiation serves (p = 2^61 - 1)
bool packed_share_tc2_supported(int k, int t, int n, size_t dim) {
    const size_t B = (dim + (size_t)k - 1) / (size_t)k;
    if ((B * (size_t)t + 7) / 8 >> 32) return false;     // a participant's keystream stays below 2^32 blocks
#define X(K, T, N) if (k == K && t == T && n == N) return true;
    SDA_TC2_SHAPES(X)
#undef X
    return false;
}

size_t packed_share_tc2_image_bytes(int k, int t, int n) {
#define X(K, T, N) if (k == K && t == T && n == N) return 2 * Shape2<K, T, N>::B_IMG;
    SDA_TC2_SHAPES(X)
#undef X
    return 0;
}

void packed_share_tc2_build_image(int k, int t, int n, const Matrix &mtx, uint64_t p, uint8_t *img) {
#define X(K, T, N) if (k == K && t == T && n == N) return build_b_image2<K, T, N>(mtx, p, img);
    SDA_TC2_SHAPES(X)
#undef X
}

size_t packed_share_tc2_slice_batches(int k, int t, int n) {
#define X(K, T, N) if (k == K && t == T && n == N) return (size_t)Shape2<K, T, N>::PASS;
    SDA_TC2_SHAPES(X)
#undef X
    return 0;
}

size_t packed_share_tc2_key_scratch_bytes(size_t P) { return P * sizeof(ChaChaPre); }

cudaError_t launch_packed_share_tc2(const LaunchCtx &lc, int rounds, int k, int t, int n, const int64_t *secrets, size_t ld,
                                    size_t P, size_t dim, size_t first_batch, size_t n_batches, const ChaChaKey *keys,
                                    uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *shares_out, unsigned *flag) {
#define X(K, T, N)                                                                                              \
    if (k == K && t == T && n == N) {                                                                           \
        *lc.kernel_name = "packed_share<" #K "," #T "," #N ">/mersenne61 tcgen05.mma.kind::i8, paired tiles";   \
        if (rounds == 8) return launch2<K, T, N, 8>(lc, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch, d_b_image, shares_out, flag); \
        if (rounds == 12) return launch2<K, T, N, 12>(lc, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch, d_b_image, shares_out, flag); \
        return launch2<K, T, N, 20>(lc, secrets, ld, P, dim, first_batch, n_batches, keys, d_key_scratch, d_b_image, shares_out, flag); \
    }
    SDA_TC2_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sda
