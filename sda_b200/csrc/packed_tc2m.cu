// packed_tc2m.cu -- mask -> share generation in one kernel (participate.rs:53-54 then :75-76): the paired-tile kernel of
// packed_tc2.cuh instantiated with MASKED for the shapes BASELINE names, p = 2^61 - 1 for both the masking and the
// sharing scheme, ChaCha20 for both streams.  The masked secrets are never written to HBM: a pass's masks are drawn and
// added to the raw secrets where they lie in shared memory, one step before they are staged as operand rows.
#include "packed_tc2.cuh"

namespace sda {

#define SDA_TC2M_SHAPES(X) X(3, 2, 5) X(5, 4, 9) X(3, 4, 7) X(3, 4, 8)

bool packed_share_tc2_masked_supported(int k, int t, int n, size_t dim, int rounds) {
    if (rounds != 20) return false;
    const size_t B = (dim + (size_t)k - 1) / (size_t)k;
    if ((B * (size_t)t + 7) / 8 >> 32) return false;     // a participant's keystreams stay below 2^32 blocks
    if (((dim + 7) / 8) >> 32) return false;
#define X(K, T, N) if (k == K && t == T && n == N) return true;
    SDA_TC2M_SHAPES(X)
#undef X
    return false;
}

// d_key_scratch: 2 packed_share_tc2_key_scratch_bytes(P) bytes; operand image: packed_share_tc2_build_image's
cudaError_t launch_packed_share_tc2_masked(const LaunchCtx &lc, int k, int t, int n, const int64_t *secrets, size_t ld, size_t P,
                                           size_t dim, const ChaChaKey *share_keys, const ChaChaKey *mask_keys,
                                           uint32_t *d_key_scratch, const uint8_t *d_b_image, int64_t *mask_out,
                                           int64_t *shares_out, unsigned *flag) {
    const size_t B = (dim + (size_t)k - 1) / (size_t)k;
#define X(K, T, N)                                                                                                          \
    if (k == K && t == T && n == N) {                                                                                       \
        *lc.kernel_name = "mask+packed_share<" #K "," #T "," #N ">/mersenne61 tcgen05.mma.kind::i8, paired tiles, masked secrets in shared memory only"; \
        return launch2<K, T, N, 20, false, true>(lc, secrets, ld, P, dim, 0, B, share_keys, d_key_scratch, d_b_image, shares_out, \
                                                 flag, N, mask_keys, mask_out);                                             \
    }
    SDA_TC2M_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace sda
