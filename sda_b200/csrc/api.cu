// api.cu -- the C ABI of include/sda_b200.h: context, scheme validation with the reference's
// error strings, host-side precomputation (reduction constants, share matrix M, reconstruction
// matrix R), staging of host buffers, and dispatch to the sm_100a kernels.  There is no CPU
// implementation of any hot-path function in this file: every entry point ends in kernel
// launches and fails with SDA_ERR_CUDA when there is no device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sda_b200.h"
#include "kernels.h"

using namespace sda;

namespace {

typedef unsigned __int128 u128;

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct sda_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // copy engines of the sliced host entry points
    std::vector<cudaEvent_t> pipe_ev;                          // their per-slice events, grown on demand
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int rounds = 20;
    uint64_t nlaunch = 0;
    const char *kernel_name = "";
    std::string err;
    DevBuf in, out, aux, scratch, draws, keys, keys_pre, mat, tc_image, tc_image_r, tc2_image, tcg_image, masked_tmp;
    int packed_path = SDA_PACKED_PATH_AUTO;
    bool debug_force_reject = false;   // SDA_B200_DEBUG_FORCE_REJECT=1 at context creation: treat the fused kernel's flag as set
    std::vector<uint64_t> tc_image_key;   // (k, t, n, matrix) the device image was built for
    std::vector<uint64_t> tc2_image_key;  // likewise for the paired-tile kernel's two images (packed_tc2.cu)
    std::vector<uint64_t> tcg_image_key;  // ... and for the run-time-shaped kernel (packed_tcg.cu)
    std::vector<uint64_t> tc_image_r_key; // (k, m', R) likewise for the reconstruction operand
    std::vector<uint64_t> r_key;          // (scheme, clerk subset) of r_cached
    Matrix r_cached;
    std::vector<uint64_t> m_key;          // (scheme, evaluation points) of m_cached
    Matrix m_cached;
    unsigned *d_flag = nullptr;    // [0] rejection flag, [1] draw_exact status
    unsigned *h_flag = nullptr;    // pinned mirror
    // deferred rejection checks (sda_ctx_set_deferred_checks): the *_dev entry points queue their flag word into h_ring and
    // return without synchronising; sda_ctx_synchronize looks at what is pending
    bool deferred = false;
    int dev_entry_depth = 0;       // > 0 inside a *_dev entry point (host entry points never defer: they hand results out)
    unsigned *h_ring = nullptr;    // pinned, DEFER_RING words
    uint64_t seq_issued = 0, seq_checked = 0;
    bool forced_pending = false;   // SDA_B200_DEBUG_FORCE_REJECT=1 seen by a deferred call
    PinBuf stage[2];               // pinned staging for pageable host buffers
    // multi-GPU: the NCCL communicator this context is a rank of, and -- for the one-process form
    // (sda_ctx_create_multi) -- the member contexts of the other devices, owned by member 0
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    std::vector<sda_ctx *> members;   // [0] = this context when non-empty
    LaunchCtx lc() { return LaunchCtx{stream, &nlaunch, sm_count, &kernel_name}; }
};

namespace {

int fail(sda_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, SDA_ERR_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define OK(call)               \
    do {                       \
        int rc_ = (call);      \
        if (rc_) return rc_;   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- host field arithmetic (parameter precomputation only) ------------------------------------
uint64_t mulmod_h(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)((u128)a * b % p); }
uint64_t submod_h(uint64_t a, uint64_t b, uint64_t p) { return a >= b ? a - b : a + p - b; }
uint64_t powmod_h(uint64_t x, uint64_t e, uint64_t p) {
    uint64_t r = 1 % p;
    x %= p;
    while (e) {
        if (e & 1) r = mulmod_h(r, x, p);
        x = mulmod_h(x, x, p);
        e >>= 1;
    }
    return r;
}
// inverse by extended Euclid; returns false when gcd(a, p) != 1
bool invmod_h(uint64_t a, uint64_t p, uint64_t *out) {
    __int128 r0 = p, r1 = a % p, t0 = 0, t1 = 1;
    while (r1 != 0) {
        __int128 q = r0 / r1, tmp = r0 - q * r1;
        r0 = r1;
        r1 = tmp;
        tmp = t0 - q * t1;
        t0 = t1;
        t1 = tmp;
    }
    if (r0 != 1) return false;
    if (t0 < 0) t0 += p;
    *out = (uint64_t)t0;
    return true;
}
uint64_t canon_h(int64_t v, uint64_t p) {
    int64_t r = (int64_t)(v % (int64_t)p);
    return (uint64_t)(r < 0 ? r + (int64_t)p : r);
}

// Lagrange basis value  l_i(x) = prod_{l != i} (x - z_l) / (z_i - z_l)  over nodes z
bool lagrange_h(const std::vector<uint64_t> &z, size_t i, uint64_t x, uint64_t p, uint64_t *out) {
    uint64_t num = 1, den = 1;
    for (size_t l = 0; l < z.size(); l++) {
        if (l == i) continue;
        num = mulmod_h(num, submod_h(x, z[l], p), p);
        den = mulmod_h(den, submod_h(z[i], z[l], p), p);
    }
    uint64_t inv;
    if (!invmod_h(den, p, &inv)) return false;
    *out = mulmod_h(num, inv, p);
    return true;
}

struct Packed {
    int k, t, n;
    uint64_t p;
    std::vector<uint64_t> a;   // secret-side points w_secrets^i, i = 0..k+t
    std::vector<uint64_t> b;   // share-side points  w_shares^j,  j = 1..n
};

const char *kind_name(int kind) { return kind == SDA_SHARING_ADDITIVE ? "Additive" : "PackedShamir"; }

int validate(sda_ctx *ctx, const sda_sharing_scheme *s, Packed *pk) {
    if (!s) return fail(ctx, SDA_ERR_INVALID, "null scheme");
    if (s->kind == SDA_SHARING_ADDITIVE) {
        if (s->share_count == 0)   // additive.rs:42 `share_count - 1` underflows
            return fail(ctx, SDA_ERR_INVALID, "attempt to subtract with overflow (additive.rs:42: share_count = 0)");
        if (s->modulus <= 0)       // rand 0.3 gen_range: assert!(low < high)
            return fail(ctx, SDA_ERR_INVALID, "Rng.gen_range called with low >= high");
        if (s->share_count > 65535)
            return fail(ctx, SDA_ERR_UNSUPPORTED, "share_count %llu > 65535", (unsigned long long)s->share_count);
        return SDA_OK;
    }
    if (s->kind != SDA_SHARING_PACKED_SHAMIR) return fail(ctx, SDA_ERR_INVALID, "unknown sharing scheme kind %d", s->kind);
    const uint64_t k = s->secret_count, t = s->privacy_threshold, n = s->share_count;
    if (k == 0 || n == 0) return fail(ctx, SDA_ERR_INVALID, "PackedShamir needs secret_count >= 1 and share_count >= 1");
    if (k > (uint64_t)MAX_K || k + t > (uint64_t)MAX_W || n > (uint64_t)MAX_N)
        return fail(ctx, SDA_ERR_UNSUPPORTED, "PackedShamir shape k=%llu t=%llu n=%llu outside k<=%d, k+t<=%d, n<=%d",
                    (unsigned long long)k, (unsigned long long)t, (unsigned long long)n, MAX_K, MAX_W, MAX_N);
    if (s->modulus < 3 || (s->modulus & 1) == 0)
        return fail(ctx, SDA_ERR_INVALID, "PackedShamir prime_modulus must be an odd prime");
    const uint64_t p = (uint64_t)s->modulus;
    if (pk) {
        pk->k = (int)k; pk->t = (int)t; pk->n = (int)n; pk->p = p;
        const uint64_t ws = canon_h(s->omega_secrets, p), wh = canon_h(s->omega_shares, p);
        pk->a.resize(k + t + 1);
        pk->b.resize(n);
        for (uint64_t i = 0; i <= k + t; i++) pk->a[i] = powmod_h(ws, i, p);
        for (uint64_t j = 1; j <= n; j++) pk->b[j - 1] = powmod_h(wh, j, p);
        for (size_t i = 0; i < pk->a.size(); i++)
            for (size_t l = 0; l < i; l++)
                if (pk->a[i] == pk->a[l])
                    return fail(ctx, SDA_ERR_INVALID, "omega_secrets has order <= secret_count + privacy_threshold: "
                                                      "interpolation points collide");
        // tss 0.2 evaluates through a radix-2 inverse FFT over k + t + 1 points and a radix-3 FFT over n + 1 points.  When
        // the sizes are FFT sizes the reference computes THAT transform whatever the roots are: only primitive roots of
        // exactly those orders make it the interpolation this library evaluates, so anything else is refused instead of
        // silently producing a different map.  (Other sizes cannot run in tss at all; they are this library's extension.)
        const uint64_t m2 = k + t + 1, m3 = n + 1;
        auto is_pow = [](uint64_t v, uint64_t b) {
            while (v % b == 0) v /= b;
            return v == 1;
        };
        if (is_pow(m2, 2) && is_pow(m3, 3) && m2 <= m3) {
            if (powmod_h(ws, m2, p) != 1 || (m2 > 1 && powmod_h(ws, m2 / 2, p) == 1))
                return fail(ctx, SDA_ERR_INVALID, "omega_secrets is not a primitive %llu-th root of unity: tss 0.2 would run its "
                                                  "radix-2 FFT with it and share a different polynomial", (unsigned long long)m2);
            if (powmod_h(wh, m3, p) != 1 || (m3 > 1 && powmod_h(wh, m3 / 3, p) == 1))
                return fail(ctx, SDA_ERR_INVALID, "omega_shares is not a primitive %llu-th root of unity: tss 0.2 would run its "
                                                  "radix-3 FFT with it and evaluate at different points", (unsigned long long)m3);
        }
        for (size_t i = 0; i < pk->b.size(); i++) {
            if (pk->b[i] == 1) return fail(ctx, SDA_ERR_INVALID, "omega_shares has order <= share_count: share point equals 1");
            for (size_t l = 0; l < i; l++)
                if (pk->b[i] == pk->b[l])
                    return fail(ctx, SDA_ERR_INVALID, "omega_shares has order <= share_count: share points collide");
        }
    }
    return SDA_OK;
}

// M[j][i] = l_{i+1}(b_j) over nodes a_0..a_{k+t}; column of a_0 dropped (value fixed to 0)
int share_matrix(sda_ctx *ctx, const Packed &pk, Matrix *M) {
    const int w = pk.k + pk.t;
    M->rows = pk.n;
    M->cols = w;
    for (int j = 0; j < pk.n; j++)
        for (int i = 0; i < w; i++)
            if (!lagrange_h(pk.a, (size_t)i + 1, pk.b[j], pk.p, &M->e[j * w + i]))
                return fail(ctx, SDA_ERR_INVALID, "prime_modulus is not prime (non-invertible difference of points)");
    return SDA_OK;
}

// the share matrix of the scheme the context used last (a participant shares many vectors under one scheme)
int share_matrix_cached(sda_ctx *ctx, const Packed &pk, Matrix *M);

// R[e][s] = lambda_{s+1}(a_{e+1}) over nodes {1} u {b_{idx_s}}; node-1 column dropped
int reconstruct_matrix(sda_ctx *ctx, const Packed &pk, const uint64_t *indices, size_t m, Matrix *R) {
    if (m > (size_t)MAX_N) return fail(ctx, SDA_ERR_UNSUPPORTED, "more than %d indexed shares", MAX_N);
    std::vector<uint64_t> z(m + 1);
    z[0] = 1;
    for (size_t s = 0; s < m; s++) {
        if (indices[s] >= (uint64_t)pk.n)
            return fail(ctx, SDA_ERR_INVALID, "share index %llu out of range for share_count %d",
                        (unsigned long long)indices[s], pk.n);
        z[s + 1] = pk.b[indices[s]];
        for (size_t l = 0; l < s; l++)
            if (indices[l] == indices[s]) return fail(ctx, SDA_ERR_INVALID, "duplicate share index %llu", (unsigned long long)indices[s]);
    }
    R->rows = pk.k;
    R->cols = (int)m;
    for (int e = 0; e < pk.k; e++)
        for (size_t s = 0; s < m; s++)
            if (!lagrange_h(z, s + 1, pk.a[e + 1], pk.p, &R->e[e * m + s]))
                return fail(ctx, SDA_ERR_INVALID, "prime_modulus is not prime (non-invertible difference of points)");
    return SDA_OK;
}

int share_matrix_cached(sda_ctx *ctx, const Packed &pk, Matrix *M) {
    std::vector<uint64_t> key{(uint64_t)pk.k, (uint64_t)pk.t, (uint64_t)pk.n, pk.p};
    key.insert(key.end(), pk.a.begin(), pk.a.end());
    key.insert(key.end(), pk.b.begin(), pk.b.end());
    if (key != ctx->m_key) {
        Matrix fresh;                              // a failure part-way must not leave a half-written matrix under the old key
        OK(share_matrix(ctx, pk, &fresh));
        ctx->m_cached = fresh;
        ctx->m_key = key;
    }
    *M = ctx->m_cached;
    return SDA_OK;
}

// rand 0.3 ChaChaRng on the host: only for the handful of seed words of the ChaCha mask scheme
// (chacha.rs:30-33 draws them from OsRng; here they come from the injected rng_seed stream)
void host_chacha_block(const ChaChaKey &key, uint64_t block, int rounds, uint32_t out[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) st[4 + i] = key.w[i];
    st[12] = (uint32_t)block;
    st[13] = (uint32_t)(block >> 32);
    st[14] = st[15] = 0;
    uint32_t x[16];
    memcpy(x, st, sizeof x);
    auto rotl = [](uint32_t v, int c) { return (v << c) | (v >> (32 - c)); };
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 16);
        x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl(x[b], 12);
        x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 8);
        x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl(x[b], 7);
    };
    for (int i = 0; i < rounds / 2; i++) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + st[i];
}

// ---- flag handling --------------------------------------------------------------------------
int clear_flags(sda_ctx *ctx) {
    CU(cudaMemsetAsync(ctx->d_flag, 0, 2 * sizeof(unsigned), ctx->stream));
    return SDA_OK;
}
int read_flags(sda_ctx *ctx, unsigned *rejected, unsigned *status) {
    CU(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (rejected) *rejected = ctx->h_flag[0];
    if (status) *status = ctx->h_flag[1];
    return SDA_OK;
}

constexpr uint64_t DEFER_RING = 4096;
struct DevEntry {                      // scope of a *_dev entry point
    sda_ctx *c;
    explicit DevEntry(sda_ctx *ctx) : c(ctx) { c->dev_entry_depth++; }
    ~DevEntry() { c->dev_entry_depth--; }
};
// look at every deferred flag word: synchronises the stream
int resolve_deferred(sda_ctx *ctx) {
    if (ctx->seq_issued == ctx->seq_checked && !ctx->forced_pending) return SDA_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    uint64_t bad = ~0ull;
    for (uint64_t q = ctx->seq_checked; q < ctx->seq_issued; q++)
        if (ctx->h_ring[q % DEFER_RING] != 0 && bad == ~0ull) bad = q;
    if (ctx->forced_pending && bad == ~0ull) bad = ctx->seq_checked;
    ctx->seq_checked = ctx->seq_issued;
    ctx->forced_pending = false;
    if (bad != ~0ull)
        return fail(ctx, SDA_ERR_REJECTED, "gen_range rejected a keystream word in deferred call #%llu (counted from the switch to "
                    "deferred checks): the outputs of that call are not the reference's; redo it with deferred checks off",
                    (unsigned long long)bad);
    return SDA_OK;
}
// the rejection flag of the call just queued: read now (synchronises), or -- inside a *_dev entry point of a context with
// deferred checks -- copied into the ring for sda_ctx_synchronize to look at, reporting "not rejected" for now
int finish_flagged(sda_ctx *ctx, unsigned *rejected, bool may_defer = true) {
    if (!(ctx->deferred && ctx->dev_entry_depth > 0 && may_defer)) {
        OK(read_flags(ctx, rejected, nullptr));
        return SDA_OK;
    }
    if (ctx->seq_issued - ctx->seq_checked >= DEFER_RING) OK(resolve_deferred(ctx));
    CU(cudaMemcpyAsync(ctx->h_ring + ctx->seq_issued % DEFER_RING, ctx->d_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->seq_issued++;
    if (ctx->debug_force_reject) ctx->forced_pending = true;
    *rejected = 0;
    return SDA_OK;
}

// exact gen_range stream of `count` samples for one key into ctx->draws (at element offset off)
int draw_exact(sda_ctx *ctx, const DrawParams &dr, int rounds, const ChaChaKey &key, size_t count, uint64_t *d_out) {
    if (count == 0) return SDA_OK;
    const double accept = (double)dr.zone / 18446744073709551616.0;
    size_t window = (size_t)((double)count / accept * 1.02) + 4096;
    for (int attempt = 0; attempt < 8; attempt++) {
        const size_t se = draw_exact_scratch_elems(window);
        CU(ctx->scratch.reserve(se * sizeof(uint64_t)));
        OK(clear_flags(ctx));
        CU(launch_draw_exact(ctx->lc(), dr, rounds, key, count, window, d_out, (uint64_t *)ctx->scratch.p,
                             ctx->d_flag + 1));
        unsigned status = 0;
        OK(read_flags(ctx, nullptr, &status));
        if (!status) return SDA_OK;
        window *= 2;
    }
    return fail(ctx, SDA_ERR_CUDA, "exact gen_range stream did not converge");
}

int upload_keys(sda_ctx *ctx, const uint8_t *seeds, size_t P) {
    std::vector<ChaChaKey> k(P);
    for (size_t p = 0; p < P; p++) k[p] = key_from_seed_bytes(seeds + 32 * p);
    CU(ctx->keys.reserve(P * sizeof(ChaChaKey)));
    CU(cudaMemcpyAsync(ctx->keys.p, k.data(), P * sizeof(ChaChaKey), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // k goes out of scope
    volatile uint32_t *wipe = reinterpret_cast<volatile uint32_t *>(k.data());   // the host copy of the keys does not linger
    for (size_t i = 0; i < P * 8; i++) wipe[i] = 0;
    return SDA_OK;
}

// ---- host <-> device staging ----------------------------------------------------------------
bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

constexpr size_t STAGE_BYTES = 32u << 20;

// pageable -> device through two pinned staging buffers (overlaps the CPU copy with the DMA);
// pinned -> device directly.  Asynchronous on ctx->stream for the pinned case.
int h2d(sda_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return SDA_OK;
    if (is_pinned(src)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return SDA_OK;
    }
    if (bytes <= (1u << 20)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return SDA_OK;
    }
    for (int i = 0; i < 2; i++) CU(ctx->stage[i].reserve(STAGE_BYTES));
    size_t off = 0;
    int i = 0;
    while (off < bytes) {
        const size_t n = std::min(STAGE_BYTES, bytes - off);
        CU(cudaEventSynchronize(ctx->ev[i]));
        memcpy(ctx->stage[i].p, (const char *)src + off, n);
        CU(cudaMemcpyAsync((char *)dst + off, ctx->stage[i].p, n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaEventRecord(ctx->ev[i], ctx->stream));
        off += n;
        i ^= 1;
    }
    return SDA_OK;
}

// device -> host; returns after the data has landed
int d2h(sda_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return SDA_OK;
    if (is_pinned(dst) || bytes <= (1u << 20)) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return SDA_OK;
    }
    for (int i = 0; i < 2; i++) CU(ctx->stage[i].reserve(STAGE_BYTES));
    size_t off = 0, pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
    int i = 0;
    CU(cudaEventSynchronize(ctx->ev[0]));
    CU(cudaEventSynchronize(ctx->ev[1]));
    while (off < bytes || pend_n[0] || pend_n[1]) {
        if (pend_n[i]) {
            CU(cudaEventSynchronize(ctx->ev[i]));
            memcpy((char *)dst + pend_off[i], ctx->stage[i].p, pend_n[i]);
            pend_n[i] = 0;
        }
        if (off < bytes) {
            const size_t n = std::min(STAGE_BYTES, bytes - off);
            CU(cudaMemcpyAsync(ctx->stage[i].p, (const char *)src + off, n, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaEventRecord(ctx->ev[i], ctx->stream));
            pend_off[i] = off;
            pend_n[i] = n;
            off += n;
        }
        i ^= 1;
    }
    return SDA_OK;
}

// ---- sliced host entry points: H2D of slice i+1, the kernel of slice i and D2H of slice i-1 run concurrently
// on the two copy engines and the SMs (pinned host buffers only; PCIe is full duplex) ------------------------
constexpr size_t PIPE_MIN_BYTES = 4u << 20;    // below this a call is latency-bound and stays on one stream
constexpr size_t PIPE_MAX_PITCH = 1u << 30;    // rows of the 2-D copies stay far below cudaDeviceProp::memPitch
#ifndef SDA_PIPE_SLICES
#define SDA_PIPE_SLICES 8       // measured: 8 slices 1.655e9, 16 1.60e9, 32 1.45e9, 64 1.33e9 secrets/s end to end
#endif                          // (each slice costs ~35 us of submission; fewer slices leave more of the first copy in exposed)
constexpr size_t PIPE_SLICES = SDA_PIPE_SLICES;

int pipe_event(sda_ctx *ctx, size_t i, cudaEvent_t *out) {
    while (ctx->pipe_ev.size() <= i) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    *out = ctx->pipe_ev[i];
    return SDA_OK;
}

// the copy streams start after everything already queued on the compute stream (earlier calls reuse ctx->in / out)
int pipe_begin(sda_ctx *ctx) {
    cudaEvent_t e;
    OK(pipe_event(ctx, 0, &e));
    CU(cudaEventRecord(e, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->h2d_stream, e, 0));
    CU(cudaStreamWaitEvent(ctx->d2h_stream, e, 0));
    return SDA_OK;
}

// the constant GEMM operand of the tensor-core share-generation kernels, resident on the device per scheme
int ensure_tc_image(sda_ctx *ctx, const Packed &pk, const Matrix &M) {
    const size_t ib = packed_share_tc_image_bytes(pk.k, pk.t, pk.n);
    std::vector<uint64_t> key{(uint64_t)pk.k, (uint64_t)pk.t, (uint64_t)pk.n, pk.p};
    key.insert(key.end(), M.e, M.e + M.rows * M.cols);
    if (key == ctx->tc_image_key) return SDA_OK;
    std::vector<uint8_t> img(ib);
    packed_share_tc_build_image(pk.k, pk.t, pk.n, M, pk.p, img.data());
    CU(ctx->tc_image.reserve(ib));
    CU(cudaMemcpyAsync(ctx->tc_image.p, img.data(), ib, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // img goes out of scope
    ctx->tc_image_key = key;
    return SDA_OK;
}

// Tensor-core share generation, whichever kernel serves the scheme: the paired-tile kernel (packed_tc2.cu) for the
// instantiated shapes over 2^61 - 1 when the vector fits its 32-bit row offsets, packed_tc.cu for those shapes over
// other primes (or when the context asks for it with SDA_PACKED_PATH_TENSOR_CORES_V1), and the run-time-shaped kernel
// (packed_tcg.cu) for every other (k, t, n).  All generate batches first_batch .. first_batch + n_batches - 1 of every
// participant; first_batch is a multiple of share_tc_slice_batches().
enum { TC_NONE = 0, TC_V1 = 1, TC_PAIRED = 2, TC_RUNTIME = 3, TC_PAIRED_RTN = 4 };
int tc_kernel_for(const sda_ctx *ctx, const Packed &pk, size_t dim) {
    if (ctx->packed_path == SDA_PACKED_PATH_CUDA_CORES) return TC_NONE;
    if (ctx->packed_path != SDA_PACKED_PATH_TENSOR_CORES_ANY_SHAPE && packed_share_tc_image_bytes(pk.k, pk.t, pk.n) != 0) {
        if (pk.p == P61 && ctx->packed_path != SDA_PACKED_PATH_TENSOR_CORES_V1 && packed_share_tc2_supported(pk.k, pk.t, pk.n, dim))
            return TC_PAIRED;
        return TC_V1;
    }
    if (pk.p == P61 && ctx->packed_path != SDA_PACKED_PATH_TENSOR_CORES_V1 &&
        packed_share_tc2n_supported(pk.k, pk.t, pk.n, dim, ctx->rounds))
        return TC_PAIRED_RTN;
    return packed_share_tcg_supported(pk.k, pk.t, pk.n) ? TC_RUNTIME : TC_NONE;
}
size_t share_tc_slice_batches(const sda_ctx *ctx, const Packed &pk, size_t dim) {
    switch (tc_kernel_for(ctx, pk, dim)) {
    case TC_PAIRED: return packed_share_tc2_slice_batches(pk.k, pk.t, pk.n);
    case TC_V1: return packed_share_tc_slice_batches(pk.k, pk.t, pk.n);
    case TC_RUNTIME: return packed_share_tcg_slice_batches(pk.k, pk.t, pk.n);
    case TC_PAIRED_RTN: return packed_share_tc2n_slice_batches(pk.k, pk.t);
    }
    return 0;
}
// the constant GEMM operand of kernel `which`, resident on the device per (kernel, scheme)
int ensure_image(sda_ctx *ctx, int which, const Packed &pk, const Matrix &M, DevBuf *buf, std::vector<uint64_t> *cached) {
    std::vector<uint64_t> key{(uint64_t)which, (uint64_t)pk.k, (uint64_t)pk.t, (uint64_t)pk.n, pk.p};
    key.insert(key.end(), M.e, M.e + M.rows * M.cols);
    if (key == *cached) return SDA_OK;
    const size_t ib = which == TC_PAIRED ? packed_share_tc2_image_bytes(pk.k, pk.t, pk.n)
                    : which == TC_PAIRED_RTN ? packed_share_tc2n_image_bytes(pk.k, pk.t, pk.n)
                                             : packed_share_tcg_image_bytes(pk.k, pk.t, pk.n);
    std::vector<uint8_t> img(ib);
    if (which == TC_PAIRED) packed_share_tc2_build_image(pk.k, pk.t, pk.n, M, pk.p, img.data());
    else if (which == TC_PAIRED_RTN) packed_share_tc2n_build_image(pk.k, pk.t, pk.n, M, pk.p, img.data());
    else packed_share_tcg_build_image(pk.k, pk.t, pk.n, M, pk.p, img.data());
    CU(buf->reserve(ib));
    CU(cudaMemcpyAsync(buf->p, img.data(), ib, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // img goes out of scope
    *cached = key;
    return SDA_OK;
}
int launch_share_tc(sda_ctx *ctx, const Packed &pk, const Matrix &M, const FieldParams &f, const DrawParams &dr,
                    const int64_t *d_secrets, size_t ld, size_t P, size_t dim, size_t first_batch, size_t n_batches,
                    const ChaChaKey *d_keys, int64_t *d_out) {
    switch (tc_kernel_for(ctx, pk, dim)) {
    case TC_V1:
        OK(ensure_tc_image(ctx, pk, M));
        CU(launch_packed_share_tc(ctx->lc(), f, dr, ctx->rounds, pk.k, pk.t, pk.n, d_secrets, ld, P, dim, first_batch, n_batches,
                                  d_keys, (const uint8_t *)ctx->tc_image.p, d_out, ctx->d_flag));
        return SDA_OK;
    case TC_PAIRED:
        OK(ensure_image(ctx, TC_PAIRED, pk, M, &ctx->tc2_image, &ctx->tc2_image_key));
        CU(ctx->keys_pre.reserve(packed_share_tc2_key_scratch_bytes(P)));
        CU(launch_packed_share_tc2(ctx->lc(), ctx->rounds, pk.k, pk.t, pk.n, d_secrets, ld, P, dim, first_batch, n_batches, d_keys,
                                   (uint32_t *)ctx->keys_pre.p, (const uint8_t *)ctx->tc2_image.p, d_out, ctx->d_flag));
        return SDA_OK;
    case TC_PAIRED_RTN:
        OK(ensure_image(ctx, TC_PAIRED_RTN, pk, M, &ctx->tc2_image, &ctx->tc2_image_key));
        CU(ctx->keys_pre.reserve(packed_share_tc2_key_scratch_bytes(P)));
        CU(launch_packed_share_tc2n(ctx->lc(), pk.k, pk.t, pk.n, d_secrets, ld, P, dim, first_batch, n_batches, d_keys,
                                    (uint32_t *)ctx->keys_pre.p, (const uint8_t *)ctx->tc2_image.p, d_out, ctx->d_flag));
        return SDA_OK;
    case TC_RUNTIME:
        OK(ensure_image(ctx, TC_RUNTIME, pk, M, &ctx->tcg_image, &ctx->tcg_image_key));
        CU(launch_packed_share_tcg(ctx->lc(), f, dr, ctx->rounds, pk.k, pk.t, pk.n, d_secrets, ld, P, dim, first_batch, n_batches,
                                   d_keys, (const uint8_t *)ctx->tcg_image.p, d_out, ctx->d_flag));
        return SDA_OK;
    }
    return fail(ctx, SDA_ERR_UNSUPPORTED, "no tensor-core kernel for this scheme");
}

// which generation of the fused share-gen -> clerk-sum kernel a shape runs on: the paired-tile one is faster on every
// instantiated shape (profiles/r02_kernels.md: 6 / 15 / 17 % on configs #3 / #4 / #5)
bool fused_prefers_paired(const Packed &pk) {
    (void)pk;
    return true;
}

// ---- core device-side operations (shared by host and device entry points) --------------------

int share_generate_core(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_secrets, size_t ld, size_t P,
                        size_t dim, const uint8_t *seeds, int64_t *d_out) {
    Packed pk;
    OK(validate(ctx, s, &pk));
    if (P == 0 || dim == 0) return SDA_OK;
    if (!seeds) return fail(ctx, SDA_ERR_INVALID, "null rng_seed");
    const bool additive = s->kind == SDA_SHARING_ADDITIVE;
    const FieldParams f = make_field((uint64_t)s->modulus);
    const int n = (int)s->share_count;
    Matrix M;
    DrawParams dr;
    size_t B, dpe;   // outputs per clerk, draws per output column
    if (additive) {
        dr = make_draw((uint64_t)s->modulus);
        B = dim;
        dpe = (size_t)n - 1;
    } else {
        if (s->modulus < 3) return fail(ctx, SDA_ERR_INVALID, "prime_modulus too small");
        dr = make_draw((uint64_t)s->modulus - 1);   // tss share(): Range::new(0, prime - 1)
        OK(share_matrix_cached(ctx, pk, &M));
        B = (dim + pk.k - 1) / pk.k;
        dpe = (size_t)pk.t;
    }
    OK(upload_keys(ctx, seeds, P));
    const ChaChaKey *d_keys = (const ChaChaKey *)ctx->keys.p;
    const bool fast = additive ? (n == 1 || additive_split_has_fast_path(n)) : packed_share_has_fast_path(pk.k, pk.t, pk.n);
    bool exact = !fast;
    // any shape, any prime: a tensor-core kernel unless the context asks for the CUDA-core ones
    const bool use_tc = !additive && tc_kernel_for(ctx, pk, dim) != TC_NONE;
    if (use_tc) {
        OK(clear_flags(ctx));
        OK(launch_share_tc(ctx, pk, M, f, dr, d_secrets, ld, P, dim, 0, B, d_keys, d_out));
        unsigned rejected = 0;
        OK(finish_flagged(ctx, &rejected));
        exact = rejected != 0;
    } else if (fast) {
        OK(clear_flags(ctx));
        if (additive) CU(ctx->keys_pre.reserve(packed_share_tc2_key_scratch_bytes(std::min<size_t>(P, 65535))));
        for (size_t p0 = 0; p0 < P; p0 += 65535) {
            const size_t pc = std::min<size_t>(65535, P - p0);
            if (additive)
                CU(launch_additive_split(ctx->lc(), f, dr, ctx->rounds, n, d_secrets + p0 * ld, ld, pc, dim, d_keys + p0,
                                         nullptr, d_out + p0 * (size_t)n * B, ctx->d_flag, (uint32_t *)ctx->keys_pre.p));
            else
                CU(launch_packed_share(ctx->lc(), f, dr, ctx->rounds, pk.k, pk.t, pk.n, M, d_secrets + p0 * ld, ld, pc,
                                       dim, d_keys + p0, nullptr, nullptr, d_out + p0 * (size_t)n * B, ctx->d_flag));
        }
        unsigned rejected = 0;
        OK(finish_flagged(ctx, &rejected));
        exact = rejected != 0;
    }
    if (!exact) return SDA_OK;
    // exact path: per participant, materialise the gen_range stream, then share from memory
    const size_t count = B * dpe;
    CU(ctx->draws.reserve(std::max<size_t>(count, 1) * sizeof(uint64_t)));
    if (!additive) {
        CU(ctx->mat.reserve(sizeof(M.e)));
        CU(cudaMemcpyAsync(ctx->mat.p, M.e, sizeof(M.e), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    for (size_t p = 0; p < P; p++) {
        const ChaChaKey key = key_from_seed_bytes(seeds + 32 * p);
        OK(draw_exact(ctx, dr, ctx->rounds, key, count, (uint64_t *)ctx->draws.p));
        if (additive)
            CU(launch_additive_split(ctx->lc(), f, dr, ctx->rounds, n, d_secrets + p * ld, ld, 1, dim, d_keys + p,
                                     (const uint64_t *)ctx->draws.p, d_out + p * (size_t)n * B, ctx->d_flag));
        else
            CU(launch_packed_share(ctx->lc(), f, dr, ctx->rounds, pk.k, pk.t, pk.n, M, d_secrets + p * ld, ld, 1, dim,
                                   d_keys + p, (const uint64_t *)ctx->draws.p, (const uint64_t *)ctx->mat.p,
                                   d_out + p * (size_t)n * B, ctx->d_flag));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SDA_OK;
}

int combine_core(sda_ctx *ctx, int64_t modulus, const int64_t *d_rows, size_t ld, size_t P, size_t L,
                 const int64_t *d_acc, int64_t *d_out) {
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "attempt to calculate the remainder with a divisor of zero or negative modulus");
    if (L == 0) return SDA_OK;
    const FieldParams f = make_field((uint64_t)modulus);
    const size_t se = combine_scratch_elems(ctx->sm_count, P, L);
    if (se) CU(ctx->scratch.reserve(se * sizeof(int64_t)));
    CU(launch_combine(ctx->lc(), f, d_rows, ld, P, L, d_acc, d_out, (int64_t *)ctx->scratch.p, se));
    return SDA_OK;
}

int reconstruct_core(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                     const int64_t *d_shares, size_t ld, size_t m, size_t B, int64_t *d_out, size_t *out_len) {
    Packed pk;
    OK(validate(ctx, s, &pk));
    if (s->kind == SDA_SHARING_ADDITIVE) {
        // additive.rs:55-73: dimension = length of the first share vector; indices ignored
        const size_t len = m ? B : 0;
        if (out_len) *out_len = len;
        return combine_core(ctx, s->modulus, d_shares, ld, m, len, nullptr, d_out);
    }
    if (out_len) *out_len = dimension;
    const size_t nb = (dimension + pk.k - 1) / pk.k;
    if (nb == 0) return SDA_OK;
    if (m < (size_t)(pk.t + pk.k)) return fail(ctx, SDA_ERR_INVALID, "Not enough shares to reconstruct");   // packed_shamir.rs:75
    if (B < nb)   // batched.rs:84 indexes out of bounds -> panic
        return fail(ctx, SDA_ERR_INVALID, "index out of bounds: the len is %zu but the index is %zu", B, B);
    if (!indices) return fail(ctx, SDA_ERR_INVALID, "null indices");
    // R depends only on the scheme and the clerk subset: a recipient revealing with the same committee again
    // (or a benchmark loop) skips the host-side Lagrange inversions
    std::vector<uint64_t> rkey{(uint64_t)pk.k, (uint64_t)pk.t, (uint64_t)pk.n, pk.p, (uint64_t)s->omega_secrets,
                               (uint64_t)s->omega_shares};
    rkey.insert(rkey.end(), indices, indices + m);
    if (rkey != ctx->r_key) {
        Matrix fresh;
        OK(reconstruct_matrix(ctx, pk, indices, m, &fresh));
        ctx->r_cached = fresh;
        ctx->r_key = rkey;
    }
    const Matrix &R = ctx->r_cached;
    const FieldParams f = make_field(pk.p);
    if (f.kind == FIELD_MERSENNE61 && ctx->packed_path != SDA_PACKED_PATH_CUDA_CORES && reveal_tc_supported(pk.k, (int)m)) {
        const size_t ib = reveal_tc_image_bytes(pk.k, (int)m);
        std::vector<uint64_t> key{(uint64_t)pk.k, (uint64_t)m};
        key.insert(key.end(), R.e, R.e + R.rows * R.cols);
        if (key != ctx->tc_image_r_key) {           // same clerk subset as the previous call: operand still resident
            std::vector<uint8_t> img(ib);
            reveal_tc_build_image(pk.k, (int)m, R, img.data());
            CU(ctx->tc_image_r.reserve(ib));
            CU(cudaMemcpyAsync(ctx->tc_image_r.p, img.data(), ib, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));   // img goes out of scope
            ctx->tc_image_r_key = key;
        }
        CU(launch_reveal_tc(ctx->lc(), pk.k, (int)m, d_shares, ld, dimension, (const uint8_t *)ctx->tc_image_r.p, d_out));
        return SDA_OK;
    }
    CU(launch_packed_reconstruct(ctx->lc(), f, pk.k, (int)m, R, d_shares, ld, dimension, d_out));
    return SDA_OK;
}

int mask_validate(sda_ctx *ctx, const sda_masking_scheme *s) {
    if (!s) return fail(ctx, SDA_ERR_INVALID, "null scheme");
    if (s->kind == SDA_MASK_NONE) return SDA_OK;
    if (s->kind != SDA_MASK_FULL && s->kind != SDA_MASK_CHACHA)
        return fail(ctx, SDA_ERR_INVALID, "unknown masking scheme kind %d", s->kind);
    if (s->modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "Rng.gen_range called with low >= high");
    if (s->kind == SDA_MASK_CHACHA && (s->seed_bitsize + 31) / 32 > 8)
        return fail(ctx, SDA_ERR_UNSUPPORTED, "ChaCha seed_bitsize %llu > 256", (unsigned long long)s->seed_bitsize);
    return SDA_OK;
}

// Full: mask = dim draws of rng(seed); ChaCha: seed words = first words of rng(seed), mask stream
// = ChaCha20(seed words).  d_mask_out: Full -> [dim]; ChaCha -> unused (seed words returned in
// seed_words_out by the caller).
// d_x != nullptr: the secrets are the fixed-point encodings (frac_bits) of the real vector d_x, produced in the same pass
int mask_core(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_secrets, size_t dim,
              const uint8_t rng_seed[32], int64_t *d_mask_out, int64_t *seed_words_out, int64_t *d_masked_out,
              const float *d_x = nullptr, int frac_bits = 0, int64_t encode_modulus = 0) {
    OK(mask_validate(ctx, s));
    if (s->kind == SDA_MASK_NONE) {   // none.rs:14-18
        if (d_x) {
            if (dim) CU(launch_fixed_encode(ctx->lc(), make_field((uint64_t)encode_modulus), frac_bits, d_x, dim, d_masked_out));
            return SDA_OK;
        }
        if (dim && d_masked_out != d_secrets)
            CU(cudaMemcpyAsync(d_masked_out, d_secrets, dim * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        return SDA_OK;
    }
    if (!rng_seed) return fail(ctx, SDA_ERR_INVALID, "null rng_seed");
    const FieldParams f = make_field((uint64_t)s->modulus);
    const DrawParams dr = make_draw((uint64_t)s->modulus);
    ChaChaKey key = key_from_seed_bytes(rng_seed);
    int rounds = ctx->rounds;
    int64_t *mask_dst = d_mask_out;
    if (s->kind == SDA_MASK_CHACHA) {
        if (s->dimension != dim)   // chacha.rs:26
            return fail(ctx, SDA_ERR_INVALID, "assertion failed: `(left == right)` (chacha.rs:26: scheme dimension %llu, %zu secrets)",
                        (unsigned long long)s->dimension, dim);
        const size_t words = (size_t)((s->seed_bitsize + 31) / 32);   // chacha.rs:30
        uint32_t blk[16];
        host_chacha_block(key, 0, ctx->rounds, blk);                  // chacha.rs:31-33, OsRng -> injected stream
        uint32_t seed[8] = {0};
        for (size_t i = 0; i < words; i++) {
            seed[i] = blk[i];
            if (seed_words_out) seed_words_out[i] = (int64_t)seed[i]; // chacha.rs:48-50
        }
        key = key_from_words(seed, words);                            // chacha.rs:36
        rounds = 20;
        mask_dst = nullptr;
    }
    if (dim == 0) return SDA_OK;
    OK(clear_flags(ctx));
    CU(launch_mask(ctx->lc(), f, dr, rounds, d_secrets, dim, key, nullptr, mask_dst, d_masked_out, ctx->d_flag, d_x, frac_bits));
    unsigned rejected = 0;
    OK(finish_flagged(ctx, &rejected));
    if (!rejected) return SDA_OK;
    if (d_x) {   // the exact path reads i64 secrets: encode first, then mask in place
        CU(launch_fixed_encode(ctx->lc(), f, frac_bits, d_x, dim, d_masked_out));
        d_secrets = d_masked_out;
    }
    CU(ctx->draws.reserve(dim * sizeof(uint64_t)));
    OK(draw_exact(ctx, dr, rounds, key, dim, (uint64_t *)ctx->draws.p));
    CU(launch_mask(ctx->lc(), f, dr, rounds, d_secrets, dim, key, (const uint64_t *)ctx->draws.p, mask_dst,
                   d_masked_out, ctx->d_flag));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDA_OK;
}

// ChaCha mask combine from host seed rows (P x mask_len i64)
int chacha_mask_combine_core(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *h_masks, size_t P,
                             size_t mask_len, int64_t *d_out) {
    const size_t dim = (size_t)s->dimension;
    if (dim == 0) return SDA_OK;
    const FieldParams f = make_field((uint64_t)s->modulus);
    const DrawParams dr = make_draw((uint64_t)s->modulus);
    std::vector<ChaChaKey> keys(P);
    for (size_t p = 0; p < P; p++) {
        uint32_t w[8] = {0};
        for (size_t i = 0; i < mask_len && i < 8; i++) w[i] = (uint32_t)h_masks[p * mask_len + i];   // chacha.rs:62-64
        keys[p] = key_from_words(w, 8);
    }
    CU(ctx->keys.reserve(std::max<size_t>(P, 1) * sizeof(ChaChaKey)));
    if (P) CU(cudaMemcpyAsync(ctx->keys.p, keys.data(), P * sizeof(ChaChaKey), cudaMemcpyHostToDevice, ctx->stream));
    const size_t se = chacha_mask_combine_scratch_elems(ctx->sm_count, P, dim);
    if (se) CU(ctx->scratch.reserve(se * sizeof(int64_t)));
    OK(clear_flags(ctx));
    CU(launch_chacha_mask_combine(ctx->lc(), f, dr, (const ChaChaKey *)ctx->keys.p, P, dim, d_out,
                                  (int64_t *)ctx->scratch.p, se, ctx->d_flag));
    unsigned rejected = 0;
    OK(finish_flagged(ctx, &rejected));
    if (!rejected) return SDA_OK;
    // exact: out = sum over seeds of the exact stream
    CU(ctx->draws.reserve(dim * sizeof(uint64_t)));
    CU(cudaMemsetAsync(d_out, 0, dim * sizeof(int64_t), ctx->stream));
    for (size_t p = 0; p < P; p++) {
        OK(draw_exact(ctx, dr, 20, keys[p], dim, (uint64_t *)ctx->draws.p));
        CU(launch_combine(ctx->lc(), f, (const int64_t *)ctx->draws.p, dim, 1, dim, d_out, d_out, nullptr, 0));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SDA_OK;
}

}  // namespace

// =================================================================================================
#pragma GCC visibility push(default)

namespace {

// ---- NCCL, bound at run time -----------------------------------------------------------------------------
// The library takes libnccl.so.2 from the process (dlopen by soname: a host application that already has NCCL
// loaded -- torch does -- shares its copy) instead of linking it, so hosts that never call the multi-GPU entry
// points do not need NCCL installed.  SDA_B200_NCCL_LIB overrides the name.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

const NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("SDA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
            api.error = dlerror();
        }
        if (!api.handle) return;
        bool ok = true;
        auto bind = [&](auto &fn, const char *sym) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, sym));
            if (!fn) {
                ok = false;
                api.error = std::string("missing symbol ") + sym;
            }
        };
        bind(api.GetVersion, "ncclGetVersion");
        bind(api.GetUniqueId, "ncclGetUniqueId");
        bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommInitAll, "ncclCommInitAll");
        bind(api.CommDestroy, "ncclCommDestroy");
        bind(api.Reduce, "ncclReduce");
        bind(api.AllGather, "ncclAllGather");
        bind(api.GroupStart, "ncclGroupStart");
        bind(api.GroupEnd, "ncclGroupEnd");
        bind(api.GetErrorString, "ncclGetErrorString");
        if (!ok) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return &api;
}

#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess)                                                                                \
            return fail(ctx, SDA_ERR_NCCL, "NCCL error %s at %s:%d", nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

int need_nccl(sda_ctx *ctx, const NcclApi **out) {
    const NcclApi *api = nccl_api();
    if (!api->handle) return fail(ctx, SDA_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", api->error.c_str());
    *out = api;
    return SDA_OK;
}

void comm_release(sda_ctx *ctx) {
    if (ctx->comm && nccl_api()->handle) nccl_api()->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_size = 1;
}

// One rank's part of the clerk-sum exchange: d_partials holds `count` canonical residues (this rank's
// column sums); on return the root's buffer holds the sums over all ranks mod m, canonical.  While comm_size * m
// fits 64 bits the exchange is one ncclReduce(sum) of u64 followed by one mod pass on the root; otherwise the
// partials are gathered and folded by the combine kernel, which adds mod m.  Call inside ncclGroupStart/End when
// several ranks live in this process (group = true defers nothing here: NCCL queues the collective).
int reduce_partials_enqueue(sda_ctx *ctx, const NcclApi *api, const FieldParams &f, int64_t *d_partials, size_t count, int root,
                            bool *needs_fold) {
    const bool sum_fits = (u128)f.m * (u128)ctx->comm_size <= ((u128)1 << 64);
    *needs_fold = !sum_fits;
    if (sum_fits) {
        NC(api->Reduce(d_partials, d_partials, count, ncclUint64, ncclSum, root, ctx->comm, ctx->stream));
        return SDA_OK;
    }
    CU(ctx->aux.reserve((size_t)ctx->comm_size * count * sizeof(int64_t)));
    NC(api->AllGather(d_partials, ctx->aux.p, count, ncclUint64, ctx->comm, ctx->stream));
    return SDA_OK;
}
int reduce_partials_finish(sda_ctx *ctx, const FieldParams &f, int64_t *d_partials, size_t count, int root, bool needs_fold) {
    if (ctx->comm_rank != root) return SDA_OK;
    if (!needs_fold) {
        CU(launch_mod_reduce(ctx->lc(), f, d_partials, count, d_partials, true));
        return SDA_OK;
    }
    return combine_core(ctx, (int64_t)f.m, (const int64_t *)ctx->aux.p, count, (size_t)ctx->comm_size, count, nullptr, d_partials);
}

}  // namespace

extern "C" {

int sda_abi_version(void) { return SDA_B200_ABI_VERSION; }

int sda_ctx_create(int device, sda_ctx **out) {
    sda_ctx *ctx = nullptr;
    if (!out) return fail(ctx, SDA_ERR_INVALID, "null out pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ctx, SDA_ERR_CUDA, "no CUDA device available (%s): libsda_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(ctx, SDA_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(ctx, SDA_ERR_CUDA, "device %d is sm_%d%d; libsda_b200 is built for sm_100a only", device, prop.major,
                    prop.minor);
    sda_ctx *c = new sda_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    ctx = c;
    cudaError_t err = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc((void **)&c->d_flag, 2 * sizeof(unsigned));
    if (err == cudaSuccess) err = cudaMallocHost((void **)&c->h_flag, 2 * sizeof(unsigned));
    for (int i = 0; i < 4 && err == cudaSuccess; i++) err = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
    if (err != cudaSuccess) {
        g_create_error = std::string("CUDA error ") + cudaGetErrorString(err) + " while creating the context";
        sda_ctx_destroy(c);
        return SDA_ERR_CUDA;
    }
    c->stream = c->own_stream;
    const char *force = getenv("SDA_B200_DEBUG_FORCE_REJECT");
    c->debug_force_reject = force && force[0] == '1';
    *out = c;
    return SDA_OK;
}

void sda_ctx_destroy(sda_ctx *ctx) {
    if (!ctx) return;
    for (size_t i = 1; i < ctx->members.size(); i++) sda_ctx_destroy(ctx->members[i]);
    ctx->members.clear();
    DeviceGuard g(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    // keys, their derived constants and materialised draw streams are secret material: zero them before the memory
    // goes back to the allocator
    for (DevBuf *b : {&ctx->keys, &ctx->keys_pre, &ctx->draws})
        if (b->p) cudaMemset(b->p, 0, b->cap);
    for (DevBuf *b : {&ctx->in, &ctx->out, &ctx->aux, &ctx->scratch, &ctx->draws, &ctx->keys, &ctx->keys_pre, &ctx->mat, &ctx->tc_image, &ctx->tc_image_r, &ctx->tc2_image, &ctx->tcg_image, &ctx->masked_tmp}) b->release();
    ctx->stage[0].release();
    ctx->stage[1].release();
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    for (int i = 0; i < 4; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (cudaEvent_t e : ctx->pipe_ev) cudaEventDestroy(e);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *sda_last_error(const sda_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sda_ctx_set_rng_rounds(sda_ctx *ctx, int rounds) {
    if (!ctx) return SDA_ERR_INVALID;
    if (rounds != 8 && rounds != 12 && rounds != 20) return fail(ctx, SDA_ERR_INVALID, "rng rounds must be 8, 12 or 20");
    ctx->rounds = rounds;
    return SDA_OK;
}
int sda_ctx_get_rng_rounds(const sda_ctx *ctx) { return ctx ? ctx->rounds : 0; }
int sda_ctx_set_packed_path(sda_ctx *ctx, int path) {
    if (!ctx) return SDA_ERR_INVALID;
    if (path != SDA_PACKED_PATH_AUTO && path != SDA_PACKED_PATH_CUDA_CORES && path != SDA_PACKED_PATH_TENSOR_CORES &&
        path != SDA_PACKED_PATH_TENSOR_CORES_V1 && path != SDA_PACKED_PATH_TENSOR_CORES_ANY_SHAPE)
        return fail(ctx, SDA_ERR_INVALID, "unknown packed-share path %d", path);
    ctx->packed_path = path;
    return SDA_OK;
}
int sda_ctx_set_stream(sda_ctx *ctx, void *cuda_stream) {
    if (!ctx) return SDA_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return SDA_OK;
}
void *sda_ctx_get_stream(const sda_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int sda_ctx_synchronize(sda_ctx *ctx) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return resolve_deferred(ctx);
}
int sda_ctx_set_deferred_checks(sda_ctx *ctx, int on) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (on && !ctx->h_ring) CU(cudaMallocHost((void **)&ctx->h_ring, DEFER_RING * sizeof(unsigned)));
    const int rc = resolve_deferred(ctx);      // nothing stays pending across a switch
    ctx->deferred = on != 0;
    return rc;
}
uint64_t sda_ctx_launch_count(const sda_ctx *ctx) { return ctx ? ctx->nlaunch : 0; }
const char *sda_ctx_last_kernel(const sda_ctx *ctx) { return ctx ? ctx->kernel_name : ""; }

int sda_host_alloc(sda_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return SDA_OK;
}
int sda_host_free(sda_ctx *ctx, void *ptr) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (ptr) CU(cudaFreeHost(ptr));
    return SDA_OK;
}

// ---- derived scheme properties (protocol/src/crypto.rs:117-155) --------------------------------
size_t sda_input_size(const sda_sharing_scheme *s) { return s->kind == SDA_SHARING_ADDITIVE ? 1 : (size_t)s->secret_count; }
size_t sda_output_size(const sda_sharing_scheme *s) { return (size_t)s->share_count; }
size_t sda_privacy_threshold(const sda_sharing_scheme *s) {
    return s->kind == SDA_SHARING_ADDITIVE ? (size_t)s->share_count - 1 : (size_t)s->privacy_threshold;
}
size_t sda_reconstruction_threshold(const sda_sharing_scheme *s) {
    return s->kind == SDA_SHARING_ADDITIVE ? (size_t)s->share_count : (size_t)(s->privacy_threshold + s->secret_count);
}
size_t sda_share_batches(const sda_sharing_scheme *s, size_t dim) {
    const size_t k = sda_input_size(s);
    return k ? (dim + k - 1) / k : 0;
}
size_t sda_mask_len(const sda_masking_scheme *s, size_t dim) {
    switch (s->kind) {
    case SDA_MASK_FULL: return dim;
    case SDA_MASK_CHACHA: return (size_t)((s->seed_bitsize + 31) / 32);
    default: return 0;
    }
}

int sda_sharing_scheme_validate(sda_ctx *ctx, const sda_sharing_scheme *s) {
    Packed pk;
    OK(validate(ctx, s, &pk));
    if (s->kind == SDA_SHARING_PACKED_SHAMIR) {
        Matrix M;
        OK(share_matrix(ctx, pk, &M));
    }
    return SDA_OK;
}

int sda_packed_share_matrix(sda_ctx *ctx, const sda_sharing_scheme *s, int64_t *out) {
    Packed pk;
    OK(validate(ctx, s, &pk));
    if (s->kind != SDA_SHARING_PACKED_SHAMIR) return fail(ctx, SDA_ERR_INVALID, "not a PackedShamir scheme");
    Matrix M;
    OK(share_matrix(ctx, pk, &M));
    for (int i = 0; i < M.rows * M.cols; i++) out[i] = (int64_t)M.e[i];
    return SDA_OK;
}

int sda_packed_reconstruct_matrix(sda_ctx *ctx, const sda_sharing_scheme *s, const uint64_t *indices, size_t m,
                                  int64_t *out) {
    Packed pk;
    OK(validate(ctx, s, &pk));
    if (s->kind != SDA_SHARING_PACKED_SHAMIR) return fail(ctx, SDA_ERR_INVALID, "not a PackedShamir scheme");
    Matrix R;
    OK(reconstruct_matrix(ctx, pk, indices, m, &R));
    for (int i = 0; i < R.rows * R.cols; i++) out[i] = (int64_t)R.e[i];
    return SDA_OK;
}

// ---- device-pointer entry points -------------------------------------------------------------------

int sda_share_generate_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_secrets, size_t secrets_ld,
                           size_t P, size_t dim, const uint8_t *seeds, int64_t *d_shares_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DevEntry dev_entry(ctx);
    if (secrets_ld < dim) return fail(ctx, SDA_ERR_INVALID, "secrets_ld < dim");
    return share_generate_core(ctx, s, d_secrets, secrets_ld, P, dim, seeds, d_shares_out);
}

// participate.rs:53-54 then :75-76 for P participants: masks drawn and added while the secrets are staged as operand
// rows (packed_tc2m.cu) where both schemes live over 2^61 - 1 and the sharing scheme has an instantiated shape; any other
// pair of schemes, and a call in which gen_range rejected a word, runs mask and share generation one after the other per
// participant with the masked secrets in a scratch vector.  Results are identical either way.
int sda_mask_share_generate_dev(sda_ctx *ctx, const sda_masking_scheme *ms, const sda_sharing_scheme *ss,
                                const int64_t *d_secrets, size_t secrets_ld, size_t P, size_t dim,
                                const uint8_t *mask_rng_seeds, const uint8_t *share_rng_seeds, int64_t *d_masks_out,
                                int64_t *d_shares_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DevEntry dev_entry(ctx);
    if (secrets_ld < dim) return fail(ctx, SDA_ERR_INVALID, "secrets_ld < dim");
    OK(mask_validate(ctx, ms));
    Packed pk;
    OK(validate(ctx, ss, &pk));
    if (P == 0) return SDA_OK;
    if (ms->kind == SDA_MASK_NONE) return share_generate_core(ctx, ss, d_secrets, secrets_ld, P, dim, share_rng_seeds, d_shares_out);
    if (!mask_rng_seeds || !share_rng_seeds) return fail(ctx, SDA_ERR_INVALID, "null rng_seed");
    if (ms->kind == SDA_MASK_CHACHA && ms->dimension != dim)   // chacha.rs:26
        return fail(ctx, SDA_ERR_INVALID, "assertion failed: `(left == right)` (chacha.rs:26: scheme dimension %llu, %zu secrets)",
                    (unsigned long long)ms->dimension, dim);
    const size_t mask_len = sda_mask_len(ms, dim);
    if (mask_len && !d_masks_out) return fail(ctx, SDA_ERR_INVALID, "null mask output");
    const size_t out_per_p = ss->kind == SDA_SHARING_PACKED_SHAMIR ? (size_t)pk.n * ((dim + pk.k - 1) / pk.k) : (size_t)ss->share_count * dim;
    const bool fused = dim > 0 && ss->kind == SDA_SHARING_PACKED_SHAMIR && (uint64_t)ms->modulus == P61 && pk.p == P61 &&
                       tc_kernel_for(ctx, pk, dim) == TC_PAIRED && packed_share_tc2_masked_supported(pk.k, pk.t, pk.n, dim, ctx->rounds);
    if (fused) {
        // keys: [0, P) the sharing streams, [P, 2P) the mask streams (Full: the participant's rng itself, full.rs:24-27;
        // ChaCha: the seed its rng draws first, chacha.rs:30-36, which is also the mask that is sent)
        std::vector<ChaChaKey> k(2 * P);
        std::vector<int64_t> words(ms->kind == SDA_MASK_CHACHA ? P * mask_len : 0);
        for (size_t p = 0; p < P; p++) {
            k[p] = key_from_seed_bytes(share_rng_seeds + 32 * p);
            ChaChaKey mk = key_from_seed_bytes(mask_rng_seeds + 32 * p);
            if (ms->kind == SDA_MASK_CHACHA) {
                uint32_t blk[16], seed[8] = {0};
                host_chacha_block(mk, 0, ctx->rounds, blk);
                for (size_t i = 0; i < mask_len; i++) {
                    seed[i] = blk[i];
                    words[p * mask_len + i] = (int64_t)seed[i];
                }
                mk = key_from_words(seed, mask_len);
            }
            k[P + p] = mk;
        }
        Matrix M;
        OK(share_matrix_cached(ctx, pk, &M));
        CU(ctx->keys.reserve(2 * P * sizeof(ChaChaKey)));
        CU(cudaMemcpyAsync(ctx->keys.p, k.data(), 2 * P * sizeof(ChaChaKey), cudaMemcpyHostToDevice, ctx->stream));
        if (!words.empty())
            CU(cudaMemcpyAsync(d_masks_out, words.data(), words.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));   // the host copies go out of scope
        volatile uint32_t *wipe = reinterpret_cast<volatile uint32_t *>(k.data());
        for (size_t i = 0; i < 2 * P * 8; i++) wipe[i] = 0;
        OK(ensure_image(ctx, TC_PAIRED, pk, M, &ctx->tc2_image, &ctx->tc2_image_key));
        CU(ctx->keys_pre.reserve(2 * packed_share_tc2_key_scratch_bytes(P)));
        OK(clear_flags(ctx));
        const ChaChaKey *d_keys = (const ChaChaKey *)ctx->keys.p;
        CU(launch_packed_share_tc2_masked(ctx->lc(), pk.k, pk.t, pk.n, d_secrets, secrets_ld, P, dim, d_keys, d_keys + P,
                                          (uint32_t *)ctx->keys_pre.p, (const uint8_t *)ctx->tc2_image.p,
                                          ms->kind == SDA_MASK_FULL ? d_masks_out : nullptr, d_shares_out, ctx->d_flag));
        unsigned rejected = 0;
        OK(finish_flagged(ctx, &rejected));
        if (ctx->debug_force_reject && !(ctx->deferred && ctx->dev_entry_depth > 0)) rejected = 1;   // test hook: exercise the redo
        if (!rejected) return SDA_OK;
    }
    // one participant at a time through a scratch vector
    CU(ctx->masked_tmp.reserve(std::max<size_t>(dim, 1) * sizeof(int64_t)));
    int64_t *d_masked = (int64_t *)ctx->masked_tmp.p;
    for (size_t p = 0; p < P; p++) {
        int64_t words[8] = {0};
        OK(mask_core(ctx, ms, d_secrets + p * secrets_ld, dim, mask_rng_seeds + 32 * p,
                     ms->kind == SDA_MASK_FULL ? d_masks_out + p * mask_len : nullptr, words, d_masked));
        if (ms->kind == SDA_MASK_CHACHA && mask_len) {
            CU(cudaMemcpyAsync(d_masks_out + p * mask_len, words, mask_len * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        OK(share_generate_core(ctx, ss, d_masked, dim, 1, dim, share_rng_seeds + 32 * p, d_shares_out + p * out_per_p));
    }
    return SDA_OK;
}

int sda_share_combine_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_shares, size_t ld, size_t P,
                          size_t L, const int64_t *d_acc_in, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(validate(ctx, s, nullptr));
    if (ld < L) return fail(ctx, SDA_ERR_INVALID, "Wrong dimension");
    return combine_core(ctx, s->modulus, d_shares, ld, P, L, d_acc_in, d_out);
}

int sda_share_generate_combine_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_secrets,
                                   size_t secrets_ld, size_t P, size_t dim, const uint8_t *seeds,
                                   const int64_t *d_acc_in, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DevEntry dev_entry(ctx);
    OK(validate(ctx, s, nullptr));
    if (secrets_ld < dim) return fail(ctx, SDA_ERR_INVALID, "secrets_ld < dim");
    const size_t n = s->share_count, B = sda_share_batches(s, dim);
    if (n * B == 0) return SDA_OK;
    if (P == 0) {
        if (d_acc_in && d_acc_in != d_out)
            CU(cudaMemcpyAsync(d_out, d_acc_in, n * B * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        else if (!d_acc_in)
            CU(cudaMemsetAsync(d_out, 0, n * B * sizeof(int64_t), ctx->stream));
        return SDA_OK;
    }
    // Mersenne-61 fast shapes: one fused kernel, the participant sum accumulated in TMEM (packed_tc.cu)
    if (s->kind == SDA_SHARING_PACKED_SHAMIR && (uint64_t)s->modulus == P61 && ctx->packed_path != SDA_PACKED_PATH_CUDA_CORES) {
        Packed pk;
        OK(validate(ctx, s, &pk));
        if (packed_share_tc_image_bytes(pk.k, pk.t, pk.n) != 0) {
            if (!seeds) return fail(ctx, SDA_ERR_INVALID, "null rng_seed");
            Matrix M;
            OK(share_matrix_cached(ctx, pk, &M));
            OK(ensure_tc_image(ctx, pk, M));
            OK(upload_keys(ctx, seeds, P));
            OK(clear_flags(ctx));
            // In place (d_acc_in == d_out) the kernel must not overwrite the running sum before the rejection flag is
            // known: a redo would start from a sum that already holds this call's (wrong) contribution.  The sums then
            // go to scratch and are copied over the accumulator only after the flag read 0.
            int64_t *d_dst = d_out;
            if (d_acc_in == d_out) {
                CU(ctx->aux.reserve(n * B * sizeof(int64_t)));
                d_dst = (int64_t *)ctx->aux.p;
            }
            CU(ctx->keys_pre.reserve(packed_share_tc2_key_scratch_bytes(P)));
            // the paired-tile generation of the fused kernel (packed_tc2f.cu) where it is the faster one; TENSOR_CORES_V1
            // keeps the first (packed_tc.cu) for side-by-side runs
            if (ctx->packed_path != SDA_PACKED_PATH_TENSOR_CORES_V1 && fused_prefers_paired(pk) &&
                packed_share_combine_tc2_supported(pk.k, pk.t, pk.n, dim)) {
                OK(ensure_image(ctx, TC_PAIRED, pk, M, &ctx->tc2_image, &ctx->tc2_image_key));
                CU(launch_packed_share_combine_tc2(ctx->lc(), ctx->rounds, pk.k, pk.t, pk.n, d_secrets, secrets_ld, P, dim,
                                                   (const ChaChaKey *)ctx->keys.p, (uint32_t *)ctx->keys_pre.p,
                                                   (const uint8_t *)ctx->tc2_image.p, d_acc_in, d_dst, ctx->d_flag));
            } else {
                CU(launch_packed_share_combine_tc(ctx->lc(), ctx->rounds, pk.k, pk.t, pk.n, d_secrets, secrets_ld, P, dim,
                                                  (const ChaChaKey *)ctx->keys.p, (const uint8_t *)ctx->tc_image.p, d_acc_in,
                                                  d_dst, ctx->d_flag, (uint32_t *)ctx->keys_pre.p));
            }
            unsigned rejected = 0;
            // (in place the copy below waits for the flag: such a call is never deferred)
            OK(finish_flagged(ctx, &rejected, d_dst == d_out));
            if (ctx->debug_force_reject && !(ctx->deferred && ctx->dev_entry_depth > 0 && d_dst == d_out)) rejected = 1;   // test hook
            if (!rejected) {
                if (d_dst != d_out)
                    CU(cudaMemcpyAsync(d_out, d_dst, n * B * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
                return SDA_OK;
            }
            // a rejected gen_range word somewhere (p ~ 2^-57 per draw): redo on the materialising path below
        }
    }
    // tiles of participants: generate [tile][n][B] into ctx->aux, fold each clerk's rows into out
    const size_t per = n * B * sizeof(int64_t);
    size_t tile = std::max<size_t>(1, std::min<size_t>(P, (size_t(2) << 30) / per));
    CU(ctx->aux.reserve(tile * per));
    int64_t *d_tile = (int64_t *)ctx->aux.p;
    const int64_t *acc = d_acc_in;
    for (size_t p0 = 0; p0 < P; p0 += tile) {
        const size_t pc = std::min(tile, P - p0);
        OK(share_generate_core(ctx, s, d_secrets + p0 * secrets_ld, secrets_ld, pc, dim, seeds + 32 * p0, d_tile));
        for (size_t r = 0; r < n; r++)
            OK(combine_core(ctx, s->modulus, d_tile + r * B, n * B, pc, B, acc ? acc + r * B : nullptr, d_out + r * B));
        acc = d_out;
    }
    return SDA_OK;
}

int sda_mod_reduce_dev(sda_ctx *ctx, int64_t modulus, const int64_t *d_in, size_t n, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    CU(launch_mod_reduce(ctx->lc(), make_field((uint64_t)modulus), d_in, n, d_out));
    return SDA_OK;
}

int sda_mod_reduce_u64_dev(sda_ctx *ctx, int64_t modulus, const uint64_t *d_in, size_t n, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    CU(launch_mod_reduce(ctx->lc(), make_field((uint64_t)modulus), (const int64_t *)d_in, n, d_out, true));
    return SDA_OK;
}

int sda_secret_reconstruct_dev(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                               const int64_t *d_shares, size_t ld, size_t m, size_t B, int64_t *d_secrets_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (ld < B) return fail(ctx, SDA_ERR_INVALID, "ld < B");
    return reconstruct_core(ctx, s, dimension, indices, d_shares, ld, m, B, d_secrets_out, nullptr);
}

int sda_mask_dev(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_secrets, size_t dim,
                 const uint8_t rng_seed[32], int64_t *d_mask_out, int64_t *d_masked_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DevEntry dev_entry(ctx);
    if (s && s->kind == SDA_MASK_CHACHA) {
        int64_t words[8] = {0};
        OK(mask_core(ctx, s, d_secrets, dim, rng_seed, nullptr, words, d_masked_out));
        const size_t nw = sda_mask_len(s, dim);
        if (nw && d_mask_out) {
            CU(cudaMemcpyAsync(d_mask_out, words, nw * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        return SDA_OK;
    }
    return mask_core(ctx, s, d_secrets, dim, rng_seed, d_mask_out, nullptr, d_masked_out);
}

int sda_fixed_encode_mask_dev(sda_ctx *ctx, const sda_masking_scheme *s, int64_t modulus, int frac_bits, const float *d_x,
                              size_t x_ld, size_t P, size_t dim, const uint8_t *seeds, int64_t *d_mask_out,
                              int64_t *d_masked_out, size_t masked_ld) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DevEntry dev_entry(ctx);
    OK(mask_validate(ctx, s));
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    if (frac_bits < 0 || frac_bits > 52) return fail(ctx, SDA_ERR_INVALID, "frac_bits must be in [0, 52]");
    if (s->kind != SDA_MASK_NONE && s->modulus != modulus)
        return fail(ctx, SDA_ERR_INVALID, "the masking scheme's modulus differs from the encoding modulus");
    if (x_ld < dim || masked_ld < dim) return fail(ctx, SDA_ERR_INVALID, "row stride shorter than the vector");
    if (P == 0 || dim == 0) return SDA_OK;
    const size_t nw = sda_mask_len(s, dim);
    if (s->kind == SDA_MASK_NONE) {
        for (size_t p = 0; p < P; p++)
            CU(launch_fixed_encode(ctx->lc(), make_field((uint64_t)modulus), frac_bits, d_x + p * x_ld, dim, d_masked_out + p * masked_ld));
        return SDA_OK;
    }
    if (!seeds) return fail(ctx, SDA_ERR_INVALID, "null rng_seed");
    if (s->kind == SDA_MASK_CHACHA && s->dimension != dim)   // chacha.rs:26
        return fail(ctx, SDA_ERR_INVALID, "assertion failed: `(left == right)` (chacha.rs:26: scheme dimension %llu, %zu secrets)",
                    (unsigned long long)s->dimension, dim);
    // every participant's kernel is queued before anything is read back: one flag read (one stream synchronisation) per
    // call instead of one per participant
    const FieldParams f = make_field((uint64_t)s->modulus);
    const DrawParams dr = make_draw((uint64_t)s->modulus);
    std::vector<int64_t> words(s->kind == SDA_MASK_CHACHA ? P * nw : 0);
    OK(clear_flags(ctx));
    for (size_t p = 0; p < P; p++) {
        ChaChaKey key = key_from_seed_bytes(seeds + 32 * p);
        int rounds = ctx->rounds;
        int64_t *mask_dst = d_mask_out ? d_mask_out + p * nw : nullptr;
        if (s->kind == SDA_MASK_CHACHA) {
            uint32_t blk[16], seed[8] = {0};
            host_chacha_block(key, 0, ctx->rounds, blk);                  // chacha.rs:31-33, OsRng -> injected stream
            for (size_t i = 0; i < nw; i++) {
                seed[i] = blk[i];
                words[p * nw + i] = (int64_t)seed[i];                     // chacha.rs:48-50
            }
            key = key_from_words(seed, nw);                               // chacha.rs:36
            rounds = 20;
            mask_dst = nullptr;
        }
        CU(launch_mask(ctx->lc(), f, dr, rounds, nullptr, dim, key, nullptr, mask_dst, d_masked_out + p * masked_ld, ctx->d_flag,
                       d_x + p * x_ld, frac_bits));
    }
    if (!words.empty() && d_mask_out)
        CU(cudaMemcpyAsync(d_mask_out, words.data(), words.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    unsigned rejected = 0;
    if (!words.empty()) CU(cudaStreamSynchronize(ctx->stream));     // `words` has been copied
    OK(finish_flagged(ctx, &rejected));
    if (!rejected) return SDA_OK;
    // a rejected gen_range word somewhere: redo participant by participant on the exact path
    for (size_t p = 0; p < P; p++) {
        int64_t w8[8] = {0};
        OK(mask_core(ctx, s, nullptr, dim, seeds + 32 * p, s->kind == SDA_MASK_CHACHA ? nullptr : (d_mask_out ? d_mask_out + p * nw : nullptr),
                     s->kind == SDA_MASK_CHACHA ? w8 : nullptr, d_masked_out + p * masked_ld, d_x + p * x_ld, frac_bits, modulus));
    }
    return SDA_OK;
}

int sda_mask_combine_dev(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_masks, size_t P, size_t mask_len,
                         int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    DevEntry dev_entry(ctx);
    OK(mask_validate(ctx, s));
    switch (s->kind) {
    case SDA_MASK_NONE:
        if (mask_len != 0) return fail(ctx, SDA_ERR_INVALID, "assertion failed: masks.iter().all(|mask| mask.len() == 0)");
        return SDA_OK;
    case SDA_MASK_FULL:
        return combine_core(ctx, s->modulus, d_masks, mask_len, P, P ? mask_len : 0, nullptr, d_out);
    default: {
        std::vector<int64_t> h(P * mask_len);
        if (!h.empty()) {
            CU(cudaMemcpyAsync(h.data(), d_masks, h.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        return chacha_mask_combine_core(ctx, s, h.data(), P, mask_len, d_out);
    }
    }
}

int sda_unmask_dev(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *d_mask, const int64_t *d_masked,
                   size_t dim, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(mask_validate(ctx, s));
    if (s->kind == SDA_MASK_NONE) {
        if (dim && d_out != d_masked)
            CU(cudaMemcpyAsync(d_out, d_masked, dim * sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        return SDA_OK;
    }
    CU(launch_submod(ctx->lc(), make_field((uint64_t)s->modulus), d_masked, d_mask, dim, d_out));   // full.rs:60-63
    return SDA_OK;
}

// ---- share wire codec ----------------------------------------------------------------------------
size_t sda_varint_max_bytes(size_t n) { return 10 * n; }

int sda_varint_encode_dev(sda_ctx *ctx, const int64_t *d_shares, size_t n, uint8_t *d_out, size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    const size_t se = varint_encode_scratch_elems(n);
    CU(ctx->scratch.reserve(se * sizeof(uint64_t)));
    CU(launch_varint_encode(ctx->lc(), d_shares, n, d_out, (uint64_t *)ctx->scratch.p));
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, (uint64_t *)ctx->scratch.p + (se - 1), sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (out_len) *out_len = (size_t)total;
    return SDA_OK;
}

int sda_varint_decode_dev(sda_ctx *ctx, const uint8_t *d_buf, size_t len, int64_t *d_shares_out, size_t cap, size_t *n) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    const size_t se = varint_decode_scratch_elems(len);
    CU(ctx->scratch.reserve(se * sizeof(uint64_t)));
    OK(clear_flags(ctx));
    CU(launch_varint_decode(ctx->lc(), d_buf, len, d_shares_out, cap, (uint64_t *)ctx->scratch.p, ctx->d_flag));
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, (uint64_t *)ctx->scratch.p + (se - 1), sizeof total, cudaMemcpyDeviceToHost, ctx->stream));
    unsigned status = 0;
    OK(read_flags(ctx, &status, nullptr));     // synchronises the stream
    if (n) *n = (size_t)total;
    if (status & 2u) return fail(ctx, SDA_ERR_INVALID, "varint stream ends inside a value");
    if (status & 1u) return fail(ctx, SDA_ERR_INVALID, "varint value longer than 10 bytes");
    if (status & 4u) return fail(ctx, SDA_ERR_INVALID, "varint stream holds %llu values, capacity %zu", (unsigned long long)total, cap);
    return SDA_OK;
}

int sda_varint_encode(sda_ctx *ctx, const int64_t *shares, size_t n, uint8_t *out, size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (out_len) *out_len = 0;
    if (n == 0) return SDA_OK;
    if (!shares || !out) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    CU(ctx->in.reserve(n * sizeof(int64_t)));
    CU(ctx->out.reserve(10 * n + 16));
    OK(h2d(ctx, ctx->in.p, shares, n * sizeof(int64_t)));
    size_t len = 0;
    OK(sda_varint_encode_dev(ctx, (const int64_t *)ctx->in.p, n, (uint8_t *)ctx->out.p, &len));
    if (out_len) *out_len = len;
    return d2h(ctx, out, ctx->out.p, len);
}

int sda_varint_decode(sda_ctx *ctx, const uint8_t *buf, size_t len, int64_t *shares_out, size_t cap, size_t *n) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (n) *n = 0;
    if (len == 0) return SDA_OK;
    if (!buf || (!shares_out && cap)) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    CU(ctx->in.reserve(len + 16));
    CU(ctx->out.reserve(std::max<size_t>(cap, 1) * sizeof(int64_t)));
    OK(h2d(ctx, ctx->in.p, buf, len));
    size_t cnt = 0;
    OK(sda_varint_decode_dev(ctx, (const uint8_t *)ctx->in.p, len, (int64_t *)ctx->out.p, cap, &cnt));
    if (n) *n = cnt;
    return d2h(ctx, shares_out, ctx->out.p, cnt * sizeof(int64_t));
}

int sda_fixed_encode_dev(sda_ctx *ctx, int64_t modulus, int frac_bits, const float *d_x, size_t n, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    if (frac_bits < 0 || frac_bits > 52) return fail(ctx, SDA_ERR_INVALID, "frac_bits must be in [0, 52]");
    CU(launch_fixed_encode(ctx->lc(), make_field((uint64_t)modulus), frac_bits, d_x, n, d_out));
    return SDA_OK;
}

int sda_fixed_decode_dev(sda_ctx *ctx, int64_t modulus, int frac_bits, uint64_t divisor, const int64_t *d_in, size_t n,
                         float *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    if (frac_bits < 0 || frac_bits > 52) return fail(ctx, SDA_ERR_INVALID, "frac_bits must be in [0, 52]");
    if (divisor == 0) return fail(ctx, SDA_ERR_INVALID, "divisor must be positive");
    CU(launch_fixed_decode(ctx->lc(), make_field((uint64_t)modulus), frac_bits, divisor, d_in, n, d_out));
    return SDA_OK;
}

int sda_synth_fill_dev(sda_ctx *ctx, uint32_t stream, int64_t modulus, uint64_t start, size_t count, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    CU(launch_synth_fill(ctx->lc(), make_field((uint64_t)modulus), stream, start, count, d_out));
    return SDA_OK;
}

// ---- host-pointer entry points ---------------------------------------------------------------------

// One participant's packed-Shamir shares with the vector walked in slices of batches: slice i is generated while
// slice i + 1 is on its way in and slice i - 1 on its way out.  Taken for the tensor-core shapes with pinned
// host buffers; *done stays false when the call should run the plain copy / kernel / copy sequence instead
// (other schemes, pageable memory, short vectors, or a gen_range rejection, whose exact path reshuffles the
// whole keystream).  ctx->in (ldp elements) and ctx->out (n B elements) are reserved by the caller.
static int share_generate_sliced(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *secrets, size_t dim, size_t ldp,
                                 const uint8_t *seed, int64_t *shares_out, bool *done) {
    *done = false;
    Packed pk;
    OK(validate(ctx, s, &pk));
    if (s->kind != SDA_SHARING_PACKED_SHAMIR || ctx->packed_path == SDA_PACKED_PATH_CUDA_CORES || s->modulus < 3 || !seed)
        return SDA_OK;
    const size_t slice_unit = share_tc_slice_batches(ctx, pk, dim);
    if (slice_unit == 0 || dim * sizeof(int64_t) < PIPE_MIN_BYTES || dim * sizeof(int64_t) > PIPE_MAX_PITCH ||
        !is_pinned(secrets) || !is_pinned(shares_out))
        return SDA_OK;
    const size_t n = (size_t)pk.n, k = (size_t)pk.k, B = (dim + k - 1) / k;
    const size_t per = ((B + PIPE_SLICES - 1) / PIPE_SLICES + slice_unit - 1) / slice_unit * slice_unit;
    const FieldParams f = make_field((uint64_t)s->modulus);
    const DrawParams dr = make_draw((uint64_t)s->modulus - 1);
    Matrix M;
    OK(share_matrix_cached(ctx, pk, &M));
    OK(upload_keys(ctx, seed, 1));
    OK(clear_flags(ctx));
    OK(pipe_begin(ctx));
    const int64_t *d_in = (const int64_t *)ctx->in.p;
    int64_t *d_out = (int64_t *)ctx->out.p;
    size_t ei = 1;
    for (size_t b0 = 0; b0 < B; b0 += per) {
        const size_t nb = std::min(per, B - b0);
        const size_t e0 = b0 * k, e1 = std::min(dim, (b0 + nb) * k);
        cudaEvent_t in_ready, out_ready;
        OK(pipe_event(ctx, ei++, &in_ready));
        OK(pipe_event(ctx, ei++, &out_ready));
        CU(cudaMemcpyAsync((void *)(d_in + e0), secrets + e0, (e1 - e0) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->h2d_stream));
        CU(cudaEventRecord(in_ready, ctx->h2d_stream));
        CU(cudaStreamWaitEvent(ctx->stream, in_ready, 0));
        OK(launch_share_tc(ctx, pk, M, f, dr, d_in, ldp, 1, dim, b0, nb, (const ChaChaKey *)ctx->keys.p, d_out));
        CU(cudaEventRecord(out_ready, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->d2h_stream, out_ready, 0));
        CU(cudaMemcpy2DAsync(shares_out + b0, B * sizeof(int64_t), d_out + b0, B * sizeof(int64_t), nb * sizeof(int64_t), n,
                             cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    unsigned rejected = 0;
    OK(read_flags(ctx, &rejected, nullptr));
    CU(cudaStreamSynchronize(ctx->d2h_stream));
    *done = rejected == 0;
    return SDA_OK;
}

int sda_share_generate(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *secrets, size_t dim,
                       const uint8_t rng_seed[32], int64_t *shares_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(validate(ctx, s, nullptr));
    const size_t n = s->share_count, B = sda_share_batches(s, dim);
    if (dim == 0) return SDA_OK;
    if (!secrets || !shares_out) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    const size_t ldp = (dim + 3) & ~(size_t)3;
    CU(ctx->in.reserve(ldp * sizeof(int64_t)));
    CU(ctx->out.reserve(n * B * sizeof(int64_t)));
    bool done = false;
    OK(share_generate_sliced(ctx, s, secrets, dim, ldp, rng_seed, shares_out, &done));
    if (done) return SDA_OK;
    OK(h2d(ctx, ctx->in.p, secrets, dim * sizeof(int64_t)));
    OK(share_generate_core(ctx, s, (const int64_t *)ctx->in.p, ldp, 1, dim, rng_seed, (int64_t *)ctx->out.p));
    return d2h(ctx, shares_out, ctx->out.p, n * B * sizeof(int64_t));
}

// participate.rs:53-54 then :75-76 on host vectors: one upload of the secrets, the fused kernel (or the two steps on the
// device), one download of the mask and of the shares -- the masked secrets never cross the link (the two trait calls
// one after the other move them down and up again: 16 of 45 bytes per secret with a Full mask)
int sda_mask_share_generate(sda_ctx *ctx, const sda_masking_scheme *ms, const sda_sharing_scheme *ss, const int64_t *secrets,
                            size_t dim, const uint8_t mask_rng_seed[32], const uint8_t share_rng_seed[32], int64_t *mask_out,
                            int64_t *shares_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(mask_validate(ctx, ms));
    OK(validate(ctx, ss, nullptr));
    if (dim == 0 && ms->kind != SDA_MASK_CHACHA) return SDA_OK;
    const size_t mask_len = sda_mask_len(ms, dim);
    const size_t out_len = ss->kind == SDA_SHARING_PACKED_SHAMIR ? (size_t)ss->share_count * sda_share_batches(ss, dim)
                                                                 : (size_t)ss->share_count * dim;
    if ((dim && (!secrets || !shares_out)) || (mask_len && !mask_out)) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    const size_t ldp = (dim + 3) & ~(size_t)3;
    CU(ctx->in.reserve(std::max<size_t>(ldp, 4) * sizeof(int64_t)));
    CU(ctx->out.reserve(std::max<size_t>(out_len, 1) * sizeof(int64_t)));
    CU(ctx->aux.reserve(std::max<size_t>(mask_len, 1) * sizeof(int64_t)));
    if (dim) OK(h2d(ctx, ctx->in.p, secrets, dim * sizeof(int64_t)));
    const bool was_deferred = ctx->deferred;      // a host entry point hands its results out: its flag is read now
    ctx->deferred = false;
    const int rc = sda_mask_share_generate_dev(ctx, ms, ss, (const int64_t *)ctx->in.p, ldp, 1, dim, mask_rng_seed, share_rng_seed,
                                               (int64_t *)ctx->aux.p, (int64_t *)ctx->out.p);
    ctx->deferred = was_deferred;
    OK(rc);
    if (mask_len) OK(d2h(ctx, mask_out, ctx->aux.p, mask_len * sizeof(int64_t)));
    if (out_len) OK(d2h(ctx, shares_out, ctx->out.p, out_len * sizeof(int64_t)));
    return SDA_OK;
}

// streamed combine of a host matrix: tiles of rows through ctx->in, running sum in ctx->out
static int combine_host(sda_ctx *ctx, int64_t modulus, const int64_t *shares, const int64_t *const *rows, size_t P,
                        size_t L, int64_t *out) {
    if (L == 0) return SDA_OK;
    const size_t ldp = (L + 3) & ~(size_t)3;
    const size_t row_bytes = ldp * sizeof(int64_t);
    const size_t tile = std::max<size_t>(1, std::min<size_t>(std::max<size_t>(P, 1), (size_t(1) << 30) / row_bytes));
    CU(ctx->in.reserve(tile * row_bytes));
    CU(ctx->out.reserve(row_bytes));
    int64_t *d_in = (int64_t *)ctx->in.p, *d_acc = (int64_t *)ctx->out.p;
    if (P == 0) CU(cudaMemsetAsync(d_acc, 0, row_bytes, ctx->stream));
    // pinned rows: walk the columns in slices so that the rows of slice i + 1 arrive while slice i is summed and
    // the sums of slice i - 1 leave (the kernel takes any column range of the staged matrix: row stride ldp)
    // out == nullptr: the sum stays on the device (ctx->out), for the multi-GPU entry points
    bool pinned = P > 0 && L * sizeof(int64_t) >= PIPE_MIN_BYTES && L * sizeof(int64_t) <= PIPE_MAX_PITCH && (!out || is_pinned(out));
    if (pinned && shares) pinned = is_pinned(shares);
    for (size_t p = 0; pinned && !shares && p < P; p++) pinned = is_pinned(rows[p]);
    if (pinned) {
        const size_t per = ((L + PIPE_SLICES - 1) / PIPE_SLICES + 1023) / 1024 * 1024;
        const size_t se = combine_scratch_elems(ctx->sm_count, std::min(tile, P), std::min(per, L));
        if (se) CU(ctx->scratch.reserve(se * sizeof(int64_t)));
        OK(pipe_begin(ctx));
        cudaEvent_t summed = nullptr;
        for (size_t p0 = 0; p0 < P; p0 += tile) {
            const size_t pc = std::min(tile, P - p0);
            const bool last = p0 + pc == P;
            if (summed) CU(cudaStreamWaitEvent(ctx->h2d_stream, summed, 0));   // the previous row tile is consumed
            size_t ei = 1;                                                     // (a wait keeps the record it was queued behind)
            for (size_t c0 = 0; c0 < L; c0 += per) {
                const size_t nc = std::min(per, L - c0);
                cudaEvent_t in_ready;
                OK(pipe_event(ctx, ei++, &in_ready));
                OK(pipe_event(ctx, ei++, &summed));
                if (shares) {
                    CU(cudaMemcpy2DAsync(d_in + c0, row_bytes, shares + p0 * L + c0, L * sizeof(int64_t), nc * sizeof(int64_t), pc,
                                         cudaMemcpyHostToDevice, ctx->h2d_stream));
                } else {
                    for (size_t p = 0; p < pc; p++)
                        CU(cudaMemcpyAsync(d_in + p * ldp + c0, rows[p0 + p] + c0, nc * sizeof(int64_t), cudaMemcpyHostToDevice,
                                           ctx->h2d_stream));
                }
                CU(cudaEventRecord(in_ready, ctx->h2d_stream));
                CU(cudaStreamWaitEvent(ctx->stream, in_ready, 0));
                OK(combine_core(ctx, modulus, d_in + c0, ldp, pc, nc, p0 ? d_acc + c0 : nullptr, d_acc + c0));
                CU(cudaEventRecord(summed, ctx->stream));
                if (last && out) {
                    CU(cudaStreamWaitEvent(ctx->d2h_stream, summed, 0));
                    CU(cudaMemcpyAsync(out + c0, d_acc + c0, nc * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->d2h_stream));
                }
            }
        }
        CU(cudaStreamSynchronize(ctx->d2h_stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return SDA_OK;
    }
    for (size_t p0 = 0; p0 < P; p0 += tile) {
        const size_t pc = std::min(tile, P - p0);
        if (shares && ldp == L) {
            OK(h2d(ctx, d_in, shares + p0 * L, pc * L * sizeof(int64_t)));
        } else if (shares) {
            CU(cudaMemcpy2DAsync(d_in, row_bytes, shares + p0 * L, L * sizeof(int64_t), L * sizeof(int64_t), pc,
                                 cudaMemcpyHostToDevice, ctx->stream));
        } else {
            for (size_t p = 0; p < pc; p++)
                OK(h2d(ctx, d_in + p * ldp, shares ? shares + (p0 + p) * L : rows[p0 + p], L * sizeof(int64_t)));
        }
        OK(combine_core(ctx, modulus, d_in, ldp, pc, L, p0 ? d_acc : nullptr, d_acc));
    }
    if (!out) {
        CU(cudaStreamSynchronize(ctx->stream));
        return SDA_OK;
    }
    return d2h(ctx, out, d_acc, L * sizeof(int64_t));
}

int sda_share_combine(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *shares, size_t P, size_t L,
                      int64_t *out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(validate(ctx, s, nullptr));
    if (P == 0 || L == 0) return SDA_OK;   // combiner.rs:17: dimension of an empty input is 0
    if (!shares || !out) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    return combine_host(ctx, s->modulus, shares, nullptr, P, L, out);
}

int sda_share_combine_rows(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *const *rows, const size_t *row_lens,
                           size_t P, int64_t *out, size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(validate(ctx, s, nullptr));
    const size_t L = P ? row_lens[0] : 0;   // combiner.rs:17
    if (out_len) *out_len = L;
    for (size_t p = 0; p < P; p++)
        if (row_lens[p] != L) return fail(ctx, SDA_ERR_INVALID, "Wrong dimension");   // combiner.rs:21
    if (L == 0) return SDA_OK;
    return combine_host(ctx, s->modulus, nullptr, rows, P, L, out);
}

static int reconstruct_host(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                            const int64_t *shares, const int64_t *const *rows, const size_t *row_lens, size_t m,
                            size_t B, int64_t *secrets_out, size_t *out_len) {
    OK(validate(ctx, s, nullptr));
    if (s->kind == SDA_SHARING_ADDITIVE) {
        // additive.rs:56-70
        const size_t len = m ? (rows ? row_lens[0] : B) : 0;
        if (out_len) *out_len = len;
        if (rows)
            for (size_t i = 0; i < m; i++)
                if (row_lens[i] != len) return fail(ctx, SDA_ERR_INVALID, "Mismatching dimension");   // additive.rs:64
        if (len == 0) return SDA_OK;
        return combine_host(ctx, s->modulus, shares, rows, m, len, secrets_out);
    }
    const size_t k = s->secret_count;
    const size_t nb = (dimension + k - 1) / k;
    if (out_len) *out_len = dimension;
    if (nb == 0) return SDA_OK;
    if (m < sda_reconstruction_threshold(s)) return fail(ctx, SDA_ERR_INVALID, "Not enough shares to reconstruct");
    size_t minlen = B;
    if (rows) {
        minlen = (size_t)-1;
        for (size_t i = 0; i < m; i++) minlen = std::min(minlen, row_lens[i]);
    }
    if (minlen < nb)
        return fail(ctx, SDA_ERR_INVALID, "index out of bounds: the len is %zu but the index is %zu", minlen, minlen);
    const size_t ldp = (nb + 3) & ~(size_t)3;
    CU(ctx->in.reserve(m * ldp * sizeof(int64_t)));
    for (size_t i = 0; i < m; i++)
        OK(h2d(ctx, (int64_t *)ctx->in.p + i * ldp, rows ? rows[i] : shares + i * B, nb * sizeof(int64_t)));
    CU(ctx->out.reserve(dimension * sizeof(int64_t)));
    OK(reconstruct_core(ctx, s, dimension, indices, (const int64_t *)ctx->in.p, ldp, m, nb, (int64_t *)ctx->out.p, nullptr));
    return d2h(ctx, secrets_out, ctx->out.p, dimension * sizeof(int64_t));
}

int sda_secret_reconstruct(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                           const int64_t *shares, size_t m, size_t B, int64_t *secrets_out, size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return reconstruct_host(ctx, s, dimension, indices, shares, nullptr, nullptr, m, B, secrets_out, out_len);
}

int sda_secret_reconstruct_rows(sda_ctx *ctx, const sda_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                                const int64_t *const *rows, const size_t *row_lens, size_t m, int64_t *secrets_out,
                                size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    return reconstruct_host(ctx, s, dimension, indices, nullptr, rows, row_lens, m, 0, secrets_out, out_len);
}

int sda_mask(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *secrets, size_t dim, const uint8_t rng_seed[32],
             int64_t *mask_out, size_t *mask_len, int64_t *masked_out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(mask_validate(ctx, s));
    const size_t ml = sda_mask_len(s, dim);
    if (mask_len) *mask_len = ml;
    if (s->kind == SDA_MASK_CHACHA && s->dimension != dim)
        return fail(ctx, SDA_ERR_INVALID, "assertion failed: `(left == right)` (chacha.rs:26: scheme dimension %llu, %zu secrets)",
                    (unsigned long long)s->dimension, dim);
    const size_t bytes = dim * sizeof(int64_t);
    CU(ctx->in.reserve(std::max<size_t>(bytes, 8)));
    CU(ctx->out.reserve(std::max<size_t>(bytes, 8)));
    CU(ctx->aux.reserve(std::max<size_t>(bytes, 8)));
    OK(h2d(ctx, ctx->in.p, secrets, bytes));
    int64_t words[8] = {0};
    OK(mask_core(ctx, s, (const int64_t *)ctx->in.p, dim, rng_seed, (int64_t *)ctx->aux.p, words, (int64_t *)ctx->out.p));
    if (s->kind == SDA_MASK_FULL) OK(d2h(ctx, mask_out, ctx->aux.p, bytes));
    if (s->kind == SDA_MASK_CHACHA)
        for (size_t i = 0; i < ml; i++) mask_out[i] = words[i];
    return d2h(ctx, masked_out, ctx->out.p, bytes);
}

int sda_mask_combine(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *masks, size_t P, size_t mask_len,
                     int64_t *out, size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(mask_validate(ctx, s));
    switch (s->kind) {
    case SDA_MASK_NONE:   // none.rs:22-25
        if (mask_len != 0) return fail(ctx, SDA_ERR_INVALID, "assertion failed: masks.iter().all(|mask| mask.len() == 0)");
        if (out_len) *out_len = 0;
        return SDA_OK;
    case SDA_MASK_FULL: {   // full.rs:38-51
        const size_t L = P ? mask_len : 0;
        if (out_len) *out_len = L;
        if (L == 0) return SDA_OK;
        return combine_host(ctx, s->modulus, masks, nullptr, P, L, out);
    }
    default: {   // chacha.rs:57-76
        const size_t dim = (size_t)s->dimension;
        if (out_len) *out_len = dim;
        if (dim == 0) return SDA_OK;
        CU(ctx->out.reserve(dim * sizeof(int64_t)));
        OK(chacha_mask_combine_core(ctx, s, masks, P, mask_len, (int64_t *)ctx->out.p));
        return d2h(ctx, out, ctx->out.p, dim * sizeof(int64_t));
    }
    }
}

int sda_unmask(sda_ctx *ctx, const sda_masking_scheme *s, const int64_t *mask, size_t mask_len, const int64_t *masked,
               size_t dim, int64_t *out) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(mask_validate(ctx, s));
    if (s->kind == SDA_MASK_NONE) {   // none.rs:29-32
        if (mask_len != 0) return fail(ctx, SDA_ERR_INVALID, "assertion failed: `(left == right)` (none.rs:30)");
        if (dim) memmove(out, masked, dim * sizeof(int64_t));
        return SDA_OK;
    }
    if (mask_len != dim)   // full.rs:58, chacha.rs:83
        return fail(ctx, SDA_ERR_INVALID, "assertion failed: `(left == right)` (full.rs:58 / chacha.rs:83: mask %zu, masked %zu)",
                    mask_len, dim);
    if (dim == 0) return SDA_OK;
    const size_t bytes = dim * sizeof(int64_t);
    CU(ctx->in.reserve(bytes));
    CU(ctx->aux.reserve(bytes));
    CU(ctx->out.reserve(bytes));
    OK(h2d(ctx, ctx->in.p, masked, bytes));
    OK(h2d(ctx, ctx->aux.p, mask, bytes));
    CU(launch_submod(ctx->lc(), make_field((uint64_t)s->modulus), (const int64_t *)ctx->in.p, (const int64_t *)ctx->aux.p,
                     dim, (int64_t *)ctx->out.p));
    return d2h(ctx, out, ctx->out.p, bytes);
}

// ---- server snapshot transpose -------------------------------------------------------------------
int sda_snapshot_transpose_dev(sda_ctx *ctx, const uint8_t *d_blobs, const uint64_t *offsets, size_t P, size_t n,
                               uint8_t *d_out, uint64_t *out_offsets) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    const size_t blobs = P * n;
    if (!offsets || !out_offsets) return fail(ctx, SDA_ERR_INVALID, "null offsets");
    // clerk-major offsets: clerk c's job is blob c of every participation, in participation order (stores.rs:95-99)
    for (size_t i = 0; i < blobs; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, SDA_ERR_INVALID, "offsets must be non-decreasing");
    uint64_t at = 0;
    for (size_t c = 0; c < n; c++)
        for (size_t p = 0; p < P; p++) {
            out_offsets[c * P + p] = at;
            at += offsets[p * n + c + 1] - offsets[p * n + c];
        }
    out_offsets[blobs] = at;
    if (blobs == 0 || at == 0) return SDA_OK;
    if (!d_blobs || !d_out) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    CU(ctx->scratch.reserve(2 * (blobs + 1) * sizeof(uint64_t)));
    uint64_t *d_in_off = (uint64_t *)ctx->scratch.p, *d_out_off = d_in_off + blobs + 1;
    CU(cudaMemcpyAsync(d_in_off, offsets, (blobs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_out_off, out_offsets, (blobs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_snapshot_transpose(ctx->lc(), d_blobs, d_in_off, P, n, d_out, d_out_off));
    CU(cudaStreamSynchronize(ctx->stream));   // the offset arrays are the caller's (pageable) memory
    return SDA_OK;
}

// ---- multi-GPU clerk sum (clerk.rs:85-86 when one box holds several GPUs) ------------------------------

int sda_nccl_unique_id(uint8_t id_out[SDA_NCCL_UNIQUE_ID_BYTES]) {
    sda_ctx *ctx = nullptr;
    if (!id_out) return fail(ctx, SDA_ERR_INVALID, "null id buffer");
    const NcclApi *api;
    OK(need_nccl(ctx, &api));
    static_assert(sizeof(ncclUniqueId) == SDA_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return SDA_OK;
}

int sda_ctx_comm_init_rank(sda_ctx *ctx, const uint8_t id[SDA_NCCL_UNIQUE_ID_BYTES], int nranks, int rank) {
    if (!ctx) return SDA_ERR_INVALID;
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, SDA_ERR_INVALID, "bad communicator arguments");
    if (!ctx->members.empty()) return fail(ctx, SDA_ERR_INVALID, "context already belongs to a one-process device group");
    const NcclApi *api;
    OK(need_nccl(ctx, &api));
    DeviceGuard g(ctx->device);
    comm_release(ctx);
    ncclUniqueId nid;
    memcpy(&nid, id, sizeof nid);
    NC(api->CommInitRank(&ctx->comm, nranks, nid, rank));
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    return SDA_OK;
}

int sda_ctx_comm_rank(const sda_ctx *ctx) { return ctx ? ctx->comm_rank : -1; }
int sda_ctx_comm_size(const sda_ctx *ctx) { return ctx ? ctx->comm_size : 0; }

int sda_partial_sums_reduce_dev(sda_ctx *ctx, int64_t modulus, int64_t *d_partials, size_t count, int root) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (modulus <= 0) return fail(ctx, SDA_ERR_INVALID, "modulus must be positive");
    if (root < 0 || root >= ctx->comm_size) return fail(ctx, SDA_ERR_INVALID, "root %d outside the communicator", root);
    if (count == 0) return SDA_OK;
    if (!ctx->members.empty()) return fail(ctx, SDA_ERR_INVALID, "one-process device groups reduce through sda_share_combine_multi*");
    if (ctx->comm_size == 1) return SDA_OK;   // a single rank's canonical partial sums are the result
    if (!ctx->comm) return fail(ctx, SDA_ERR_INVALID, "no communicator: call sda_ctx_comm_init_rank first");
    const NcclApi *api;
    OK(need_nccl(ctx, &api));
    const FieldParams f = make_field((uint64_t)modulus);
    bool fold = false;
    OK(reduce_partials_enqueue(ctx, api, f, d_partials, count, root, &fold));
    return reduce_partials_finish(ctx, f, d_partials, count, root, fold);
}

int sda_share_combine_ranks_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *d_shares, size_t ld, size_t P_local,
                                size_t L, const int64_t *d_acc_in, int64_t *d_out, int root) {
    if (!ctx) return SDA_ERR_INVALID;
    DeviceGuard g(ctx->device);
    OK(validate(ctx, s, nullptr));
    if (ld < L) return fail(ctx, SDA_ERR_INVALID, "Wrong dimension");
    if (L == 0) return SDA_OK;
    if (P_local == 0 && !d_acc_in) CU(cudaMemsetAsync(d_out, 0, L * sizeof(int64_t), ctx->stream));   // a rank without rows adds 0
    else OK(combine_core(ctx, s->modulus, d_shares, ld, P_local, L, d_acc_in, d_out));
    return sda_partial_sums_reduce_dev(ctx, s->modulus, d_out, L, root);
}

int sda_ctx_create_multi(const int *devices, int ndev, sda_ctx **out) {
    sda_ctx *ctx = nullptr;
    if (!out) return SDA_ERR_INVALID;
    *out = nullptr;
    if (!devices || ndev < 1) return fail(ctx, SDA_ERR_INVALID, "need at least one device");
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return fail(ctx, SDA_ERR_INVALID, "device %d listed twice", devices[i]);
    std::vector<sda_ctx *> m((size_t)ndev, nullptr);
    auto undo = [&]() {
        for (int i = ndev - 1; i >= 0; i--)
            if (m[i]) {
                m[i]->members.clear();
                sda_ctx_destroy(m[i]);
            }
    };
    for (int i = 0; i < ndev; i++) {
        const int rc = sda_ctx_create(devices[i], &m[i]);
        if (rc != SDA_OK) {
            undo();
            return rc;   // message already in the thread's create error
        }
    }
    if (ndev > 1) {
        const NcclApi *api = nccl_api();
        std::vector<ncclComm_t> comms((size_t)ndev, nullptr);
        ncclResult_t r = api->handle ? api->CommInitAll(comms.data(), ndev, devices) : ncclSystemError;
        if (r != ncclSuccess) {
            const std::string why = api->handle ? api->GetErrorString(r) : "libnccl.so.2 could not be loaded: " + api->error;
            undo();
            return fail(ctx, SDA_ERR_NCCL, "NCCL communicator over %d devices: %s", ndev, why.c_str());
        }
        for (int i = 0; i < ndev; i++) {
            m[i]->comm = comms[i];
            m[i]->comm_rank = i;
            m[i]->comm_size = ndev;
        }
    }
    m[0]->members = m;
    *out = m[0];
    return SDA_OK;
}

int sda_ctx_multi_count(const sda_ctx *ctx) { return ctx ? (ctx->members.empty() ? 1 : (int)ctx->members.size()) : 0; }
sda_ctx *sda_ctx_multi_member(sda_ctx *ctx, int i) {
    if (!ctx) return nullptr;
    if (ctx->members.empty()) return i == 0 ? ctx : nullptr;
    return i >= 0 && i < (int)ctx->members.size() ? ctx->members[(size_t)i] : nullptr;
}

// member i's canonical partial sums (d_part[i], `count` elements on device i) -> totals mod m in d_part[0]
static int multi_reduce(sda_ctx *ctx, int64_t modulus, const std::vector<int64_t *> &d_part, size_t count) {
    const size_t G = ctx->members.size();
    if (G <= 1) return SDA_OK;
    const NcclApi *api;
    OK(need_nccl(ctx, &api));
    const FieldParams f = make_field((uint64_t)modulus);
    std::vector<char> fold(G, 0);
    NC(api->GroupStart());
    int rc = SDA_OK;
    for (size_t i = 0; i < G && rc == SDA_OK; i++) {
        sda_ctx *mi = ctx->members[i];
        DeviceGuard g(mi->device);
        bool fl = false;
        rc = reduce_partials_enqueue(mi, api, f, d_part[i], count, 0, &fl);
        fold[i] = fl;
        if (rc != SDA_OK && mi != ctx) ctx->err = mi->err;
    }
    NC(api->GroupEnd());
    OK(rc);
    DeviceGuard g(ctx->device);
    OK(reduce_partials_finish(ctx, f, d_part[0], count, 0, fold[0] != 0));
    for (size_t i = 1; i < G; i++) {       // the members' streams have nothing left but the collective
        DeviceGuard gi(ctx->members[i]->device);
        CU(cudaStreamSynchronize(ctx->members[i]->stream));
    }
    return SDA_OK;
}

int sda_share_combine_multi_dev(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *const *d_shares, size_t ld,
                                const size_t *P_per_device, size_t L, int64_t *const *d_partials, int64_t *d_out) {
    if (!ctx) return SDA_ERR_INVALID;
    OK(validate(ctx, s, nullptr));
    if (!d_shares || !P_per_device || !d_partials || !d_out) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    if (ld < L) return fail(ctx, SDA_ERR_INVALID, "Wrong dimension");
    if (L == 0) return SDA_OK;
    const size_t G = (size_t)sda_ctx_multi_count(ctx);
    std::vector<int64_t *> part(G);
    for (size_t i = 0; i < G; i++) {
        sda_ctx *mi = sda_ctx_multi_member(ctx, (int)i);
        DeviceGuard g(mi->device);
        part[i] = i == 0 ? d_out : d_partials[i];
        int rc = SDA_OK;
        if (P_per_device[i] == 0) {
            if (cudaMemsetAsync(part[i], 0, L * sizeof(int64_t), mi->stream) != cudaSuccess) rc = fail(mi, SDA_ERR_CUDA, "cudaMemsetAsync failed");
        } else {
            rc = combine_core(mi, s->modulus, d_shares[i], ld, P_per_device[i], L, nullptr, part[i]);
        }
        if (rc != SDA_OK) {
            if (mi != ctx) ctx->err = mi->err;
            return rc;
        }
    }
    return multi_reduce(ctx, s->modulus, part, L);
}

int sda_share_combine_rows_multi(sda_ctx *ctx, const sda_sharing_scheme *s, const int64_t *const *rows, const size_t *row_lens,
                                 size_t P, int64_t *out, size_t *out_len) {
    if (!ctx) return SDA_ERR_INVALID;
    OK(validate(ctx, s, nullptr));
    const size_t L = P ? row_lens[0] : 0;   // combiner.rs:17
    if (out_len) *out_len = L;
    for (size_t p = 0; p < P; p++)
        if (row_lens[p] != L) return fail(ctx, SDA_ERR_INVALID, "Wrong dimension");   // combiner.rs:21
    if (L == 0) return SDA_OK;
    if (!rows || !out) return fail(ctx, SDA_ERR_INVALID, "null buffer");
    const size_t G = (size_t)sda_ctx_multi_count(ctx);
    // contiguous blocks of participants per device, each summed by its own context on its own host thread
    // (every device has its own PCIe link; one context per thread is the library's concurrency model)
    std::vector<int> rcs(G, SDA_OK);
    std::vector<std::thread> th;
    const size_t per = (P + G - 1) / G;
    for (size_t i = 0; i < G; i++) {
        th.emplace_back([&, i]() {
            sda_ctx *mi = sda_ctx_multi_member(ctx, (int)i);
            const size_t p0 = std::min(P, i * per), pc = std::min(per, P - p0);
            cudaSetDevice(mi->device);
            rcs[i] = combine_host(mi, s->modulus, nullptr, rows + p0, pc, L, nullptr);   // sum stays in mi->out
        });
    }
    for (auto &t : th) t.join();
    std::vector<int64_t *> part(G);
    for (size_t i = 0; i < G; i++) {
        sda_ctx *mi = sda_ctx_multi_member(ctx, (int)i);
        if (rcs[i] != SDA_OK) {
            if (mi != ctx) ctx->err = mi->err;
            return rcs[i];
        }
        part[i] = (int64_t *)mi->out.p;
    }
    OK(multi_reduce(ctx, s->modulus, part, L));
    DeviceGuard g(ctx->device);
    return d2h(ctx, out, part[0], L * sizeof(int64_t));
}

}  // extern "C"
#pragma GCC visibility pop
