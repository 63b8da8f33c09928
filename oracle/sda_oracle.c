/*
 * sda_oracle.c -- CPU ORACLE (test infrastructure only; see sda_oracle.h for the rules).
 *
 * Plain C restatement of the reference's hot path.  Citations are paths relative to
 * /root/reference (snipsco/sda @ 8cf97f2).  Arithmetic is signed i64 with truncating
 * remainder exactly like Rust's `%`; products are widened to __int128 so that the same
 * code is also defined for the 61-bit primes of BASELINE.json configs #2-#5, where the
 * reference itself (i64 products) overflows.  For p < 2^31 widening changes nothing.
 * Build with -fwrapv (Rust release builds wrap on i64 add/sub overflow).
 */
#define _GNU_SOURCE
#include "sda_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>

typedef __int128 i128;
typedef unsigned __int128 u128;

static __thread char g_err[256];
const char *sdao_last_error(void) { return g_err; }
static int fail(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 1;
}

/* truncating remainder of a widened product: Rust `(a * b) % m` without the overflow */
static inline int64_t mulrem(int64_t a, int64_t b, int64_t m) { return (int64_t)(((i128)a * b) % m); }

/* ======================================================================================
 * rand 0.3  (external crate; model of ChaChaRng / OsRng / Rng::gen_range)
 *   call sites: client/src/crypto/sharing/additive.rs:4,17,43
 *               client/src/crypto/masking/full.rs:5,16,25
 *               client/src/crypto/masking/chacha.rs:5,29,32,36,38,67,69
 * ====================================================================================== */

#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d)                 \
    do {                               \
        a += b; d ^= a; d = ROTL32(d, 16); \
        c += d; b ^= c; b = ROTL32(b, 12); \
        a += b; d ^= a; d = ROTL32(d, 8);  \
        c += d; b ^= c; b = ROTL32(b, 7);  \
    } while (0)

/* rand-0.3 chacha.rs `core`: `rounds`/2 double rounds, then add the input state. */
void sdao_chacha_block(const uint32_t state[16], int rounds, uint32_t out[16]) {
    uint32_t x[16];
    memcpy(x, state, sizeof x);
    for (int i = 0; i < rounds / 2; i++) {
        QR(x[0], x[4], x[8], x[12]);
        QR(x[1], x[5], x[9], x[13]);
        QR(x[2], x[6], x[10], x[14]);
        QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]);
        QR(x[1], x[6], x[11], x[12]);
        QR(x[2], x[7], x[8], x[13]);
        QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + state[i];
}

/* ChaChaRng::from_seed(&[u32]): "expa""nd 3""2-by""te k" constants, up to 8 key words taken
 * from the seed (rest zero), 128-bit block counter = 0, no nonce. */
void sdao_rng_from_seed_rounds(sdao_rng *r, const uint32_t *seed, size_t n, int rounds) {
    memset(r, 0, sizeof *r);
    r->kind = 0;
    r->rounds = rounds;
    r->state[0] = 0x61707865u; r->state[1] = 0x3320646eu;
    r->state[2] = 0x79622d32u; r->state[3] = 0x6b206574u;
    for (size_t i = 0; i < n && i < 8; i++) r->state[4 + i] = seed[i];
    r->idx = 16;
}
void sdao_rng_from_seed(sdao_rng *r, const uint32_t *seed, size_t n) {
    sdao_rng_from_seed_rounds(r, seed, n, 20);
}
void sdao_rng_os(sdao_rng *r) {
    memset(r, 0, sizeof *r);
    r->kind = 1;
}

uint32_t sdao_rng_next_u32(sdao_rng *r) {
    if (r->kind == 1) { /* OsRng: one OS request per value */
        uint32_t v;
        if (getrandom(&v, sizeof v, 0) != (ssize_t)sizeof v) abort();
        return v;
    }
    if (r->idx == 16) { /* ChaChaRng::update: emit block, then 128-bit counter += 1 */
        sdao_chacha_block(r->state, r->rounds, r->buf);
        r->idx = 0;
        if (++r->state[12] == 0 && ++r->state[13] == 0 && ++r->state[14] == 0) ++r->state[15];
    }
    return r->buf[r->idx++];
}

/* Rng::next_u64 default: first word is the HIGH half */
uint64_t sdao_rng_next_u64(sdao_rng *r) {
    if (r->kind == 1) {
        uint64_t v;
        if (getrandom(&v, sizeof v, 0) != (ssize_t)sizeof v) abort();
        return v;
    }
    uint64_t hi = sdao_rng_next_u32(r);
    uint64_t lo = sdao_rng_next_u32(r);
    return (hi << 32) | lo;
}

/* Rng::gen_range(low, high) for i64 == Range::new(low, high).ind_sample:
 * range = high - low (as u64); zone = MAX - MAX % range; loop { v = next_u64; accept if v < zone } */
int64_t sdao_gen_range(sdao_rng *r, int64_t low, int64_t high) {
    if (!(low < high)) abort(); /* assert!(low < high, "Rng.gen_range called with low >= high") */
    uint64_t range = (uint64_t)high - (uint64_t)low;
    uint64_t zone = UINT64_MAX - UINT64_MAX % range;
    for (;;) {
        uint64_t v = sdao_rng_next_u64(r);
        r->draws++;
        if (v < zone) return (int64_t)((uint64_t)low + v % range);
        r->rejections++;
    }
}

/* ======================================================================================
 * protocol/src/crypto.rs:117-155  derived scheme properties
 * ====================================================================================== */
size_t sdao_input_size(const sdao_sharing_scheme *s) {
    return s->kind == SDAO_SHARING_ADDITIVE ? 1 : (size_t)s->secret_count;
}
size_t sdao_output_size(const sdao_sharing_scheme *s) { return (size_t)s->share_count; }
size_t sdao_privacy_threshold(const sdao_sharing_scheme *s) {
    return s->kind == SDAO_SHARING_ADDITIVE ? (size_t)s->share_count - 1 : (size_t)s->privacy_threshold;
}
size_t sdao_reconstruction_threshold(const sdao_sharing_scheme *s) {
    return s->kind == SDAO_SHARING_ADDITIVE ? (size_t)s->share_count
                                            : (size_t)(s->privacy_threshold + s->secret_count);
}

/* ======================================================================================
 * threshold-secret-sharing 0.2 (external crate `tss`): numtheory + packed
 *   call sites: client/src/crypto/sharing/packed_shamir.rs:14-21,42,55-62,75,76
 * ====================================================================================== */

/* numtheory::mod_pow: square-and-multiply, results are signed representatives */
int64_t sdao_mod_pow(int64_t x, uint64_t e, int64_t p) {
    int64_t acc = 1;
    while (e > 0) {
        if (e % 2 == 0) { x = mulrem(x, x, p); e /= 2; }
        else            { acc = mulrem(acc, x, p); e -= 1; }
    }
    return acc;
}

/* numtheory::gcd (extended Euclid), iterative form of the same recurrence */
static void egcd(int64_t a, int64_t b, int64_t *g, int64_t *x, int64_t *y) {
    if (b == 0) { *g = a; *x = 1; *y = 0; return; }
    int64_t g1, x1, y1;
    egcd(b, a % b, &g1, &x1, &y1);
    *g = g1; *x = y1; *y = x1 - y1 * (a / b);
}
/* numtheory::mod_inverse */
int64_t sdao_mod_inverse(int64_t k, int64_t p) {
    int64_t k2 = k % p, g, x, y;
    int64_t r;
    if (k2 < 0) { egcd(p, -k2, &g, &x, &y); r = -y; }
    else        { egcd(p, k2, &g, &x, &y); r = y; }
    return (p + r) % p;
}

/* numtheory::fft2: A(x) = B(x^2) + x C(x^2), recursive, a Vec per level */
void sdao_fft2(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out) {
    if (n == 1) { out[0] = a[0]; return; }
    size_t h = n / 2;
    int64_t *b = malloc(sizeof(int64_t) * 4 * h), *c = b + h, *bp = c + h, *cp = bp + h;
    for (size_t i = 0; i < h; i++) { b[i] = a[2 * i]; c[i] = a[2 * i + 1]; }
    int64_t w2 = mulrem(omega, omega, p);
    sdao_fft2(b, h, w2, p, bp);
    sdao_fft2(c, h, w2, p, cp);
    for (size_t i = 0; i < h; i++) {
        int64_t x = sdao_mod_pow(omega, i, p); /* re-evaluated inside the butterfly, as tss does */
        int64_t xc = mulrem(x, cp[i], p);
        out[i] = (bp[i] + xc) % p;
        out[i + h] = (bp[i] - xc) % p;
    }
    free(b);
}
void sdao_fft2_inverse(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out) {
    int64_t winv = sdao_mod_inverse(omega, p);
    int64_t ninv = sdao_mod_inverse((int64_t)n, p);
    sdao_fft2(a, n, winv, p, out);
    for (size_t i = 0; i < n; i++) out[i] = mulrem(out[i], ninv, p);
}
/* numtheory::fft3: A(x) = B(x^3) + x C(x^3) + x^2 D(x^3) */
void sdao_fft3(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out) {
    if (n == 1) { out[0] = a[0]; return; }
    size_t h = n / 3;
    int64_t *b = malloc(sizeof(int64_t) * 6 * h), *c = b + h, *d = c + h;
    int64_t *bp = d + h, *cp = bp + h, *dp = cp + h;
    for (size_t i = 0; i < h; i++) { b[i] = a[3 * i]; c[i] = a[3 * i + 1]; d[i] = a[3 * i + 2]; }
    int64_t w3 = mulrem(mulrem(omega, omega, p), omega, p);
    sdao_fft3(b, h, w3, p, bp);
    sdao_fft3(c, h, w3, p, cp);
    sdao_fft3(d, h, w3, p, dp);
    for (size_t i = 0; i < h; i++) {
        for (size_t q = 0; q < 3; q++) {
            size_t j = i + q * h;
            int64_t x = sdao_mod_pow(omega, j, p);
            int64_t xx = mulrem(x, x, p);
            int64_t v = (bp[i] + mulrem(x, cp[i], p)) % p;
            out[j] = (v + mulrem(xx, dp[i], p)) % p;
        }
    }
    free(b);
}
void sdao_fft3_inverse(const int64_t *a, size_t n, int64_t omega, int64_t p, int64_t *out) {
    int64_t winv = sdao_mod_inverse(omega, p);
    int64_t ninv = sdao_mod_inverse((int64_t)n, p);
    sdao_fft3(a, n, winv, p, out);
    for (size_t i = 0; i < n; i++) out[i] = mulrem(out[i], ninv, p);
}

/* numtheory::compute_newton_coefficients: divided differences, one mod_inverse each */
static void newton_coefficients(const int64_t *pts, const int64_t *vals, size_t n, int64_t p, int64_t *coef) {
    memcpy(coef, vals, n * sizeof(int64_t));
    for (size_t j = 1; j < n; j++) {
        for (size_t i = n - 1; i >= j; i--) {
            int64_t pd = (pts[i] - pts[i - j]) % p;
            int64_t pdinv = sdao_mod_inverse(pd, p);
            int64_t cd = (coef[i] - coef[i - 1]) % p;
            coef[i] = mulrem(cd, pdinv, p);
        }
    }
}
/* numtheory::newton_evaluate */
static int64_t newton_evaluate(const int64_t *pts, const int64_t *coef, size_t n, int64_t x, int64_t p) {
    int64_t np = 1, acc = 0;
    for (size_t i = 0; i < n; i++) {
        acc = (acc + mulrem(coef[i], np, p)) % p;
        if (i + 1 < n) np = mulrem(np, (x - pts[i]) % p, p);
    }
    return acc;
}

static int is_pow(size_t v, size_t base) {
    if (v == 0) return 0;
    while (v % base == 0) v /= base;
    return v == 1;
}

/* packed::PackedSecretSharing::share, randomness passed in (the crate draws it from OsRng).
 *   values = [0] ++ secrets ++ randomness            (recover_polynomial)
 *   coef   = fft2_inverse(values, omega_secrets)     (k+t+1 must be 2^a)
 *   coef  ++= zeros up to share_count+1              (must be 3^b)
 *   points = fft3(coef, omega_shares); assert points[0]==0; shares = points[1..]
 * General sizes (spec extension, SURVEY item 6): Newton-interpolate through
 * (omega_secrets^i, values[i]) and evaluate at omega_shares^j, j=1..n -- the same map. */
int sdao_tss_share_with_randomness(const sdao_sharing_scheme *s, const int64_t *secrets,
                                   const int64_t *randomness, int force_general,
                                   int64_t *poly_out, int64_t *shares_out) {
    size_t k = s->secret_count, t = s->privacy_threshold, n = s->share_count;
    int64_t p = s->modulus;
    size_t m = k + t + 1;
    int64_t *values = malloc(sizeof(int64_t) * m);
    values[0] = 0;
    for (size_t i = 0; i < k; i++) values[1 + i] = secrets[i];
    for (size_t i = 0; i < t; i++) values[1 + k + i] = randomness[i];

    int fft_ok = !force_general && is_pow(m, 2) && is_pow(n + 1, 3) && m <= n + 1;
    if (fft_ok) {
        int64_t *coef = calloc(n + 1, sizeof(int64_t));
        int64_t *points = malloc(sizeof(int64_t) * (n + 1));
        sdao_fft2_inverse(values, m, s->omega_secrets, p, coef);
        if (poly_out) memcpy(poly_out, coef, m * sizeof(int64_t));
        sdao_fft3(coef, n + 1, s->omega_shares, p, points);
        int bad = (points[0] % p) != 0;
        memcpy(shares_out, points + 1, n * sizeof(int64_t));
        free(coef); free(points); free(values);
        if (bad) return fail("tss: share polynomial does not vanish at 1");
        return 0;
    }
    int64_t *pts = malloc(sizeof(int64_t) * 2 * m), *coef = pts + m;
    for (size_t i = 0; i < m; i++) pts[i] = sdao_mod_pow(s->omega_secrets, i, p);
    newton_coefficients(pts, values, m, p, coef);
    if (poly_out) memset(poly_out, 0, m * sizeof(int64_t)); /* monomial form not produced here */
    for (size_t j = 1; j <= n; j++) {
        int64_t x = sdao_mod_pow(s->omega_shares, j, p);
        shares_out[j - 1] = newton_evaluate(pts, coef, m, x, p);
    }
    free(pts); free(values);
    return 0;
}

/* packed::PackedSecretSharing::reconstruct:
 *   points = [1] ++ [omega_shares^(i+1) for i in indices]; values = [0] ++ shares
 *   Newton interpolation; evaluate at omega_secrets^e, e = 1..k */
int sdao_tss_reconstruct(const sdao_sharing_scheme *s, const uint64_t *indices,
                         const int64_t *shares, size_t m, int64_t *secrets_out) {
    size_t k = s->secret_count;
    int64_t p = s->modulus;
    int64_t *pts = malloc(sizeof(int64_t) * 3 * (m + 1)), *vals = pts + m + 1, *coef = vals + m + 1;
    pts[0] = 1; vals[0] = 0;
    for (size_t i = 0; i < m; i++) {
        pts[1 + i] = sdao_mod_pow(s->omega_shares, indices[i] + 1, p);
        vals[1 + i] = shares[i];
    }
    newton_coefficients(pts, vals, m + 1, p, coef);
    for (size_t e = 1; e <= k; e++) {
        int64_t x = sdao_mod_pow(s->omega_secrets, e, p);
        secrets_out[e - 1] = newton_evaluate(pts, coef, m + 1, x, p);
    }
    free(pts);
    return 0;
}

/* ======================================================================================
 * client/src/crypto/sharing
 * ====================================================================================== */

/* additive.rs:32-51  AdditiveSecretSharing::generate_for_batch */
static int additive_generate_for_batch(const sdao_sharing_scheme *s, const int64_t *batch, size_t blen,
                                       sdao_rng *rng, int64_t *shares) {
    if (blen != 1) return fail("Batch input wrong length");
    int64_t secret = batch[0];
    size_t n = s->share_count;
    for (size_t j = 0; j + 1 < n; j++) shares[j] = sdao_gen_range(rng, 0, s->modulus); /* :42-44 */
    int64_t last = secret;
    for (size_t j = 0; j + 1 < n; j++) last = (last - shares[j]) % s->modulus;         /* :47 */
    shares[n - 1] = last;
    return 0;
}

/* packed_shamir.rs:40-43 -> tss share(): t draws from Range::new(0, prime - 1) on the rng */
static int packed_generate_for_batch(const sdao_sharing_scheme *s, const int64_t *batch, size_t blen,
                                     sdao_rng *rng, int64_t *shares) {
    if (blen != s->secret_count) return fail("Sharing failed for packed secret sharing scheme");
    int64_t rnd[64];
    size_t t = s->privacy_threshold;
    if (t > 64) return fail("oracle: privacy_threshold > 64 unsupported");
    for (size_t i = 0; i < t; i++) rnd[i] = sdao_gen_range(rng, 0, s->modulus - 1);
    return sdao_tss_share_with_randomness(s, batch, rnd, 0, NULL, shares);
}

/* batched.rs:18-53  impl<G: BatchShareGenerator> ShareGenerator for G */
int sdao_share_generate(const sdao_sharing_scheme *s, const int64_t *secrets, size_t dim,
                        sdao_rng *rng, int64_t *out) {
    size_t k = sdao_input_size(s), n = sdao_output_size(s);
    if (n == 0 || k == 0) return fail("oracle: degenerate scheme");
    size_t B = (dim + k - 1) / k;                                   /* :23 */
    for (size_t b = 0; b < B; b++) {
        /* generate_for_batch returns a fresh Vec per batch (:35,:41); kept as a malloc */
        int64_t *shares = malloc(sizeof(int64_t) * n);
        int64_t *padded = NULL;
        const int64_t *batch;
        if ((b + 1) * k <= dim) {
            batch = secrets + b * k;                                /* :33-35 */
        } else {
            padded = calloc(k, sizeof(int64_t));                    /* :38-41 zero padding */
            memcpy(padded, secrets + b * k, (dim - b * k) * sizeof(int64_t));
            batch = padded;
        }
        int rc = s->kind == SDAO_SHARING_ADDITIVE
                     ? additive_generate_for_batch(s, batch, k, rng, shares)
                     : packed_generate_for_batch(s, batch, k, rng, shares);
        if (!rc)
            for (size_t r = 0; r < n; r++) out[r * B + b] = shares[r]; /* :46-48 */
        free(shares); free(padded);
        if (rc) return rc;
    }
    return 0;
}

/* Same linear map through a precomputed matrix (columns = shares of unit vectors, built
 * with the Newton/FFT machinery above) -- the fair CPU baseline.  Canonical outputs. */
int sdao_share_generate_matrix(const sdao_sharing_scheme *s, const int64_t *secrets, size_t dim,
                               sdao_rng *rng, int64_t *out) {
    if (s->kind == SDAO_SHARING_ADDITIVE) return sdao_share_generate(s, secrets, dim, rng, out);
    size_t k = s->secret_count, t = s->privacy_threshold, n = s->share_count, w = k + t;
    int64_t p = s->modulus;
    if (w > 64) return fail("oracle: k+t > 64 unsupported");
    size_t B = (dim + k - 1) / k;
    uint64_t *M = malloc(sizeof(uint64_t) * n * w);
    int64_t unit[64], col[256];
    if (n > 256) { free(M); return fail("oracle: n > 256 unsupported"); }
    for (size_t i = 0; i < w; i++) {
        memset(unit, 0, sizeof unit);
        unit[i] = 1;
        sdao_tss_share_with_randomness(s, unit, unit + k, 0, NULL, col);
        for (size_t j = 0; j < n; j++) M[j * w + i] = (uint64_t)(((col[j] % p) + p) % p);
    }
    uint64_t x[64];
    /* w * p^2 < 2^128  =>  one reduction per share instead of one per term */
    int lazy = ((uint64_t)p >> 61) == 0 && w <= 32;
    for (size_t b = 0; b < B; b++) {
        for (size_t i = 0; i < k; i++) {
            int64_t v = (b * k + i < dim) ? secrets[b * k + i] : 0;
            x[i] = (uint64_t)(((v % p) + p) % p);
        }
        for (size_t i = 0; i < t; i++) x[k + i] = (uint64_t)sdao_gen_range(rng, 0, p - 1);
        for (size_t j = 0; j < n; j++) {
            u128 acc = 0;
            if (lazy) for (size_t i = 0; i < w; i++) acc += (u128)M[j * w + i] * x[i];
            else      for (size_t i = 0; i < w; i++) acc += (u128)M[j * w + i] * x[i] % (uint64_t)p;
            out[j * B + b] = (int64_t)(acc % (uint64_t)p);
        }
    }
    free(M);
    return 0;
}

/* combiner.rs:15-29  Combiner::combine  (also full.rs:37-52 and additive.rs:55-73 bodies) */
int sdao_share_combine(int64_t modulus, const int64_t *shares, size_t P, size_t L, size_t ld, int64_t *out) {
    for (size_t i = 0; i < L; i++) out[i] = 0;                      /* :19 */
    for (size_t p = 0; p < P; p++) {
        const int64_t *row = shares + p * ld;
        for (size_t i = 0; i < L; i++) {
            out[i] += row[i];                                       /* :23 */
            out[i] %= modulus;                                      /* :24 */
        }
    }
    return 0;
}

/* additive.rs:55-73 | batched.rs:68-97 + packed_shamir.rs:73-77 */
int sdao_secret_reconstruct(const sdao_sharing_scheme *s, size_t dimension, const uint64_t *indices,
                            const int64_t *shares, size_t m, size_t B, int64_t *out, size_t *out_len) {
    if (s->kind == SDAO_SHARING_ADDITIVE) {
        /* dimension = length of the first share vector; indices ignored (additive.rs:56-59) */
        size_t len = m ? B : 0;
        sdao_share_combine(s->modulus, shares, m, len, B, out);
        if (out_len) *out_len = len;
        return 0;
    }
    size_t k = s->secret_count;
    size_t nb = (dimension + k - 1) / k;                            /* batched.rs:77 */
    if (nb > 0 && B < nb) return fail("oracle: share vector shorter than batch count (reference panics, batched.rs:84)");
    int64_t *batch = malloc(sizeof(int64_t) * (m + k));
    int64_t *sec = batch + m;
    size_t w = 0;
    for (size_t b = 0; b < nb; b++) {
        for (size_t si = 0; si < m; si++) batch[si] = shares[si * B + b]; /* :83-85 */
        /* packed_shamir.rs:74-75 */
        if (m < sdao_reconstruction_threshold(s)) { free(batch); return fail("Not enough shares to reconstruct"); }
        sdao_tss_reconstruct(s, indices, batch, m, sec);
        for (size_t e = 0; e < k; e++)
            if (w < dimension) out[w++] = sec[e];                   /* :88-90, truncate :94 */
    }
    free(batch);
    if (out_len) *out_len = dimension;
    return 0;
}

/* ======================================================================================
 * client/src/crypto/masking
 * ====================================================================================== */

int sdao_mask(const sdao_masking_scheme *s, const int64_t *secrets, size_t dim, sdao_rng *rng,
              int64_t *mask_out, size_t *mask_len, int64_t *masked_out) {
    switch (s->kind) {
    case SDAO_MASK_NONE:                                            /* none.rs:14-18 */
        *mask_len = 0;
        memcpy(masked_out, secrets, dim * sizeof(int64_t));
        return 0;
    case SDAO_MASK_FULL:                                            /* full.rs:22-34 */
        for (size_t i = 0; i < dim; i++) mask_out[i] = sdao_gen_range(rng, 0, s->modulus);
        for (size_t i = 0; i < dim; i++) masked_out[i] = (secrets[i] + mask_out[i]) % s->modulus;
        *mask_len = dim;
        return 0;
    case SDAO_MASK_CHACHA: {                                        /* chacha.rs:25-53 */
        if (s->dimension != dim) return fail("assertion failed: `(left == right)` (chacha.rs:26)");
        size_t words = (s->seed_bitsize + 31) / 32;                 /* :30 */
        uint32_t seed[64];
        if (words > 64) return fail("oracle: seed too long");
        for (size_t i = 0; i < words; i++) seed[i] = sdao_rng_next_u32(rng); /* :31-33 (OsRng) */
        sdao_rng gen;
        sdao_rng_from_seed(&gen, seed, words);                      /* :36 */
        for (size_t i = 0; i < dim; i++) {
            int64_t mk = sdao_gen_range(&gen, 0, s->modulus);       /* :37-39 */
            masked_out[i] = (secrets[i] + mk) % s->modulus;         /* :42-45 */
        }
        for (size_t i = 0; i < words; i++) mask_out[i] = (int64_t)seed[i]; /* :48-50 */
        *mask_len = words;
        return 0;
    }
    }
    return fail("oracle: unknown masking scheme");
}

int sdao_mask_combine(const sdao_masking_scheme *s, const int64_t *masks, size_t P, size_t mask_len,
                      int64_t *out, size_t *out_len) {
    switch (s->kind) {
    case SDAO_MASK_NONE:                                            /* none.rs:22-25 */
        if (mask_len != 0) return fail("assertion failed: masks.iter().all(|mask| mask.len() == 0)");
        *out_len = 0;
        return 0;
    case SDAO_MASK_FULL:                                            /* full.rs:38-51 */
        *out_len = P ? mask_len : 0;
        return sdao_share_combine(s->modulus, masks, P, *out_len, mask_len, out);
    case SDAO_MASK_CHACHA: {                                        /* chacha.rs:57-76 */
        size_t dim = s->dimension;
        for (size_t i = 0; i < dim; i++) out[i] = 0;
        for (size_t p = 0; p < P; p++) {
            uint32_t seed[64];
            size_t words = mask_len < 64 ? mask_len : 64;
            for (size_t i = 0; i < words; i++) seed[i] = (uint32_t)masks[p * mask_len + i]; /* :62-64 */
            sdao_rng gen;
            sdao_rng_from_seed(&gen, seed, words);
            for (size_t i = 0; i < dim; i++) {
                int64_t mk = sdao_gen_range(&gen, 0, s->modulus);
                out[i] += mk;
                out[i] %= s->modulus;
            }
        }
        *out_len = dim;
        return 0;
    }
    }
    return fail("oracle: unknown masking scheme");
}

int sdao_unmask(const sdao_masking_scheme *s, const int64_t *mask, size_t mask_len,
                const int64_t *masked, size_t dim, int64_t *out) {
    (void)mask;
    if (s->kind == SDAO_MASK_NONE) {                                /* none.rs:29-32 */
        if (mask_len != 0) return fail("assertion failed: `(left == right)` (none.rs:30)");
        memcpy(out, masked, dim * sizeof(int64_t));
        return 0;
    }
    if (mask_len != dim) return fail("assertion failed: `(left == right)` (full.rs:58 / chacha.rs:83)");
    for (size_t i = 0; i < dim; i++) out[i] = (masked[i] - mask[i]) % s->modulus; /* full.rs:60-63 */
    return 0;
}

/* client/src/receive.rs:13-21 */
void sdao_positive(int64_t m, int64_t *v, size_t n) {
    for (size_t i = 0; i < n; i++) if (v[i] < 0) v[i] += m;
}
void sdao_canonical(int64_t m, int64_t *v, size_t n) {
    /* receive.rs:13-21: `if v < 0 { v + m }` on the truncated remainder (no (r + m) % m: that
     * overflows i64 for m > 2^62) */
    for (size_t i = 0; i < n; i++) { int64_t r = v[i] % m; v[i] = r < 0 ? r + m : r; }
}

/* ======================================================================================
 * integer-encoding 1.0  VarInt for i64: zig-zag then LEB128
 *   call sites: client/src/crypto/encryption/sodium.rs:36-41 (encode), :83-90 (decode)
 * ====================================================================================== */
size_t sdao_varint_encode(const int64_t *values, size_t n, uint8_t *out) {
    size_t w = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t z = ((uint64_t)values[i] << 1) ^ (uint64_t)(values[i] >> 63);
        while (z >= 0x80) { out[w++] = (uint8_t)(z | 0x80); z >>= 7; }
        out[w++] = (uint8_t)z;
    }
    return w;
}
size_t sdao_varint_decode(const uint8_t *buf, size_t len, int64_t *out, size_t max_out) {
    size_t r = 0, cnt = 0;
    while (r < len) {
        uint64_t z = 0;
        int shift = 0, done = 0;
        while (r < len && shift <= 63) {
            uint8_t b = buf[r++];
            z |= (uint64_t)(b & 0x7f) << shift;
            shift += 7;
            if (!(b & 0x80)) { done = 1; break; }
        }
        if (!done) return (size_t)-1;
        if (cnt >= max_out) return (size_t)-1;
        out[cnt++] = (int64_t)(z >> 1) ^ -(int64_t)(z & 1);
    }
    return cnt;
}

/* ======================================================================================
 * fixed-point codec for real-valued vectors.  NOT in the
 * reference, whose API takes Vec<i64> (client/src/participate.rs:10,25; motive README.md:12-13):
 * an adjacent addition with this definition, shared with the CUDA side.
 *   encode: q = rint(x * 2^frac_bits) (ties to even, exact in double), residue q mod p in [0, p)
 *   decode: centred lift c in (-p/2, p/2], value (double)c / 2^frac_bits / divisor, rounded to float
 * ====================================================================================== */
#include <math.h>
void sdao_fixed_encode(const float *x, size_t n, int frac_bits, int64_t p, int64_t *out) {
    const double scale = ldexp(1.0, frac_bits);
    for (size_t i = 0; i < n; i++) {
        int64_t q = llrint((double)x[i] * scale);
        int64_t r = q % p;
        out[i] = r < 0 ? r + p : r;
    }
}
void sdao_fixed_decode(const int64_t *in, size_t n, int frac_bits, int64_t p, uint64_t divisor, float *out) {
    const double scale = ldexp(1.0, frac_bits);
    for (size_t i = 0; i < n; i++) {
        int64_t r = in[i] % p;
        if (r < 0) r += p;
        int64_t c = r > p / 2 ? r - p : r;
        out[i] = (float)((double)c / scale / (double)divisor);
    }
}

/* ======================================================================================
 * synthetic inputs (definition shared with the CUDA side; not from the reference)
 * ====================================================================================== */
void sdao_synth_fill(uint32_t stream, int64_t modulus, uint64_t start, size_t count, int64_t *out) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    const char *tag = "sda-b200-synthetic-v1";
    uint8_t key[32] = {0};
    memcpy(key, tag, strlen(tag));
    for (int i = 0; i < 8; i++)
        st[4 + i] = (uint32_t)key[4 * i] | (uint32_t)key[4 * i + 1] << 8 | (uint32_t)key[4 * i + 2] << 16 |
                    (uint32_t)key[4 * i + 3] << 24;
    st[11] = stream;
    uint32_t blk[16];
    uint64_t cur = UINT64_MAX;
    for (size_t i = 0; i < count; i++) {
        uint64_t e = start + i, b = e / 8;
        if (b != cur) {
            st[12] = (uint32_t)b; st[13] = (uint32_t)(b >> 32); st[14] = st[15] = 0;
            sdao_chacha_block(st, 20, blk);
            cur = b;
        }
        unsigned w = (unsigned)(e % 8) * 2;
        uint64_t v = ((uint64_t)blk[w] << 32) | blk[w + 1];
        out[i] = (int64_t)(v % (uint64_t)modulus);
    }
}

/* ======================================================================================
 * parameter search (not from the reference: its parameters are hard-coded in tests)
 * ====================================================================================== */
int64_t sdao_find_root_of_order(int64_t p, uint64_t q) {
    if (q == 0 || (uint64_t)(p - 1) % q) return 0;
    uint64_t fac[16]; int nf = 0;
    uint64_t r = q;
    for (uint64_t f = 2; f * f <= r; f++)
        if (r % f == 0) { fac[nf++] = f; while (r % f == 0) r /= f; }
    if (r > 1) fac[nf++] = r;
    for (int64_t c = 2; c < p; c++) {
        int64_t w = sdao_mod_pow(c, (uint64_t)(p - 1) / q, p);
        w = ((w % p) + p) % p;
        if (w == 1 && q != 1) continue;
        int ok = ((sdao_mod_pow(w, q, p) % p) + p) % p == 1;
        for (int i = 0; ok && i < nf; i++)
            if (((sdao_mod_pow(w, q / fac[i], p) % p) + p) % p == 1) ok = 0;
        if (ok) return w;
    }
    return 0;
}
