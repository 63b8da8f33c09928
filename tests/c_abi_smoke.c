/* c_abi_smoke.c -- the drop-in boundary from plain C, no Python: the reference's own end-to-end vectors
 * (integration-tests/tests/full_loop.rs:11-27,113,148: two participants holding [1,2,3,4] aggregate to [2,4,6,8];
 * README.md:157: three participants, additive mod 433) driven through include/sda_b200.h exactly as a Rust shim
 * would: mask -> share -> per-clerk combine -> reconstruct -> mask combine -> unmask.
 *
 *   gcc -std=c99 -Wall -Iinclude tests/c_abi_smoke.c -o c_abi_smoke -Lsda_b200 -lsda_b200 -Wl,-rpath,$PWD/sda_b200
 *
 * exit 0: all vectors reproduced; 77: no CUDA device (the library has no CPU fallback); 1: mismatch or error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sda_b200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != SDA_OK) {                                                                     \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, sda_last_error(ctx));                  \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

static void seed_for(uint8_t seed[32], int tag, int who) {
    for (int i = 0; i < 32; i++) seed[i] = (uint8_t)(17 * tag + 31 * who + i);
}

/* P participants, each holding `dim` secrets; returns 0 when the revealed aggregate equals `expect` */
static int full_loop(sda_ctx *ctx, const char *name, const sda_sharing_scheme *ss, const sda_masking_scheme *ms,
                     const int64_t *secrets /* [P][dim] */, size_t P, size_t dim, const int64_t *expect) {
    const size_t n = sda_output_size(ss), B = sda_share_batches(ss, dim), ml = sda_mask_len(ms, dim);
    int64_t *shares = calloc(P * n * B + 1, sizeof(int64_t));     /* [P][n][B] */
    int64_t *masks = calloc(P * ml + 1, sizeof(int64_t));         /* [P][ml] */
    int64_t *masked = calloc(dim + 1, sizeof(int64_t));
    int64_t *clerk = calloc(n * B + 1, sizeof(int64_t));          /* [n][B] */
    int64_t *revealed = calloc(dim + B + 1, sizeof(int64_t)), *mask_sum = calloc(dim + ml + 1, sizeof(int64_t));
    int64_t *out = calloc(dim + 1, sizeof(int64_t));
    uint8_t seed[32];
    for (size_t p = 0; p < P; p++) {
        size_t got_ml = 0;
        seed_for(seed, 1, (int)p);
        CHECK(sda_mask(ctx, ms, secrets + p * dim, dim, seed, masks + p * ml, &got_ml, masked));       /* participate.rs:53-54 */
        if (got_ml != ml) return fprintf(stderr, "%s: mask length %zu != %zu\n", name, got_ml, ml), 1;
        seed_for(seed, 2, (int)p);
        CHECK(sda_share_generate(ctx, ss, masked, dim, seed, shares + p * n * B));                    /* participate.rs:75-76 */
        /* the same two steps through the one-call fast path: identical mask and shares */
        int64_t *mask2 = calloc(ml + 1, sizeof(int64_t)), *shares2 = calloc(n * B + 1, sizeof(int64_t));
        uint8_t mseed[32];
        seed_for(mseed, 1, (int)p);
        CHECK(sda_mask_share_generate(ctx, ms, ss, secrets + p * dim, dim, mseed, seed, mask2, shares2));
        int differ = memcmp(mask2, masks + p * ml, ml * sizeof(int64_t)) != 0 ||
                     memcmp(shares2, shares + p * n * B, n * B * sizeof(int64_t)) != 0;
        free(mask2);
        free(shares2);
        if (differ) return fprintf(stderr, "%s: sda_mask_share_generate differs from mask + share_generate\n", name), 1;
    }
    for (size_t c = 0; c < n; c++) {                                                                  /* clerk.rs:85-86 */
        const int64_t **rows = malloc(P * sizeof *rows);
        size_t *lens = malloc(P * sizeof *lens), len = 0;
        for (size_t p = 0; p < P; p++) {
            rows[p] = shares + (p * n + c) * B;
            lens[p] = B;
        }
        CHECK(sda_share_combine_rows(ctx, ss, rows, lens, P, clerk + c * B, &len));
        free(rows);
        free(lens);
        if (len != B) return fprintf(stderr, "%s: combined length %zu != %zu\n", name, len, B), 1;
    }
    uint64_t *indices = malloc(n * sizeof *indices);
    for (size_t c = 0; c < n; c++) indices[c] = c;
    size_t rlen = 0, mlen = 0;
    CHECK(sda_secret_reconstruct(ctx, ss, dim, indices, clerk, n, B, revealed, &rlen));               /* receive.rs:113-116 */
    CHECK(sda_mask_combine(ctx, ms, masks, P, ml, mask_sum, &mlen));                                  /* receive.rs:140-144 */
    CHECK(sda_unmask(ctx, ms, mask_sum, mlen, revealed, dim, out));                                   /* receive.rs:149-152 */
    int bad = rlen != dim;
    for (size_t i = 0; i < dim && !bad; i++) bad = out[i] != expect[i];
    printf("%-44s %s\n", name, bad ? "MISMATCH" : "ok");
    free(shares); free(masks); free(masked); free(clerk); free(revealed); free(mask_sum); free(out); free(indices);
    return bad;
}

int main(void) {
    sda_ctx *ctx = NULL;
    if (sda_abi_version() != SDA_B200_ABI_VERSION) return fprintf(stderr, "ABI version mismatch\n"), 1;
    int rc = sda_ctx_create(0, &ctx);
    if (rc == SDA_ERR_CUDA) {
        printf("no CUDA device: %s\n", sda_last_error(NULL));
        return 77;
    }
    if (rc != SDA_OK) return fprintf(stderr, "sda_ctx_create -> %d: %s\n", rc, sda_last_error(NULL)), 1;

    const sda_sharing_scheme additive = {SDA_SHARING_ADDITIVE, 3, 0, 0, 433, 0, 0};
    const sda_sharing_scheme packed = {SDA_SHARING_PACKED_SHAMIR, 8, 3, 4, 433, 354, 150};            /* full_loop.rs:57-64 */
    const sda_masking_scheme none = {SDA_MASK_NONE, 0, 0, 0}, full = {SDA_MASK_FULL, 433, 0, 0};
    const sda_masking_scheme chacha = {SDA_MASK_CHACHA, 433, 4, 128};                                 /* full_loop.rs:42-52 */
    const int64_t two[8] = {1, 2, 3, 4, 1, 2, 3, 4}, two_sum[4] = {2, 4, 6, 8};
    int bad = 0;
    bad |= full_loop(ctx, "full_loop.rs additive(3), no mask", &additive, &none, two, 2, 4, two_sum);
    bad |= full_loop(ctx, "full_loop.rs additive(3) + Full{433}", &additive, &full, two, 2, 4, two_sum);
    bad |= full_loop(ctx, "full_loop.rs additive(3) + ChaCha{433,4,128}", &additive, &chacha, two, 2, 4, two_sum);
    bad |= full_loop(ctx, "full_loop.rs PackedShamir{3,8,4,433,354,150}", &packed, &none, two, 2, 4, two_sum);
    /* README.md:86,105-107,157 / docs/simple-cli-example.sh: 0..9, all zero, 0 1 0 1 ... -> 0 2 2 4 4 6 6 8 8 10 */
    int64_t cli[30], cli_sum[10] = {0, 2, 2, 4, 4, 6, 6, 8, 8, 10};
    for (int i = 0; i < 10; i++) {
        cli[i] = i;
        cli[10 + i] = 0;
        cli[20 + i] = i & 1;
    }
    bad |= full_loop(ctx, "README.md walkthrough additive(3), dim 10", &additive, &none, cli, 3, 10, cli_sum);
    /* error strings of the reference come back verbatim (combiner.rs:21) */
    {
        const int64_t r0[3] = {1, 2, 3}, r1[2] = {1, 2};
        const int64_t *rows[2] = {r0, r1};
        const size_t lens[2] = {3, 2};
        int64_t out[3];
        size_t len = 0;
        const int e = sda_share_combine_rows(ctx, &additive, rows, lens, 2, out, &len);
        const int ok = e == SDA_ERR_INVALID && strcmp(sda_last_error(ctx), "Wrong dimension") == 0;
        printf("%-44s %s\n", "combiner.rs:21 \"Wrong dimension\"", ok ? "ok" : "MISMATCH");
        bad |= !ok;
    }
    printf("kernels launched: %llu\n", (unsigned long long)sda_ctx_launch_count(ctx));
    bad |= sda_ctx_launch_count(ctx) == 0;
    sda_ctx_destroy(ctx);
    return bad ? 1 : 0;
}
