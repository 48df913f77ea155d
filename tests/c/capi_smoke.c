/* tests/c/capi_smoke.c -- the C ABI from plain C99: the header must compile without C++, every entry
 * point must link, and without a GPU gsn_ctx_create must fail with GSN_ERR_NO_DEVICE (no CPU fallback).
 * With a GPU it runs one 8-point 32-bit transform and checks it against the DFT definition. */
#include <stdio.h>
#include <string.h>

#include "gpusnarks_b200.h"

int main(void) {
    gsn_ctx *ctx = NULL;
    int cnt = -1;
    gsn_device_count(&cnt);
    int rc = gsn_ctx_create(&ctx, 0);
    if (cnt <= 0) {
        if (rc != GSN_ERR_NO_DEVICE || ctx != NULL) { printf("expected GSN_ERR_NO_DEVICE, got %d\n", rc); return 1; }
        if (!strstr(gsn_last_error(), "no CPU fallback")) { printf("unexpected message: %s\n", gsn_last_error()); return 1; }
        printf("capi ok (no device): %s\n", gsn_last_error());
        return 0;
    }
    if (rc != GSN_OK) { printf("gsn_ctx_create: %s\n", gsn_last_error()); return 1; }
    {
        const uint32_t p = 97, w = 64; /* 64 has order 8 modulo 97 */
        uint32_t a[8] = {1, 2, 3, 4, 5, 6, 7, 8}, ref[8];
        for (int i = 0; i < 8; ++i) {
            unsigned long long acc = 0, wi = 1;
            for (int k = 0; k < i; ++k) wi = wi * w % p;
            unsigned long long x = 1;
            for (int j = 0; j < 8; ++j) { acc = (acc + a[j] * x) % p; x = x * wi % p; }
            ref[i] = (uint32_t)acc;
        }
        rc = gsn_ntt32_host(ctx, a, 8, w, p, 0);
        if (rc != GSN_OK) { printf("gsn_ntt32_host: %s\n", gsn_last_error()); return 1; }
        if (memcmp(a, ref, sizeof(a)) != 0) { printf("mismatch\n"); return 1; }
        if (gsn_ntt32_host(ctx, a, 12, w, p, 0) != GSN_ERR_NOT_POW2) { printf("expected NOT_POW2\n"); return 1; }
    }
    gsn_ctx_destroy(ctx);
    printf("capi ok (device)\n");
    return 0;
}
