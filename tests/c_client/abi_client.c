/* A plain C client of the drop-in boundary (include/pfc.h -> libpfc_b200.so): what a non-Python binding (cgo, JNI, a C
 * trainer) would compile against.  Runs without a GPU: it calls the entry points that need none --
 *   pfc_version / pfc_error_string / the shape helpers, and
 *   pfc_host_mt19937_uniform on a generator state built with the PUBLISHED MT19937 seeding (Matsumoto & Nishimura,
 *   init_genrand: s[0] = seed, s[j] = 1812433253 * (s[j-1] ^ (s[j-1] >> 30)) + j), checked against the published
 *   known answers: first output of seed 5489 is 3499211612 and the 10000th is 4123659995 (ISO C++ [rand.predef]).
 * The library turns outputs into float32 as (word & (2^24 - 1)) * 2^-24, so the low 24 bits are what is compared.
 * Built and run by tests/test_library_abi.py::test_plain_c_client. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pfc.h"

static void seed_blob(uint8_t* blob, size_t bytes, uint32_t seed) {
    uint64_t s[624], seed64 = seed, next = 0;
    int32_t left = 1, seeded = 1;
    int j;
    memset(blob, 0, bytes);
    s[0] = seed;
    for (j = 1; j < 624; ++j) s[j] = (uint32_t)(1812433253u * ((uint32_t)s[j - 1] ^ ((uint32_t)s[j - 1] >> 30)) + (uint32_t)j);
    memcpy(blob, &seed64, 8);
    memcpy(blob + 8, &left, 4);
    memcpy(blob + 12, &seeded, 4);
    memcpy(blob + 16, &next, 8);
    memcpy(blob + 24, s, sizeof s);
}

int main(void) {
    const size_t bytes = pfc_host_mt19937_state_bytes();
    uint8_t* blob = (uint8_t*)malloc(bytes);
    float* out = (float*)malloc(10000 * sizeof(float));
    int rc, fails = 0;
    if (pfc_version() < 200) { printf("FAIL version %d\n", pfc_version()); return 1; }
    if (strlen(pfc_error_string(PFC_ERR_SHAPE)) == 0) { printf("FAIL error string\n"); return 1; }
    if (pfc_padded_classes(93431) != 93440 || pfc_padded_batch(1000) != 1024) { printf("FAIL shape helpers\n"); return 1; }
    seed_blob(blob, bytes, 5489u);
    rc = pfc_host_mt19937_uniform(blob, bytes, out, 10000);
    if (rc != PFC_OK) { printf("FAIL rc %d\n", rc); return 1; }
    if (out[0] != (float)(3499211612u & 0xffffffu) / 16777216.0f) { printf("FAIL first output %.9g\n", out[0]); ++fails; }
    if (out[9999] != (float)(4123659995u & 0xffffffu) / 16777216.0f) { printf("FAIL 10000th output %.9g\n", out[9999]); ++fails; }
    /* the state continues: 5000 + 5000 draws equal 10000 draws */
    {
        float* two = (float*)malloc(10000 * sizeof(float));
        seed_blob(blob, bytes, 5489u);
        rc = pfc_host_mt19937_uniform(blob, bytes, two, 5000);
        rc |= pfc_host_mt19937_uniform(blob, bytes, two + 5000, 5000);
        if (rc != PFC_OK || memcmp(two, out, 10000 * sizeof(float)) != 0) { printf("FAIL split draw\n"); ++fails; }
        free(two);
    }
    /* a blob that is not a seeded generator state is refused */
    memset(blob, 0, bytes);
    if (pfc_host_mt19937_uniform(blob, bytes, out, 4) == PFC_OK) { printf("FAIL unseeded blob accepted\n"); ++fails; }
    free(blob);
    free(out);
    if (fails) return 1;
    printf("abi_client: ok (libpfc_b200 version %d)\n", pfc_version());
    return 0;
}
