/*
 * ORACLE build shim (test infrastructure): the reference needs libsodium for exactly one call,
 * randombytes_buf (src/groth16.cpp:216-217).  libsodium is not in this image; the kernel CSPRNG
 * (getrandom(2)) supplies the same contract.  ORACLE_FIXED_RS=<hex seed> makes r,s deterministic
 * so the reference's final proof can be compared byte-for-byte in tests.
 */
#ifndef ORACLE_SHIM_SODIUM_H
#define ORACLE_SHIM_SODIUM_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <sys/random.h>

static inline void randombytes_buf(void *const buf, const size_t size)
{
    static uint64_t ctr = 0;
    const char *fixed = getenv("ORACLE_FIXED_RS");
    uint8_t *p = (uint8_t *)buf;
    if (fixed) {
        /* splitmix64 stream keyed by the env value; call order r then s as in groth16.cpp */
        uint64_t x = strtoull(fixed, NULL, 16) + 0x9E3779B97F4A7C15ULL * (++ctr);
        for (size_t i = 0; i < size; i++) {
            if ((i & 7) == 0) {
                x += 0x9E3779B97F4A7C15ULL;
                uint64_t z = x;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
                z = z ^ (z >> 31);
                for (size_t k = 0; k < 8 && i + k < size; k++) p[i + k] = (uint8_t)(z >> (8 * k));
            }
        }
        return;
    }
    size_t off = 0;
    while (off < size) {
        ssize_t n = getrandom(p + off, size - off, 0);
        if (n > 0) off += (size_t)n;
    }
}
#endif
