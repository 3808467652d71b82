// randombytes_buf for callers written against libsodium (src/groth16.cpp:213-217 draws the blinding factors with it):
// the kernel's CSPRNG.  B200_R / B200_S (64 hex digits, big-endian, test only) fix the two draws of a proof.
#ifndef B200_ENGINE_SODIUM_H
#define B200_ENGINE_SODIUM_H
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>
static inline void randombytes_buf(void *const buf, const size_t size) {
    static int draw = 0;
    const char *fixed = getenv((draw++ & 1) ? "B200_S" : "B200_R");
    if (fixed && strlen(fixed) <= 64) {
        unsigned char *o = (unsigned char *)buf;
        memset(o, 0, size);
        size_t n = strlen(fixed);
        for (size_t i = 0; i < n && i / 2 < size; i++) {
            char ch = fixed[n - 1 - i];
            int v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : (ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : 0;
            o[i / 2] |= (unsigned char)(v << (4 * (i & 1)));
        }
        return;
    }
    size_t done = 0;
    while (done < size) {
        ssize_t k = getrandom((char *)buf + done, size - done, 0);
        if (k <= 0) abort();
        done += (size_t)k;
    }
}
#endif
