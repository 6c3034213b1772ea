/* TEST INFRASTRUCTURE. Prints XXH64(le64(value), seed) using the reference's vendored xxHash
 * (external/xxHash/xxhash.c, v0.8.2) exactly as double_bloomfilter.hpp:276-281 calls it.
 * usage: xxh64_kat <value-hex> <seed-dec> ...   (pairs) */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "xxhash.h"
int main(int argc, char** argv) {
    for (int i = 1; i + 1 < argc; i += 2) {
        uint64_t v = strtoull(argv[i], NULL, 16);
        uint64_t s = strtoull(argv[i + 1], NULL, 10);
        printf("%016llx %llu %016llx\n", (unsigned long long)v, (unsigned long long)s,
               (unsigned long long)XXH64(&v, sizeof(v), s));
    }
    return 0;
}
