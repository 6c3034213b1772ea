// linedigest.c -- order-independent digest of a k-mer count file (TEST INFRASTRUCTURE).
//
// The reference checks two outputs by sorting both and comparing line by line (pytools/compare_outputs.py:1-33, which
// also forgets to compare the line counts).  At full size (10^8 lines, 5-26 GB of text) a sort costs minutes, so this
// tool reads "KMER COUNT\n" lines from stdin -- typically a FIFO the counter writes to with -o -- and prints
//   {"lines": n, "sum": Σ h(line) mod 2^64, "xor": ⊕ h(line), "count_sum": Σ COUNT, "bytes": b, "bad": malformed lines}
// h = a 64-bit hash of the line's bytes (multiply-xorshift over 8-byte words, murmur3 finaliser).  Two files have the
// same digest iff (up to 2^-64-ish collisions) they hold the same multiset of lines: the same canonical k-mers with
// the same counts, whatever the order.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t fmix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static inline uint64_t hash_line(const unsigned char* p, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ (uint64_t)n;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        h = (h ^ w) * 0x9FB21C651E98DF25ULL;
        h ^= h >> 29;
        p += 8; n -= 8;
    }
    uint64_t w = 0;
    memcpy(&w, p, n);
    h = (h ^ w) * 0xD6E8FEB86659FD93ULL;
    return fmix64(h);
}

int main(void) {
    const size_t CAP = 64u << 20;
    unsigned char* buf = (unsigned char*)malloc(CAP);
    size_t have = 0;
    uint64_t lines = 0, sum = 0, x = 0, count_sum = 0, bytes = 0, bad = 0;
    for (;;) {
        const size_t got = fread(buf + have, 1, CAP - have, stdin);
        const size_t end = have + got;
        bytes += got;
        size_t start = 0;
        for (;;) {
            unsigned char* nl = (unsigned char*)memchr(buf + start, '\n', end - start);
            if (!nl) break;
            const size_t len = (size_t)(nl - (buf + start));
            const uint64_t h = hash_line(buf + start, len);
            lines++; sum += h; x ^= h;
            // COUNT = the decimal digits after the last space
            size_t sp = len;
            while (sp > 0 && buf[start + sp - 1] != ' ') sp--;
            if (sp == 0 || sp == len) bad++;
            else {
                uint64_t c = 0;
                for (size_t i = sp; i < len; i++) {
                    const unsigned d = (unsigned)buf[start + i] - '0';
                    if (d > 9) { bad++; c = 0; break; }
                    c = c * 10 + d;
                }
                count_sum += c;
            }
            start += len + 1;
        }
        have = end - start;
        if (have == CAP) { fprintf(stderr, "linedigest: line longer than %zu bytes\n", CAP); return 2; }
        memmove(buf, buf + start, have);
        if (got == 0) break;
    }
    if (have) bad++;   // trailing bytes without a newline
    printf("{\"lines\": %llu, \"sum\": %llu, \"xor\": %llu, \"count_sum\": %llu, \"bytes\": %llu, \"bad\": %llu}\n",
           (unsigned long long)lines, (unsigned long long)sum, (unsigned long long)x, (unsigned long long)count_sum,
           (unsigned long long)bytes, (unsigned long long)bad);
    return 0;
}
