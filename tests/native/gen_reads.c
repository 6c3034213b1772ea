// gen_reads.c -- deterministic synthetic inputs for the full-size parity cases (TEST INFRASTRUCTURE).
//
// The same binary, built from this file with gcc, produces byte-identical files in the build container (where the
// unmodified reference binary mints the golden digests, tests/golden/make_fullsize_digests.py) and on the GPU box
// (where tests/test_gpu_fullsize_reference.py feeds them to the CUDA path).  Shapes follow SURVEY.md section 8d:
//   reads   G n_reads L err seed wrap   uniform random genome; read start uniform in [0, G-L]; every base substituted
//                                       with probability err by a uniformly different base; every read reverse-
//                                       complemented with probability 1/2; FASTA, header ">r<i>", sequence wrapped
//                                       at `wrap` columns (0 = one line)
//   genome  G seed wrap n_repeats       one record holding the genome itself with n_repeats 1 kbp segments copied
//                                       elsewhere (the declared stand-in for the missing example/ecoli1x.fasta)
// RNG: xoshiro256** seeded through splitmix64 (public-domain algorithms by Blackman & Vigna), no libc rand().
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t next(void) {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
}
static void seed_rng(uint64_t x) {
    for (int i = 0; i < 4; i++) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        s[i] = z ^ (z >> 31);
    }
}
static inline uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }

static uint8_t* make_genome(uint64_t G) {
    uint8_t* g = (uint8_t*)malloc(G + 32);
    if (!g) { fprintf(stderr, "gen_reads: out of memory\n"); exit(2); }
    for (uint64_t i = 0; i < G; i += 32) {
        uint64_t r = next();
        for (int j = 0; j < 32; j++) g[i + j] = (uint8_t)((r >> (2 * j)) & 3);
    }
    return g;
}

#define OBUF (16u << 20)
static char* obuf;
static size_t opos;
static FILE* out;
static inline void flush_out(void) { if (opos) { if (fwrite(obuf, 1, opos, out) != opos) { perror("gen_reads: write"); exit(2); } opos = 0; } }
static inline void put(const char* p, size_t n) {
    if (opos + n > OBUF) flush_out();
    memcpy(obuf + opos, p, n);
    opos += n;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: gen_reads OUT reads G n_reads L err seed wrap\n       gen_reads OUT genome G seed wrap n_repeats\n");
        return 1;
    }
    out = strcmp(argv[1], "-") ? fopen(argv[1], "wb") : stdout;
    if (!out) { perror(argv[1]); return 2; }
    obuf = (char*)malloc(OBUF);
    static const char ASCII[4] = {'A', 'C', 'G', 'T'};
    if (!strcmp(argv[2], "genome") && argc == 7) {
        const uint64_t G = strtoull(argv[3], 0, 10), seed = strtoull(argv[4], 0, 10);
        const uint64_t wrap = strtoull(argv[5], 0, 10), reps = strtoull(argv[6], 0, 10);
        seed_rng(seed);
        uint8_t* g = make_genome(G);
        for (uint64_t r = 0; r < reps; r++) {
            const uint64_t a = below(G - 1000), b = below(G - 1000);
            memmove(g + b, g + a, 1000);
        }
        char hdr[128];
        int n = snprintf(hdr, sizeof hdr, ">ecoli1x_standin seed=%llu G=%llu\n", (unsigned long long)seed, (unsigned long long)G);
        put(hdr, (size_t)n);
        char line[4096];
        for (uint64_t i = 0; i < G; i += wrap) {
            const uint64_t m = G - i < wrap ? G - i : wrap;
            for (uint64_t j = 0; j < m; j++) line[j] = ASCII[g[i + j]];
            line[m] = '\n';
            put(line, (size_t)m + 1);
        }
    } else if (!strcmp(argv[2], "reads") && argc == 9) {
        const uint64_t G = strtoull(argv[3], 0, 10), n_reads = strtoull(argv[4], 0, 10), L = strtoull(argv[5], 0, 10);
        const double err = atof(argv[6]);
        const uint64_t seed = strtoull(argv[7], 0, 10), wrap = strtoull(argv[8], 0, 10);
        const uint64_t err_thr = (uint64_t)(err * 4294967296.0);   // P(substitution) = err_thr / 2^32
        seed_rng(seed);
        uint8_t* g = make_genome(G);
        uint8_t* r = (uint8_t*)malloc(L);
        char* line = (char*)malloc(L + L / (wrap ? wrap : L) + 64);
        for (uint64_t i = 0; i < n_reads; i++) {
            const uint64_t p = below(G - L + 1);
            const int rc = (int)(next() >> 63);
            if (rc) for (uint64_t j = 0; j < L; j++) r[j] = (uint8_t)(3 - g[p + L - 1 - j]);
            else memcpy(r, g + p, L);
            if (err_thr) {
                for (uint64_t j = 0; j < L; j++) {
                    const uint64_t x = next();
                    if ((x & 0xFFFFFFFFu) < err_thr) r[j] = (uint8_t)((r[j] + 1 + ((x >> 32) % 3)) & 3);
                }
            }
            char hdr[32];
            int n = snprintf(hdr, sizeof hdr, ">r%llu\n", (unsigned long long)i);
            put(hdr, (size_t)n);
            size_t o = 0;
            for (uint64_t j = 0; j < L; j++) {
                line[o++] = ASCII[r[j]];
                if (wrap && (j + 1) % wrap == 0 && j + 1 < L) line[o++] = '\n';
            }
            line[o++] = '\n';
            put(line, o);
        }
    } else {
        fprintf(stderr, "gen_reads: bad arguments\n");
        return 1;
    }
    flush_out();
    if (out != stdout) fclose(out);
    return 0;
}
