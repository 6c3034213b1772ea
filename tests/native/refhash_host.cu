// Host-side check of csrc/kg_refhash.cuh (the reference's own hash functions, restated __host__ __device__).
// stdin : u32 0, u32 n, n x (u64 value, u64 seed)  ->  stdout: XXH64(le64(value), seed) per line;   or
// stdin : u32 k, u32 n, n bytes of 2-bit codes (one maximal run)          stdout (text), one line per window:
//   "<Hf mod 2^54> <Hb mod 2^54> <root> <XXH64(root, seed_0)> ... <XXH64(root, seed_15)>"
// The first window's hashes come from the Horner routine over packed keys (forward and reverse complement), every
// later window from the O(1) rolling update -- the two code paths the kernels will use.
#include <cstdio>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/csrc/kg_refhash.cuh"

static void pack(const std::vector<unsigned char>& c, size_t at, unsigned k, bool rc, std::vector<unsigned long long>& key) {
    const unsigned W = (k + 31) / 32;
    key.assign(W, 0);
    for (unsigned j = 0; j < k; j++) {
        const unsigned code = rc ? 3u - c[at + k - 1 - j] : c[at + j];
        const unsigned pos = k - 1 - j;
        key[W - 1 - pos / 32] |= (unsigned long long)code << (2 * (pos % 32));
    }
}

int main() {
    unsigned hdr[2];
    if (fread(hdr, 4, 2, stdin) != 2) return 2;
    if (hdr[0] == 0) {   // XXH64 mode: n (value, seed) pairs of u64 -> one hash per line
        for (unsigned i = 0; i < hdr[1]; i++) {
            unsigned long long vs[2];
            if (fread(vs, 8, 2, stdin) != 2) return 2;
            printf("%llu\n", (unsigned long long)kg_xxh64_8(vs[0], vs[1]));
        }
        return 0;
    }
    const unsigned k = hdr[0], n = hdr[1], W = (k + 31) / 32;
    std::vector<unsigned char> c(n);
    if (n && fread(c.data(), 1, n, stdin) != n) return 2;
    if (n < k) return 0;
    std::vector<unsigned long long> f, r;
    pack(c, 0, k, false, f);
    pack(c, 0, k, true, r);
    uint64_t hf = kg_b5_horner(f.data(), W, k), hb = kg_b5_horner(r.data(), W, k);
    const uint64_t p5k1 = kg_pow5(k - 1);
    for (unsigned j = 0;; j++) {
        const uint64_t root = kg_ref_root(hf, hb);
        printf("%llu %llu %llu", (unsigned long long)(hf & KG_REF_MASK54), (unsigned long long)(hb & KG_REF_MASK54), (unsigned long long)root);
        for (unsigned i = 0; i < 16; i++) printf(" %llu", (unsigned long long)kg_xxh64_8(root, kg_ref_seed(i)));
        printf("\n");
        if (j + k >= n) break;
        kg_b5_roll(hf, hb, c[j + k], c[j], p5k1);
    }
    return 0;
}
