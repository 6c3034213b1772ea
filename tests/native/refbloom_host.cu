// Host-side run of the per-window logic of csrc/kg_refbloom.cuh (the bit-exact Bloom emulation): the
// SAME kg_rb_window<1|2|3> / kg_rb_admits functions the kernels call, driven over the windows of one run of bases in a
// scrambled order, with plain minimum in place of atomicMin.
// stdin : u32 k, u32 n, u32 log2(m), u32 nh, u32 nh2, u32 first_ordinal, n bytes of 2-bit codes
// stdout: line 1 "new_in_first new_in_second"; line 2 the set bits of filter 2 (ascending); line 3 one 0/1 per window
//         (admitted in pass 2), in input order
#include <cstdio>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/csrc/kg_refbloom.cuh"

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
    unsigned hdr[6];
    if (fread(hdr, 4, 6, stdin) != 6) return 2;
    const unsigned k = hdr[0], n = hdr[1], W = (k + 31) / 32;
    std::vector<unsigned char> c(n);
    if (n && fread(c.data(), 1, n, stdin) != n) return 2;
    const size_t m = (size_t)1 << hdr[2];
    std::vector<u32> T1(m, KG_RB_NEVER), T2(m, KG_RB_NEVER);
    KgRefBloom rb{T1.data(), T2.data(), (u64)m - 1, hdr[3], hdr[4]};
    const unsigned nw = n >= k ? n - k + 1 : 0;
    std::vector<unsigned long long> roots(nw), f, r;
    for (unsigned j = 0; j < nw; j++) {
        pack(c, j, k, false, f);
        pack(c, j, k, true, r);
        roots[j] = kg_ref_root(kg_b5_horner(f.data(), W, k), kg_b5_horner(r.data(), W, k));
    }
    // a fixed scrambled visiting order (the sweeps must not care)
    std::vector<unsigned> order(nw);
    for (unsigned j = 0; j < nw; j++) order[j] = j;
    unsigned long long s = 88172645463325252ULL;
    for (unsigned j = nw; j > 1; j--) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; std::swap(order[j - 1], order[s % j]); }
    u32 n1 = 0, n2 = 0;
    u64 pos[KG_RB_MAX_NH];
    for (int sweep = 1; sweep <= 3; sweep++)
        for (unsigned j : order) {
            kg_rb_positions_from_root(roots[j], rb, pos);
            const u32 ord = hdr[5] + j + k - 1;                     // global end position of window j
            if (sweep == 1) kg_rb_window<1>(rb, pos, ord, n1, n2);
            else if (sweep == 2) kg_rb_window<2>(rb, pos, ord, n1, n2);
            else kg_rb_window<3>(rb, pos, ord, n1, n2);
        }
    printf("%u %u\n", n1, n2);
    for (size_t b = 0; b < m; b++) if (T2[b] != KG_RB_NEVER) printf("%zu ", b);
    printf("\n");
    for (unsigned j = 0; j < nw; j++) { kg_rb_positions_from_root(roots[j], rb, pos); printf("%d", kg_rb_admits(rb, pos) ? 1 : 0); }
    printf("\n");
    return 0;
}
