// Host-side run of the position-independent code of csrc/kg_skm.cuh (minimizer bucketing): the SAME
// kg_skm_hash_word / kg_skm_segment_word / kg_window_at / kg_key_bucket functions the kernels call, driven block by
// block the way kg_skm_scatter drives them, and checked against a brute-force model written from the definition.
// stdin : u32 k, u32 nb, u32 pl, u32 C (carried bases), u32 nruns, then per run: u32 len, len bytes of 2-bit codes
// stdout: "OK windows=<n> descriptors=<n> max_bucket_share=<x>" or "FAIL: <what>"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/csrc/kg_skm.cuh"

struct Desc { u32 b, j0, n, hp; };

template <int W>
static bool check_windows(const std::vector<u64>& words, const std::vector<unsigned char>& bases, const std::vector<Desc>& descs,
                          u32 k, u32 m, u32 nb) {
    const KgKGeom g = kg_geom(k);
    for (const Desc& d : descs)
        for (u32 o = 0; o < d.n; o++) {
            const u32 e = d.j0 + o;
            u64 f[W], want[W], rc[W];
            kg_window_at<W>(words.data(), e, g, f);
            for (int i = 0; i < W; i++) { want[i] = 0; rc[i] = 0; }
            for (u32 j = 0; j < k; j++) {
                const u32 pos = k - 1 - j;
                want[W - 1 - pos / 32] |= (u64)bases[e + 1 - k + j] << (2 * (pos % 32));
                rc[W - 1 - pos / 32] |= (u64)(3u - bases[e - j]) << (2 * (pos % 32));
            }
            if (memcmp(f, want, sizeof f) != 0) { printf("FAIL: kg_window_at differs at end position %u\n", e); return false; }
            u64 wreg[W + 1], f2[W];
            kg_load_window_words<W>(words.data(), d.j0, wreg);       // loaded once per descriptor, as the insert kernel does
            kg_window_regs<W>(wreg, e, g, f2);
            if (memcmp(f2, want, sizeof f2) != 0) { printf("FAIL: kg_window_regs differs at end position %u\n", e); return false; }
            if (kg_key_bucket<W>(want, k, m, nb) != d.b || kg_key_bucket<W>(rc, k, m, nb) != d.b) {
                printf("FAIL: kg_key_bucket differs from the scatter's bucket at end position %u\n", e);
                return false;
            }
            if (o > 0 && e - 1 != d.j0 + o - 1) return false;
        }
    return true;
}

int main() {
    u32 hdr[5];
    if (fread(hdr, 4, 5, stdin) != 5) return 2;
    const u32 k = hdr[0], nb = hdr[1], pl = hdr[2], C = hdr[3], nruns = hdr[4];
    const u32 m = kg_skm_m(k), W = (k + 31) / 32;
    std::vector<unsigned char> bases;
    std::vector<unsigned char> starts;
    for (u32 r = 0; r < nruns; r++) {
        u32 len;
        if (fread(&len, 4, 1, stdin) != 1) return 2;
        const size_t at = bases.size();
        bases.resize(at + len);
        starts.resize(at + len, 0);
        if (len && fread(bases.data() + at, 1, len, stdin) != len) return 2;
        if (len) starts[at] = 1;
    }
    const u32 T = (u32)bases.size();
    const u32 nwords = T / 32 + W + 4;
    std::vector<u64> words(nwords + 1, 0);
    std::vector<u32> brk(nwords + 1, 0);
    for (u32 i = 0; i < T; i++) {
        words[i / 32] |= (u64)bases[i] << (62 - 2 * (i % 32));
        if (starts[i]) brk[i / 32] |= 1u << (31 - i % 32);
    }
    if (T) brk[0] |= 0x80000000u;
    // ---- the kernel's loop structure, block by block
    const u32 halo = (k - m + 31) / 32;
    std::vector<u32> H(KG_SKM_CHUNKS * 32), S(KG_SKM_CHUNKS * 32), M(KG_SKM_CHUNKS);
    std::vector<Desc> descs;
    u64 n_windows = 0;
    const u32 nblocks = (nwords + KG_SKM_TPB - 1) / KG_SKM_TPB;
    for (u32 blk = 0; blk < nblocks; blk++) {
        if ((u64)blk * KG_SKM_TPB * 32u >= T) continue;
        const long long first = (long long)blk * KG_SKM_TPB - (long long)halo;
        for (u32 lw = 0; lw < KG_SKM_TPB + halo; lw++) {
            const long long gc = first + lw;
            u32 mn = 0xFFFFFFFFu;
            if (gc < (long long)nwords) mn = kg_skm_hash_word(words.data(), gc, m, H.data(), S.data(), lw);
            else for (u32 i = 0; i < 32; i++) { H[KG_SKM_AT(lw, i)] = 0xFFFFFFFFu; S[KG_SKM_AT(lw, i)] = 0xFFFFFFFFu; }
            M[lw] = mn;
        }
        for (u32 tid = 0; tid < KG_SKM_TPB; tid++) {
            const u32 gw = blk * KG_SKM_TPB + tid;
            if (gw >= nwords) continue;
            n_windows += kg_skm_segment_word(words.data(), brk.data(), T, C, k, m, nb, gw, H.data(), S.data(), M.data(), tid + halo, W,
                                             [&](u32 b, u32 j0, u32 n, u32 hp) { descs.push_back(Desc{b, j0, n, hp}); });
        }
    }
    // ---- brute-force model from the definition
    std::vector<u32> run(T, 0), want_bucket(T, 0xFFFFFFFFu);
    for (u32 i = 0; i < T; i++) run[i] = (i == 0 || starts[i]) ? 1 : run[i - 1] + 1;
    u64 want_windows = 0;
    const u64 mmask = m == 32 ? ~0ULL : ((1ULL << (2 * m)) - 1ULL);
    for (u32 e = 0; e < T; e++) {
        if (run[e] < k || e < C) continue;
        want_windows++;
        u32 mv = 0xFFFFFFFFu;
        for (u32 me = e + 1 - k + m - 1; me <= e; me++) {            // m-mer ending at me
            u64 f = 0, r = 0;
            for (u32 j = 0; j < m; j++) {
                f = (f << 2) | bases[me + 1 - m + j];
                r = (r << 2) | (u64)(3u - bases[me - j]);
            }
            f &= mmask; r &= mmask;
            mv = std::min(mv, kg_mmer_hash(std::min(f, r)));
        }
        want_bucket[e] = kg_min_to_bucket(mv, nb);
    }
    if (n_windows != want_windows) { printf("FAIL: %llu windows, expected %llu\n", n_windows, want_windows); return 0; }
    std::vector<unsigned char> seen(T, 0);
    std::vector<u64> share(nb, 0);
    for (const Desc& d : descs) {
        if (d.n < 1 || d.n > KG_SKM_MAXRUN || d.b >= nb) { printf("FAIL: malformed descriptor\n"); return 0; }
        if (d.j0 / 32 != (d.j0 + d.n - 1) / 32) { printf("FAIL: descriptor crosses a packed word\n"); return 0; }
        if (d.hp != (run[d.j0] > k ? 1u : 0u)) { printf("FAIL: has_pred flag at %u\n", d.j0); return 0; }
        const u64 dd = kg_skm_desc(d.j0, 5, d.n, d.hp, d.b % pl, d.b / pl);
        if (KG_SKM_J0(dd) != d.j0 || KG_SKM_SRC(dd) != 5 || KG_SKM_N(dd) != d.n || KG_SKM_HP(dd) != d.hp ||
            KG_SKM_PART(dd) != d.b % pl || KG_SKM_OWNER(dd) != d.b / pl) { printf("FAIL: descriptor encoding\n"); return 0; }
        for (u32 o = 0; o < d.n; o++) {
            const u32 e = d.j0 + o;
            if (e >= T || seen[e] || want_bucket[e] != d.b) { printf("FAIL: window ending at %u (seen %d, bucket %u want %u)\n", e, e < T ? seen[e] : -1, d.b, e < T ? want_bucket[e] : 0); return 0; }
            seen[e] = 1;
        }
        share[d.b] += d.n;
    }
    for (u32 e = 0; e < T; e++)
        if ((want_bucket[e] != 0xFFFFFFFFu) != (seen[e] != 0)) { printf("FAIL: window ending at %u not covered exactly once\n", e); return 0; }
    bool ok = true;
    switch (W) {
        case 1: ok = check_windows<1>(words, bases, descs, k, m, nb); break;
        case 2: ok = check_windows<2>(words, bases, descs, k, m, nb); break;
        case 3: ok = check_windows<3>(words, bases, descs, k, m, nb); break;
        case 4: ok = check_windows<4>(words, bases, descs, k, m, nb); break;
        case 5: ok = check_windows<5>(words, bases, descs, k, m, nb); break;
        case 6: ok = check_windows<6>(words, bases, descs, k, m, nb); break;
        case 7: ok = check_windows<7>(words, bases, descs, k, m, nb); break;
        case 8: ok = check_windows<8>(words, bases, descs, k, m, nb); break;
    }
    if (!ok) return 0;
    const u64 mx = share.empty() ? 0 : *std::max_element(share.begin(), share.end());
    printf("OK windows=%llu descriptors=%zu max_bucket_share=%.4f\n", n_windows, descs.size(),
           n_windows ? (double)mx / (double)n_windows : 0.0);
    return 0;
}
