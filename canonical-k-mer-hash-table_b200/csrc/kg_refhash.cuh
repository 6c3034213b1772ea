// kg_refhash.cuh -- the reference's OWN hash functions, for the bit-exact double-Bloom-filter emulation mode
// (DESIGN.md section 9).  The product path hashes canonical keys with kg_hash_key; to
// reproduce the reference's filter contents bit for bit the same numbers the reference computes are needed:
//   root(w)  = min(Hf, Hb) mod 2^54,  Hf = sum c_i 5^(k-1-i),  Hb = sum (3-c_i) 5^i      hash_functions.cpp:102-232
//   h_i(w)   = XXH64(le64(root), seed_i) & (m-1)                                          double_bloomfilter.hpp:276-281
// Hb is the same polynomial read over the reverse complement, so both come from one Horner routine over a packed
// right-aligned key; while a window slides one base they are updated in O(1) (kg_b5_roll).  Arithmetic is mod 2^64
// and masked to 54 bits at the end: 2^54 divides 2^64, and 5 is invertible mod 2^64, so +, -, x and the division by
// 5 of the reference's `di = 5^-1 mod q` commute with the mask.
// Plain integer code, __host__ __device__: tests/native/refhash_host.cu checks it on a CPU against known answers
// minted from the reference's own objects (tests/golden/kats.json).  The kernels that use it are next round's work.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KG_RH_HD __host__ __device__ __forceinline__
#else
#define KG_RH_HD inline
#endif

#define KG_REF_MASK54 ((1ULL << 54) - 1ULL)
#define KG_INV5_MOD64 0xCCCCCCCCCCCCCCCDULL   // 5 * 0xCCCCCCCCCCCCCCCD == 1 (mod 2^64)

KG_RH_HD uint64_t kg_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// XXH64 of exactly 8 bytes (xxHash specification, "input length < 32" path with one 8-byte lane)
KG_RH_HD uint64_t kg_xxh64_8(uint64_t value, uint64_t seed) {
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL,
                   P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint64_t h = seed + P5 + 8ULL;
    h ^= kg_rotl64(value * P2, 31) * P1;
    h = kg_rotl64(h, 27) * P1 + P4;
    h ^= h >> 33; h *= P2;
    h ^= h >> 29; h *= P3;
    h ^= h >> 32;
    return h;
}

// the first 16 seeds of double_bloomfilter.hpp:434-444 (nh = ceil(h) <= 16 for every fpr the CLI accepts)
KG_RH_HD uint64_t kg_ref_seed(uint32_t i) {
    const uint16_t s[16] = {2411, 3253, 1061, 1129, 2269, 7309, 3491, 8237, 6359, 8779, 6553, 5443, 2447, 8999, 8623, 5779};
    return s[i & 15u];
}

// sum c_i 5^(k-1-i) mod 2^64 over a right-aligned 2-bit key (W = ceil(k/32) words, word 0 most significant)
KG_RH_HD uint64_t kg_b5_horner(const unsigned long long* key, uint32_t W, uint32_t k) {
    uint64_t h = 0;
    uint32_t nchar = k - 32u * (W - 1u);
    for (uint32_t w = 0; w < W; w++) {
        const uint64_t v = key[w];
        for (int c = (int)nchar - 1; c >= 0; c--) h = h * 5ULL + ((v >> (2 * c)) & 3ULL);
        nchar = 32u;
    }
    return h;
}

// the window slides one base: c_out leaves on the left, c_in enters on the right; p5k1 = 5^(k-1) mod 2^64
KG_RH_HD void kg_b5_roll(uint64_t& hf, uint64_t& hb, uint32_t c_in, uint32_t c_out, uint64_t p5k1) {
    hf = (hf - (uint64_t)c_out * p5k1) * 5ULL + (uint64_t)c_in;
    hb = (hb - (uint64_t)(3u - c_out)) * KG_INV5_MOD64 + (uint64_t)(3u - c_in) * p5k1;
}

KG_RH_HD uint64_t kg_pow5(uint32_t e) {
    uint64_t r = 1, b = 5;
    while (e) { if (e & 1u) r *= b; b *= b; e >>= 1; }
    return r;
}

// parallel_parser.hpp:2889-2894: root = min(Hb, Hf) of the mod-2^54 hashes
KG_RH_HD uint64_t kg_ref_root(uint64_t hf, uint64_t hb) {
    const uint64_t a = hf & KG_REF_MASK54, b = hb & KG_REF_MASK54;
    return a < b ? a : b;
}
