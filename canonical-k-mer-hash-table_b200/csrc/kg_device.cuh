// kg_device.cuh -- device-side building blocks shared by the kernels of libkaarme_gpu.so (sm_100a).
//
// Data layout in HBM:
//   packed base stream   u64 words, 32 bases per word, first base in bits 63:62 (so numeric order ==
//                        lexicographic order, as in KMerFactoryCanonical2BC, kmer_factory.cpp:172-205)
//   break mask           u32 words, bit (31 - i%32) of word i/32 set <=> base i starts a new run
//                        (after a header, a non-ACGT byte, or the start of a stream)
//   count table          open addressing, linear probing (+1 slot, what parallel_parser.hpp:711-789
//                        effectively does), slot = [meta u64][key words W][first-pos u64 if Kaarme],
//                        stride padded to 16 B (W=1) or a multiple of 32 B so a probe touches whole sectors
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define KG_LOCKED 0xFFFFFFFFu

// ---- statistics / error word written by the kernels -------------------------------------------------
struct KgStats {
    u64 input_kmers;     // complete windows seen
    u64 inserted;        // occurrences inserted into the local table
    u64 distinct;        // slots claimed
    u64 new_in_first;    // Bloom counters (double_bloomfilter.hpp:389-396)
    u64 new_in_second;
    u64 bloom_rejected;  // pass 2: windows not admitted
    u32 table_full;      // an insert exhausted the probe budget
    u32 pad;
};

// ---- stream state carried from batch to batch (lives in device memory; no host round trip) ----------
struct KgStream {
    u32 in_header;     // header state at the first byte of the next batch (text_chunk::broken_header)
    u32 pending_break; // a break event happened after the last packed base
    u32 carry_bases;   // bases at the head of the packed stream that belong to the previous batch (C)
    u32 total_bases;   // bases in the packed stream of the current batch, carry included (T)
    u64 bases_seen;    // bases packed so far in this stream, before the current batch (global ordinal base)
};

// ---- 64-bit mixers ---------------------------------------------------------------------------------
__device__ __forceinline__ u64 kg_fmix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// hash of a canonical key (replaces RollingHasherDual, hash_functions.cpp:102-232: the parity contract
// is the k-mer multiset, not the slot order, so any well-mixed function of the canonical key serves)
template <int W>
__device__ __forceinline__ u64 kg_hash_key(const u64 (&key)[W]) {
    // multiply-xorshift chain: one 64-bit multiply per key word and one to finish.  Consumers take the HIGH bits (range
    // partition by multiply-high), which is where a product mixes best; the xorshift between the multiplies brings the high
    // bits of every word back down so that they reach the high bits of the result as well.  (Two earlier versions ran a
    // full fmix64 per word, then fmix64 once at the end: 2W resp. W + 4 multiplies; the hash was ~30 % of the
    // instructions of the insert kernel.)
    u64 h = 0x9E3779B97F4A7C15ULL;
#pragma unroll
    for (int i = 0; i < W; i++) {
        h = (h ^ key[i]) * (0xD6E8FEB86659FD93ULL + 2ULL * (u64)i);
        h ^= h >> 32;
    }
    h *= 0x9FB21C651E98DF25ULL;
    return h ^ (h >> 29);
}

// Range partitioning of the 64-bit hash: owner shard = floor(h * world / 2^64); inside the shard the
// fractional part (low 64 bits of h * world) is uniform again, and slot = floor(frac * nslots / 2^64).
// A partition p of `nb` (single GPU) is floor(frac * nb / 2^64), so partition p maps to the CONTIGUOUS slot
// range [p*nslots/nb, (p+1)*nslots/nb): inserting one partition at a time keeps the live table region in L2.
__device__ __forceinline__ u32 kg_owner(u64 h, u32 world) { return (u32)__umul64hi(h, (u64)world); }
__device__ __forceinline__ u64 kg_local_hash(u64 h, u32 world) { return h * (u64)world; }
// position of hash h inside a range of n slots starting at lo: floor(h * n / 2^64), with one 32-bit multiply when the
// range has fewer than 2^32 slots (every partition of the bucketed path; uses the high 32 bits of h)
__device__ __forceinline__ u64 kg_place(u64 h, u64 lo, u64 n) {
    return lo + ((n >> 32) ? __umul64hi(h, n) : (((h >> 32) * (n & 0xFFFFFFFFULL)) >> 32));
}
__device__ __forceinline__ u64 kg_slot(u64 h, u64 nslots, u32 world) { return kg_place(kg_local_hash(h, world), 0, nslots); }

// ---- strong (L2-coherent, L1-bypassing) accesses used on the table -----------------------------------
__device__ __forceinline__ u32 kg_ld_u32(const void* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 kg_ld_u64(const void* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// acquire loads: what follows (the key words of a slot) is ordered after the observation of a published `meta`
__device__ __forceinline__ u32 kg_ld_acquire_u32(const void* p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 kg_ld_acquire_u64(const void* p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void kg_ld_v2(const void* p, u64& a, u64& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void kg_st_u64(void* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void kg_st_release_u32(void* p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- table geometry ----------------------------------------------------------------------------------
struct KgTable {
    u64* slots;      // nslots * stride words
    u64 nslots;      // next_prime3mod4(...)
    u32 stride;      // u64 words per slot
    u32 kaarme;      // slot carries a first-occurrence word at [1+W]
    u32 world;       // shards the hash space is split into (slot uses the in-shard fraction of the hash)
    u32 packed_tb;   // != 0: PACKED 16-byte slots (W == 2 only): word 0 = count << packed_tb | key word 0,
                     // word 1 = key word 1; packed_tb = bits of key word 0 in use (2k - 64)
    const u32* full_flag;  // KgStats::table_full of the pass: a probe sequence that runs long gives up once it is set
};

// Probe budget.  The reference probes the whole table before it reports "full" (parallel_parser.hpp:742-746); here a
// probe sequence may run over min(nslots, 2^20) slots -- at any load factor below ~0.9999 no cluster is that long -- and
// once ONE insert has exhausted its budget every other long-running probe gives up within 256 steps (full_flag), so a
// full table is reported in about a second instead of hanging the device.
#define KG_MAX_PROBE (1ull << 20)
#define KG_COUNT_MARGIN 0x100000u   // counters stop this far below their limit: more threads than this never race on one slot

__host__ __device__ inline u32 kg_slot_stride_words(u32 W, bool kaarme) {
    u32 need = 1 + W + (kaarme ? 1u : 0u);
    if (need <= 2) return 2;           // 16 B: two slots per 32-byte sector
    return (need + 3u) & ~3u;          // multiple of 32 B
}

// find-or-insert `key` starting at `slot`, count += 1.  Returns the slot index, or ~0 when the probe budget is exhausted.
//   meta low 32 bits: 0 = empty, KG_LOCKED = being written, otherwise the occurrence count.
//   Writer: CAS 0->LOCKED, store key words, release-store count=1.  Reader: an ACQUIRE load of meta, then the key words
//   (W > 1; relaxed loads ordered after the acquire -- PTX does not order loads by control dependency).  W == 1 reads
//   meta and the key in one aligned 16-byte access.  Key words never change once published.
// The probe loops below have ONE exit and no side effect other than the claim: lanes of a warp whose probe is over wait at
// the loop's end for the others, and the count update (RED) and the statistics run once, converged, after it.  (With
// `break`s inside, the compiler kept the lanes apart and replayed the whole tail once per loop iteration: the probe code
// was 55 % of the insert kernel's instructions at a third of a warp's width.)
#define KG_PROBE_SEARCH 0
#define KG_PROBE_FOUND 1
#define KG_PROBE_NEW 2
#define KG_PROBE_FAIL 3
template <int W>
__device__ __forceinline__ u64 kg_table_add(const KgTable& t, const u64 (&key)[W], u64 slot, bool& is_new) {
    u64* p = t.slots + slot * t.stride;
    u64* const end = t.slots + t.nslots * t.stride;
    u32 budget = (u32)(t.nslots < KG_MAX_PROBE ? t.nslots : KG_MAX_PROBE);
    u32 m = 0;
    int state = KG_PROBE_SEARCH;
    do {
        u64 meta, k0 = 0;
        bool k0_valid = false;
        if (W == 1) { kg_ld_v2(p, meta, k0); k0_valid = true; }
        else meta = kg_ld_acquire_u64(p);
        m = (u32)meta;
        if (m == 0) {
            const u32 old = atomicCAS((u32*)p, 0u, KG_LOCKED);
            if (old == 0) {
#pragma unroll
                for (int i = 0; i < W; i++) kg_st_u64(p + 1 + i, key[i]);
                kg_st_release_u32(p, 1u);
                state = KG_PROBE_NEW;
            }
            m = KG_LOCKED;                     // somebody else owns or owned it: observe it again, with acquire
        }
        if (state == KG_PROBE_SEARCH) {
            if (m == KG_LOCKED) {
                do { m = kg_ld_acquire_u32(p); } while (m == KG_LOCKED);
                k0_valid = false;
            }
            // slot is published: compare
            if (!k0_valid) k0 = kg_ld_u64(p + 1);
            bool same = (k0 == key[0]);
            if (same) {
#pragma unroll
                for (int i = 1; i < W; i++) same = same && (kg_ld_u64(p + 1 + i) == key[i]);
            }
            if (same) {
                state = KG_PROBE_FOUND;
            } else {
                p += t.stride;
                if (p == end) p = t.slots;
                if (--budget == 0 || ((budget & 255u) == 0 && t.full_flag && kg_ld_u32(t.full_flag))) state = KG_PROBE_FAIL;
            }
        }
    } while (state == KG_PROBE_SEARCH);
    is_new = state == KG_PROBE_NEW;
    if (state == KG_PROBE_FOUND && m < 0xFFFFFFFFu - KG_COUNT_MARGIN) atomicAdd((u32*)p, 1u);   // saturates below KG_LOCKED
    return state == KG_PROBE_FAIL ? ~0ULL : (u64)(p - t.slots) / t.stride;
}

// ---- packed 16-byte slots (W == 2, 2k - 64 + count bits <= 64) ---------------------------------------------------
// One 16-byte load decides a probe, one 128-bit CAS claims an empty slot with its count already 1, one 64-bit RED
// bumps the count: no lock state, no fence, two L2 transactions on the hit path instead of four.
__device__ __forceinline__ void kg_cas128(u64* p, u64 cmp0, u64 cmp1, u64 val0, u64 val1, u64& old0, u64& old1) {
    asm volatile(
        "{\n\t"
        ".reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}"
        : "=l"(old0), "=l"(old1)
        : "l"(cmp0), "l"(cmp1), "l"(val0), "l"(val1), "l"(p)
        : "memory");
}
__device__ __forceinline__ void kg_red_add_u64(u64* p, u64 v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 kg_table_add_packed(const KgTable& t, const u64 (&key)[2], u64 slot, bool& is_new) {
    const u32 tb = t.packed_tb;
    const u64 keymask = (1ULL << tb) - 1, one = 1ULL << tb;
    const u64 sat = ((~0ULL) >> tb) - (u64)KG_COUNT_MARGIN;   // stop counting before the field could wrap into the key
    u64* p = t.slots + slot * 2;
    u64* const end = t.slots + t.nslots * 2;
    u32 budget = (u32)(t.nslots < KG_MAX_PROBE ? t.nslots : KG_MAX_PROBE);
    u64 w0, w1;
    int state = KG_PROBE_SEARCH;
    do {
        kg_ld_v2(p, w0, w1);
        if ((w0 | w1) == 0) {
            kg_cas128(p, 0, 0, one | key[0], key[1], w0, w1);     // w0, w1 = what was there
            if ((w0 | w1) == 0) state = KG_PROBE_NEW;
        }
        if (state == KG_PROBE_SEARCH) {
            if ((w0 & keymask) == key[0] && w1 == key[1]) {
                state = KG_PROBE_FOUND;
            } else {
                p += 2;
                if (p == end) p = t.slots;
                if (--budget == 0 || ((budget & 255u) == 0 && t.full_flag && kg_ld_u32(t.full_flag))) state = KG_PROBE_FAIL;
            }
        }
    } while (state == KG_PROBE_SEARCH);
    is_new = state == KG_PROBE_NEW;
    if (state == KG_PROBE_FOUND && (w0 >> tb) < sat) kg_red_add_u64(p, one);
    return state == KG_PROBE_FAIL ? ~0ULL : (u64)(p - t.slots) >> 1;
}

// ---- one probe at a time (the bucketed insert, kg_skm.cuh) --------------------------------------------------------------
// kg_probe_step looks at ONE slot.  FOUND / NEW: the occurrence is counted, p is the k-mer's slot.  SEARCH: not decided
// yet -- p has moved on to the next slot, or stays where it is when the slot is being written by somebody else right now
// (no spinning: the caller parks the key in its warp's queue and comes back with a full warp of such keys).
template <int W>
__device__ __forceinline__ int kg_probe_step(const KgTable& t, const u64 (&key)[W], u64*& p) {
    if constexpr (W == 2) {
        if (t.packed_tb) {
            const u32 tb = t.packed_tb;
            const u64 keymask = (1ULL << tb) - 1, one = 1ULL << tb;
            u64 w0, w1;
            kg_ld_v2(p, w0, w1);
            if ((w0 | w1) == 0) {
                kg_cas128(p, 0, 0, one | key[0], key[1], w0, w1);     // w0, w1 = what was there
                if ((w0 | w1) == 0) return KG_PROBE_NEW;
            }
            if ((w0 & keymask) == key[0] && w1 == key[1]) {
                if ((w0 >> tb) < ((~0ULL) >> tb) - (u64)KG_COUNT_MARGIN) kg_red_add_u64(p, one);
                return KG_PROBE_FOUND;
            }
            p += 2;
            if (p == t.slots + t.nslots * 2) p = t.slots;
            return KG_PROBE_SEARCH;
        }
    }
    u64 meta, k0 = 0;
    if (W == 1) kg_ld_v2(p, meta, k0);                 // one aligned 16-byte access: key word 0 comes with a published meta
    else meta = kg_ld_acquire_u64(p);
    const u32 m = (u32)meta;
    if (m == 0) {
        if (atomicCAS((u32*)p, 0u, KG_LOCKED) == 0u) {
#pragma unroll
            for (int i = 0; i < W; i++) kg_st_u64(p + 1 + i, key[i]);
            kg_st_release_u32(p, 1u);
            return KG_PROBE_NEW;
        }
        return KG_PROBE_SEARCH;                        // claimed by somebody else this instant: look again later
    }
    if (m == KG_LOCKED) return KG_PROBE_SEARCH;        // being written: look again later
    if (W != 1) k0 = kg_ld_u64(p + 1);
    bool same = (k0 == key[0]);
    if (same) {
#pragma unroll
        for (int i = 1; i < W; i++) same = same && (kg_ld_u64(p + 1 + i) == key[i]);
    }
    if (same) {
        if (m < 0xFFFFFFFFu - KG_COUNT_MARGIN) atomicAdd((u32*)p, 1u);
        return KG_PROBE_FOUND;
    }
    p += t.stride;
    if (p == t.slots + t.nslots * t.stride) p = t.slots;
    return KG_PROBE_SEARCH;
}

// read-only lookup starting at `slot` (compaction / decode); returns slot or ~0
template <int W>
__device__ __forceinline__ u64 kg_table_find(const KgTable& t, const u64 (&key)[W], u64 slot) {
    const u64 max_probe = t.nslots < KG_MAX_PROBE ? t.nslots : KG_MAX_PROBE;
    for (u64 probe = 0; probe < max_probe; probe++) {
        const u64* p = t.slots + slot * t.stride;
        if ((u32)p[0] == 0) return ~0ULL;
        bool same = true;
#pragma unroll
        for (int i = 0; i < W; i++) same = same && (p[1 + i] == key[i]);
        if (same) return slot;
        slot = slot + 1 == t.nslots ? 0 : slot + 1;
    }
    return ~0ULL;
}

// ---- 2-bit helpers -----------------------------------------------------------------------------------
// reverse the order of the 32 2-bit characters of a word
__device__ __forceinline__ u64 kg_rev2(u64 x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

template <int W>
struct KgKmerWindow {
    u64 f[W];  // forward window, right-aligned, word 0 most significant
    u64 r[W];  // reverse complement
};

// masks for a k-mer of W words
struct KgKGeom {
    u32 k;
    u32 topbits;  // bits of word 0 in use: 2k - 64(W-1), in (0,64]
    u64 topmask;
};
__host__ __device__ inline KgKGeom kg_geom(u32 k) {
    KgKGeom g;
    u32 W = (k + 31) / 32;
    g.k = k;
    g.topbits = 2 * k - 64 * (W - 1);
    g.topmask = g.topbits == 64 ? ~0ULL : ((1ULL << g.topbits) - 1);
    return g;
}

template <int W>
__device__ __forceinline__ void kg_push(KgKmerWindow<W>& w, const KgKGeom& g, u32 c) {
#pragma unroll
    for (int i = 0; i < W - 1; i++) w.f[i] = (w.f[i] << 2) | (w.f[i + 1] >> 62);
    w.f[W - 1] = (w.f[W - 1] << 2) | (u64)c;
    w.f[0] &= g.topmask;
#pragma unroll
    for (int i = W - 1; i > 0; i--) w.r[i] = (w.r[i] >> 2) | (w.r[i - 1] << 62);
    w.r[0] = (w.r[0] >> 2) | ((u64)(3u - c) << (g.topbits - 2));
}

// reverse complement of a right-aligned 2k-bit value
template <int W>
__device__ __forceinline__ void kg_revcomp(const u64 (&f)[W], u64 (&r)[W], const KgKGeom& g) {
    // reverse all 32*W characters, complement, then shift right by (64W - 2k) bits
    u64 t[W];
#pragma unroll
    for (int i = 0; i < W; i++) t[i] = ~kg_rev2(f[W - 1 - i]);
    const u32 sh = 64 - g.topbits;  // 0..62, even
    if (sh == 0) {
#pragma unroll
        for (int i = 0; i < W; i++) r[i] = t[i];
    } else {
#pragma unroll
        for (int i = W - 1; i > 0; i--) r[i] = (t[i] >> sh) | (t[i - 1] << (64 - sh));
        r[0] = t[0] >> sh;
    }
}

// forward is canonical when it is <= reverse complement (ties keep forward: parallel_parser.hpp:686-702)
template <int W>
__device__ __forceinline__ bool kg_forward_is_canonical(const KgKmerWindow<W>& w) {
    bool fwd = true;  // equal => forward
#pragma unroll
    for (int i = W - 1; i >= 0; i--) {
        if (w.f[i] != w.r[i]) fwd = w.f[i] < w.r[i];
    }
    return fwd;
}
