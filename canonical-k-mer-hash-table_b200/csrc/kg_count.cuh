// kg_count.cuh -- K1b/K3 "window + insert": rolling forward / reverse-complement window over the packed
// base stream, canonical selection, hash, and find-or-insert into the count table.
//
// Replaces, per base of input:
//   window shift + reverse complement + lexicographic min   parallel_parser.hpp:643-702 (-m 0),
//                                                           kmer_factory.cpp:172-239 (-m 2 / Bloom)
//   rolling hash                                            hash_functions.cpp:102-232
//   find-or-insert + count                                  parallel_parser.hpp:711-789,
//                                                           kmer_hash_table.cpp:2207-2567 (process_kmer_MT)
// Mapping: one thread per packed word (32 consecutive k-mer END positions).  The thread rebuilds the
// window of the k-1 preceding bases straight from the packed words (O(W) funnel shifts + one bit-reversal
// for the reverse complement) instead of warming up base by base, then rolls 32 steps.
#pragma once
#include "kg_device.cuh"

enum KgSink : int {
    KG_SINK_TABLE = 0,   // insert into the local table
    KG_SINK_BLOOM1 = 1,  // Bloom pass 1: test-and-set F1/F2
    KG_SINK_BLOOM2 = 2   // Bloom pass 2: insert only if F2 admits
};

// Blocked double Bloom filter.  Filter bit h of the reference (double_bloomfilter.hpp:303-368: bit 2h = filter 1, bit
// 2h+1 = filter 2 of one interleaved array) becomes: all nh bit positions of a k-mer fall into ONE 256-bit block per filter,
// and the two blocks sit side by side [F1: 4 words][F2: 4 words] (64 B, two sectors).  m = bits per filter
// (main.cpp:404-410), nblocks = m / 256, total 2m bits as in the reference.
// Why 256 bits: the false-positive rate of a blocked filter grows as the blocks shrink (a block holding more k-mers than
// average is fuller than average).  Measured against the reference's 2 028 admitted singletons on C3 (-u 8e7, 56 M distinct
// 31-mers): 64-bit blocks admitted 27 461 (13.5x), the model gives 5.0x for 128-bit and 2.6x for 256-bit blocks.
// Exactness: pass 1 must decide "first / second / later occurrence" atomically per k-mer, or two concurrent first
// occurrences both believe they are first and the k-mer never reaches filter 2 (a false NEGATIVE; the reference papers over
// this with "a bit was set by someone else meanwhile => also put it into filter 2", double_bloomfilter.hpp:401-411, which on
// a GPU fires for unrelated k-mers all the time).  No atomic spans 256 bits, so bit 255 of the F1 block is a LOCK: the
// test-and-set of both filters runs under it.  Occurrences of k-mers already in filter 2 (most of them at coverage >= 3)
// never take the lock: an unlocked look at F2 can at worst see a half-written block and take the slow path.
struct KgBloom {
    u64* bits;     // nblocks * 8 words: [F1 x4][F2 x4]
    u64 nblocks;   // 256-bit blocks per filter
    u32 nh;        // ceil(h) (main.cpp:417)
    u32 world;     // (unused since placement went by partition; kept for the struct layout)
};
#define KG_BLOOM_LOCK 0x8000000000000000ULL   // bit 63 of F1 word 3 = bit 255 of the block; never a k-mer's position

// the nh bit positions of a k-mer inside its 256-bit block (positions 0..254), as four word masks
__device__ __forceinline__ void kg_bloom_mask(u64 h, u32 nh, u64 (&m)[4]) {
    u64 g = kg_fmix64(h ^ 0xA24BAED4963EE407ULL);
    m[0] = m[1] = m[2] = m[3] = 0;
    for (u32 i = 0; i < nh; i++) {
        if (i == 8) g = kg_fmix64(g + 0x9FB21C651E98DF25ULL);    // 8 positions per 64-bit draw
        const u32 pos = (u32)((((g >> (8 * (i & 7u))) & 255ULL) * 255ULL) >> 8);   // 0..254
        const u64 bit = 1ULL << (pos & 63u);
        const u32 w = pos >> 6;
        m[0] |= w == 0 ? bit : 0ULL; m[1] |= w == 1 ? bit : 0ULL; m[2] |= w == 2 ? bit : 0ULL; m[3] |= w == 3 ? bit : 0ULL;
    }
}
__device__ __forceinline__ u64 kg_atom_or_acquire(u64* p, u64 v) {
    u64 old;
    asm volatile("atom.acquire.gpu.global.or.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
__device__ __forceinline__ void kg_red_and_release(u64* p, u64 v) {
    asm volatile("red.release.gpu.global.and.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool kg_bloom_has(const u64* blk, const u64 (&m)[4]) {
    u64 a, b, c, d;
    kg_ld_v2(blk, a, b);
    kg_ld_v2(blk + 2, c, d);
    return ((a & m[0]) == m[0]) & ((b & m[1]) == m[1]) & ((c & m[2]) == m[2]) & ((d & m[3]) == m[3]);
}

// pass 1 (insertion_process, double_bloomfilter.hpp:371-413): in F2 -> done; in F1 -> into F2; else into F1
__device__ __forceinline__ void kg_bloom_insert(const KgBloom& bf, u64 h, u64 block, u32& new1, u32& new2) {
    u64 m[4];
    kg_bloom_mask(h, bf.nh, m);
    u64* blk = bf.bits + block * 8;
    if (kg_bloom_has(blk + 4, m)) return;                             // already in the second filter (no lock needed)
    bool done = false;
    while (!done) {                                                  // (the lock holder finishes inside its own iteration)
        if ((kg_atom_or_acquire(blk + 3, KG_BLOOM_LOCK) & KG_BLOOM_LOCK) == 0) {
            u64* dst = blk;                                          // where the k-mer goes: F1, else F2, else nowhere
            u32* counter = &new1;
            if (kg_bloom_has(blk, m)) {
                dst = blk + 4; counter = &new2;
                if (kg_bloom_has(blk + 4, m)) dst = nullptr;
            }
            if (dst) {
                (*counter)++;
#pragma unroll
                for (int i = 0; i < 4; i++) if (m[i]) atomicOr(dst + i, m[i]);   // (RED: F1 word 3 also carries the lock bit)
            }
            kg_red_and_release(blk + 3, ~KG_BLOOM_LOCK);
            done = true;
        }
    }
}

// pass 2 admission (second_contains, double_bloomfilter.hpp:319-337; parallel_parser.hpp:2021-2026)
__device__ __forceinline__ bool kg_bloom_admits(const KgBloom& bf, u64 h, u64 block) {
    u64 m[4];
    kg_bloom_mask(h, bf.nh, m);
    const ulonglong2* f2 = reinterpret_cast<const ulonglong2*>(bf.bits + block * 8 + 4);
    const ulonglong2 lo = __ldg(f2), hi = __ldg(f2 + 1);
    return ((lo.x & m[0]) == m[0]) & ((lo.y & m[1]) == m[1]) & ((hi.x & m[2]) == m[2]) & ((hi.y & m[3]) == m[3]);
}

struct KgCountArgs {
    const u64* words;
    const u32* brk;
    const KgStream* st;
    KgTable table;
    KgBloom bloom;
    KgStats* stats;
    u32 k;
    u32 rank, world;
};

// What the Kaarme representation needs to know about one occurrence of a k-mer (kmer.hpp:108-123 flags):
// packed as (global_end_position << 4) | has_pred << 3 | self_forward_canonical << 2 | dropped_char
struct KgOcc {
    u64 word;
};
__device__ __forceinline__ KgOcc kg_make_occ(u64 global_pos, bool has_pred, bool fwd, u32 c_out) {
    KgOcc o;
    o.word = (global_pos << 4) | ((u64)has_pred << 3) | ((u64)fwd << 2) | (u64)c_out;
    return o;
}

// Walk the k-mer windows that END inside packed word t: calls f(key, hash, occ) once per complete window whose
// end position is >= C (positions below C belong to the previous batch).  Returns the number of windows.
template <int W, typename F>
__device__ __forceinline__ u32 kg_for_each_window(const u64* __restrict__ words, const u32* __restrict__ brk,
                                                  u32 T, u32 C, u32 k, u32 t, u64 pos0, F&& f,
                                                  u32 j0 = 0, u32 j1 = 32) {
    if ((u64)t * 32u + j0 >= T) return 0;
    const KgKGeom g = kg_geom(k);
    const u64 myword = words[t];
    const u32 mybrk = brk[t];
    const u32 jend = min(j1, T - t * 32u);
    // run length at the base just before position 32t + j0: distance back to the most recent run start
    u32 run = 0;
    {
        const u32 head = j0 ? (mybrk >> (32 - j0)) : 0u;         // break bits of bases 0..j0-1 of this word
        if (head) {
            run = __ffs(head);
        } else {
            run = j0;
            bool found = false;
#pragma unroll 1
            for (int i = 1; i <= W + 1 && !found; i++) {
                if ((int)t - i < 0) break;                          // position 0 always carries a break bit
                u32 b = brk[t - i];
                if (b) { run += __ffs(b); found = true; }           // lowest set bit = most recent base of that word
                else run += 32;
            }
        }
    }
    KgKmerWindow<W> w;
    // forward window = the k bases before position 32t + j0, right-aligned (the oldest one is the base that drops
    // out at the first step: the predecessor's first base, needed by the Kaarme occurrence record)
    if (j0 == 0) {
#pragma unroll
        for (int i = 0; i < W; i++) {
            int src = (int)t - 1 - i;
            w.f[W - 1 - i] = src >= 0 ? words[src] : 0ULL;
        }
    } else {
        const u32 s = 64 - 2 * j0;                                   // unread bits of this word, 2..62
        u64 lo = myword;
#pragma unroll
        for (int i = 0; i < W; i++) {
            int src = (int)t - 1 - i;
            const u64 hi = src >= 0 ? words[src] : 0ULL;
            w.f[W - 1 - i] = (hi << (64 - s)) | (lo >> s);
            lo = hi;
        }
    }
    w.f[0] &= g.topmask;
    kg_revcomp<W>(w.f, w.r, g);
    const u32 base_pos = t * 32u;
    u32 n_windows = 0;
#pragma unroll 1
    for (u32 j = j0; j < jend; j++) {
        const u32 c = (u32)(myword >> (62 - 2 * j)) & 3u;
        const u32 c_out = (u32)(w.f[0] >> (g.topbits - 2)) & 3u;   // base leaving the window = first base of the predecessor
        kg_push<W>(w, g, c);
        run = ((mybrk >> (31 - j)) & 1u) ? 1u : run + 1u;
        if (run >= k && base_pos + j >= C) {
            n_windows++;
            u64 key[W];
            const bool fwd = kg_forward_is_canonical<W>(w);
#pragma unroll
            for (int i = 0; i < W; i++) key[i] = fwd ? w.f[i] : w.r[i];
            f(key, kg_hash_key<W>(key), kg_make_occ(pos0 + base_pos + j, run > k, fwd, c_out), j);
        }
    }
    return n_windows;
}

#define KG_WARP_ADD(stats, var, field)                                            \
    {                                                                             \
        u32 v_ = var;                                                             \
        for (int d = 16; d; d >>= 1) v_ += __shfl_xor_sync(0xffffffffu, v_, d);   \
        if ((threadIdx.x & 31u) == 0 && v_) atomicAdd(&(stats)->field, (u64)v_);  \
    }

// what to do with one canonical k-mer on the shard that owns it
template <int W, int SINK>
struct KgConsume {
    KgTable table;
    KgBloom bloom;
    // placement: slot0 = t_lo + floor(h' * t_n / 2^64), Bloom word = b_lo + floor(h' * b_n / 2^64), h' = h * hmul.
    // Defaults (set by init()) = the whole table / filter of the shard with the in-shard part of the hash (hmul = world);
    // the minimizer-bucketed insert (kg_skm.cuh) sets the partition's range per window and hmul = 1.
    u64 t_lo = 0, t_n = 0, b_lo = 0, b_n = 0;
    u32 hmul = 1;
    u32 n_new = 0, n_ins = 0, n_b1 = 0, n_b2 = 0, n_rej = 0;
    bool full = false;
    __device__ __forceinline__ void init(const KgTable& t, const KgBloom& b) {
        table = t; bloom = b;
        t_lo = 0; t_n = t.nslots; b_lo = 0; b_n = b.nblocks; hmul = t.world;
    }
    __device__ __forceinline__ void operator()(const u64 (&key)[W], u64 h, KgOcc occ, u32 = 0) {
        const u64 hl = h * (u64)hmul;
        if (SINK == KG_SINK_BLOOM1) {
            kg_bloom_insert(bloom, h, kg_place(hl, b_lo, b_n), n_b1, n_b2);
            return;
        }
        if (SINK == KG_SINK_BLOOM2 && !kg_bloom_admits(bloom, h, kg_place(hl, b_lo, b_n))) { n_rej++; return; }
        bool is_new;
        u64 slot;
        const u64 slot0 = kg_place(hl, t_lo, t_n);
        if constexpr (W == 2) slot = table.packed_tb ? kg_table_add_packed(table, key, slot0, is_new) : kg_table_add<W>(table, key, slot0, is_new);
        else slot = kg_table_add<W>(table, key, slot0, is_new);
        if (slot == ~0ULL) { full = true; return; }
        n_ins++;
        n_new += is_new ? 1u : 0u;
        // Kaarme mode: remember the EARLIEST occurrence (atomicMax of the complement; the table starts zeroed).
        // Same sector as the slot that was just touched, so it stays an L2 hit.
        if (table.kaarme) atomicMax(table.slots + slot * table.stride + 1 + W, ~occ.word);
    }
    __device__ __forceinline__ void flush(KgStats* stats) {
        if (SINK == KG_SINK_TABLE || SINK == KG_SINK_BLOOM2) {
            KG_WARP_ADD(stats, n_ins, inserted)
            KG_WARP_ADD(stats, n_new, distinct)
        }
        if (SINK == KG_SINK_BLOOM1) {
            KG_WARP_ADD(stats, n_b1, new_in_first)
            KG_WARP_ADD(stats, n_b2, new_in_second)
        }
        if (SINK == KG_SINK_BLOOM2) { KG_WARP_ADD(stats, n_rej, bloom_rejected) }
        if (full) stats->table_full = 1;
    }
};

template <int W, int SINK>
__global__ void __launch_bounds__(256) kg_count_kernel(KgCountArgs a) {
    const u32 T = a.st->total_bases;
    const u32 C = a.st->carry_bases;
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    KgConsume<W, SINK> sink;
    sink.init(a.table, a.bloom);
    u32 n_windows = kg_for_each_window<W>(a.words, a.brk, T, C, a.k, t, a.st->bases_seen, sink);
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
    sink.flush(a.stats);
}

// block-level add of a per-thread counter: one atomic per counter per block (one per warp would put millions of RMWs on
// a single address in the persistent insert kernel)
__device__ __forceinline__ void kg_block_add(u32 v, u64* dst, u32* smem /*8 words*/) {
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31u) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (u32 w = 0; w < blockDim.x / 32; w++) t += smem[w];
        if (t) atomicAdd(dst, (u64)t);
    }
    __syncthreads();
}

#define KG_MAX_BUCKETS 1024

// ---- K5 export: stream-compact slots whose reported count >= min_abundance -------------------------------
// (replaces the table scan of write_kmers, kmer_hash_table.cpp:2013-2050)
__device__ __forceinline__ u32 kg_reported_count(u32 n, int count_mode, int table_mode) {
    if (count_mode == 0) return n;
    if (table_mode == 0) return n & 0xFFFFu;          // uint16 wrap, parallel_parser.hpp:720-734
    return n > 16383u ? 16383u : n;                   // 14-bit saturation, kmer.cpp:699-714
}

template <int W>
__global__ void __launch_bounds__(256) kg_export_kernel(KgTable table, u64 slot_begin, u64 slot_end, u64 min_abundance,
                                                        int count_mode, int table_mode, u64* __restrict__ out_keys,
                                                        u32* __restrict__ out_counts, u32* out_n) {
    const u64 s = slot_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool emit = false;
    u32 rep = 0;
    const u64* p = nullptr;
    u64 pk0 = 0, pk1 = 0;
    if (s < slot_end) {
        p = table.slots + s * table.stride;
        u32 n;
        if (W == 2 && table.packed_tb) {
            pk0 = p[0] & ((1ULL << table.packed_tb) - 1);
            pk1 = p[1];
            const u64 n64 = p[0] >> table.packed_tb;             // up to 62 bits wide: clamp, do not truncate
            n = n64 > 0xFFFFFFFEULL ? 0xFFFFFFFEu : (u32)n64;
        } else {
            n = (u32)p[0];
        }
        if (n != 0 && n != KG_LOCKED) {
            rep = kg_reported_count(n, count_mode, table_mode);
            emit = min_abundance > 0 && (u64)rep >= min_abundance;
        }
    }
    const u32 ballot = __ballot_sync(0xffffffffu, emit);
    if (ballot == 0) return;
    const u32 lane = threadIdx.x & 31u;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(out_n, (u32)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (emit) {
        const u32 idx = base + __popc(ballot & ((1u << lane) - 1u));
        if (W == 2 && table.packed_tb) {
            out_keys[(u64)idx * 2] = pk0;
            out_keys[(u64)idx * 2 + 1] = pk1;
        } else {
#pragma unroll
            for (int i = 0; i < W; i++) out_keys[(u64)idx * W + i] = p[1 + i];
        }
        out_counts[idx] = rep;
    }
}
