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

// Word-blocked double Bloom filter.  Filter bit h of the reference (double_bloomfilter.hpp:303-368: bit 2h = filter
// 1, bit 2h+1 = filter 2 of one interleaved array) becomes: all nh bit positions of a k-mer fall into ONE 64-bit word
// per filter, the two words sit side by side [F1 word][F2 word] (16 B, one sector).  m = bits per filter
// (main.cpp:404-410), nwords = m / 64.
// Why one word: a single atomicOr then is an exact, atomic test-and-set of the whole k-mer.  Of any number of
// concurrent occurrences of a k-mer exactly one sees "not all bits were set" -- no false negative is possible and
// new_in_first / new_in_second are exact without the reference's race heuristic ("a bit was set by someone else
// meanwhile => also put it into filter 2", :401-411), which with ~300k threads working on one L2-resident filter
// region fired for unrelated k-mers and inflated the false-positive rate to 2.6 % (a 256-bit-block version of this
// filter; profiles/r01_fullsize_reference_parity.txt shows the admitted singletons).
struct KgBloom {
    u64* bits;     // nwords * 2 words: [F1][F2] pairs
    u64 nblocks;   // nwords
    u32 nh;        // ceil(h) (main.cpp:417)
    u32 world;     // shards (the word index uses the in-shard part of the hash)
};

// the nh bit positions of a k-mer inside its 64-bit word (6 hash bits each)
__device__ __forceinline__ u64 kg_bloom_mask(u64 h, u32 nh) {
    u64 g = kg_fmix64(h ^ 0xA24BAED4963EE407ULL);
    u64 mask = 0;
    for (u32 i = 0; i < nh; i++) {
        if (i == 10) g = kg_fmix64(g + 0x9FB21C651E98DF25ULL);   // 10 probes per 64-bit draw
        mask |= 1ULL << ((g >> (6 * (i % 10))) & 63ULL);
    }
    return mask;
}
// word index = range partition of the in-shard hash (like the table slot), so a bucket of the partitioned path
// touches one contiguous region of the filter; the bit positions inside the word come from an independent mix
__device__ __forceinline__ u64 kg_bloom_block(u64 h, u64 nwords, u32 world) {
    return __umul64hi(kg_local_hash(h, world), nwords);
}

// pass 1 (insertion_process, double_bloomfilter.hpp:371-413): in F2 -> done; in F1 -> into F2; else into F1
__device__ __forceinline__ void kg_bloom_insert(const KgBloom& bf, u64 h, u32& new1, u32& new2) {
    const u64 mask = kg_bloom_mask(h, bf.nh);
    u64* w = bf.bits + kg_bloom_block(h, bf.nblocks, bf.world) * 2;
    u64 f1, f2;
    kg_ld_v2(w, f1, f2);
    if ((f2 & mask) == mask) return;                              // already in the second filter
    if ((f1 & mask) != mask) {
        const u64 old1 = atomicOr(w, mask);                       // exact test-and-set of the whole k-mer
        if ((old1 & mask) != mask) { new1++; return; }            // I am its first occurrence
    }
    const u64 old2 = atomicOr(w + 1, mask);
    if ((old2 & mask) != mask) new2++;                            // first to complete it in filter 2
}

// pass 2 admission (second_contains, double_bloomfilter.hpp:319-337; parallel_parser.hpp:2021-2026)
__device__ __forceinline__ bool kg_bloom_admits(const KgBloom& bf, u64 h) {
    const u64 mask = kg_bloom_mask(h, bf.nh);
    const u64 f2 = __ldg(bf.bits + kg_bloom_block(h, bf.nblocks, bf.world) * 2 + 1);
    return (f2 & mask) == mask;
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
    u32 n_new = 0, n_ins = 0, n_b1 = 0, n_b2 = 0, n_rej = 0;
    bool full = false;
    __device__ __forceinline__ void operator()(const u64 (&key)[W], u64 h, KgOcc occ, u32 = 0) {
        if (SINK == KG_SINK_BLOOM1) {
            kg_bloom_insert(bloom, h, n_b1, n_b2);
            return;
        }
        if (SINK == KG_SINK_BLOOM2 && !kg_bloom_admits(bloom, h)) { n_rej++; return; }
        bool is_new;
        u64 slot;
        if constexpr (W == 2) slot = table.packed_tb ? kg_table_add_packed(table, key, h, is_new) : kg_table_add<W>(table, key, h, is_new);
        else slot = kg_table_add<W>(table, key, h, is_new);
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
    sink.table = a.table;
    sink.bloom = a.bloom;
    u32 n_windows = kg_for_each_window<W>(a.words, a.brk, T, C, a.k, t, a.st->bases_seen, sink);
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
    sink.flush(a.stats);
}

// insert keys from a key array (received from other shards, or this GPU's partition-major bucket buffer).
// Blocks walk KG_CHUNK-key chunks (coalesced loads); statistics leave the block as one atomic per
// counter (one atomic per warp would put millions of RMWs on a single address).
#define KG_CHUNK 1024u            // keys per work item of the persistent insert kernels

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

template <int W, int SINK>
__global__ void __launch_bounds__(256) kg_insert_keys_kernel(const u64* __restrict__ keys, u64 n, const u32* n_dev,
                                                             KgTable table, KgBloom bloom, KgStats* stats, u32* work) {
    __shared__ u32 sm[8];
    if (n_dev) n = *n_dev;
    KgConsume<W, SINK> sink;
    sink.table = table;
    sink.bloom = bloom;
    // Persistent blocks pull KG_CHUNK-key chunks from a global work counter, so whatever their relative speed the
    // resident blocks always work at the FRONT of the (partition-major) key array: the keys in flight span about
    // grid * KG_CHUNK keys and the table region they hit stays L2-resident even when a partition holds only
    // ~1 M keys.  (A static block -> 4096-key mapping spread the in-flight window over ~5 M keys, several regions;
    // a plain grid-stride loop lets slow blocks fall behind and is worse still.)
    __shared__ u32 s_chunk;
    for (;;) {
        if (threadIdx.x == 0) s_chunk = atomicAdd(work, 1u);
        __syncthreads();
        const u64 first = (u64)s_chunk * KG_CHUNK;
        __syncthreads();
        if (first >= n) break;
#pragma unroll 1
        for (u32 j = 0; j < KG_CHUNK / 256; j++) {
            const u64 i = first + (u64)j * 256u + threadIdx.x;
            if (i < n) {
                u64 key[W];
                if constexpr (W % 2 == 0) {                                    // streamed once: evict-first, 16 B loads
#pragma unroll
                    for (int q = 0; q < W; q += 2) {
                        const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(keys + i * W + q));
                        key[q] = v.x; key[q + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < W; q++) key[q] = __ldcs(keys + i * W + q);
                }
                KgOcc none; none.word = ~0ULL;
                sink(key, kg_hash_key<W>(key), none);
            }
        }
    }
    if (SINK == KG_SINK_TABLE || SINK == KG_SINK_BLOOM2) {
        kg_block_add(sink.n_ins, &stats->inserted, sm);
        kg_block_add(sink.n_new, &stats->distinct, sm);
    }
    if (SINK == KG_SINK_BLOOM1) {
        kg_block_add(sink.n_b1, &stats->new_in_first, sm);
        kg_block_add(sink.n_b2, &stats->new_in_second, sm);
    }
    if (SINK == KG_SINK_BLOOM2) kg_block_add(sink.n_rej, &stats->bloom_rejected, sm);
    if (sink.full) stats->table_full = 1;
}

// Same, but the keys are gathered through a segment table: logical key i lives at
// keys[seg_src[j] + (i - seg_start[j])] for the segment j with seg_start[j] <= i < seg_start[j+1].
// The receiver of an exchange orders the segments partition-major across senders, so that one launch still
// walks the table region by region although every sender delivered its own partition-sorted run.
template <int W, int SINK>
__global__ void __launch_bounds__(256) kg_insert_segs_kernel(const u64* __restrict__ keys, const u64* __restrict__ seg_start,
                                                             const u64* __restrict__ seg_src, u32 nseg,
                                                             KgTable table, KgBloom bloom, KgStats* stats, u32* work) {
    __shared__ u32 sm[8];
    const u64 n = seg_start[nseg];
    KgConsume<W, SINK> sink;
    sink.table = table;
    sink.bloom = bloom;
    __shared__ u32 s_chunk, s_seg0;
    for (;;) {
        if (threadIdx.x == 0) {
            const u32 ch = atomicAdd(work, 1u);
            s_chunk = ch;
            // segment holding the first key of the chunk (one binary search per chunk, not per key)
            const u64 f = (u64)ch * KG_CHUNK;
            u32 lo = 0, hi = nseg;
            if (f < n) while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (seg_start[mid] <= f) lo = mid; else hi = mid; }
            s_seg0 = lo;
        }
        __syncthreads();
        const u64 first = (u64)s_chunk * KG_CHUNK;
        u32 seg = s_seg0;
        __syncthreads();
        if (first >= n) break;
#pragma unroll 1
        for (u32 j = 0; j < KG_CHUNK / 256; j++) {
            const u64 i = first + (u64)j * 256u + threadIdx.x;
            if (i < n) {
                while (seg + 1 < nseg && seg_start[seg + 1] <= i) seg++;     // segments are ~10^5 keys: rarely advances
                const u64 src = seg_src[seg] + (i - seg_start[seg]);
                u64 key[W];
                if constexpr (W % 2 == 0) {
#pragma unroll
                    for (int q = 0; q < W; q += 2) {
                        const ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(keys + src * W + q));
                        key[q] = v.x; key[q + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < W; q++) key[q] = __ldcs(keys + src * W + q);
                }
                KgOcc none; none.word = ~0ULL;
                sink(key, kg_hash_key<W>(key), none);
            }
        }
    }
    if (SINK == KG_SINK_TABLE || SINK == KG_SINK_BLOOM2) {
        kg_block_add(sink.n_ins, &stats->inserted, sm);
        kg_block_add(sink.n_new, &stats->distinct, sm);
    }
    if (SINK == KG_SINK_BLOOM1) {
        kg_block_add(sink.n_b1, &stats->new_in_first, sm);
        kg_block_add(sink.n_b2, &stats->new_in_second, sm);
    }
    if (SINK == KG_SINK_BLOOM2) kg_block_add(sink.n_rej, &stats->bloom_rejected, sm);
    if (sink.full) stats->table_full = 1;
}

// ---- multi-GPU bucketing: exact, deterministic layout without global atomics --------------------------------
//   kg_owner_hist    per block: how many of its k-mers go to each owner shard      -> blk_hist[block][owner]
//   kg_bucket_scan   one block: bucket totals, bucket offsets, per-block bases     -> blk_base[block][owner]
//   kg_owner_scatter per block: write each key at blk_base[owner] + (shared cursor)++   -> send buffer
// hist and scatter MUST be launched with the same grid (a block sees the same k-mers in both).
#define KG_MAX_BUCKETS 1024

struct KgBucketArgs {
    const u64* words;
    const u32* brk;
    const KgStream* st;
    u32* blk_hist;      // [nblocks][nb]
    u32* blk_base;      // [nblocks][nb], relative to the bucket start
    const u32* bucket_offs;  // [nb+1] bucket start in the send buffer
    u64* out_keys;      // send buffer, W words per key
    KgStats* stats;
    u32 k;
    u32 nb;             // number of buckets: world (multi-GPU) or partitions (single GPU)
    u32 world;          // > 1: bucket = owner shard; 1: bucket = partition of the local hash
};

// bucket = floor(h * nb / 2^64) with nb = world * local_partitions: the high part is the owner shard
// (floor(h*world/2^64)), the low part the partition of the in-shard hash -- buckets are owner-major, and the
// buckets of one owner map to consecutive, contiguous regions of that owner's table.
__device__ __forceinline__ u32 kg_bucket_of(u64 h, u32 world, u32 nb) {
    (void)world;
    return (u32)__umul64hi(h, (u64)nb);
}

template <int W>
__global__ void __launch_bounds__(128) kg_owner_hist(KgBucketArgs a) {   // launched with KgBucketGeom<W>::WPB threads
    __shared__ u32 s_hist[KG_MAX_BUCKETS];
    for (u32 i = threadIdx.x; i < a.nb; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const u32 T = a.st->total_bases, C = a.st->carry_bases;
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 nb = a.nb;
    u32 n_windows = kg_for_each_window<W>(a.words, a.brk, T, C, a.k, t, 0,
                                          [&](const u64 (&key)[W], u64 h, KgOcc, u32) { (void)key; atomicAdd(&s_hist[kg_bucket_of(h, a.world, nb)], 1u); });
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
    __syncthreads();
    for (u32 i = threadIdx.x; i < nb; i += blockDim.x) a.blk_hist[(u64)blockIdx.x * nb + i] = s_hist[i];
}

// Column scan, one block per bucket b: blk_base[block][b] = keys of bucket b held by earlier blocks
// (exclusive, RELATIVE to the bucket start); counts_out[b] = bucket total.
__global__ void __launch_bounds__(1024) kg_bucket_colscan(const u32* __restrict__ blk_hist, u32* __restrict__ blk_base,
                                                          u32 nblocks, u32 nb, u32* counts_out) {
    __shared__ u32 sm[1024];
    const u32 b = blockIdx.x;
    const u32 per = (nblocks + 1023) / 1024;
    const u32 b0 = threadIdx.x * per, b1 = min(b0 + per, nblocks);
    u32 mine = 0;
    for (u32 i = b0; i < b1; i++) mine += blk_hist[(u64)i * nb + b];
    sm[threadIdx.x] = mine;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {          // inclusive Hillis-Steele scan of the 1024 partial sums
        u32 v = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    u32 cur = sm[threadIdx.x] - mine;
    for (u32 i = b0; i < b1; i++) { u32 h = blk_hist[(u64)i * nb + b]; blk_base[(u64)i * nb + b] = cur; cur += h; }
    if (threadIdx.x == 1023) counts_out[b] = sm[1023];
}

// offs_out[b] = first key index of bucket b in the send buffer; offs_out[nb] = total keys
__global__ void __launch_bounds__(1024) kg_bucket_offsets(const u32* __restrict__ counts, u32 nb, u32* offs_out) {
    __shared__ u32 sm[1024];
    const u32 v = threadIdx.x < nb ? counts[threadIdx.x] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {
        u32 t = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
        __syncthreads();
        sm[threadIdx.x] += t;
        __syncthreads();
    }
    if (threadIdx.x < nb) offs_out[threadIdx.x] = sm[threadIdx.x] - v;
    if (threadIdx.x == 1023) offs_out[nb] = sm[1023];
}

// eight consecutive windows (positions 32t+j0 .. +7), fully unrolled so per-window state can live in registers
template <int W, typename F>
__device__ __forceinline__ u32 kg_window8(const u64* __restrict__ words, const u32* __restrict__ brk,
                                          u32 T, u32 C, u32 k, u32 t, u32 j0, F&& f) {
    if ((u64)t * 32u + j0 >= T) return 0;
    const KgKGeom g = kg_geom(k);
    const u64 myword = words[t];
    const u32 mybrk = brk[t];
    const u32 jend = min(j0 + 8u, T - t * 32u);
    u32 run = 0;
    {
        const u32 head = j0 ? (mybrk >> (32 - j0)) : 0u;
        if (head) {
            run = __ffs(head);
        } else {
            run = j0;
            bool found = false;
#pragma unroll 1
            for (int i = 1; i <= W + 1 && !found; i++) {
                if ((int)t - i < 0) break;
                u32 b = brk[t - i];
                if (b) { run += __ffs(b); found = true; }
                else run += 32;
            }
        }
    }
    KgKmerWindow<W> w;
    if (j0 == 0) {
#pragma unroll
        for (int i = 0; i < W; i++) {
            int src = (int)t - 1 - i;
            w.f[W - 1 - i] = src >= 0 ? words[src] : 0ULL;
        }
    } else {
        const u32 s = 64 - 2 * j0;
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
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const u32 j = j0 + q;
        if (j < jend) {
            const u32 c = (u32)(myword >> (62 - 2 * j)) & 3u;
            kg_push<W>(w, g, c);
            run = ((mybrk >> (31 - j)) & 1u) ? 1u : run + 1u;
            if (run >= k && base_pos + j >= C) {
                n_windows++;
                u64 key[W];
                const bool fwd = kg_forward_is_canonical<W>(w);
#pragma unroll
                for (int i = 0; i < W; i++) key[i] = fwd ? w.f[i] : w.r[i];
                f(q, key, kg_hash_key<W>(key));
            }
        }
    }
    return n_windows;
}

// Geometry shared by hist and scatter: a block owns KgBucketGeom<W>::WPB packed words (32*WPB k-mer end positions).
// hist runs one thread per word; scatter runs four threads per word (8 positions each) and stages the block's
// keys in shared memory so that every bucket leaves the block as ONE contiguous, coalesced run.  (Writing each key
// straight from its thread costs two 8-byte partial-sector stores per k-mer and runs at ~25 G keys/s; see
// profiles/r01_partitioned_scatter_insert_ncu.txt and profiles/r01_scatter_probe_result.txt.)
template <int W>
struct KgBucketGeom {
    static constexpr int WPB = W <= 2 ? 128 : (W <= 4 ? 64 : 32);   // words per block: <= 64 KiB of staged keys
    static constexpr int TPB = 4 * WPB;                             // scatter threads per block
    static constexpr int KEYS = 32 * WPB;                           // staged keys per block (upper bound)
    static constexpr size_t smem_bytes(u32 nb) {
        return (size_t)KEYS * W * 8 + (size_t)KEYS * 2 /*bucket of a staged key*/ + (size_t)TPB * 8 * 4 /*ranks*/ +
               (size_t)nb * 4 * 3 /*count, offset, global base*/ + 64;
    }
};

// Peer exchange (kg_peer_connect): instead of a local send buffer the runs go STRAIGHT into the owners' receive
// buffers over NVLink (plain stores to peer memory, visible at kernel end).  peer[d] = owner d's receive buffer for
// this round as mapped into this process (d == own rank: a local buffer), remote_base[b] = first key index of this
// rank's run for bucket b in that buffer (kg_exchange_plan.hpp), pl = local partitions per owner (owner = b / pl).
struct KgPeerArgs {
    u64* const* peer;
    const u64* remote_base;
    u32 pl;
};

template <int W, bool PEER>
__device__ __forceinline__ void kg_owner_scatter_body(const KgBucketArgs& a, const KgPeerArgs& pa) {
    using G = KgBucketGeom<W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* s_keys = reinterpret_cast<u64*>(smem_raw);                                  // KEYS * W
    u32* s_rank = reinterpret_cast<u32*>(s_keys + (size_t)G::KEYS * W);                // TPB * 8
    u32* s_cnt = s_rank + G::TPB * 8;                                                  // nb
    u32* s_off = s_cnt + a.nb;                                                         // nb
    u32* s_gbase = s_off + a.nb;                                                       // nb
    unsigned short* s_kb = reinterpret_cast<unsigned short*>(s_gbase + a.nb);          // KEYS
    __shared__ u32 s_warp[32];
    const u32 nb = a.nb, tid = threadIdx.x;
    for (u32 i = tid; i < nb; i += G::TPB) s_cnt[i] = 0;
    __syncthreads();
    const u32 T = a.st->total_bases, C = a.st->carry_bases;
    const u32 t = blockIdx.x * G::WPB + (tid >> 2);
    const u32 j0 = (tid & 3u) * 8u;
    // pass 1: bucket and rank-in-bucket of each of my (<= 8) windows.  W <= 4: the keys stay in registers and
    // the windows are computed once; wider keys recompute them in pass 2 (register budget).
    constexpr bool kInRegs = W <= 4;
    u64 kreg[kInRegs ? 8 : 1][W];
    u32 br[8];
#pragma unroll
    for (int q = 0; q < 8; q++) br[q] = 0xFFFFFFFFu;
    if constexpr (kInRegs) {
        kg_window8<W>(a.words, a.brk, T, C, a.k, t, j0, [&](int q, const u64 (&key)[W], u64 h) {
            const u32 b = kg_bucket_of(h, a.world, nb);
            const u32 r = atomicAdd(&s_cnt[b], 1u);
            br[q] = (b << 16) | r;
#pragma unroll
            for (int i = 0; i < W; i++) kreg[q][i] = key[i];
        });
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) s_rank[q * G::TPB + tid] = 0xFFFFFFFFu;
        kg_for_each_window<W>(a.words, a.brk, T, C, a.k, t, 0, [&](const u64 (&key)[W], u64 h, KgOcc, u32 j) {
            (void)key;
            const u32 b = kg_bucket_of(h, a.world, nb);
            const u32 r = atomicAdd(&s_cnt[b], 1u);
            s_rank[(j - j0) * G::TPB + tid] = (b << 16) | r;
        }, j0, j0 + 8);
    }
    __syncthreads();
    // exclusive scan of the bucket counts (nb <= 1024), and the global base of every bucket run
    {
        const u32 per = (nb + G::TPB - 1) / G::TPB;
        const u32 b0 = tid * per, b1 = min(b0 + per, nb);
        u32 mine = 0;
        for (u32 i = b0; i < b1; i++) mine += s_cnt[i];
        u32 incl = mine;
        const u32 lane = tid & 31u, warp = tid >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (u32)d) incl += o; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        u32 pre = 0;
        for (u32 w2 = 0; w2 < warp; w2++) pre += s_warp[w2];
        u32 cur = pre + incl - mine;
        for (u32 i = b0; i < b1; i++) {
            s_off[i] = cur;
            cur += s_cnt[i];
            if constexpr (PEER) s_gbase[i] = (u32)pa.remote_base[i] + a.blk_base[(u64)blockIdx.x * nb + i];   // key index in the owner's buffer (< 2^32)
            else s_gbase[i] = a.bucket_offs[i] + a.blk_base[(u64)blockIdx.x * nb + i];
        }
    }
    __syncthreads();
    // pass 2: drop each key at its staged position (keys grouped by bucket)
    if constexpr (kInRegs) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (br[q] != 0xFFFFFFFFu) {
                const u32 b = br[q] >> 16, idx = s_off[b] + (br[q] & 0xFFFFu);
#pragma unroll
                for (int i = 0; i < W; i++) s_keys[(size_t)idx * W + i] = kreg[q][i];
                s_kb[idx] = (unsigned short)b;
            }
        }
    } else {
        kg_for_each_window<W>(a.words, a.brk, T, C, a.k, t, 0, [&](const u64 (&key)[W], u64, KgOcc, u32 j) {
            const u32 brj = s_rank[(j - j0) * G::TPB + tid];
            const u32 b = brj >> 16, idx = s_off[b] + (brj & 0xFFFFu);
#pragma unroll
            for (int i = 0; i < W; i++) s_keys[(size_t)idx * W + i] = key[i];
            s_kb[idx] = (unsigned short)b;
        }, j0, j0 + 8);
    }
    __syncthreads();
    // coalesced write-out: consecutive staged keys of one bucket go to consecutive global addresses
    const u32 n = s_off[nb - 1] + s_cnt[nb - 1];
    for (u32 i = tid; i < n; i += G::TPB) {
        const u32 b = s_kb[i];
        u64* dst;
        if constexpr (PEER) dst = pa.peer[b / pa.pl] + (u64)(s_gbase[b] + (i - s_off[b])) * W;
        else dst = a.out_keys + (u64)(s_gbase[b] + (i - s_off[b])) * W;
        if (W % 2 == 0) {
#pragma unroll
            for (int q = 0; q < W; q += 2)
                *reinterpret_cast<ulonglong2*>(dst + q) = make_ulonglong2(s_keys[(size_t)i * W + q], s_keys[(size_t)i * W + q + 1]);
        } else {
#pragma unroll
            for (int q = 0; q < W; q++) dst[q] = s_keys[(size_t)i * W + q];
        }
    }
}

template <int W>
__global__ void __launch_bounds__(KgBucketGeom<W>::TPB) kg_owner_scatter(KgBucketArgs a) {
    kg_owner_scatter_body<W, false>(a, KgPeerArgs{nullptr, nullptr, 1});
}
// fused bucket -> peer-store kernel: the exchange IS the scatter's write-out
template <int W>
__global__ void __launch_bounds__(KgBucketGeom<W>::TPB) kg_owner_scatter_peer(KgBucketArgs a, KgPeerArgs pa) {
    kg_owner_scatter_body<W, true>(a, pa);
}

// ---- single-GPU one-pass bucketing: reserve, don't count ----------------------------------------------------------
// On one GPU the bucket layout need not be exact (nothing is sent anywhere), so the histogram pass is dropped:
// a block computes its keys ONCE (kept in registers), ranks them with shared atomics, reserves a run in every
// bucket's fixed-capacity region with one global atomic per (block, bucket), stages the keys in shared memory
// and writes coalesced runs.  A key that does not fit its region (pathological skew) is inserted directly.
// One window pass instead of three (hist + two in kg_owner_scatter).

struct KgReserveArgs {
    const u64* words;
    const u32* brk;
    const KgStream* st;
    u32* cursors;       // [nb] keys reserved so far in every bucket region (may run past cap: overflow)
    u64* out_keys;      // nb regions of cap keys
    KgStats* stats;
    KgTable table;      // overflow keys are inserted directly
    KgBloom bloom;
    u32 k;
    u32 nb;
    u32 cap;            // keys per bucket region
};

template <int W, int SINK>
__global__ void __launch_bounds__(KgBucketGeom<W>::TPB) kg_scatter_reserve(KgReserveArgs a) {
    using G = KgBucketGeom<W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* s_keys = reinterpret_cast<u64*>(smem_raw);                                  // KEYS * W
    u32* s_cnt = reinterpret_cast<u32*>(s_keys + (size_t)G::KEYS * W);                 // nb
    u32* s_off = s_cnt + a.nb;                                                         // nb
    u32* s_gbase = s_off + a.nb;                                                       // nb
    unsigned short* s_kb = reinterpret_cast<unsigned short*>(s_gbase + a.nb);          // KEYS
    __shared__ u32 s_warp[32];
    __shared__ u32 s_stat[8];
    const u32 nb = a.nb, tid = threadIdx.x;
    for (u32 i = tid; i < nb; i += G::TPB) s_cnt[i] = 0;
    __syncthreads();
    const u32 T = a.st->total_bases, C = a.st->carry_bases;
    const u32 t = blockIdx.x * G::WPB + (tid >> 2);
    const u32 j0 = (tid & 3u) * 8u;
    u64 kreg[8][W];
    u32 br[8];
#pragma unroll
    for (int q = 0; q < 8; q++) br[q] = 0xFFFFFFFFu;
    const u32 n_windows = kg_window8<W>(a.words, a.brk, T, C, a.k, t, j0, [&](int q, const u64 (&key)[W], u64 h) {
        const u32 b = (u32)__umul64hi(h, (u64)nb);
        const u32 r = atomicAdd(&s_cnt[b], 1u);
        br[q] = (b << 16) | r;
#pragma unroll
        for (int i = 0; i < W; i++) kreg[q][i] = key[i];
    });
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
    __syncthreads();
    {   // exclusive scan of the bucket counts + one global reservation per non-empty bucket
        const u32 per = (nb + G::TPB - 1) / G::TPB;
        const u32 b0 = tid * per, b1 = min(b0 + per, nb);
        u32 mine = 0;
        for (u32 i = b0; i < b1; i++) mine += s_cnt[i];
        u32 incl = mine;
        const u32 lane = tid & 31u, warp = tid >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (u32)d) incl += o; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        u32 pre = 0;
        for (u32 w2 = 0; w2 < warp; w2++) pre += s_warp[w2];
        u32 cur = pre + incl - mine;
        for (u32 i = b0; i < b1; i++) {
            const u32 c = s_cnt[i];
            s_off[i] = cur;
            cur += c;
            s_gbase[i] = c ? atomicAdd(&a.cursors[i], c) : 0u;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (br[q] != 0xFFFFFFFFu) {
            const u32 b = br[q] >> 16, idx = s_off[b] + (br[q] & 0xFFFFu);
#pragma unroll
            for (int i = 0; i < W; i++) s_keys[(size_t)idx * W + i] = kreg[q][i];
            s_kb[idx] = (unsigned short)b;
        }
    }
    __syncthreads();
    const u32 n = s_off[nb - 1] + s_cnt[nb - 1];
    KgConsume<W, SINK> sink;
    sink.table = a.table;
    sink.bloom = a.bloom;
    for (u32 i = tid; i < n; i += G::TPB) {
        const u32 b = s_kb[i];
        const u32 rel = s_gbase[b] + (i - s_off[b]);
        if (rel < a.cap) {
            u64* dst = a.out_keys + ((u64)b * a.cap + rel) * W;
            if (W % 2 == 0) {
#pragma unroll
                for (int q = 0; q < W; q += 2)
                    *reinterpret_cast<ulonglong2*>(dst + q) = make_ulonglong2(s_keys[(size_t)i * W + q], s_keys[(size_t)i * W + q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < W; q++) dst[q] = s_keys[(size_t)i * W + q];
            }
        } else {                                   // region full: insert this key here and now
            u64 key[W];
#pragma unroll
            for (int q = 0; q < W; q++) key[q] = s_keys[(size_t)i * W + q];
            KgOcc none; none.word = ~0ULL;
            sink(key, kg_hash_key<W>(key), none);
        }
    }
    if (SINK == KG_SINK_TABLE || SINK == KG_SINK_BLOOM2) {
        kg_block_add(sink.n_ins, &a.stats->inserted, s_stat);
        kg_block_add(sink.n_new, &a.stats->distinct, s_stat);
    }
    if (SINK == KG_SINK_BLOOM1) {
        kg_block_add(sink.n_b1, &a.stats->new_in_first, s_stat);
        kg_block_add(sink.n_b2, &a.stats->new_in_second, s_stat);
    }
    if (SINK == KG_SINK_BLOOM2) kg_block_add(sink.n_rej, &a.stats->bloom_rejected, s_stat);
    if (sink.full) a.stats->table_full = 1;
}

// segment table over the bucket regions: seg b = keys [b*cap, b*cap + min(cursor[b], cap))
__global__ void __launch_bounds__(1024) kg_seg_from_cursors(const u32* __restrict__ cursors, u32 nb, u32 cap,
                                                            u64* __restrict__ seg_start, u64* __restrict__ seg_src) {
    __shared__ u64 sm[1024];
    const u64 v = threadIdx.x < nb ? (u64)min(cursors[threadIdx.x], cap) : 0ULL;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {
        u64 t = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
        __syncthreads();
        sm[threadIdx.x] += t;
        __syncthreads();
    }
    if (threadIdx.x < nb) { seg_start[threadIdx.x] = sm[threadIdx.x] - v; seg_src[threadIdx.x] = (u64)threadIdx.x * cap; }
    if (threadIdx.x == 1023) seg_start[nb] = sm[1023];
}

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
            n = (u32)(p[0] >> table.packed_tb);
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
