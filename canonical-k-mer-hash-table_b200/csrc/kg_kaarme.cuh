// kg_kaarme.cuh -- K4 "compact": build the Kaarme representation from the counted table, and decode it.
//
// The reference builds its 8-byte-per-k-mer structure WHILE counting (PointerHashTableCanonicalAV::
// process_kmer_MT, kmer_hash_table.cpp:2207-2567): each k-mer stores one or two characters plus a pointer to
// the k-mer that preceded it in the read where it was first seen; k-mers without a predecessor ("roots") are
// stored in full in a secondary array (:2239-2282).  About 18 dependent cache misses per insertion.
// Here counting runs on the plain table (one sector per k-mer) and records, per k-mer, its EARLIEST occurrence
// (position, orientation, dropped base).  The structure is then built in one pass:
//     predecessor(X) = the k-mer that preceded X at X's earliest occurrence.
// Its own earliest occurrence is strictly earlier, so chains are acyclic by construction (the reference needs
// check_for_cycle, kmer_hash_table.cpp:3098-3244).
//
// Slot word (bit-compatible with kmer.hpp:108-123):
//   [63:26] predecessor index | root index   [25:12] count, saturating at 16383 (kmer.cpp:699-714)
//   [11:10] C[0]   [9:8] C[k-1]   bit5 predecessor forward-canonical   bit4 self forward-canonical
//   bit1 predecessor exists   bit0 occupied          (C = canonical string of the k-mer)
// Indices are DENSE (rank among occupied table slots), so the array holds exactly one word per k-mer.
#pragma once
#include "kg_skm.cuh"

struct KgKaarme {
    u64* slots;     // n_kmers words
    u64* roots;     // n_roots * W words
    u64 n_kmers;
    u64 n_roots;
};

struct KgCompactStats {
    u64 roots;
    u64 max_chain;
    u64 chain_sum;
    u64 bad;        // malformed chains seen by the decoder (must stay 0)
};

// ---- dense numbering of the occupied slots ------------------------------------------------------------------
// bitmap word w covers slots [32w, 32w+32); word_count[w] = popcount
__global__ void __launch_bounds__(256) kg_occupancy_bitmap(KgTable t, u32* __restrict__ bitmap, u32* __restrict__ word_count) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool occ = false;
    if (s < t.nslots) {
        u32 m = (u32)t.slots[s * t.stride];
        occ = m != 0 && m != KG_LOCKED;
    }
    const u32 b = __ballot_sync(0xffffffffu, occ);
    if ((threadIdx.x & 31u) == 0) {      // bitmap / word_count hold gridDim.x * 8 words: every warp writes its own
        const u64 w = s >> 5;
        bitmap[w] = b;
        word_count[w] = __popc(b);
    }
}

// exclusive scan of word_count in three steps (block sums -> scan of block sums -> add back)
__global__ void __launch_bounds__(1024) kg_scan_blocks(const u32* __restrict__ in, u64* __restrict__ block_sum, u64 n) {
    __shared__ u64 sm[32];
    const u64 i = (u64)blockIdx.x * 1024 + threadIdx.x;
    u64 v = i < n ? in[i] : 0;
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31u) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        u64 t = sm[threadIdx.x];
        for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (threadIdx.x == 0) block_sum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) kg_scan_block_sums(u64* block_sum, u64 nblocks, u64* total) {
    __shared__ u64 sm[1024];
    const u64 per = (nblocks + 1023) / 1024;
    const u64 b0 = threadIdx.x * per, b1 = min(b0 + per, nblocks);
    u64 mine = 0;
    for (u64 i = b0; i < b1; i++) mine += block_sum[i];
    sm[threadIdx.x] = mine;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {
        u64 v = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
        __syncthreads();
        sm[threadIdx.x] += v;
        __syncthreads();
    }
    u64 cur = sm[threadIdx.x] - mine;
    for (u64 i = b0; i < b1; i++) { u64 v = block_sum[i]; block_sum[i] = cur; cur += v; }
    if (threadIdx.x == 1023) *total = sm[1023];
}
__global__ void __launch_bounds__(1024) kg_scan_finish(const u32* __restrict__ in, const u64* __restrict__ block_sum,
                                                       u64* __restrict__ out, u64 n) {
    __shared__ u64 sm[1024];
    const u64 i = (u64)blockIdx.x * 1024 + threadIdx.x;
    const u64 v = i < n ? in[i] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {
        u64 t = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
        __syncthreads();
        sm[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = block_sum[blockIdx.x] + sm[threadIdx.x] - v;
}

__device__ __forceinline__ u64 kg_dense_index(const u32* bitmap, const u64* word_prefix, u64 slot) {
    const u64 w = slot >> 5;
    return word_prefix[w] + __popc(bitmap[w] & ((1u << (slot & 31u)) - 1u));
}

// ---- character access on right-aligned multiword keys (no dynamic register indexing) ----------------------------
template <int W>
__device__ __forceinline__ u32 kg_get_char(const u64 (&key)[W], u32 k, u32 pos_from_left) {
    const u32 p = k - 1 - pos_from_left;   // position from the right end
    const u32 word = W - 1 - p / 32, sh = 2 * (p % 32);
    u64 v = 0;
#pragma unroll
    for (int i = 0; i < W; i++) if ((u32)i == word) v = key[i];
    return (u32)(v >> sh) & 3u;
}
template <int W>
__device__ __forceinline__ void kg_set_char(u64 (&key)[W], u32 k, u32 pos_from_left, u32 c) {
    const u32 p = k - 1 - pos_from_left;
    const u32 word = W - 1 - p / 32, sh = 2 * (p % 32);
#pragma unroll
    for (int i = 0; i < W; i++) if ((u32)i == word) key[i] |= (u64)c << sh;
}

// where a k-mer lives: owner shard and table partition follow from its minimizer bucket (kg_skm.cuh), the slot inside
// the partition from its own hash.  nb == 1: one GPU, direct insert, the whole table is one partition.
struct KgPlacement {
    const u64* part_lo;   // [pl + 1]
    u32 pl, nb, m, rank;
};

// ---- build: one thread per table slot ------------------------------------------------------------------------------
// roots == nullptr: only count the roots (sizing pass).  A predecessor that lives on ANOTHER shard (its minimizer
// differs and hashes to a bucket of another owner: a few per cent of the k-mers) cannot be pointed at: the k-mer
// becomes a root, so every shard's structure is self-contained and decodes on its own.
template <int W>
__global__ void __launch_bounds__(256) kg_kaarme_build(KgTable t, u32 k, const u32* __restrict__ bitmap,
                                                       const u64* __restrict__ word_prefix, KgKaarme out,
                                                       u64* root_counter, KgPlacement pm) {
    const u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t.nslots) return;
    const u64* p = t.slots + s * t.stride;
    const u32 n = (u32)p[0];
    if (n == 0 || n == KG_LOCKED) return;
    const KgKGeom g = kg_geom(k);
    u64 C[W];
#pragma unroll
    for (int i = 0; i < W; i++) C[i] = p[1 + i];
    const u64 occ_inv = p[1 + W];
    bool has_pred = false, self_fwd = true, pred_fwd = true;
    u64 pred_slot = ~0ULL;
    if (occ_inv != 0) {
        const u64 occ = ~occ_inv;
        has_pred = (occ >> 3) & 1ULL;
        self_fwd = (occ >> 2) & 1ULL;
        const u32 c_out = (u32)(occ & 3ULL);
        if (has_pred) {
            // occurrence-orientation string S of this k-mer, then the predecessor occurrence c_out + S[0..k-2]
            u64 S[W], R[W];
            if (self_fwd) {
#pragma unroll
                for (int i = 0; i < W; i++) S[i] = C[i];
            } else {
                kg_revcomp<W>(C, S, g);
            }
            u64 P[W];
#pragma unroll
            for (int i = W - 1; i > 0; i--) P[i] = (S[i] >> 2) | (S[i - 1] << 62);
            P[0] = (S[0] >> 2) | ((u64)c_out << (g.topbits - 2));
            kg_revcomp<W>(P, R, g);
            KgKmerWindow<W> w;
#pragma unroll
            for (int i = 0; i < W; i++) { w.f[i] = P[i]; w.r[i] = R[i]; }
            pred_fwd = kg_forward_is_canonical<W>(w);
            u64 key[W];
#pragma unroll
            for (int i = 0; i < W; i++) key[i] = pred_fwd ? P[i] : R[i];
            const u32 bkt = pm.nb > 1 ? kg_key_bucket<W>(key, k, pm.m, pm.nb) : 0u;
            if (bkt / pm.pl != (pm.nb > 1 ? pm.rank : 0u)) {
                has_pred = false;                       // predecessor belongs to another shard -> root
            } else {
                const u64 lo = pm.part_lo[bkt % pm.pl], n_part = pm.part_lo[bkt % pm.pl + 1] - lo;
                pred_slot = kg_table_find<W>(t, key, kg_place(kg_hash_key<W>(key), lo, n_part));
                if (pred_slot == ~0ULL) has_pred = false;   // predecessor was not admitted (Bloom) -> this k-mer is a root
            }
        }
    }
    u64 ptr;
    if (has_pred) {
        ptr = kg_dense_index(bitmap, word_prefix, pred_slot);
    } else {
        ptr = atomicAdd(root_counter, 1ULL);
        if (out.roots) {
#pragma unroll
            for (int i = 0; i < W; i++) out.roots[ptr * W + i] = C[i];
        }
    }
    if (!out.slots) return;
    const u32 cnt = n > 16383u ? 16383u : n;
    const u32 left = kg_get_char<W>(C, k, 0), right = kg_get_char<W>(C, k, k - 1);
    u64 word = (ptr << 26) | ((u64)cnt << 12) | ((u64)left << 10) | ((u64)right << 8) |
               ((u64)pred_fwd << 5) | ((u64)self_fwd << 4) | ((u64)has_pred << 1) | 1ULL;
    out.slots[kg_dense_index(bitmap, word_prefix, s)] = word;
}

// ---- decode: one thread per Kaarme slot (reconstruct_kmer_in_slot, kmer_hash_table.cpp:3848-4058) ------------------
#define KS_PTR(d) ((d) >> 26)
#define KS_COUNT(d) ((u32)((d) >> 12) & 16383u)
#define KS_LEFT(d) ((u32)((d) >> 10) & 3u)
#define KS_RIGHT(d) ((u32)((d) >> 8) & 3u)
#define KS_PRED_FWD(d) ((u32)((d) >> 5) & 1u)
#define KS_SELF_FWD(d) ((u32)((d) >> 4) & 1u)
#define KS_HAS_PRED(d) ((u32)((d) >> 1) & 1u)

template <int W>
__device__ __forceinline__ bool kg_kaarme_decode(const KgKaarme& ks, u32 k, u64 idx, u64 (&key)[W], u64& hops_out) {
#pragma unroll
    for (int i = 0; i < W; i++) key[i] = 0;
    int L = 0, R = (int)k - 1, Lc = 0, Rc = (int)k - 1;
    bool pir = false;
    u64 pos = idx, hops = 0;
    u64 d = ks.slots[pos];
    for (;;) {
        if (!(d & 1ULL)) return false;
        if (!KS_HAS_PRED(d)) break;
        if (L == Lc) { kg_set_char<W>(key, k, (u32)L, pir ? 3u - KS_RIGHT(d) : KS_LEFT(d)); L++; if (L > R) { hops_out = hops; return true; } }
        if (R == Rc) { kg_set_char<W>(key, k, (u32)R, pir ? 3u - KS_LEFT(d) : KS_RIGHT(d)); R--; if (L > R) { hops_out = hops; return true; } }
        const u32 s = KS_SELF_FWD(d), p = KS_PRED_FWD(d);
        int shift = s ? -1 : 1;
        if (pir) shift = -shift;
        Lc += shift; Rc += shift;
        if (s != p) pir = !pir;
        pos = KS_PTR(d);
        if (pos >= ks.n_kmers || ++hops > ks.n_kmers) return false;
        d = ks.slots[pos];
    }
    // root: remaining characters come from the stored string
    const u64 r = KS_PTR(d);
    if (r >= ks.n_roots) return false;
    u64 root[W];
#pragma unroll
    for (int i = 0; i < W; i++) root[i] = ks.roots[r * W + i];
    const int Ls = L - Lc;
    if (!pir) {
        for (int a = L, b = Ls; a <= R; a++, b++) {
            if (b < 0 || b >= (int)k) return false;
            kg_set_char<W>(key, k, (u32)a, kg_get_char<W>(root, k, (u32)b));
        }
    } else {
        for (int a = L, b = (int)k - Ls - 1; a <= R; a++, b--) {
            if (b < 0 || b >= (int)k) return false;
            kg_set_char<W>(key, k, (u32)a, 3u - kg_get_char<W>(root, k, (u32)b));
        }
    }
    hops_out = hops;
    return true;
}

// export from the compact structure: decode every k-mer in [begin, end) whose count passes the threshold
template <int W>
__global__ void __launch_bounds__(256) kg_kaarme_export(KgKaarme ks, u32 k, u64 begin, u64 end, u64 min_abundance,
                                                        u64* __restrict__ out_keys, u32* __restrict__ out_counts, u32* out_n,
                                                        KgCompactStats* cs) {
    const u64 i = begin + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool emit = false;
    u64 key[W];
    u32 cnt = 0;
    if (i < end) {
        const u64 d = ks.slots[i];
        cnt = KS_COUNT(d);
        if (min_abundance > 0 && (u64)cnt >= min_abundance) {
            u64 hops = 0;
            emit = kg_kaarme_decode<W>(ks, k, i, key, hops);
            if (!emit) atomicAdd(&cs->bad, 1ULL);
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
#pragma unroll
        for (int q = 0; q < W; q++) out_keys[(u64)idx * W + q] = key[q];
        out_counts[idx] = cnt;
    }
}

// chain statistics (max / mean hops): decode everything once
template <int W>
__global__ void __launch_bounds__(256) kg_kaarme_chain_stats(KgKaarme ks, u32 k, KgCompactStats* cs) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 hops = 0;
    bool ok = true;
    if (i < ks.n_kmers) {
        u64 key[W];
        ok = kg_kaarme_decode<W>(ks, k, i, key, hops);
    }
    u64 mx = hops, sum = hops;
    for (int d = 16; d; d >>= 1) {
        u64 o = __shfl_xor_sync(0xffffffffu, mx, d);
        mx = o > mx ? o : mx;
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
    }
    if ((threadIdx.x & 31u) == 0) { atomicMax(&cs->max_chain, mx); atomicAdd(&cs->chain_sum, sum); }
    if (!ok) atomicAdd(&cs->bad, 1ULL);
}

// ---- order-independent checksum of the counted k-mers (kg_checksum) ---------------------------------------------------
// out[0] = k-mers with reported count >= min_abundance, out[1] = sum of their counts, out[2] = sum of g(key),
// out[3] = sum of g(key) * count (mod 2^64), g = a 64-bit mix of the key words that is independent of the table hash.
// Sharded runs add the four words over the shards; equal words <=> the same multiset of (k-mer, count) as one table.
template <int W>
__device__ __forceinline__ u64 kg_checksum_mix(const u64 (&key)[W]) {
    u64 h = 0x2545F4914F6CDD1DULL;
#pragma unroll
    for (int i = 0; i < W; i++) h = kg_fmix64(h ^ key[i]) + 0x9E3779B97F4A7C15ULL * (u64)(i + 1);
    return kg_fmix64(h);
}
__device__ __forceinline__ void kg_checksum_commit(u64 n, u64 c, u64 g, u64 gc, u64* out) {
    for (int d = 16; d; d >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, d);
        c += __shfl_xor_sync(0xffffffffu, c, d);
        g += __shfl_xor_sync(0xffffffffu, g, d);
        gc += __shfl_xor_sync(0xffffffffu, gc, d);
    }
    if ((threadIdx.x & 31u) == 0 && n) { atomicAdd(out, n); atomicAdd(out + 1, c); atomicAdd(out + 2, g); atomicAdd(out + 3, gc); }
}
template <int W>
__global__ void __launch_bounds__(256) kg_checksum_table(KgTable table, u64 min_abundance, int count_mode, int table_mode, u64* out) {
    u64 n = 0, c = 0, g = 0, gc = 0;
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < table.nslots; s += (u64)gridDim.x * blockDim.x) {
        const u64* p = table.slots + s * table.stride;
        u64 key[W];
        u64 cnt;
        if (W == 2 && table.packed_tb) {
            key[0] = p[0] & ((1ULL << table.packed_tb) - 1);
            key[W - 1] = p[1];
            cnt = p[0] >> table.packed_tb;
        } else {
            cnt = (u32)p[0];
            if (cnt == KG_LOCKED) cnt = 0;
#pragma unroll
            for (int i = 0; i < W; i++) key[i] = p[1 + i];
        }
        if (cnt == 0) continue;
        const u32 c32 = cnt > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (u32)cnt;
        const u64 rep = count_mode == 0 ? (u64)c32 : (table_mode == 0 ? (u64)(c32 & 0xFFFFu) : (u64)(c32 > 16383u ? 16383u : c32));
        if (min_abundance == 0 || rep < min_abundance) continue;
        const u64 h = kg_checksum_mix<W>(key);
        n++; c += rep; g += h; gc += h * rep;
    }
    kg_checksum_commit(n, c, g, gc, out);
}
template <int W>
__global__ void __launch_bounds__(256) kg_checksum_kaarme(KgKaarme ks, u32 k, u64 min_abundance, u64* out, KgCompactStats* cs) {
    u64 n = 0, c = 0, g = 0, gc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ks.n_kmers; i += (u64)gridDim.x * blockDim.x) {
        const u64 d = ks.slots[i];
        const u64 rep = KS_COUNT(d);
        if (min_abundance == 0 || rep < min_abundance) continue;
        u64 key[W], hops = 0;
        if (!kg_kaarme_decode<W>(ks, k, i, key, hops)) { atomicAdd(&cs->bad, 1ULL); continue; }
        const u64 h = kg_checksum_mix<W>(key);
        n++; c += rep; g += h; gc += h * rep;
    }
    kg_checksum_commit(n, c, g, gc, out);
}
