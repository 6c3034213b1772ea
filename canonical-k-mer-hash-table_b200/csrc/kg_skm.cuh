// kg_skm.cuh -- minimizer ("super-k-mer") bucketing: the L2-blocked / hash-sharded insert without shipping keys.
//
// What the reference does per window is find-or-insert into ONE big table (parallel_parser.hpp:711-789,
// kmer_hash_table.cpp:2207-2567): a random DRAM access per k-mer.  The first B200 version bucketed every canonical key
// (8W bytes per k-mer written, read again by the insert, and sent over NVLink between shards).  Consecutive k-mers of a
// read share k-1 bases, so this version buckets RUNS of windows instead:
//   * owner bucket of a window = hash of its MINIMIZER: the canonical m-mer (m <= 21) with the smallest hash among the
//     k-m+1 m-mers of the window.  A k-mer and its reverse complement contain the same canonical m-mers, so the bucket
//     is a function of the canonical k-mer -- what sharding needs -- and neighbouring windows almost always agree.
//   * a run of consecutive windows of one bucket inside one packed word (<= 32 windows) is ONE 8-byte descriptor
//     {end position of the first window, number of windows, sender, partition, owner}: ~0.7 B per k-mer at k = 51
//     instead of 16 B.  The packed 2-bit stream of the batch (0.25 B per base) stays where it is (one GPU) or is
//     copied once per peer by the copy engines (several GPUs).
//   * bucket b of `nb = world * pl` belongs to shard b / pl and to its table partition p = b % pl, the CONTIGUOUS slot
//     range [part_lo[p], part_lo[p+1]); inside it the slot is a range partition of the k-mer's own hash.  Inserting
//     bucket after bucket keeps the live table region L2-resident, exactly as before.
//   * kg_skm_insert rebuilds each window from the packed stream (a few funnel shifts; stateless, so the windows of a
//     warp's 32 descriptors are spread evenly over its lanes) and does the find-or-insert.  The dependent chain per
//     k-mer is "ALU -> one L2 slot load", no longer "DRAM key load -> L2 slot load".
// Everything position-independent is __host__ __device__ so that tests/native/skm_host.cu runs the same code on the
// CPU against a brute-force model written from the definition.
#pragma once
#include "kg_count.cuh"

#define KG_SKM_TPB 128                       // scatter threads per block = packed words (32 positions each) per block
#define KG_SKM_MAXHALO 8                     // words of m-mer hashes needed to the left of a block: ceil((k-m)/32) <= 8
#define KG_SKM_CHUNKS (KG_SKM_TPB + KG_SKM_MAXHALO)
#define KG_SKM_MAXRUN 32u
#define KG_SKM_SUB 16u                       // sub-regions per bucket: a scatter block appends to sub-region blockIdx % 16, so the
                                             // reservations of a bucket spread over 16 counters in 16 different sectors
#define KG_SKM_META 73728                    // bytes at the head of a batch slot: header (64 B) + dense descriptor counts
#define KG_SKM_MAXSEG (1024 * 16 + 64)        // (partition, sender, sub-region) segments + one overflow list per sender

// minimizer length: short enough that runs are long (k-m+1 m-mers per window), long enough that buckets balance
__host__ __device__ inline u32 kg_skm_m(u32 k) {
    if (k <= 11) return k;
    if (k < 27) return 11;
    return k - 16 > 21 ? 21 : k - 16;
}
__host__ __device__ inline u32 kg_mmer_hash(u64 canon) {
    u64 x = canon * 0x9E3779B97F4A7C15ULL;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    return (u32)(x >> 32);
}
// bucket from the minimum m-mer hash of a window.  The minimum of ~30 uniform values is not uniform, so it goes through
// a bijective scramble (murmur3 fmix32) before the range partition.
__host__ __device__ inline u32 kg_min_to_bucket(u32 minv, u32 nb) {
    u32 h = minv;
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return (u32)(((u64)h * (u64)nb) >> 32);
}

// ---- descriptor ---------------------------------------------------------------------------------------------------
// [63:32] end position (in the sender's packed stream of this batch) of the first window
// [31:26] sender rank   [25:21] windows - 1   [20] the first window has a predecessor window in its read
// [19:10] partition of the owner's table   [9:4] owner rank   [3:0] zero
__host__ __device__ inline u64 kg_skm_desc(u32 j0, u32 src, u32 n, u32 has_pred, u32 part, u32 owner) {
    return ((u64)j0 << 32) | ((u64)src << 26) | ((u64)(n - 1u) << 21) | ((u64)(has_pred & 1u) << 20) | ((u64)part << 10) | ((u64)owner << 4);
}
#define KG_SKM_J0(d) ((u32)((d) >> 32))
#define KG_SKM_SRC(d) ((u32)((d) >> 26) & 63u)
#define KG_SKM_N(d) (((u32)((d) >> 21) & 31u) + 1u)
#define KG_SKM_HP(d) ((u32)((d) >> 20) & 1u)
#define KG_SKM_PART(d) ((u32)((d) >> 10) & 1023u)
#define KG_SKM_OWNER(d) ((u32)((d) >> 4) & 63u)

// ---- window extraction: the k bases ENDING at position e (inclusive) of a packed stream, right-aligned --------------
template <int W>
__host__ __device__ inline void kg_window_at(const u64* __restrict__ words, u32 e, const KgKGeom& g, u64 (&f)[W]) {
    // t = word of the last base, j = bases of that word that belong to the window (1..32).  Branch-free: the part taken
    // from the older word is shifted by 2j in two steps (2j - 1, then 1), so that j = 32 shifts it out entirely.
    const u32 t = e >> 5, j = (e & 31u) + 1u;
    u64 lo = words[t];
#pragma unroll
    for (int i = 0; i < W; i++) {
        const int src = (int)t - 1 - i;
        const u64 hi = src >= 0 ? words[src] : 0ULL;
        f[W - 1 - i] = ((hi << (2u * j - 1u)) << 1) | (lo >> (64u - 2u * j));
        lo = hi;
    }
    f[0] &= g.topmask;
}
// the same from W + 1 packed words held in registers: wreg[i] = word (e >> 5) - i (a descriptor never crosses a packed
// word, so all its windows use the same W + 1 words)
template <int W>
__host__ __device__ inline void kg_window_regs(const u64 (&wreg)[W + 1], u32 e, const KgKGeom& g, u64 (&f)[W]) {
    const u32 j = (e & 31u) + 1u;
#pragma unroll
    for (int i = 0; i < W; i++) f[W - 1 - i] = ((wreg[i + 1] << (2u * j - 1u)) << 1) | (wreg[i] >> (64u - 2u * j));
    f[0] &= g.topmask;
}
template <int W>
__host__ __device__ inline void kg_load_window_words(const u64* __restrict__ words, u32 e, u64 (&wreg)[W + 1]) {
    const u32 t = e >> 5;
#pragma unroll
    for (int i = 0; i <= W; i++) wreg[i] = (int)t - i >= 0 ? words[t - i] : 0ULL;
}
__host__ __device__ inline u32 kg_base_at(const u64* __restrict__ words, u32 pos) {
    return (u32)(words[pos >> 5] >> (62 - 2 * (pos & 31u))) & 3u;
}

#ifndef __CUDA_ARCH__
static inline u64 kg_rev2_host(u64 x) {   // host twin of kg_rev2 (no __brevll on the CPU)
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
}
#endif
__host__ __device__ inline u64 kg_rev2_hd(u64 x) {
#ifdef __CUDA_ARCH__
    return kg_rev2(x);
#else
    return kg_rev2_host(x);
#endif
}
// reverse complement of one right-aligned m-mer (m <= 32)
__host__ __device__ inline u64 kg_revcomp1(u64 f, u32 m) { return (~kg_rev2_hd(f)) >> (64 - 2 * m); }

// ---- scatter, phase 1: m-mer hashes of one packed word (32 end positions) ----------------------------------------------
// The block's arrays are TRANSPOSED: position i of local word lw lives at [i * KG_SKM_CHUNKS + lw], so that the threads of
// a warp (consecutive words, same i) touch consecutive shared-memory banks.  (Word-major -- [lw * 32 + i] -- puts all 32
// lanes on one bank: the first version of this kernel spent its time in 32-way bank conflicts.)
#define KG_SKM_AT(lw, i) ((i) * KG_SKM_CHUNKS + (lw))
// H[(lw,i)] = hash of the canonical m-mer ENDING at position 32*gc + i; sfx[(lw,i)] = min(H[(lw,i..31)]); returns
// min(H[(lw,0..31)]).  Words before the stream (gc < 0) and m-mers that reach before position 0 give 0xFFFFFFFF (no valid
// window holds them).
__host__ __device__ inline u32 kg_skm_hash_word(const u64* __restrict__ words, long long gc, u32 m, u32* H, u32* sfx, u32 lw) {
    if (gc < 0) {
        for (u32 i = 0; i < 32; i++) { H[KG_SKM_AT(lw, i)] = 0xFFFFFFFFu; sfx[KG_SKM_AT(lw, i)] = 0xFFFFFFFFu; }
        return 0xFFFFFFFFu;
    }
    const u64 mmask = m == 32 ? ~0ULL : ((1ULL << (2 * m)) - 1ULL);
    u64 f = gc > 0 ? (words[gc - 1] & mmask) : 0ULL;                  // the m bases before position 32*gc
    u64 r = kg_revcomp1(f, m);
    const u64 w = words[gc];
    const u32 first_full = gc > 0 ? 0u : m - 1u;                       // in word 0 the first m-1 positions hold no m-mer
    for (u32 i = 0; i < 32; i++) {
        const u64 c = (w >> (62 - 2 * i)) & 3ULL;
        f = ((f << 2) | c) & mmask;
        r = (r >> 2) | ((3ULL - c) << (2 * m - 2));
        H[KG_SKM_AT(lw, i)] = i >= first_full ? kg_mmer_hash(f < r ? f : r) : 0xFFFFFFFFu;
    }
    u32 run = 0xFFFFFFFFu;
    for (int i = 31; i >= 0; i--) { const u32 h = H[KG_SKM_AT(lw, i)]; run = h < run ? h : run; sfx[KG_SKM_AT(lw, i)] = run; }
    return run;
}

// ---- scatter, phase 2: windows of one packed word -> descriptors ------------------------------------------------------
// H / sfx (transposed, KG_SKM_AT) / cmin are indexed by LOCAL word: the caller passes the arrays of its block (halo words
// first) and the local index lw of the word being processed.  emit(bucket, j0, n, has_pred).
__host__ __device__ inline u32 kg_ctz32(u32 x) {       // x != 0
#ifdef __CUDA_ARCH__
    return (u32)__ffs((int)x) - 1u;
#else
    u32 n = 0;
    while (!((x >> n) & 1u)) n++;
    return n;
#endif
}

// The walk over the 32 positions only RECORDS where descriptors start (bit masks; their buckets in a small register
// queue); the descriptors are emitted afterwards, one loop iteration per descriptor.  (Emitting inside the walk made
// nearly every one of its 32 iterations pay for the emit path -- an atomic and a store -- on behalf of one or two lanes.)
template <typename E>
__host__ __device__ inline u32 kg_skm_segment_word(const u64* __restrict__ words, const u32* __restrict__ brk, u32 T, u32 C,
                                                   u32 k, u32 m, u32 nb, u32 gw, const u32* H, const u32* sfx, const u32* cmin,
                                                   u32 lw, u32 kwords, E&& emit) {
    (void)words;
    if ((u64)gw * 32u >= T) return 0;
    const u32 wlen = k - m + 1u;                                       // m-mers per window
    const u32 mybrk = brk[gw];
    // run length at the base just before this word
    u32 run = 0;
    for (u32 i = 1; i <= kwords + 1u; i++) {
        if (gw < i) break;                                             // position 0 always carries a break bit
        const u32 b = brk[gw - i];
        if (b) { run += kg_ctz32(b) + 1u; break; }                     // lowest set bit = most recent run start of that word
        run += 32u;
    }
    const u32 jend = T - gw * 32u < 32u ? T - gw * 32u : 32u;
    const u32 base = lw * 32u;                                         // local position of this word's first base
    // bit j of vmask = a counted window ends at position j, of smask = a descriptor starts there, of pmask = that window
    // has a predecessor window in its read; pend = starts not emitted yet, their buckets queued in bq (10 bits each)
    u32 vmask = 0, smask = 0, pmask = 0, pend = 0, nq = 0, n_windows = 0;
    u64 bq = 0;
    auto flush = [&]() {
        u32 todo = pend, idx = 0;
        while (todo) {
            const u32 j = kg_ctz32(todo);
            todo &= todo - 1u;
            const u32 stop = (smask | ~vmask) & ~((2u << j) - 1u);     // next start, or first position without a window, above j
            const u32 end = stop ? kg_ctz32(stop) : 32u;
            const u32 b = (u32)(bq >> (10u * (nq - 1u - idx))) & 1023u;
            idx++;
            emit(b, gw * 32u + j, end - j, (pmask >> j) & 1u);
            n_windows += end - j;
        }
        pend = 0; nq = 0; bq = 0;
    };
    u32 own = 0xFFFFFFFFu, prev_b = 0xFFFFFFFFu;
    for (u32 j = 0; j < jend; j++) {
        run = ((mybrk >> (31u - j)) & 1u) ? 1u : run + 1u;
        const u32 h = H[KG_SKM_AT(lw, j)];
        own = h < own ? h : own;                                       // min over this word's m-mers up to j
        const u32 pos = gw * 32u + j;
        if (run >= k && pos >= C) {
            u32 mv;
            if (j + 1u >= wlen) {                                      // all m-mers of the window end inside this word
                mv = 0xFFFFFFFFu;
                for (u32 i = j + 1u - wlen; i <= j; i++) { const u32 v = H[KG_SKM_AT(lw, i)]; mv = v < mv ? v : mv; }
            } else {                                                   // head in earlier words + own prefix
                const u32 s = base + j + 1u - wlen;                    // local position of the first m-mer end (>= 0 by the halo)
                const u32 sv = sfx[KG_SKM_AT(s >> 5, s & 31u)];
                mv = sv < own ? sv : own;
                for (u32 cw = (s >> 5) + 1u; cw < lw; cw++) { const u32 v = cmin[cw]; mv = v < mv ? v : mv; }
            }
            const u32 b = kg_min_to_bucket(mv, nb);
            if (b != prev_b) {                                         // a descriptor starts here
                if (nq == 6u) flush();                                 // (rare) queue full: everything pending ends before j
                smask |= 1u << j;
                pend |= 1u << j;
                bq = (bq << 10) | (u64)b;
                nq++;
                if (run > k) pmask |= 1u << j;
            }
            vmask |= 1u << j;
            prev_b = b;
        } else {
            prev_b = 0xFFFFFFFFu;
        }
    }
    flush();
    return n_windows;
}

// minimizer bucket of an arbitrary k-mer given as a right-aligned key (Kaarme build: where does my predecessor live?)
template <int W>
__host__ __device__ inline u32 kg_key_bucket(const u64 (&key)[W], u32 k, u32 m, u32 nb) {
    const u64 mmask = m == 32 ? ~0ULL : ((1ULL << (2 * m)) - 1ULL);
    u64 f = 0, r = 0;
    u32 mv = 0xFFFFFFFFu;
    for (u32 i = 0; i < k; i++) {
        const u32 p = k - 1u - i, word = (u32)W - 1u - p / 32u, sh = 2u * (p % 32u);
        u64 v = 0;
#pragma unroll
        for (int q = 0; q < W; q++) if ((u32)q == word) v = key[q];
        const u64 c = (v >> sh) & 3ULL;
        f = ((f << 2) | c) & mmask;
        r = (r >> 2) | ((3ULL - c) << (2 * m - 2));
        if (i + 1u >= m) { const u32 h = kg_mmer_hash(f < r ? f : r); mv = h < mv ? h : mv; }
    }
    return kg_min_to_bucket(mv, nb);
}

#ifdef __CUDACC__
// ---- scatter kernel ---------------------------------------------------------------------------------------------------
struct KgSkmScatterArgs {
    const u64* words;
    const u32* brk;
    const KgStream* st;
    u32* cursors;       // [(nb * KG_SKM_SUB + 1) * 8]: one reservation counter per (bucket, sub-region), each in its own
                        // 32-byte sector; the last one is the overflow list's
    u64* regions;       // nb * KG_SKM_SUB sub-regions of cap descriptors
    u64* ovf;           // overflow list (descriptors whose region was full), ovf_cap entries
    u64* hdr;           // [0] = global ordinal of position 0 of this batch (KgStream::bases_seen), [1] = T
    KgStats* stats;
    u32 k, m, nb, pl, cap, src;
    u32 ovf_cap;
    u32 nwords;         // packed words this launch covers (upper bound on ceil(T/32))
};

__global__ void __launch_bounds__(KG_SKM_TPB) kg_skm_scatter(KgSkmScatterArgs a) {
    __shared__ u32 sH[KG_SKM_CHUNKS * 32];
    __shared__ u32 sS[KG_SKM_CHUNKS * 32];
    __shared__ u32 sM[KG_SKM_CHUNKS];
    const u32 T = a.st->total_bases, C = a.st->carry_bases;
    const u32 tid = threadIdx.x;
    const u32 halo = (a.k - a.m + 31u) / 32u;                           // words of history a window's m-mers can reach into
    const long long first = (long long)blockIdx.x * KG_SKM_TPB - (long long)halo;   // global word of local word 0
    if (blockIdx.x == 0 && tid == 0) { a.hdr[0] = a.st->bases_seen; a.hdr[1] = T; }
    if ((u64)blockIdx.x * KG_SKM_TPB * 32u >= T) return;                // block-uniform: nothing to do, no barrier below is reached by anyone
    for (u32 lw = tid; lw < KG_SKM_TPB + halo; lw += KG_SKM_TPB) {
        const long long gc = first + lw;
        u32 mn = 0xFFFFFFFFu;
        if (gc < (long long)a.nwords) mn = kg_skm_hash_word(a.words, gc, a.m, sH, sS, lw);
        else for (u32 i = 0; i < 32; i++) { sH[KG_SKM_AT(lw, i)] = 0xFFFFFFFFu; sS[KG_SKM_AT(lw, i)] = 0xFFFFFFFFu; }
        sM[lw] = mn;
    }
    __syncthreads();
    const u32 gw = blockIdx.x * KG_SKM_TPB + tid;
    u32 n_windows = 0;
    if (gw < a.nwords) {
        const u32 kwords = (a.k + 31u) / 32u;
        n_windows = kg_skm_segment_word(a.words, a.brk, T, C, a.k, a.m, a.nb, gw, sH, sS, sM, tid + halo, kwords,
                                        [&](u32 b, u32 j0, u32 n, u32 hp) {
                                            const u64 d = kg_skm_desc(j0, a.src, n, hp, b % a.pl, b / a.pl);
                                            const u32 r = b * KG_SKM_SUB + (blockIdx.x & (KG_SKM_SUB - 1u));
                                            const u32 idx = atomicAdd(&a.cursors[r * 8u], 1u);
                                            if (idx < a.cap) a.regions[(u64)r * a.cap + idx] = d;
                                            else {
                                                const u32 o = atomicAdd(&a.cursors[a.nb * KG_SKM_SUB * 8u], 1u);
                                                if (o < a.ovf_cap) a.ovf[o] = d;
                                                else a.stats->table_full = 2;   // cannot happen: ovf_cap = every position of a batch
                                            }
                                        });
    }
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
}

// dense descriptor counts of a slot, written next to its header where the peers pull them from:
// counts[r] = descriptors in sub-region r (r < nb * KG_SKM_SUB), counts[nb * KG_SKM_SUB] = descriptors in the overflow list
__global__ void __launch_bounds__(256) kg_skm_pack_counts(const u32* __restrict__ cursors, u32 nregions, u32 cap, u32 ovf_cap,
                                                          u32* __restrict__ counts) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nregions) counts[i] = min(cursors[i * 8u], cap);
    else if (i == nregions) counts[i] = min(cursors[i * 8u], ovf_cap);
}

// Where the receiver finds every sender's descriptor counts (local copy) and descriptors (in place: own slot, or a peer's
// slot mapped over NVLink).
struct KgSkmPeers {
    const u32* counts[64];
    const u64* desc[64];
};

// Segment table of one round on owner `me`: the descriptors of my partitions, PARTITION-MAJOR across senders and their
// sub-regions (so one insert launch still walks my table region by region), followed by every sender's overflow list (not
// owner-sorted: the insert skips what belongs to other owners).
// seg id = (p * world + s) * KG_SKM_SUB + sub for regions, pl * world * KG_SKM_SUB + s for the overflow lists.
__global__ void __launch_bounds__(1024) kg_skm_segments(const KgSkmPeers* __restrict__ peers, u32 world, u32 me, u32 pl, u32 nb,
                                                        u32 cap, u64* __restrict__ seg_start, const u64** __restrict__ seg_ptr) {
    __shared__ u64 sm[1024];
    const u32 nreg = pl * world * KG_SKM_SUB, nseg = nreg + world, tid = threadIdx.x;
    const u32 per = (nseg + 1023u) / 1024u;
    const u32 i0 = tid * per, i1 = min(i0 + per, nseg);
    u64 mine = 0;
    for (u32 i = i0; i < i1; i++) {
        if (i < nreg) {
            const u32 sub = i % KG_SKM_SUB, ps = i / KG_SKM_SUB, p = ps / world, s = ps % world;
            mine += (u64)peers->counts[s][(me * pl + p) * KG_SKM_SUB + sub];
        } else {
            mine += (u64)peers->counts[i - nreg][nb * KG_SKM_SUB];
        }
    }
    sm[tid] = mine;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {
        const u64 t = tid >= d ? sm[tid - d] : 0;
        __syncthreads();
        sm[tid] += t;
        __syncthreads();
    }
    u64 cur = sm[tid] - mine;
    for (u32 i = i0; i < i1; i++) {
        u64 n;
        const u64* ptr;
        if (i < nreg) {
            const u32 sub = i % KG_SKM_SUB, ps = i / KG_SKM_SUB, p = ps / world, s = ps % world;
            const u32 r = (me * pl + p) * KG_SKM_SUB + sub;
            n = (u64)peers->counts[s][r];
            ptr = peers->desc[s] + (u64)r * cap;
        } else {
            const u32 s = i - nreg;
            n = (u64)peers->counts[s][nb * KG_SKM_SUB];
            ptr = peers->desc[s] + (u64)nb * KG_SKM_SUB * cap;
        }
        seg_start[i] = cur;
        seg_ptr[i] = ptr;
        cur += n;
    }
    if (tid == 1023) seg_start[nseg] = sm[1023];
}

// ---- insert kernel ------------------------------------------------------------------------------------------------------
struct KgSkmSources {                   // one entry per sender rank (one GPU: entry 0)
    const u64* words[64];
    const u64* hdr[64];                 // hdr[r][0] = global ordinal of position 0 of rank r's batch
};
struct KgSkmInsertArgs {
    const u64* seg_start;               // [nseg + 1] logical index of the first descriptor of each segment
    const u64* const* seg_ptr;          // [nseg] where the segment's descriptors are (local, or a peer's slot over NVLink)
    u32 nseg;
    u32 nparts, segs_per_part;          // the first nparts * segs_per_part segments are the partitions' regions, in table order
    u32 my_rank;                        // descriptors of another owner are skipped (overflow lists are not owner-sorted)
    const KgSkmSources* src;
    const u64* part_lo;                 // [pl + 1] first table slot of each partition
    const u64* bpart_lo;                // [pl + 1] first Bloom word of each partition
    KgTable table;
    KgBloom bloom;
    KgStats* stats;
    u32* work;
    u32 k;
};

// Probe queue: unresolved probes of a warp wait in shared memory until 32 of them are together.  A probe sequence of a
// k-mer takes 1.4 steps on average but 4+ for the unluckiest of 32 lanes; looping per window until every lane is done ran
// the probe code at a third of a warp's width (45 % of this kernel's instructions).  Instead every window gets ONE probe
// with the whole warp; what is not decided is parked -- key, next slot, probes so far, occurrence record -- and the queue is
// drained 32 entries at a time, again one probe each at full width.  The queue lives across the warp's groups (entries
// hold absolute slot pointers); only the end of the kernel drains it to the last entry.
#define KG_SKM_QCAP 64u

template <int W, int SINK, int MINB>
__global__ void __launch_bounds__(256, MINB) kg_skm_insert(KgSkmInsertArgs a) {
    __shared__ u32 sm[8];
    __shared__ u64 s_desc[8][32];       // the 32 descriptors of the warp's current group
    __shared__ u32 s_excl[8][33];       // exclusive prefix of their window counts ([32] = total)
    __shared__ u64 s_qkey[SINK == KG_SINK_BLOOM1 ? 1 : 8][SINK == KG_SINK_BLOOM1 ? 1 : KG_SKM_QCAP][W];
    __shared__ u64 s_qptr[SINK == KG_SINK_BLOOM1 ? 1 : 8][SINK == KG_SINK_BLOOM1 ? 1 : KG_SKM_QCAP];
    __shared__ u64 s_qocc[SINK == KG_SINK_BLOOM1 ? 1 : 8][SINK == KG_SINK_BLOOM1 ? 1 : KG_SKM_QCAP];
    __shared__ u32 s_qcnt[SINK == KG_SINK_BLOOM1 ? 1 : 8][SINK == KG_SINK_BLOOM1 ? 1 : KG_SKM_QCAP];
    const u64 n_desc = a.seg_start[a.nseg];
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    const KgKGeom g = kg_geom(a.k);
    const u32 max_probe = (u32)(a.table.nslots < KG_MAX_PROBE ? a.table.nslots : KG_MAX_PROBE);
    KgConsume<W, SINK> sink;
    sink.init(a.table, a.bloom);
    u32 qn = 0;                         // entries in this warp's queue (the same value in every lane)

    // one probe for the lanes with `live`; counts what is decided, parks what is not
    auto probe_round = [&](bool live, u64 (&key)[W], u64* p, u32 cnt, u64 occw) {
        int st = -1;
        if (live) {
            st = kg_probe_step<W>(a.table, key, p);
            if (st != KG_PROBE_SEARCH) {
                sink.n_ins++;
                sink.n_new += st == KG_PROBE_NEW ? 1u : 0u;
                if (a.table.kaarme) atomicMax(p + 1 + W, ~occw);     // earliest occurrence (table starts zeroed)
            } else if (++cnt >= max_probe || ((cnt & 255u) == 0 && kg_ld_u32(a.table.full_flag))) {
                sink.full = true;                                    // probe budget exhausted: the table is full
                st = KG_PROBE_FAIL;
            }
        }
        const u32 park = __ballot_sync(0xffffffffu, st == KG_PROBE_SEARCH);
        if (st == KG_PROBE_SEARCH) {
            const u32 at = qn + __popc(park & lt);
#pragma unroll
            for (int x = 0; x < W; x++) s_qkey[warp][at][x] = key[x];
            s_qptr[warp][at] = (u64)(uintptr_t)p;
            s_qocc[warp][at] = occw;
            s_qcnt[warp][at] = cnt;
        }
        qn += __popc(park);
        __syncwarp();
    };
    auto drain = [&](u32 keep) {        // one probe for the youngest 32 entries, until at most `keep` are left
        while (qn > keep) {
            const u32 take = qn < 32u ? qn : 32u;
            const u32 at = qn - take + lane;
            const bool live = lane < take;
            u64 key[W];
            u64* p = nullptr;
            u32 cnt = 0;
            u64 occw = 0;
            if (live) {
#pragma unroll
                for (int x = 0; x < W; x++) key[x] = s_qkey[warp][at][x];
                p = (u64*)(uintptr_t)s_qptr[warp][at];
                occw = s_qocc[warp][at];
                cnt = s_qcnt[warp][at];
            }
            __syncwarp();
            qn -= take;
            probe_round(live, key, p, cnt, occw);
        }
    };

    for (;;) {
        // Persistent warps pull groups of 32 descriptors (a few hundred windows) from one global counter: whatever their
        // relative speed, the warps in flight work at the FRONT of the partition-major descriptor array, so the table
        // region they hit stays L2-resident.  (Claiming 128 descriptors per warp put ~13 M windows = a dozen partitions
        // in flight and the table went to DRAM for every second probe.)
        u32 claim = 0;
        if (lane == 0) claim = atomicAdd(a.work, 1u);
        claim = __shfl_sync(0xffffffffu, claim, 0);
        const u64 first = (u64)claim * 32u;
        if (first >= n_desc) break;
        const u64 i = first + lane;
        u64 d = 0;
        u32 n = 0;
        u32 seg = 0;
        {   // segment of the group's first descriptor: the same search in every lane (uniform loads), then each lane walks on
            u32 hi = a.nseg;
            while (hi - seg > 1) { const u32 mid = (seg + hi) >> 1; if (a.seg_start[mid] <= first) seg = mid; else hi = mid; }
        }
        // Prefetch into L2 the slice of the NEXT partition's table region that corresponds to this group's place in its
        // own partition.  A batch touches every sector of a region about twice, so half of the probes would be first
        // touches served from DRAM; by the time the walk reaches the next partition its region is L2-resident already.
        if (SINK != KG_SINK_BLOOM1 && a.segs_per_part && seg < a.nparts * a.segs_per_part) {
            const u32 part = seg / a.segs_per_part;
            if (part + 1 < a.nparts) {
                const u64 p0 = a.seg_start[part * a.segs_per_part], p1 = a.seg_start[(part + 1) * a.segs_per_part];
                const u32 ngrp = (u32)((p1 - p0 + 31u) >> 5), j = (u32)((first - p0) >> 5);
                const u64 s0 = __ldg(a.part_lo + part + 1), s1 = __ldg(a.part_lo + part + 2);
                const u32 lines = (u32)(((s1 - s0) * a.table.stride * 8u + 127u) >> 7);      // 128-byte lines of the region
                const u32 per_grp = lines / ngrp + 1u;
                const char* base = (const char*)(a.table.slots + s0 * a.table.stride);
                const u32 l1 = min(lines, (j + 1u) * per_grp);
                for (u32 l = j * per_grp + lane; l < l1; l += 32u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (u64)l * 128u));
            }
        }
        if (i < n_desc) {
            while (a.seg_start[seg + 1] <= i) seg++;  // (i < n_desc = seg_start[nseg]: stops inside the table)
            d = __ldcs(a.seg_ptr[seg] + (i - a.seg_start[seg]));
            n = KG_SKM_OWNER(d) == a.my_rank ? KG_SKM_N(d) : 0u;
        }
        u32 incl = n;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= (u32)s) incl += o; }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        s_desc[warp][lane] = d;
        s_excl[warp][lane] = incl - n;
        if (lane == 31) s_excl[warp][32] = total;
        __syncwarp();
        // every lane takes a CONTIGUOUS share of the group's windows: equal work for all lanes however the windows are
        // spread over the descriptors, and the walk from one descriptor to the next is incremental (no search per window)
        const u32 per = (total + 31u) >> 5;
        u32 gi = lane * per;
        const u32 gend = min(total, gi + per);
        u32 q = 0, e = 0, next = 0;
        u64 dq = 0;
        const u64* __restrict__ words = nullptr;
        u64 wreg[W + 1];                                             // the packed words the current descriptor's windows use
        u64 t_lo = 0, t_n = 0, b_lo = 0, b_n = 0;                    // its partition's slot range / Bloom block range
        auto enter = [&]() {                                         // per-descriptor state (not per window)
            words = a.src->words[KG_SKM_SRC(dq)];
            kg_load_window_words<W>(words, e, wreg);
            const u32 part = KG_SKM_PART(dq);
            if (SINK != KG_SINK_BLOOM1) { t_lo = __ldg(a.part_lo + part); t_n = __ldg(a.part_lo + part + 1) - t_lo; }
            if (SINK != KG_SINK_TABLE) { b_lo = __ldg(a.bpart_lo + part); b_n = __ldg(a.bpart_lo + part + 1) - b_lo; }
        };
#pragma unroll
        for (int x = 0; x <= W; x++) wreg[x] = 0;
        if (gi < gend) {
            // descriptor holding window gi: largest q with excl[q] <= gi (it has windows: excl[q + 1] > gi)
#pragma unroll
            for (u32 s = 16; s; s >>= 1) if (s_excl[warp][q + s] <= gi) q += s;
            dq = s_desc[warp][q];
            e = KG_SKM_J0(dq) + (gi - s_excl[warp][q]);              // end position of window gi
            next = s_excl[warp][q + 1];                              // first window of the next descriptor
            enter();
        }
#pragma unroll 1
        for (u32 it = 0; it < per; it++) {                           // the same trip count in every lane (collectives inside)
            bool live = gi < gend;
            u64 key[W];
            u64* p = nullptr;
            KgOcc occ; occ.word = ~0ULL;
            if (live) {
                KgKmerWindow<W> win;
                kg_window_regs<W>(wreg, e, g, win.f);
                kg_revcomp<W>(win.f, win.r, g);
                const bool fwd = kg_forward_is_canonical<W>(win);
#pragma unroll
                for (int x = 0; x < W; x++) key[x] = fwd ? win.f[x] : win.r[x];
                const u64 h = kg_hash_key<W>(key);
                if (SINK == KG_SINK_BLOOM1 || SINK == KG_SINK_BLOOM2) {
                    if (SINK == KG_SINK_BLOOM1) {
                        kg_bloom_insert(a.bloom, h, kg_place(h, b_lo, b_n), sink.n_b1, sink.n_b2);
                        live = false;
                    } else if (!kg_bloom_admits(a.bloom, h, kg_place(h, b_lo, b_n))) {
                        sink.n_rej++;
                        live = false;
                    }
                }
                if (SINK != KG_SINK_BLOOM1 && live) {
                    p = a.table.slots + kg_place(h, t_lo, t_n) * a.table.stride;
                    if (a.table.kaarme) {
                        const u32 srcr = KG_SKM_SRC(dq);
                        const bool hp = e > KG_SKM_J0(dq) || KG_SKM_HP(dq);
                        const u32 c_out = hp ? kg_base_at(words, e - a.k) : 0u;
                        occ = kg_make_occ(((u64)srcr << 48) | (a.src->hdr[srcr][0] + (u64)e), hp, fwd, c_out);
                    }
                }
                // on to my next window
                if (++gi < gend) {
                    if (gi == next) {                                // next descriptor that holds windows
                        do { q++; next = s_excl[warp][q + 1]; } while (next == gi);
                        dq = s_desc[warp][q];
                        e = KG_SKM_J0(dq);
                        enter();
                    } else {
                        e++;
                    }
                }
            }
            if (SINK != KG_SINK_BLOOM1) {
                probe_round(live, key, p, 0u, occ.word);
                if (qn > KG_SKM_QCAP - 32u) drain(KG_SKM_QCAP - 64u + 31u);    // keep room for the next round's 32
            }
        }
        __syncwarp();                                               // the group's shared arrays are reused by the next claim
    }
    if (SINK != KG_SINK_BLOOM1) drain(0u);                           // what is still parked when the work runs out
    if (SINK == KG_SINK_TABLE || SINK == KG_SINK_BLOOM2) {
        kg_block_add(sink.n_ins, &a.stats->inserted, sm);
        kg_block_add(sink.n_new, &a.stats->distinct, sm);
    }
    if (SINK == KG_SINK_BLOOM1) {
        kg_block_add(sink.n_b1, &a.stats->new_in_first, sm);
        kg_block_add(sink.n_b2, &a.stats->new_in_second, sm);
    }
    if (SINK == KG_SINK_BLOOM2) kg_block_add(sink.n_rej, &a.stats->bloom_rejected, sm);
    if (sink.full) a.stats->table_full = 1;
}
#endif  // __CUDACC__
