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
#define KG_SKM_META 8192                     // bytes at the head of a batch slot: header (64 B) + bucket cursors
#define KG_SKM_MAXSEG (1024 + 64)             // partition-major segments across senders + one overflow list per sender

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
    const u32 q = e + 1u, t = q >> 5, j0 = q & 31u;
    if (j0 == 0) {
#pragma unroll
        for (int i = 0; i < W; i++) {
            const int src = (int)t - 1 - i;
            f[W - 1 - i] = src >= 0 ? words[src] : 0ULL;
        }
    } else {
        const u32 s = 64 - 2 * j0;
        u64 lo = words[t];
#pragma unroll
        for (int i = 0; i < W; i++) {
            const int src = (int)t - 1 - i;
            const u64 hi = src >= 0 ? words[src] : 0ULL;
            f[W - 1 - i] = (hi << (64 - s)) | (lo >> s);
            lo = hi;
        }
    }
    f[0] &= g.topmask;
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
// H[i] = hash of the canonical m-mer ENDING at position 32*gc + i; sfx[i] = min(H[i..31]); returns min(H[0..31]).
// Words before the stream (gc < 0) and m-mers that reach before position 0 give 0xFFFFFFFF (no valid window holds them).
__host__ __device__ inline u32 kg_skm_hash_word(const u64* __restrict__ words, long long gc, u32 m, u32* H, u32* sfx) {
    if (gc < 0) {
        for (int i = 0; i < 32; i++) { H[i] = 0xFFFFFFFFu; sfx[i] = 0xFFFFFFFFu; }
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
        H[i] = i >= first_full ? kg_mmer_hash(f < r ? f : r) : 0xFFFFFFFFu;
    }
    u32 run = 0xFFFFFFFFu;
    for (int i = 31; i >= 0; i--) { run = H[i] < run ? H[i] : run; sfx[i] = run; }
    return run;
}

// ---- scatter, phase 2: windows of one packed word -> descriptors ------------------------------------------------------
// H / sfx / cmin are indexed by LOCAL word (local word lw covers positions 32*(gw - lw_of_gw) ...): the caller passes the
// arrays of its block (halo words first) and the local index of the word being processed.  emit(bucket, j0, n, has_pred).
template <typename E>
__host__ __device__ inline u32 kg_skm_segment_word(const u64* __restrict__ words, const u32* __restrict__ brk, u32 T, u32 C,
                                                   u32 k, u32 m, u32 nb, u32 gw, const u32* H, const u32* sfx, const u32* cmin,
                                                   u32 lw, u32 kwords, E&& emit) {
    if ((u64)gw * 32u >= T) return 0;
    const u32 wlen = k - m + 1u;                                       // m-mers per window
    const u32 mybrk = brk[gw];
    // run length at the base just before this word
    u32 run = 0;
    for (u32 i = 1; i <= kwords + 1u; i++) {
        if (gw < i) break;                                             // position 0 always carries a break bit
        const u32 b = brk[gw - i];
        if (b) {
            u32 low = 0;
            while (!((b >> low) & 1u)) low++;                          // lowest set bit = most recent run start of that word
            run += low + 1u;
            break;
        }
        run += 32u;
    }
    const u32 jend = T - gw * 32u < 32u ? T - gw * 32u : 32u;
    const u32 base = lw * 32u;                                         // local position of this word's first base
    u32 n_windows = 0, cur_n = 0, cur_b = 0, cur_j0 = 0, cur_hp = 0, own = 0xFFFFFFFFu;
    for (u32 j = 0; j < jend; j++) {
        run = ((mybrk >> (31u - j)) & 1u) ? 1u : run + 1u;
        const u32 h = H[base + j];
        own = h < own ? h : own;                                       // min over this word's m-mers up to j
        const u32 pos = gw * 32u + j;
        if (run >= k && pos >= C) {
            u32 mv;
            if (j + 1u >= wlen) {                                      // all m-mers of the window end inside this word
                mv = 0xFFFFFFFFu;
                for (u32 i = j + 1u - wlen; i <= j; i++) { const u32 v = H[base + i]; mv = v < mv ? v : mv; }
            } else {                                                   // head in earlier words + own prefix
                const u32 s = base + j + 1u - wlen;                    // local position of the first m-mer end (base + j >= wlen - 1 by halo)
                mv = sfx[s] < own ? sfx[s] : own;
                for (u32 cw = (s >> 5) + 1u; cw < lw; cw++) { const u32 v = cmin[cw]; mv = v < mv ? v : mv; }
            }
            const u32 b = kg_min_to_bucket(mv, nb);
            n_windows++;
            if (cur_n && b == cur_b) {
                cur_n++;
            } else {
                if (cur_n) emit(cur_b, cur_j0, cur_n, cur_hp);
                cur_b = b; cur_j0 = pos; cur_n = 1; cur_hp = run > k ? 1u : 0u;
            }
        } else if (cur_n) {
            emit(cur_b, cur_j0, cur_n, cur_hp);
            cur_n = 0;
        }
    }
    if (cur_n) emit(cur_b, cur_j0, cur_n, cur_hp);
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
    u32* cursors;       // [nb + 1] descriptors reserved per bucket region (may run past cap); [nb] = overflow cursor
    u64* regions;       // nb regions of cap descriptors
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
        if (gc < (long long)a.nwords) mn = kg_skm_hash_word(a.words, gc, a.m, sH + lw * 32u, sS + lw * 32u);
        else for (int i = 0; i < 32; i++) { sH[lw * 32u + i] = 0xFFFFFFFFu; sS[lw * 32u + i] = 0xFFFFFFFFu; }
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
                                            const u32 idx = atomicAdd(&a.cursors[b], 1u);
                                            if (idx < a.cap) a.regions[(u64)b * a.cap + idx] = d;
                                            else {
                                                const u32 o = atomicAdd(&a.cursors[a.nb], 1u);
                                                if (o < a.ovf_cap) a.ovf[o] = d;
                                                else a.stats->table_full = 2;   // cannot happen: ovf_cap = every position of a batch
                                            }
                                        });
    }
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
}

// Where the receiver finds every sender's bucket cursors (local copy) and descriptors (in place: own slot, or a peer's slot
// mapped over NVLink).
struct KgSkmPeers {
    const u32* cursors[64];
    const u64* desc[64];
};

// Segment table of one round on owner `me`: the descriptors of my partitions, PARTITION-MAJOR across senders (so one
// insert launch still walks my table region by region), followed by every sender's overflow list (not owner-sorted:
// the insert skips what belongs to other owners).  seg id = p * world + s for regions, pl * world + s for overflow.
__global__ void __launch_bounds__(1024) kg_skm_segments(const KgSkmPeers* __restrict__ peers, u32 world, u32 me, u32 pl, u32 nb,
                                                        u32 cap, u32 ovf_cap, u64* __restrict__ seg_start,
                                                        const u64** __restrict__ seg_ptr) {
    __shared__ u64 sm[1024];
    const u32 nseg = pl * world + world, tid = threadIdx.x;
    const u32 per = (nseg + 1023u) / 1024u;
    const u32 i0 = tid * per, i1 = min(i0 + per, nseg);
    u64 mine = 0;
    for (u32 i = i0; i < i1; i++) {
        if (i < pl * world) { const u32 p = i / world, s = i % world; mine += (u64)min(peers->cursors[s][me * pl + p], cap); }
        else { const u32 s = i - pl * world; mine += (u64)min(peers->cursors[s][nb], ovf_cap); }
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
        if (i < pl * world) {
            const u32 p = i / world, s = i % world;
            n = (u64)min(peers->cursors[s][me * pl + p], cap);
            ptr = peers->desc[s] + (u64)(me * pl + p) * cap;
        } else {
            const u32 s = i - pl * world;
            n = (u64)min(peers->cursors[s][nb], ovf_cap);
            ptr = peers->desc[s] + (u64)nb * cap;
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

#define KG_SKM_CLAIM 4u                 // groups of 32 descriptors a warp claims with one atomic

template <int W, int SINK>
__global__ void __launch_bounds__(256) kg_skm_insert(KgSkmInsertArgs a) {
    __shared__ u32 sm[8];
    const u64 n_desc = a.seg_start[a.nseg];
    const u32 lane = threadIdx.x & 31u;
    const KgKGeom g = kg_geom(a.k);
    KgConsume<W, SINK> sink;
    sink.init(a.table, a.bloom);
    for (;;) {
        // persistent warps pull groups from one global counter: whatever their relative speed, the warps in flight
        // work at the FRONT of the (partition-major) descriptor array, so the live table region stays L2-resident
        u32 claim = 0;
        if (lane == 0) claim = atomicAdd(a.work, 1u);
        claim = __shfl_sync(0xffffffffu, claim, 0);
        const u64 first = (u64)claim * (KG_SKM_CLAIM * 32u);
        if (first >= n_desc) break;
        // segment of the first descriptor of the claim (one binary search per claim, done redundantly by every lane)
        u32 seg;
        {
            u32 lo = 0, hi = a.nseg;
            while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (a.seg_start[mid] <= first) lo = mid; else hi = mid; }
            seg = lo;
        }
#pragma unroll 1
        for (u32 grp = 0; grp < KG_SKM_CLAIM; grp++) {
            const u64 i = first + (u64)grp * 32u + lane;
            u64 d = 0;
            u32 n = 0;
            if (i < n_desc) {
                while (seg + 1 < a.nseg && a.seg_start[seg + 1] <= i) seg++;
                d = __ldcs(a.seg_ptr[seg] + (i - a.seg_start[seg]));
                n = KG_SKM_OWNER(d) == a.my_rank ? KG_SKM_N(d) : 0u;
            }
            // inclusive prefix of the window counts over the warp; windows are then dealt round-robin to the lanes
            u32 incl = n;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= (u32)s) incl += o; }
            const u32 total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll 1
            for (u32 w0 = 0; w0 < total; w0 += 32u) {
                const u32 gi = w0 + lane;
                // smallest lane q with incl[q] > gi
                u32 q = 0;
#pragma unroll
                for (u32 s = 16; s; s >>= 1) { const u32 v = __shfl_sync(0xffffffffu, incl, (q + s - 1u) & 31u); if (v <= gi) q += s; }
                q &= 31u;
                const u32 before = __shfl_sync(0xffffffffu, incl - n, q);
                const u64 dq = __shfl_sync(0xffffffffu, d, q);
                if (gi < total) {
                    const u32 off = gi - before;
                    const u32 e = KG_SKM_J0(dq) + off;
                    const u32 srcr = KG_SKM_SRC(dq), part = KG_SKM_PART(dq);
                    const u64* __restrict__ words = a.src->words[srcr];
                    KgKmerWindow<W> win;
                    kg_window_at<W>(words, e, g, win.f);
                    kg_revcomp<W>(win.f, win.r, g);
                    const bool fwd = kg_forward_is_canonical<W>(win);
                    u64 key[W];
#pragma unroll
                    for (int x = 0; x < W; x++) key[x] = fwd ? win.f[x] : win.r[x];
                    if (SINK == KG_SINK_BLOOM1 || SINK == KG_SINK_BLOOM2) {
                        sink.b_lo = __ldg(a.bpart_lo + part);
                        sink.b_n = __ldg(a.bpart_lo + part + 1) - sink.b_lo;
                    }
                    if (SINK != KG_SINK_BLOOM1) {
                        sink.t_lo = __ldg(a.part_lo + part);
                        sink.t_n = __ldg(a.part_lo + part + 1) - sink.t_lo;
                    }
                    KgOcc occ; occ.word = ~0ULL;
                    if (SINK != KG_SINK_BLOOM1 && a.table.kaarme) {
                        const bool hp = off > 0 || KG_SKM_HP(dq);
                        const u32 c_out = hp ? kg_base_at(words, e - a.k) : 0u;
                        const u64 gpos = ((u64)srcr << 48) | (a.src->hdr[srcr][0] + (u64)e);
                        occ = kg_make_occ(gpos, hp, fwd, c_out);
                    }
                    sink(key, kg_hash_key<W>(key), occ);
                }
            }
        }
    }
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
