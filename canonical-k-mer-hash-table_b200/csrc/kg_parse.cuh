// kg_parse.cuh -- K1 "parse": raw FASTA / one-string-per-line bytes -> packed 2-bit base stream + break
// mask.  Replaces the per-byte scanner of the reference functors (parallel_parser.hpp:597-638 FASTA,
// :391-400 PLAIN; char codec functions_strings.cpp:56-70) and the chunk-overlap logic of
// text_reader.h:91-226 (here: device-resident carry of the last k-1 bases).
//
// Semantics restated:
//   FASTA  '>' anywhere starts a header that runs through the next '\n'; a header resets the window;
//          '\n' outside a header is skipped (multi-line records are one sequence)
//   PLAIN  '\n' is just an invalid byte (each line its own sequence)
//   both   AaCcGgTt -> 0,1,2,3; any other byte resets the window
// Both "am I inside a header" and "did a reset happen since the previous base" are last-writer-wins
// scans, so the batch is processed in 4 KiB tiles with two tiny single-block scans in between:
//   kg_hdr_summary -> kg_hdr_scan -> kg_tile_count -> kg_tile_scan -> kg_tile_pack
// All kernels are HBM-streaming: 3 reads of the raw bytes + 0.375 B/base written; the per-byte work is SWAR
// (kg_masks16 / kg_parse16 below).
#pragma once
#include "kg_device.cuh"

#define KG_PT 256                    // threads per tile block
#define KG_BPT 16                    // bytes per thread (one 128-bit load)
#define KG_TILE (KG_PT * KG_BPT)     // 4096 bytes per tile

// last-writer-wins effects
#define KG_EFF_NONE 0u
#define KG_EFF_SET 1u    // header: enter header   | pending: a break happened after the last base
#define KG_EFF_CLEAR 2u  // header: leave header   | pending: a base was emitted after the last break

__device__ __forceinline__ uint4 kg_load_tile16(const uint8_t* in, size_t n, size_t off) {
    // 16 bytes at `off` (16-byte aligned base pointer + multiple-of-16 offset); bytes past n read as '\n'
    // in the sense of "no effect": callers mask by position.
    uint4 v = make_uint4(0, 0, 0, 0);
    if (off + 16 <= n) {
        v = *reinterpret_cast<const uint4*>(in + off);
    } else if (off < n) {
        uint8_t tmp[16];
#pragma unroll
        for (int i = 0; i < 16; i++) tmp[i] = (off + i < n) ? in[off + i] : 0;
        v = *reinterpret_cast<uint4*>(tmp);
    }
    return v;
}
// TMA-staged variant (opt-in, KG_PARSE_TMA=1): ONE elected thread issues a 1-D bulk copy (cp.async.bulk, the TMA unit;
// SASS UBLKCP) of the block's whole 4 KiB tile into shared memory and arms an mbarrier with the byte count; every
// thread waits on the barrier's phase and takes its 16 bytes with one LDS.128 -- 256 LDG.128 per block become one
// asynchronous copy that does not occupy the LSU.  Only full tiles (the source must be 16-byte aligned and the size a
// multiple of 16); the ragged last tile of a batch takes the LDG path.  `tile_base` is block-uniform.
template <bool TMA>
__device__ __forceinline__ uint4 kg_fetch16(const uint8_t* in, size_t n, size_t tile_base, size_t off) {
    if constexpr (TMA) {
        __shared__ __align__(128) uint8_t s_tile[KG_TILE];
        __shared__ __align__(8) unsigned long long s_mbar;
        if (tile_base + KG_TILE <= n) {                         // uniform over the block
            const u32 mbar = (u32)__cvta_generic_to_shared(&s_mbar);
            const u32 dst = (u32)__cvta_generic_to_shared(s_tile);
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"((u32)KG_TILE) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(in + tile_base), "r"((u32)KG_TILE), "r"(mbar) : "memory");
            }
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "KG_TMA_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
                "@p bra KG_TMA_DONE;\n\t"
                "bra KG_TMA_WAIT;\n\t"
                "KG_TMA_DONE:\n\t"
                "}" ::"r"(mbar) : "memory");
            return *reinterpret_cast<const uint4*>(s_tile + (off - tile_base));
        }
    }
    return kg_load_tile16(in, n, off);
}

// combine for last-writer-wins: later non-NONE effect overrides
__device__ __forceinline__ u32 kg_lww(u32 earlier, u32 later) { return later != KG_EFF_NONE ? later : earlier; }

// inclusive last-writer-wins scan over the block's threads; returns the EXCLUSIVE value for this thread
// (effect of all earlier threads) and the block total in *block_total.
__device__ __forceinline__ u32 kg_block_lww_exclusive(u32 mine, u32* smem /*>= 8 words*/, u32* block_total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl = kg_lww(o, incl);
    }
    u32 excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = KG_EFF_NONE;
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    u32 pre = KG_EFF_NONE, tot = KG_EFF_NONE;
    for (u32 w = 0; w < KG_PT / 32; w++) {
        u32 e = smem[w];
        if (w < warp) pre = kg_lww(pre, e);
        tot = kg_lww(tot, e);
    }
    __syncthreads();
    *block_total = tot;
    return kg_lww(pre, excl);
}

__device__ __forceinline__ u32 kg_block_sum_exclusive(u32 mine, u32* smem, u32* block_total) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (u32)d) incl += o;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    u32 pre = 0, tot = 0;
    for (u32 w = 0; w < KG_PT / 32; w++) {
        u32 e = smem[w];
        if (w < warp) pre += e;
        tot += e;
    }
    __syncthreads();
    *block_total = tot;
    return pre + incl - mine;
}

// ---- SWAR classification of 16 bytes ----------------------------------------------------------------------
// Per-byte compares run four at a time (__vcmpeq4); the results are folded into 16-bit position masks
// (bit i = byte i) and the 2-bit codes into one 32-bit word (bits 2i+1:2i = code of byte i), so the per-byte
// state machine of the reference's scanner (parallel_parser.hpp:597-638) becomes a handful of bit operations
// per EVENT (header, newline, invalid byte) instead of a branchy loop per byte.
struct KgMasks16 {
    u32 base;   // A C G T a c g t
    u32 nl;     // '\n'
    u32 gt;     // '>'
    u32 codes;  // 2-bit code of every byte (garbage where base is 0)
};
__device__ __forceinline__ u32 kg_movemask4(u32 m) {          // 0xFF/0x00 per byte -> 4 bits, bit i = byte i
    return ((m & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ KgMasks16 kg_masks16(const uint4& v, int nvalid) {
    const u32 w[4] = {v.x, v.y, v.z, v.w};
    KgMasks16 m;
    m.base = m.nl = m.gt = m.codes = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u32 x = w[i], t = x & 0xDFDFDFDFu;                 // fold case
        const u32 b4 = __vcmpeq4(t, 0x41414141u) | __vcmpeq4(t, 0x43434343u) | __vcmpeq4(t, 0x47474747u) | __vcmpeq4(t, 0x54545454u);
        u32 c2 = (t >> 1) & 0x03030303u;                         // A0 C1 G3 T2
        c2 ^= (c2 >> 1) & 0x01010101u;                           // A0 C1 G2 T3
        m.base |= kg_movemask4(b4) << (4 * i);
        m.nl |= kg_movemask4(__vcmpeq4(x, 0x0A0A0A0Au)) << (4 * i);
        m.gt |= kg_movemask4(__vcmpeq4(x, 0x3E3E3E3Eu)) << (4 * i);
        m.codes |= ((c2 * 0x01041040u) >> 24) << (8 * i);
    }
    const u32 live = nvalid >= 16 ? 0xFFFFu : ((1u << nvalid) - 1u);
    m.base &= live; m.nl &= live; m.gt &= live;
    return m;
}

// header effect of 16 bytes: '>' => SET (from this byte on), '\n' => CLEAR (from the next byte on); last one wins
__device__ __forceinline__ u32 kg_hdr_effect16(const KgMasks16& m) {
    if ((m.gt | m.nl) == 0) return KG_EFF_NONE;
    return m.gt > m.nl ? KG_EFF_SET : KG_EFF_CLEAR;              // the higher top bit is the later byte
}

// ---- pass A (FASTA only): per-tile header effect ------------------------------------------------------
template <bool TMA>
__global__ void __launch_bounds__(KG_PT) kg_hdr_summary(const uint8_t* __restrict__ in, size_t n,
                                                        u32* __restrict__ tile_hdr_eff) {
    __shared__ u32 sm[8];
    const size_t off = (size_t)blockIdx.x * KG_TILE + (size_t)threadIdx.x * KG_BPT;
    int nvalid = off >= n ? 0 : (n - off >= 16 ? 16 : (int)(n - off));
    uint4 v = kg_fetch16<TMA>(in, n, (size_t)blockIdx.x * KG_TILE, off);
    const KgMasks16 m = kg_masks16(v, nvalid);
    u32 tot;
    kg_block_lww_exclusive(kg_hdr_effect16(m), sm, &tot);
    if (threadIdx.x == 0) tile_hdr_eff[blockIdx.x] = tot;
}

// ---- single-block exclusive last-writer-wins scan over tiles ------------------------------------------
// in_eff[t] -> out_state[t] = state (0/1) at the start of tile t; *state_io: initial state in, final out.
__global__ void __launch_bounds__(1024) kg_lww_scan(const u32* __restrict__ in_eff, u32* __restrict__ out_state,
                                                    u32 ntiles, u32* state_io) {
    __shared__ u32 sw[32];
    const u32 per = (ntiles + 1023) / 1024;
    const u32 b = threadIdx.x * per, e = min(b + per, ntiles);
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    u32 mine = KG_EFF_NONE;
    for (u32 t = b; t < e; t++) mine = kg_lww(mine, in_eff[t]);
    // inclusive last-writer-wins scan across the 1024 threads: shuffles inside a warp, then across the 32 warps
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (u32)d) incl = kg_lww(o, incl); }
    if (lane == 31) sw[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 v = sw[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= (u32)d) v = kg_lww(o, v); }
        sw[lane] = v;
    }
    __syncthreads();
    const u32 init = *state_io ? KG_EFF_SET : KG_EFF_CLEAR;
    u32 excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = KG_EFF_NONE;
    const u32 pre = kg_lww(kg_lww(init, warp ? sw[warp - 1] : KG_EFF_NONE), excl);
    u32 cur = pre;
    for (u32 t = b; t < e; t++) {
        out_state[t] = cur == KG_EFF_SET ? 1u : 0u;
        cur = kg_lww(cur, in_eff[t]);
    }
    __syncthreads();
    if (threadIdx.x == 1023) *state_io = (kg_lww(pre, mine) == KG_EFF_SET) ? 1u : 0u;
}

// ---- classification of one thread's 16 bytes given its incoming header state ---------------------------
struct KgThreadParse {
    u32 nbases;     // bases emitted
    u32 bits;       // 2-bit codes, first base in the TOP bits of a 32-bit word
    u32 brk;        // break flag per emitted base, first base in bit 15 (of 16), *excluding* incoming pending
    u32 first_needs_pending;  // 1 if the first base's break flag is (still) decided by the incoming pending
    u32 pend_eff;   // effect on "pending break" after these bytes
};

__device__ __forceinline__ u32 kg_bits_from(u32 lo) { return ~((1u << lo) - 1u); }          // bits lo..31
__device__ __forceinline__ u32 kg_bits_upto(u32 hi) { return hi >= 31 ? ~0u : ((2u << hi) - 1u); }   // bits 0..hi

template <bool FASTA>
__device__ __forceinline__ KgThreadParse kg_parse16(const KgMasks16& m, int nvalid, u32 in_header) {
    KgThreadParse r;
    const u32 live = nvalid >= 16 ? 0xFFFFu : ((1u << nvalid) - 1u);
    u32 B = m.base;                    // bytes that are bases
    u32 E;                             // break events (bit = position of the byte that breaks the window)
    if (FASTA) {
        // header regions: from a '>' (or from byte 0 when we start inside a header) through the next '\n'
        u32 hdrmask = 0, starts = 0, pos = 0, hdr = in_header;
        if (hdr) {
            const u32 nlb = m.nl;
            starts = 1u;
            if (nlb) { const u32 e = __ffs(nlb) - 1; hdrmask = kg_bits_upto(e); pos = e + 1; hdr = 0; }
            else { hdrmask = live; pos = 16; }
        }
        while (pos < 16) {
            const u32 g = m.gt & kg_bits_from(pos);
            if (!g) break;
            const u32 sb = __ffs(g) - 1;
            starts |= 1u << sb;
            const u32 nlb = m.nl & kg_bits_from(sb);
            if (nlb) { const u32 e = __ffs(nlb) - 1; hdrmask |= kg_bits_from(sb) & kg_bits_upto(e); pos = e + 1; }
            else { hdrmask |= kg_bits_from(sb) & live; pos = 16; }
        }
        B &= ~hdrmask;
        // a header resets the window (event at its first byte); other non-base bytes outside headers are events,
        // except '\n', which is skipped inside a record
        E = (starts | (live & ~(m.base | m.nl) & ~hdrmask)) & live;
    } else {
        E = live & ~m.base;           // PLAIN: every non-ACGT byte, newline included, ends the string
    }
    r.nbases = __popc(B);
    // break flag of a base = an event since the previous base; only the FIRST base after each event gets one
    u32 brk = 0, ev = E;
    while (ev) {
        const u32 e = __ffs(ev) - 1;
        const u32 after = B & kg_bits_from(e);      // (an event byte is never a base)
        if (!after) break;
        const u32 fb = __ffs(after) - 1;
        brk |= 1u << fb;
        ev &= kg_bits_from(fb);                     // events up to that base are spent
    }
    const u32 lowB = B ? (u32)__ffs(B) - 1 : 32u;
    r.first_needs_pending = (B && (E & kg_bits_upto(lowB)) == 0) ? 1u : 0u;   // no event before the first base
    if (B) {
        const u32 top = 31u - __clz(B);
        r.pend_eff = (E & kg_bits_from(top)) ? KG_EFF_SET : KG_EFF_CLEAR;      // events after the last base
    } else {
        r.pend_eff = E ? KG_EFF_SET : KG_EFF_NONE;
    }
    // compact codes and break bits of the base positions (drop every non-base position below the top base)
    u32 codes = m.codes, holes = B ? (~B & kg_bits_upto(31u - __clz(B))) : 0u;
    while (holes) {
        const u32 h = 31u - __clz(holes);           // highest hole first: lower indices stay valid
        const u32 lowc = (1u << (2 * h)) - 1u, lowb = (1u << h) - 1u;
        codes = (codes & lowc) | ((codes >> 2) & ~lowc);
        brk = (brk & lowb) | ((brk >> 1) & ~lowb);
        holes &= ~(1u << h);
    }
    // to the output convention: first base in the TOP bits
    u32 rev = __brev(codes);
    rev = ((rev >> 1) & 0x55555555u) | ((rev & 0x55555555u) << 1);
    r.bits = r.nbases ? (rev & ~((r.nbases >= 16) ? 0u : ((1u << (32 - 2 * r.nbases)) - 1u))) : 0u;
    r.brk = (__brev(brk) >> 16) & 0xFFFFu;
    if (r.first_needs_pending == 0 && r.nbases) { /* first base's flag already set by its event */ }
    return r;
}

// ---- pass B: per-tile base count and pending-break effect -----------------------------------------------
template <bool FASTA, bool TMA>
__global__ void __launch_bounds__(KG_PT) kg_tile_count(const uint8_t* __restrict__ in, size_t n,
                                                       const u32* __restrict__ tile_hdr_in,
                                                       u32* __restrict__ tile_nbases, u32* __restrict__ tile_pend_eff) {
    __shared__ u32 sm[8];
    const size_t off = (size_t)blockIdx.x * KG_TILE + (size_t)threadIdx.x * KG_BPT;
    int nvalid = off >= n ? 0 : (n - off >= 16 ? 16 : (int)(n - off));
    uint4 v = kg_fetch16<TMA>(in, n, (size_t)blockIdx.x * KG_TILE, off);
    const KgMasks16 m = kg_masks16(v, nvalid);
    u32 hdr_in = 0, tot;
    if (FASTA) {
        u32 pre = kg_block_lww_exclusive(kg_hdr_effect16(m), sm, &tot);
        u32 tile_in = tile_hdr_in[blockIdx.x] ? KG_EFF_SET : KG_EFF_CLEAR;
        hdr_in = kg_lww(tile_in, pre) == KG_EFF_SET;
    }
    KgThreadParse p = kg_parse16<FASTA>(m, nvalid, hdr_in);
    u32 total_bases, pend_tot;
    kg_block_sum_exclusive(p.nbases, sm, &total_bases);
    kg_block_lww_exclusive(p.pend_eff, sm, &pend_tot);
    if (threadIdx.x == 0) { tile_nbases[blockIdx.x] = total_bases; tile_pend_eff[blockIdx.x] = pend_tot; }
}

// ---- single-block exclusive sum over tiles; also finalises the stream state ------------------------------
__global__ void __launch_bounds__(1024) kg_tile_scan(const u32* __restrict__ tile_nbases, u32* __restrict__ tile_off,
                                                     u32 ntiles, KgStream* st) {
    __shared__ u32 sw[32];
    const u32 per = (ntiles + 1023) / 1024;
    const u32 b = threadIdx.x * per, e = min(b + per, ntiles);
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    u32 mine = 0;
    for (u32 t = b; t < e; t++) mine += tile_nbases[t];
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (u32)d) incl += o; }
    if (lane == 31) sw[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 v = sw[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= (u32)d) v += o; }
        sw[lane] = v;
    }
    __syncthreads();
    const u32 pre = st->carry_bases + (warp ? sw[warp - 1] : 0u) + incl - mine;
    u32 cur = pre;
    for (u32 t = b; t < e; t++) { tile_off[t] = cur; cur += tile_nbases[t]; }
    __syncthreads();
    if (threadIdx.x == 1023) st->total_bases = pre + mine;
}

// ---- pass C: pack bases and break bits ---------------------------------------------------------------------
// words/brk must be zero beyond the carried head; tile boundary words are merged with atomicOr.
template <bool FASTA, bool TMA>
__global__ void __launch_bounds__(KG_PT) kg_tile_pack(const uint8_t* __restrict__ in, size_t n,
                                                      const u32* __restrict__ tile_hdr_in,
                                                      const u32* __restrict__ tile_off,
                                                      const u32* __restrict__ tile_pend_in,
                                                      u64* __restrict__ words, u32* __restrict__ brk) {
    __shared__ u32 sm[8];
    __shared__ u64 sw[KG_TILE / 32 + 2];
    __shared__ u32 sb[KG_TILE / 32 + 2];
    const size_t off = (size_t)blockIdx.x * KG_TILE + (size_t)threadIdx.x * KG_BPT;
    int nvalid = off >= n ? 0 : (n - off >= 16 ? 16 : (int)(n - off));
    uint4 v = kg_fetch16<TMA>(in, n, (size_t)blockIdx.x * KG_TILE, off);
    const KgMasks16 m = kg_masks16(v, nvalid);
    for (u32 i = threadIdx.x; i < KG_TILE / 32 + 2; i += KG_PT) { sw[i] = 0; sb[i] = 0; }
    u32 hdr_in = 0, tot;
    if (FASTA) {
        u32 pre = kg_block_lww_exclusive(kg_hdr_effect16(m), sm, &tot);
        u32 tile_in = tile_hdr_in[blockIdx.x] ? KG_EFF_SET : KG_EFF_CLEAR;
        hdr_in = kg_lww(tile_in, pre) == KG_EFF_SET;
    } else {
        __syncthreads();
    }
    KgThreadParse p = kg_parse16<FASTA>(m, nvalid, hdr_in);
    u32 total_bases, pend_tot;
    u32 base_pre = kg_block_sum_exclusive(p.nbases, sm, &total_bases);
    u32 pend_pre = kg_block_lww_exclusive(p.pend_eff, sm, &pend_tot);
    const u32 t_off = tile_off[blockIdx.x];
    if (p.nbases) {
        u32 tile_pend = tile_pend_in[blockIdx.x] ? KG_EFF_SET : KG_EFF_CLEAR;
        u32 pend_in = kg_lww(tile_pend, pend_pre) == KG_EFF_SET;
        u32 bmask = p.brk;
        if (p.first_needs_pending && pend_in) bmask |= 1u << 15;
        // local ordinal relative to the first word this tile touches
        u32 lo = (t_off & 31u) + base_pre;
        u32 wi = lo >> 5, bo = lo & 31u;                // word index, base offset inside the word
        // bits: p.nbases characters in the top of a 32-bit value -> place at base offset bo of a 64-bit word
        u64 chunk = (u64)p.bits << 32;                  // first base at bits 63:62
        u32 bchunk = bmask << 16;                       // first base at bit 31
        atomicOr(&sw[wi], chunk >> (2 * bo));
        atomicOr(&sb[wi], bchunk >> bo);
        if (bo + p.nbases > 32) {                       // spills into the next word
            u32 used = 32 - bo;                         // bases that fitted
            atomicOr(&sw[wi + 1], chunk << (2 * used));
            atomicOr(&sb[wi + 1], bchunk << used);
        }
    }
    __syncthreads();
    const u32 nwords = total_bases ? (((t_off & 31u) + total_bases + 31u) >> 5) : 0;
    const u32 gw0 = t_off >> 5;
    for (u32 i = threadIdx.x; i < nwords; i += KG_PT) {
        if (i == 0 || i + 1 == nwords) {
            if (sw[i]) atomicOr(&words[gw0 + i], sw[i]);
            if (sb[i]) atomicOr(&brk[gw0 + i], sb[i]);
        } else {
            words[gw0 + i] = sw[i];
            brk[gw0 + i] = sb[i];
        }
    }
}

// ---- carry: move the tail of the finished batch to the head of the next one -------------------------------
// Keeps whole words: the carried region starts at word floor((T-k)/32), so C = T - 32*that >= k (or everything when
// T < k).  k bases, not k-1: the first window counted in the next batch then sees a run of k+1 whenever its read really
// continues to the left, so "this window has a predecessor" (what the Kaarme structure records) does not depend on where
// a batch happened to end.  Runs AFTER the count kernel of the batch, single block.
__global__ void kg_carry_save(const u64* __restrict__ words, const u32* __restrict__ brk, KgStream* st,
                              u64* __restrict__ carry_words, u32* __restrict__ carry_brk, u32 k, u32 max_words) {
    const u32 T = st->total_bases;
    u32 w0 = T >= k ? (T - k) >> 5 : 0;
    u32 nw = ((T + 31) >> 5) - w0;
    if (nw > max_words) nw = max_words;  // cannot happen (max_words = W+3)
    for (u32 i = threadIdx.x; i < max_words; i += blockDim.x) {
        carry_words[i] = i < nw ? words[w0 + i] : 0;
        carry_brk[i] = i < nw ? brk[w0 + i] : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->bases_seen += (u64)(w0) * 32u;   // global ordinal of the first carried base
        st->carry_bases = T - w0 * 32u;
    }
}

// Runs BEFORE packing a batch: zero-initialised words/brk get the carried head; position 0 is a run start.
__global__ void kg_carry_restore(u64* __restrict__ words, u32* __restrict__ brk, const KgStream* st,
                                 const u64* __restrict__ carry_words, const u32* __restrict__ carry_brk,
                                 u32 max_words) {
    const u32 C = st->carry_bases;
    const u32 nw = (C + 31) >> 5;
    for (u32 i = threadIdx.x; i < max_words; i += blockDim.x) {
        if (i < nw) {
            words[i] = carry_words[i];
            u32 b = carry_brk[i];
            if (i == 0) b |= 0x80000000u;
            brk[i] = b;
        }
    }
}

// pending-break input for every tile: exclusive last-writer-wins scan seeded by the stream state
// (reuses kg_lww_scan with state_io = &st->pending_break)
