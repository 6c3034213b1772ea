// kg_text.cuh -- K5b "text dump": format exported (key, count) records as the reference's output lines ON THE GPU.
//
// Replaces the single-threaded, char-by-char `ofstream <<` loop of write_kmers (kmer_hash_table.cpp:2013-2050) and
// write_kmers_on_disk_separately_even_faster (:4318-4524): "<k characters ACGT> <decimal count>\n" per k-mer.
// After the GPU speed-up of the count pass that writer dominates the wall clock, so the
// lines are produced by a kernel and the host only write(2)s whole buffers.
//
// The per-record formatting (kg_ndigits, kg_format_line) is plain integer/byte code, compiled for host AND device,
// so that tests/native/text_format_host.cu can check it on a CPU against the reference writer format.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KG_HD __host__ __device__ __forceinline__
#else
#define KG_HD inline
#endif

#define KG_TEXT_MAX_COUNT_DIGITS 10u   // counts are 32-bit
// upper bound on the bytes of one line: k characters, ' ', <= 10 digits, '\n'
KG_HD uint32_t kg_line_bound(uint32_t k) { return k + 2u + KG_TEXT_MAX_COUNT_DIGITS; }

KG_HD uint32_t kg_ndigits(uint32_t v) {
    uint32_t n = 1;
    if (v >= 100000000u) { n += 8; v /= 100000000u; }
    if (v >= 10000u) { n += 4; v /= 10000u; }
    if (v >= 100u) { n += 2; v /= 100u; }
    if (v >= 10u) n += 1;
    return n;
}

// Writes one output line at dst and returns its length (= k + 2 + kg_ndigits(count)).
//   key: W = ceil(k/32) words, the 2k-bit string right-aligned, word 0 most significant
//        (KMerFactoryCanonical2BC layout, kmer_factory.cpp:31-33): word 0 holds the first k - 32(W-1) characters.
//   Characters: 0,1,2,3 -> A,C,G,T (functions_strings.cpp:72-95).
KG_HD uint32_t kg_format_line(const unsigned long long* key, uint32_t W, uint32_t k, uint32_t count, char* dst) {
    char* p = dst;
    uint32_t nchar = k - 32u * (W - 1u);                  // characters held by word 0 (1..32)
    for (uint32_t w = 0; w < W; w++) {
        const unsigned long long v = key[w];
        for (int c = (int)nchar - 1; c >= 0; c--) {
            const uint32_t code = (uint32_t)(v >> (2 * c)) & 3u;
            *p++ = (char)((0x54474341u >> (8u * code)) & 0xFFu);   // "ACGT"[code]
        }
        nchar = 32u;
    }
    *p++ = ' ';
    const uint32_t nd = kg_ndigits(count);
    uint32_t v = count;
    for (int i = (int)nd - 1; i >= 0; i--) { p[i] = (char)('0' + v % 10u); v /= 10u; }
    p += nd;
    *p++ = '\n';
    return (uint32_t)(p - dst);
}

#if defined(__CUDACC__)
// One thread per exported record (n = *n_dev of them, compacted by kg_export_kernel / kg_kaarme_export).
// A block: line lengths -> exclusive scan -> ONE atomicAdd reserves the block's byte range in `out` (line order
// is unspecified, as in the reference) -> the 256 lines are staged in shared memory at the same offset modulo 16
// as their destination -> coalesced 16-byte stores (byte stores only for the ragged head and tail).
// Dynamic shared memory: 16 + blockDim.x * kg_line_bound(k) bytes.
#define KG_TEXT_TPB 256
__global__ void __launch_bounds__(KG_TEXT_TPB) kg_format_text(const unsigned long long* __restrict__ keys,
                                                              const uint32_t* __restrict__ counts,
                                                              const uint32_t* __restrict__ n_dev, uint32_t W, uint32_t k,
                                                              char* __restrict__ out, unsigned long long* cursor) {
    extern __shared__ __align__(16) unsigned char s_text[];
    __shared__ uint32_t s_warp[KG_TEXT_TPB / 32];
    __shared__ unsigned long long s_base;
    const uint32_t n = *n_dev;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if ((unsigned long long)blockIdx.x * KG_TEXT_TPB >= n) return;          // uniform over the block
    const uint32_t i = blockIdx.x * KG_TEXT_TPB + tid;
    uint32_t count = 0, len = 0;
    if (i < n) { count = counts[i]; len = k + 2u + kg_ndigits(count); }
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t pre = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < KG_TEXT_TPB / 32; w++) { const uint32_t s = s_warp[w]; if (w < warp) pre += s; total += s; }
    if (tid == 0) s_base = atomicAdd(cursor, (unsigned long long)total);
    __syncthreads();
    const unsigned long long base = s_base;
    const uint32_t pad = (uint32_t)(base & 15ULL);
    if (i < n) kg_format_line(keys + (unsigned long long)i * W, W, k, count, reinterpret_cast<char*>(s_text) + pad + (pre + incl - len));
    __syncthreads();
    char* gdst = out + (base - pad);                     // 16-byte aligned; s_text[b] <-> gdst[b]
    const uint32_t end = pad + total;
    for (uint32_t lo = tid * 16u; lo < end; lo += KG_TEXT_TPB * 16u) {
        const uint32_t hi = lo + 16u;
        if (lo >= pad && hi <= end) {
            *reinterpret_cast<uint4*>(gdst + lo) = *reinterpret_cast<const uint4*>(s_text + lo);
        } else {
            const uint32_t b0 = lo > pad ? lo : pad, b1 = hi < end ? hi : end;
            for (uint32_t b = b0; b < b1; b++) gdst[b] = (char)s_text[b];
        }
    }
}
#endif
