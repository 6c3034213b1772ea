// kg_refbloom.cuh -- bit-exact emulation of the reference's double Bloom filter.
// OPT-IN (kg_config.reserved & KG_CFG_REFERENCE_BLOOM) and NOT YET RUN ON HARDWARE: written after this round's GPU
// budget was spent.  The formulation is proven on the CPU (tests/test_bloom_order_free_model_cpu.py), the hash
// functions are checked on the CPU (kg_refhash.cuh); the kernels below are their transcription.
//
// The reference's pass 1 (insertion_process, double_bloomfilter.hpp:371-413) looks sequential -- which filter a window
// lands in depends on every insertion before it -- but after window s ALL of its bit positions B_s are set in filter 1
// whichever branch it took, so filter 1 just before window t is the union of B_s over s < t.  With
//     T1[b] = min { ord(t) : b in B_t }          ord = the window's global end position (input order)
// "all my filter-1 bits were set before me" is max_b T1[b] < ord(t); the duplicate-position quirk (:401-411: two of the
// nh hash values coincide on a still-unset bit, the insert "fails", the k-mer is ALSO put into filter 2) is "a
// duplicated b has T1[b] == ord(t)".  The promoted windows define T2 the same way; bit b of the squeezed filter 2 is
// T2[b] != never.  Three order-free sweeps over the (resident) packed base stream:
//     sweep 1  atomicMin(T1[b], ord) for every bit of every window                       (while the input is fed)
//     sweep 2  promoted windows: atomicMin(T2[b], ord)
//     sweep 3  new_in_first / new_in_second (:389-396), exactly as a single worker counts them
// Pass 2 admits a window iff its first floor(h) positions are in filter 2 (main.cpp:472, parallel_parser.hpp:2021-2026).
// Speed is not the point of this mode (the root is recomputed per window from the canonical key: 2k multiply-adds).
#pragma once
#include "kg_count.cuh"
#include "kg_refhash.cuh"

#define KG_RB_NEVER 0xFFFFFFFFu
#define KG_RB_MAX_NH 16

struct KgRefBloom {
    u32* T1;       // [m] ordinal of the first window that set bit b of filter 1
    u32* T2;       // [m] same for filter 2
    u64 mask;      // m - 1 (m = bits per filter, a power of two: main.cpp:404-410)
    u32 nh;        // ceil(h): hash functions of pass 1   (main.cpp:417)
    u32 nh2;       // floor(h): hash functions tested in pass 2 (main.cpp:472)
};

// ---- per-window logic, __host__ __device__: tests/native/refbloom_host.cu runs exactly this code on a CPU, windows
// in a scrambled order, against the sequential semantics ---------------------------------------------------------
KG_RH_HD void kg_rb_min(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}

KG_RH_HD void kg_rb_positions_from_root(u64 root, const KgRefBloom& rb, u64 (&pos)[KG_RB_MAX_NH]) {
    for (int i = 0; i < KG_RB_MAX_NH; i++) pos[i] = (u32)i < rb.nh ? (kg_xxh64_8(root, kg_ref_seed((u32)i)) & rb.mask) : 0;
}

// "the insert into this filter is reported as failed": a position hit twice by this window that nobody set before it
KG_RH_HD bool kg_rb_dup_on_fresh_bit(const u32* T, const u64 (&pos)[KG_RB_MAX_NH], u32 nh, u32 ord) {
    bool dup = false;
    for (u32 i = 1; i < nh; i++)
        for (u32 j = 0; j < i; j++)
            if (pos[i] == pos[j] && T[pos[i]] == ord) dup = true;
    return dup;
}
KG_RH_HD bool kg_rb_all_before(const u32* T, const u64 (&pos)[KG_RB_MAX_NH], u32 n, u32 ord) {
    bool all = true;
    for (u32 i = 0; i < n; i++) all = all && (T[pos[i]] < ord);       // KG_RB_NEVER is never < ord
    return all;
}

// what sweep SWEEP does with one window (ord = its global end position, pos = its nh bit positions)
template <int SWEEP>
KG_RH_HD void kg_rb_window(const KgRefBloom& rb, const u64 (&pos)[KG_RB_MAX_NH], u32 ord, u32& n1, u32& n2) {
    if (SWEEP == 1) {
        for (u32 i = 0; i < rb.nh; i++) kg_rb_min(&rb.T1[pos[i]], ord);
        return;
    }
    const bool promoted = kg_rb_all_before(rb.T1, pos, rb.nh, ord) || kg_rb_dup_on_fresh_bit(rb.T1, pos, rb.nh, ord);
    if (SWEEP == 2) {
        if (promoted) for (u32 i = 0; i < rb.nh; i++) kg_rb_min(&rb.T2[pos[i]], ord);
        return;
    }
    // SWEEP 3: the counters (double_bloomfilter.hpp:389-396)
    if (kg_rb_all_before(rb.T2, pos, rb.nh, ord)) return;              // in filter 2 when it arrives: nothing happens
    if (promoted) n2 += kg_rb_dup_on_fresh_bit(rb.T2, pos, rb.nh, ord) ? 0u : 1u;
    else n1 += 1u;
}

// pass 2 admission (second_contains with floor(h) hashes, parallel_parser.hpp:2021-2026)
KG_RH_HD bool kg_rb_admits(const KgRefBloom& rb, const u64 (&pos)[KG_RB_MAX_NH]) {
    bool admit = true;
    for (u32 i = 0; i < rb.nh2; i++) admit = admit && (rb.T2[pos[i]] != KG_RB_NEVER);
    return admit;
}

#if defined(__CUDACC__)
// bit positions of a canonical k-mer: root = min(Hf, Hb) is orientation-free, so it can be taken from the canonical key
template <int W>
__device__ __forceinline__ void kg_rb_positions(const u64 (&key)[W], u32 k, const KgRefBloom& rb, u64 (&pos)[KG_RB_MAX_NH]) {
    const KgKGeom g = kg_geom(k);
    u64 rc[W];
    kg_revcomp<W>(key, rc, g);
    kg_rb_positions_from_root(kg_ref_root(kg_b5_horner(key, (u32)W, k), kg_b5_horner(rc, (u32)W, k)), rb, pos);
}

template <int W, int SWEEP>
__global__ void __launch_bounds__(256) kg_refbloom_sweep(const u64* __restrict__ words, const u32* __restrict__ brk,
                                                         const KgStream* __restrict__ st, KgRefBloom rb, KgStats* stats, u32 k) {
    const u32 T = st->total_bases, C = st->carry_bases;
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 n1 = 0, n2 = 0;
    const u32 n_windows = kg_for_each_window<W>(words, brk, T, C, k, t, st->bases_seen,
        [&](const u64 (&key)[W], u64, KgOcc occ, u32) {
            u64 pos[KG_RB_MAX_NH];
            kg_rb_positions<W>(key, k, rb, pos);
            kg_rb_window<SWEEP>(rb, pos, (u32)(occ.word >> 4), n1, n2);    // ordinal = global end position of the window
        });
    if (SWEEP == 1) { KG_WARP_ADD(stats, n_windows, input_kmers) }
    if (SWEEP == 3) {
        KG_WARP_ADD(stats, n1, new_in_first)
        KG_WARP_ADD(stats, n2, new_in_second)
    }
}

// count pass of the emulation mode: admission by the squeezed filter 2, then the ordinary find-or-insert
template <int W>
__global__ void __launch_bounds__(256) kg_refbloom_count(KgCountArgs a, KgRefBloom rb) {
    const u32 T = a.st->total_bases, C = a.st->carry_bases;
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    u32 n_ins = 0, n_new = 0, n_rej = 0;
    bool full = false;
    const KgTable table = a.table;
    const u32 n_windows = kg_for_each_window<W>(a.words, a.brk, T, C, a.k, t, a.st->bases_seen,
        [&](const u64 (&key)[W], u64 h, KgOcc occ, u32) {
            u64 pos[KG_RB_MAX_NH];
            kg_rb_positions<W>(key, a.k, rb, pos);
            if (!kg_rb_admits(rb, pos)) { n_rej++; return; }
            bool is_new;
            u64 slot;
            const u64 slot0 = kg_slot(h, table.nslots, table.world);
            if constexpr (W == 2) slot = table.packed_tb ? kg_table_add_packed(table, key, slot0, is_new) : kg_table_add<W>(table, key, slot0, is_new);
            else slot = kg_table_add<W>(table, key, slot0, is_new);
            if (slot == ~0ULL) { full = true; return; }
            n_ins++;
            n_new += is_new ? 1u : 0u;
            if (table.kaarme) atomicMax(table.slots + slot * table.stride + 1 + W, ~occ.word);
        });
    KG_WARP_ADD(a.stats, n_windows, input_kmers)
    KG_WARP_ADD(a.stats, n_ins, inserted)
    KG_WARP_ADD(a.stats, n_new, distinct)
    KG_WARP_ADD(a.stats, n_rej, bloom_rejected)
    if (full) a.stats->table_full = 1;
}
#endif
