// kaarme_gpu.cu -- implementation of the C ABI in include/kaarme_gpu.h (libkaarme_gpu.so, sm_100a only).
// Host-side orchestration: pinned double-buffered H2D on a copy stream, parse + bucketing kernels on a compute
// stream, the insert on its own stream, device-resident stream state (no host round trip per batch), chunked export.
// There is no CPU fallback anywhere in this file: every data-path step is a kernel launch.
#include "../../include/kaarme_gpu.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>   // types and prototypes only: the library is bound at run time (see kg_nccl below)

#include "kg_count.cuh"
#include "kg_device.cuh"
#include "kg_kaarme.cuh"
#include "kg_parse.cuh"
#include "kg_refbloom.cuh"
#include "kg_skm.cuh"
#include "kg_text.cuh"

#define KG_MAX_W 8
#define KG_MAX_WORLD 64
#define KG_DISPATCH_W(W_, fn, ...)                     \
    switch (W_) {                                      \
        case 1: fn<1>(__VA_ARGS__); break;             \
        case 2: fn<2>(__VA_ARGS__); break;             \
        case 3: fn<3>(__VA_ARGS__); break;             \
        case 4: fn<4>(__VA_ARGS__); break;             \
        case 5: fn<5>(__VA_ARGS__); break;             \
        case 6: fn<6>(__VA_ARGS__); break;             \
        case 7: fn<7>(__VA_ARGS__); break;             \
        case 8: fn<8>(__VA_ARGS__); break;             \
    }


#define KG_DEFAULT_BATCH (128ull << 20)
#define KG_MAX_BATCH (1ull << 30)

// the slot header must hold the dense descriptor counts of the largest bucket set, and the segment table one entry per
// (partition, sender, sub-region) plus one overflow list per sender
static_assert(64 + 4 * ((size_t)KG_MAX_BUCKETS * KG_SKM_SUB + 1) <= KG_SKM_META, "KG_SKM_META too small for the descriptor counts");
static_assert((size_t)KG_MAX_BUCKETS * KG_SKM_SUB + KG_MAX_WORLD <= KG_SKM_MAXSEG, "KG_SKM_MAXSEG too small");
static_assert(KG_MAX_BUCKETS <= 1024 && KG_MAX_WORLD <= 64, "descriptor fields: 10 bits of partition, 6 bits of owner / sender");

// One of the two batch slots of the minimizer-bucketed path.  A slot is ONE device allocation
//   [ header 64 B | descriptor counts | pad to KG_SKM_META | packed 2-bit words of the batch | descriptor regions + overflow list ]
// so that a peer can map it with one IPC handle and pull header + counts + words with one copy.
struct SkmSlot {
    uint8_t* slab = nullptr;
    u64* hdr = nullptr;           // [0] global ordinal of position 0 of the batch, [1] bases in the packed stream
    u32* counts = nullptr;        // dense descriptor counts per (bucket, sub-region) + overflow (kg_skm_pack_counts)
    u64* words = nullptr;         // packed stream of the batch (parse writes it here)
    u64* desc = nullptr;          // nb regions of skm_cap descriptors, then the overflow list
    KgSkmSources* d_src = nullptr;    // where the insert of this slot finds every sender's words + header
    KgSkmPeers* d_peers = nullptr;    // where kg_skm_segments finds every sender's counts + descriptors
    u64* d_seg_start = nullptr;
    const u64** d_seg_ptr = nullptr;
    cudaEvent_t ev_ready = nullptr;   // s_compute: the slot holds this rank's batch of the round
    cudaEvent_t ev_free = nullptr;    // the slot may be overwritten (one GPU: its insert is done; several: every rank's is)
    uint8_t* peer_slab[KG_MAX_WORLD] = {};   // every rank's slot of this parity as mapped into this process
    void* peer_opened[KG_MAX_WORLD] = {};    // IPC mappings to close at destroy
    uint8_t* r_buf[KG_MAX_WORLD] = {};       // local copies of the peers' header + counts + words (pulled every round)
};

struct kg_ctx {
    kg_config cfg;
    int W = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    size_t batch_bytes = 0;
    uint32_t max_tiles = 0;
    // raw input double buffer
    uint8_t* d_raw[2] = {nullptr, nullptr};
    cudaEvent_t ev_raw_free[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy_done[2] = {nullptr, nullptr};
    uint8_t* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};
    int raw_idx = 0;
    // parse scratch
    u32 *d_tile_hdr_eff = nullptr, *d_tile_hdr_in = nullptr, *d_tile_nbases = nullptr, *d_tile_pend_eff = nullptr,
        *d_tile_pend_in = nullptr, *d_tile_off = nullptr;
    u64* d_words_direct = nullptr;   // packed stream of the direct path (the bucketed path packs into the batch slot)
    u32* d_brk = nullptr;
    size_t words_cap = 0;  // in words
    u64* d_carry_words = nullptr;
    u32* d_carry_brk = nullptr;
    u32 carry_max_words = 0;
    KgStream* d_stream = nullptr;
    KgStats* d_stats = nullptr;
    KgTable table{nullptr, 0, 0, 0, 1, 0, nullptr};
    size_t table_bytes = 0;
    KgBloom bloom{nullptr, 0, 0, 1};
    size_t bloom_bytes = 0;
    uint64_t bloom_m = 0;
    int pass = 0;
    bool stream_open = false;
    bool bloom_done = false;
    uint64_t new_in_second = 0;
    // timing
    cudaEvent_t ev_pass_begin = nullptr, ev_pass_end = nullptr;
    std::vector<cudaEvent_t> ev_pool;   // [parse_begin, parse_end(=count_begin), count_end] per batch
    size_t ev_used = 0;
    std::vector<cudaEvent_t> ins_pool;  // [insert_begin, insert_end] pairs around the insert kernels of the pass
    size_t ins_used = 0;
    uint64_t ins_launches = 0;
    uint64_t raw_bytes_pass = 0;
    uint64_t launches = 0;
    // export staging
    u64* d_out_keys[2] = {nullptr, nullptr};
    u32* d_out_counts[2] = {nullptr, nullptr};
    u32* d_out_n[2] = {nullptr, nullptr};
    u64* h_out_keys[2] = {nullptr, nullptr};
    u32* h_out_counts[2] = {nullptr, nullptr};
    u32* h_out_n = nullptr;
    size_t out_cap_dev = 0, out_cap_host = 0;  // records the device / pinned export buffers can hold
    cudaEvent_t ev_out[2] = {nullptr, nullptr};
    // GPU-side text dump (kg_export_text): formatted lines, double-buffered
    char* d_text[2] = {nullptr, nullptr};
    char* h_text[2] = {nullptr, nullptr};
    u64* d_text_cur[2] = {nullptr, nullptr};   // bytes written into d_text[i] by the format kernel
    u64* h_text_cur = nullptr;                 // pinned, 2 words
    size_t text_cap = 0;                       // bytes per text buffer
    bool text_configured = false;
    // minimizer-bucketed path (kg_skm.cuh): several GPUs, or partitions > 1 on one GPU
    bool bucketed = false;              // this context may bucket (slots / streams exist)
    bool pass_bucketed = false;         // the current pass buckets its batches
    u32 nb = 0;                         // buckets of the current pass = world * local partitions
    u32 pl = 1;                         // local partitions of the current pass
    u32 skm_m = 0;                      // minimizer length for this k
    u32 skm_cap = 0;                    // descriptors per (bucket, sub-region) in the current pass
    u32* d_cursors = nullptr;           // reservation counters of the scatter, one per 32-byte sector
    u32 skm_ovf_cap = 0;                // descriptors the overflow list can hold (= every position of a batch)
    u64 skm_region_total = 0, skm_desc_cap = 0;
    size_t skm_words_bytes = 0, skm_slab_bytes = 0;
    SkmSlot slot[2];
    u64 *d_part_lo = nullptr, *d_bpart_lo = nullptr;   // [pl + 1] partition bounds of the table / the Bloom filter
    ncclComm_t comm = nullptr;          // carries only the per-round one-word all-reduce (and the handle exchange)
    cudaStream_t s_insert = nullptr;    // all-reduce, pulls, kg_skm_insert
    u32* d_round = nullptr;             // [0] = 0, [1] = 1 (send values), [4 + parity] = sum of the round
    u32* h_round_sum = nullptr;         // pinned
    cudaEvent_t ev_pass_ready = nullptr, ev_tail = nullptr;
    uint64_t round = 0;
    u32* d_work = nullptr;              // work counter of the persistent insert kernel
    u32 sm_count = 148;
    u32 insert_grid_auto = 0;           // resident blocks per SM of the insert kernel in use (occupancy API, first launch)
    u32 insert_grid_env = 0;            // KG_INSERT_GRID: fewer blocks per SM than that (leaves room for the other stream)
    int insert_occ = 1;                 // KG_INSERT_OCC: register budget of kg_skm_insert (1 = unconstrained, 6 = six blocks per SM)
    // bit-exact emulation of the reference's double Bloom filter (kg_refbloom.cuh)
    bool ref_bloom = false;
    KgRefBloom rb{nullptr, nullptr, 0, 0, 0};
    struct RbBatch { u64* words; u32* brk; KgStream* st; u32 nthreads; };
    std::vector<RbBatch> rb_log;        // packed stream of every batch of the Bloom pass, kept for sweeps 2 and 3
    int rb_streams = 0;                 // streams begun in the current pass (the emulation needs exactly one)
    bool feed_prefetch = true;          // pipelined H2D in kg_feed (KG_FEED_PREFETCH=0 switches it off)
    bool parse_tma = true;              // parse tiles staged by TMA bulk copies (KG_PARSE_TMA=0: plain 128-bit loads)
    // Kaarme representation (after kg_compact)
    KgKaarme kaarme{nullptr, nullptr, 0, 0};
    KgCompactStats* d_cstats = nullptr;
    bool compacted = false, counted = false;
    std::string err;
};

static thread_local std::string g_err;

// KG_TRACE=1: progress of the collective steps on stderr (which call does a hung multi-GPU run sit in?)
static bool kg_trace_on() { static const bool on = getenv("KG_TRACE") != nullptr; return on; }
static bool kg_trace_sync() { static const bool on = getenv("KG_TRACE") && atoi(getenv("KG_TRACE")) >= 2; return on; }
// KG_TRACE=2: additionally wait for the stream after every step of a round (serialises everything; debugging only)
#define KG_TRACE_SYNC(ctx, stream, what)                                                       \
    do {                                                                                       \
        if (kg_trace_sync()) {                                                                 \
            KG_TRACE(ctx, "  ... %s queued", what);                                            \
            cudaError_t e_ = cudaStreamSynchronize(stream);                                    \
            KG_TRACE(ctx, "  ... %s done (%s)", what, cudaGetErrorString(e_));                 \
        }                                                                                      \
    } while (0)
#define KG_TRACE(ctx, ...)                                                                     \
    do {                                                                                       \
        if (kg_trace_on()) {                                                                   \
            fprintf(stderr, "[kg rank %d/%d] ", (ctx)->cfg.rank, (ctx)->cfg.world);            \
            fprintf(stderr, __VA_ARGS__);                                                      \
            fprintf(stderr, "\n");                                                             \
            fflush(stderr);                                                                    \
        }                                                                                      \
    } while (0)

// NCCL is resolved with dlopen at first use instead of a link-time dependency: inside a Python process torch
// has already loaded its own (newer) libnccl.so.2, and a DT_NEEDED on the system copy would either shadow it
// (breaking `import torch`) or be shadowed by it.  dlopen by soname returns whichever copy is already mapped.
struct KgNccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
static KgNccl& kg_nccl() {
    static KgNccl n;
    if (!n.handle) {
        n.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!n.handle) n.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) {
#define KG_SYM(name) n.name = (decltype(n.name))dlsym(n.handle, "nccl" #name)
            KG_SYM(GetUniqueId); KG_SYM(CommInitRank); KG_SYM(CommDestroy); KG_SYM(AllGather); KG_SYM(AllReduce); KG_SYM(Send);
            KG_SYM(Recv); KG_SYM(GroupStart); KG_SYM(GroupEnd); KG_SYM(GetErrorString);
#undef KG_SYM
            n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.AllReduce && n.Send && n.Recv &&
                   n.GroupStart && n.GroupEnd && n.GetErrorString;
        }
    }
    return n;
}

#define KG_CUDA(ctx, call)                                                                            \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char buf_[512];                                                                           \
            snprintf(buf_, sizeof(buf_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            if (ctx) (ctx)->err = buf_;                                                               \
            g_err = buf_;                                                                             \
            return e_ == cudaErrorMemoryAllocation ? KG_ENOMEM : KG_ECUDA;                           \
        }                                                                                             \
    } while (0)

// ---- host math: functions_math.cpp:53-96 (table sizing is part of the drop-in contract) --------------------
static uint64_t next_prime3mod4(uint64_t at_least) {
    uint64_t c = at_least;
    if (c <= 2) return 2;
    if ((c & 1) == 0) c += 1;
    for (;;) {
        bool prime = true;
        for (uint64_t d = 3; d * d <= c; d += 2)
            if (c % d == 0) { prime = false; break; }
        if (prime && (c % 4 == 3)) return c;
        c += 2;
    }
}

// main.cpp:401-418
static void bloom_params(uint64_t U, double fpr, uint64_t* m, uint32_t* nh) {
    double bits_min = (-(double)U * std::log(fpr)) / std::pow(std::log(2.0), 2.0);
    double h = (bits_min / (double)U) * std::log(2.0);
    uint64_t p2 = 2;
    while (p2 < (uint64_t)bits_min) p2 *= 2;
    *m = p2;
    *nh = (uint32_t)std::ceil(h);
}

extern "C" int kg_abi_version(void) { return KG_ABI_VERSION; }

extern "C" const char* kg_strerror(int s) {
    switch (s) {
        case KG_OK: return "ok";
        case KG_EBADARG: return "bad argument or call order";
        case KG_ECUDA: return "CUDA error / no usable sm_100 device";
        case KG_ETABLE_FULL: return "hash table is full";
        case KG_ENCCL: return "NCCL error";
        case KG_ENOMEM: return "out of device or pinned memory";
        case KG_ESINK: return "export sink aborted";
        default: return "unknown status";
    }
}

extern "C" const char* kg_last_error(const kg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int kg_device_count(int* count) {
    if (!count) return KG_EBADARG;
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); cudaGetLastError(); return KG_ECUDA; }
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) (*count)++;
    }
    return KG_OK;
}

extern "C" int kg_host_alloc(size_t bytes, void** out) {
    if (!out) return KG_EBADARG;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); cudaGetLastError(); return KG_ENOMEM; }
    return KG_OK;
}
extern "C" int kg_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return KG_ECUDA; }
    return KG_OK;
}

// -----------------------------------------------------------------------------------------------------------
static void free_slot(kg_ctx* c, SkmSlot& s) {
    for (int r = 0; r < KG_MAX_WORLD; r++) {
        if (s.peer_opened[r]) cudaIpcCloseMemHandle(s.peer_opened[r]);
        s.peer_opened[r] = nullptr;
        cudaFree(s.r_buf[r]);
        s.r_buf[r] = nullptr;
    }
    cudaFree(s.slab); cudaFree(s.d_src); cudaFree(s.d_peers); cudaFree(s.d_seg_start); cudaFree(s.d_seg_ptr);
    if (s.ev_ready) cudaEventDestroy(s.ev_ready);
    if (s.ev_free) cudaEventDestroy(s.ev_free);
    s = SkmSlot();
    (void)c;
}

static void free_all(kg_ctx* c) {
    cudaSetDevice(c->cfg.device);
    for (cudaStream_t st : {c->s_compute, c->s_copy, c->s_insert}) if (st) cudaStreamSynchronize(st);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_raw[i]);
        if (c->h_stage[i]) cudaFreeHost(c->h_stage[i]);
        if (c->ev_raw_free[i]) cudaEventDestroy(c->ev_raw_free[i]);
        if (c->ev_copy_done[i]) cudaEventDestroy(c->ev_copy_done[i]);
        if (c->ev_stage_free[i]) cudaEventDestroy(c->ev_stage_free[i]);
        cudaFree(c->d_out_keys[i]); cudaFree(c->d_out_counts[i]); cudaFree(c->d_out_n[i]);
        if (c->h_out_keys[i]) cudaFreeHost(c->h_out_keys[i]);
        if (c->h_out_counts[i]) cudaFreeHost(c->h_out_counts[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
        cudaFree(c->d_text[i]); cudaFree(c->d_text_cur[i]);
        if (c->h_text[i]) cudaFreeHost(c->h_text[i]);
        free_slot(c, c->slot[i]);
    }
    if (c->h_out_n) cudaFreeHost(c->h_out_n);
    if (c->h_text_cur) cudaFreeHost(c->h_text_cur);
    if (c->h_round_sum) cudaFreeHost(c->h_round_sum);
    cudaFree(c->kaarme.slots); cudaFree(c->kaarme.roots); cudaFree(c->d_cstats); cudaFree(c->d_work);
    cudaFree(c->d_part_lo); cudaFree(c->d_bpart_lo); cudaFree(c->d_round); cudaFree(c->d_cursors);
    for (auto& b : c->rb_log) { cudaFree(b.words); cudaFree(b.brk); cudaFree(b.st); }
    c->rb_log.clear();
    cudaFree(c->rb.T1); cudaFree(c->rb.T2);
    if (c->comm) { kg_nccl().CommDestroy(c->comm); c->comm = nullptr; }
    if (c->ev_tail) cudaEventDestroy(c->ev_tail);
    if (c->ev_pass_ready) cudaEventDestroy(c->ev_pass_ready);
    if (c->s_insert) cudaStreamDestroy(c->s_insert);
    cudaFree(c->d_tile_hdr_eff); cudaFree(c->d_tile_hdr_in); cudaFree(c->d_tile_nbases);
    cudaFree(c->d_tile_pend_eff); cudaFree(c->d_tile_pend_in); cudaFree(c->d_tile_off);
    cudaFree(c->d_words_direct); cudaFree(c->d_brk); cudaFree(c->d_carry_words); cudaFree(c->d_carry_brk);
    cudaFree(c->d_stream); cudaFree(c->d_stats); cudaFree(c->table.slots); cudaFree(c->bloom.bits);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto e : c->ins_pool) cudaEventDestroy(e);
    if (c->ev_pass_begin) cudaEventDestroy(c->ev_pass_begin);
    if (c->ev_pass_end) cudaEventDestroy(c->ev_pass_end);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    cudaGetLastError();
}

// Sizes of the minimizer-bucketing buffers of one batch slot.  A batch of n raw bytes holds at most n bases
// (+ the carried ones), hence at most that many windows and -- worst case, one window per descriptor -- descriptors.
//   regions   nb fixed-capacity bucket regions.  A run averages ~(k-m+2)/2 windows cut at packed-word boundaries; the
//             regions together hold positions/4 * 1.25 descriptors when a window has >= 8 m-mers (positions * 1.25
//             otherwise) + 256 per bucket: comfortably more than a balanced batch needs.
//   overflow  whatever does not fit its region (skewed input) goes to one shared list that can hold EVERY position of
//             the batch, so no input can overflow it; the insert walks it last, without L2 blocking.
static void skm_geometry(kg_ctx* c) {
    const u64 positions = c->batch_bytes + 32ull * c->carry_max_words + 64;
    const u32 wlen = c->cfg.k - c->skm_m + 1;
    c->skm_region_total = (wlen >= 8 ? positions / 4 : positions) * 5 / 4 + (u64)KG_MAX_BUCKETS * KG_SKM_SUB * 64;
    c->skm_ovf_cap = (u32)positions;
    c->skm_words_bytes = ((c->words_cap * sizeof(u64)) + 255) / 256 * 256;
    c->skm_desc_cap = c->skm_region_total + c->skm_ovf_cap + 64;
    c->skm_slab_bytes = KG_SKM_META + c->skm_words_bytes + c->skm_desc_cap * sizeof(u64);
}

static int alloc_slots(kg_ctx* c) {
    if (c->slot[0].slab) return KG_OK;
    for (int b = 0; b < 2; b++) {
        SkmSlot& s = c->slot[b];
        KG_CUDA(c, cudaMalloc(&s.slab, c->skm_slab_bytes));
        KG_CUDA(c, cudaMemset(s.slab, 0, KG_SKM_META));
        s.hdr = (u64*)s.slab;
        s.counts = (u32*)(s.slab + 64);
        s.words = (u64*)(s.slab + KG_SKM_META);
        s.desc = (u64*)(s.slab + KG_SKM_META + c->skm_words_bytes);
        KG_CUDA(c, cudaMalloc(&s.d_src, sizeof(KgSkmSources)));
        KG_CUDA(c, cudaMalloc(&s.d_peers, sizeof(KgSkmPeers)));
        KG_CUDA(c, cudaMalloc(&s.d_seg_start, sizeof(u64) * (KG_SKM_MAXSEG + 2)));
        KG_CUDA(c, cudaMalloc(&s.d_seg_ptr, sizeof(u64*) * (KG_SKM_MAXSEG + 2)));
        KG_CUDA(c, cudaEventCreateWithFlags(&s.ev_ready, cudaEventDisableTiming));
        KG_CUDA(c, cudaEventCreateWithFlags(&s.ev_free, cudaEventDisableTiming));
        s.peer_slab[c->cfg.rank] = s.slab;
    }
    return KG_OK;
}

// the device-side tables that say where every rank's words / counts / descriptors of a slot are to be found
static int publish_slot_tables(kg_ctx* c) {
    const int world = c->cfg.world, me = c->cfg.rank;
    for (int b = 0; b < 2; b++) {
        SkmSlot& s = c->slot[b];
        KgSkmSources src;
        KgSkmPeers peers;
        memset(&src, 0, sizeof(src));
        memset(&peers, 0, sizeof(peers));
        for (int r = 0; r < world; r++) {
            const uint8_t* meta = r == me ? s.slab : s.r_buf[r];                 // local (copied) header + counts + packed words
            src.hdr[r] = (const u64*)meta;
            src.words[r] = (const u64*)(meta + KG_SKM_META);
            peers.counts[r] = (const u32*)(meta + 64);
            peers.desc[r] = (const u64*)(s.peer_slab[r] + KG_SKM_META + c->skm_words_bytes);   // read in place (NVLink for r != me)
        }
        KG_CUDA(c, cudaMemcpy(s.d_src, &src, sizeof(src), cudaMemcpyHostToDevice));
        KG_CUDA(c, cudaMemcpy(s.d_peers, &peers, sizeof(peers), cudaMemcpyHostToDevice));
    }
    return KG_OK;
}

extern "C" int kg_create(const kg_config* cfg, kg_ctx** out) {
    if (!cfg || !out) return KG_EBADARG;
    *out = nullptr;
    if (cfg->abi_version != KG_ABI_VERSION) { g_err = "abi_version mismatch"; return KG_EBADARG; }
    if (cfg->k < 1 || (cfg->k + 31) / 32 > KG_MAX_W) { g_err = "k must be in [1, 256]"; return KG_EBADARG; }
    if (cfg->table_mode != KG_TABLE_PLAIN && cfg->table_mode != KG_TABLE_KAARME) { g_err = "table_mode must be 0 or 2"; return KG_EBADARG; }
    if (cfg->input_mode != KG_INPUT_FASTA && cfg->input_mode != KG_INPUT_PLAIN) { g_err = "input_mode must be 0 or 2"; return KG_EBADARG; }
    if (cfg->use_bloom) {
        if (cfg->expected_unique == 0 || !(cfg->fpr > 0.0 && cfg->fpr < 1.0)) { g_err = "bloom needs expected_unique > 0 and 0 < fpr < 1"; return KG_EBADARG; }
    } else if (cfg->min_slots == 0) { g_err = "min_slots must be > 0"; return KG_EBADARG; }
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) { g_err = "bad rank/world"; return KG_EBADARG; }
    if (cfg->world > 64 || (uint64_t)cfg->partitions * (uint64_t)cfg->world > KG_MAX_BUCKETS) { g_err = "world must be <= 64 and world * partitions <= 1024"; return KG_EBADARG; }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        g_err = e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range";
        cudaGetLastError();
        return KG_ECUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) {
        g_err = "device is not compute capability 10.x (this library ships sm_100a code only; there is no fallback)";
        cudaGetLastError();
        return KG_ECUDA;
    }
    kg_ctx* c = new kg_ctx();
    c->cfg = *cfg;
    c->sm_count = (u32)prop.multiProcessorCount;
    if (const char* e = getenv("KG_INSERT_GRID")) c->insert_grid_env = (u32)atoi(e);
    // register budget and resident blocks of the persistent insert kernel (profiles/r02_overlap_sweep.txt): two-word keys
    // run best capped at 48 registers with 4 of the 5 possible blocks per SM -- the rest of the SM then takes blocks of
    // the next batch's parse + bucketing kernels; wider keys would spill under that cap
    c->W = (int)((cfg->k + 31) / 32);
    c->insert_occ = c->W <= 2 ? 5 : 1;
    c->insert_grid_env = c->W <= 2 ? 4 : 0;
    if (const char* e = getenv("KG_INSERT_OCC")) c->insert_occ = atoi(e);
    if (const char* e = getenv("KG_PARSE_TMA")) c->parse_tma = atoi(e) != 0;
    if (const char* e = getenv("KG_FEED_PREFETCH")) c->feed_prefetch = atoi(e) != 0;
    c->W = (int)((cfg->k + 31) / 32);
    c->skm_m = kg_skm_m(cfg->k);
    c->batch_bytes = cfg->batch_bytes ? cfg->batch_bytes : KG_DEFAULT_BATCH;
    if (c->batch_bytes > KG_MAX_BATCH) c->batch_bytes = KG_MAX_BATCH;
    c->batch_bytes = (c->batch_bytes + KG_TILE - 1) / KG_TILE * KG_TILE;
    c->max_tiles = (uint32_t)(c->batch_bytes / KG_TILE);
    c->carry_max_words = (u32)c->W + 3;
    c->words_cap = c->batch_bytes / 32 + c->carry_max_words + 8;

#define KG_TRY(call)                              \
    do {                                          \
        int s_ = [&]() -> int { KG_CUDA(c, call); return KG_OK; }(); \
        if (s_ != KG_OK) { g_err = c->err; free_all(c); delete c; return s_; } \
    } while (0)

    KG_TRY(cudaSetDevice(cfg->device));
    {   // parse + bucketing of the NEXT batch get the SM slots the persistent insert leaves free: higher priority
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        KG_TRY(cudaStreamCreateWithPriority(&c->s_compute, cudaStreamNonBlocking, hi));
    }
    KG_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        KG_TRY(cudaMalloc(&c->d_raw[i], c->batch_bytes));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_raw_free[i], cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_copy_done[i], cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_stage_free[i], cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
    KG_TRY(cudaMalloc(&c->d_tile_hdr_eff, sizeof(u32) * c->max_tiles));
    KG_TRY(cudaMalloc(&c->d_tile_hdr_in, sizeof(u32) * c->max_tiles));
    KG_TRY(cudaMalloc(&c->d_tile_nbases, sizeof(u32) * c->max_tiles));
    KG_TRY(cudaMalloc(&c->d_tile_pend_eff, sizeof(u32) * c->max_tiles));
    KG_TRY(cudaMalloc(&c->d_tile_pend_in, sizeof(u32) * c->max_tiles));
    KG_TRY(cudaMalloc(&c->d_tile_off, sizeof(u32) * c->max_tiles));
    KG_TRY(cudaMalloc(&c->d_brk, sizeof(u32) * c->words_cap));
    KG_TRY(cudaMalloc(&c->d_carry_words, sizeof(u64) * c->carry_max_words));
    KG_TRY(cudaMalloc(&c->d_carry_brk, sizeof(u32) * c->carry_max_words));
    KG_TRY(cudaMalloc(&c->d_stream, sizeof(KgStream)));
    KG_TRY(cudaMalloc(&c->d_stats, sizeof(KgStats)));
    KG_TRY(cudaMalloc(&c->d_work, sizeof(u32) * 4));
    KG_TRY(cudaMalloc(&c->d_part_lo, sizeof(u64) * (KG_MAX_BUCKETS + 2)));
    KG_TRY(cudaMalloc(&c->d_bpart_lo, sizeof(u64) * (KG_MAX_BUCKETS + 2)));
    KG_TRY(cudaMemset(c->d_stream, 0, sizeof(KgStream)));
    KG_TRY(cudaMemset(c->d_stats, 0, sizeof(KgStats)));
    KG_TRY(cudaEventCreate(&c->ev_pass_begin));
    KG_TRY(cudaEventCreate(&c->ev_pass_end));
    c->ref_bloom = (cfg->reserved & KG_CFG_REFERENCE_BLOOM) != 0;
    if (c->ref_bloom) {
        uint64_t m; uint32_t nh;
        bloom_params(cfg->expected_unique, cfg->fpr, &m, &nh);
        const double h = (-(double)cfg->expected_unique * std::log(cfg->fpr)) / std::pow(std::log(2.0), 2.0) / (double)cfg->expected_unique * std::log(2.0);
        if (!cfg->use_bloom || cfg->world != 1 || m > (1ull << 31) || nh < 1 || nh > KG_RB_MAX_NH) {
            g_err = "KG_CFG_REFERENCE_BLOOM needs use_bloom, one GPU, m <= 2^31 bits and at most 16 hash functions";
            free_all(c); delete c; return KG_EBADARG;
        }
        c->bloom_m = m;
        c->bloom.nh = nh;
        c->rb.mask = m - 1;
        c->rb.nh = nh;
        c->rb.nh2 = (u32)std::floor(h);                    // main.cpp:472 passes the double; the test loop truncates it
        KG_TRY(cudaMalloc(&c->rb.T1, sizeof(u32) * m));
        KG_TRY(cudaMalloc(&c->rb.T2, sizeof(u32) * m));
    } else if (cfg->use_bloom) {
        uint64_t m; uint32_t nh;
        bloom_params(cfg->expected_unique, cfg->fpr, &m, &nh);
        if (nh < 1) nh = 1;
        if (nh > 16) nh = 16;
        // every shard holds m/world bits per filter (rounded up to whole 256-bit blocks, at least one)
        uint64_t m_local = (m + (uint64_t)cfg->world - 1) / (uint64_t)cfg->world;
        uint64_t nblocks = (m_local + 255) / 256;
        if (nblocks == 0) nblocks = 1;
        c->bloom_m = m;
        c->bloom.nblocks = nblocks;
        c->bloom.nh = nh;
        c->bloom.world = (u32)cfg->world;
        c->bloom_bytes = nblocks * 64;                // [F1 block][F2 block] pairs: 2m bits, as the reference's interleaved array
        KG_TRY(cudaMalloc(&c->bloom.bits, c->bloom_bytes));
    }
    // partitions: 0 = choose per pass from the table / filter size, 1 = never bucket on one GPU, > 1 = as given
    c->bucketed = cfg->world > 1 || (cfg->partitions != 1 && !c->ref_bloom);
    if (c->bucketed) {
        skm_geometry(c);
        KG_TRY(cudaStreamCreateWithFlags(&c->s_insert, cudaStreamNonBlocking));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_pass_ready, cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_tail, cudaEventDisableTiming));
        KG_TRY(cudaMalloc(&c->d_round, sizeof(u32) * 8));
        KG_TRY(cudaMalloc(&c->d_cursors, sizeof(u32) * 8 * ((size_t)KG_MAX_BUCKETS * KG_SKM_SUB + 1)));
        {   // d_round[0] = 0, d_round[1] = 1: the two possible contributions to a round's "ranks with a batch" sum
            const u32 init[8] = {0, 1, 0, 0, 0, 0, 0, 0};
            KG_TRY(cudaMemcpy(c->d_round, init, sizeof(init), cudaMemcpyHostToDevice));
        }
        KG_TRY(cudaHostAlloc((void**)&c->h_round_sum, sizeof(u32) * 2, cudaHostAllocDefault));
        if (cfg->world == 1) {   // one GPU: slots now; several GPUs: kg_comm_init allocates and maps them
            int rc = alloc_slots(c);
            if (rc == KG_OK) rc = publish_slot_tables(c);
            if (rc != KG_OK) { g_err = c->err; free_all(c); delete c; return rc; }
        }
    } else {
        KG_TRY(cudaMalloc(&c->d_words_direct, sizeof(u64) * c->words_cap));
    }
#undef KG_TRY
    *out = c;
    return KG_OK;
}

extern "C" int kg_destroy(kg_ctx* c) {
    if (!c) return KG_OK;
    free_all(c);
    delete c;
    return KG_OK;
}

#define KG_NCCL(ctx, call)                                                                           \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) {                                                                      \
            char buf_[512];                                                                           \
            snprintf(buf_, sizeof(buf_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, kg_nccl().GetErrorString(r_)); \
            if (ctx) (ctx)->err = buf_;                                                               \
            g_err = buf_;                                                                             \
            return KG_ENCCL;                                                                          \
        }                                                                                             \
    } while (0)

extern "C" int kg_comm_unique_id(void* id_out) {
    if (!id_out) return KG_EBADARG;
    static_assert(sizeof(ncclUniqueId) <= KG_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    kg_ctx* none = nullptr;
    if (!kg_nccl().ok) { g_err = "libnccl.so.2 could not be loaded"; return KG_ENCCL; }
    KG_NCCL(none, kg_nccl().GetUniqueId(&id));
    memset(id_out, 0, KG_UNIQUE_ID_BYTES);
    memcpy(id_out, &id, sizeof(id));
    return KG_OK;
}

// Load every kernel a round launches NOW (cudaFuncGetAttributes forces the load).  CUDA loads kernels lazily, at their
// first launch, and loading may have to synchronise the whole process; with several GPUs driven by threads of ONE process
// that first launch can come while another GPU's NCCL kernel spins waiting for this thread -- a deadlock (seen on
// 2 B200s: `kaarme --gpus 2` hung in round 0 unless CUDA_MODULE_LOADING=EAGER).  After this, and one warm-up all-reduce,
// nothing is loaded inside a round any more.
template <int W>
static void preload_insert(int occ) {
    cudaFuncAttributes at;
    if (occ >= 6) {
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_TABLE, 6>);
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_BLOOM1, 6>);
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_BLOOM2, 6>);
    } else if (occ == 5) {
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_TABLE, 5>);
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_BLOOM1, 5>);
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_BLOOM2, 5>);
    } else {
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_TABLE, 1>);
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_BLOOM1, 1>);
        cudaFuncGetAttributes(&at, kg_skm_insert<W, KG_SINK_BLOOM2, 1>);
    }
}
static void preload_round_kernels(kg_ctx* c) {
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, kg_hdr_summary<true>); cudaFuncGetAttributes(&at, kg_hdr_summary<false>);
    cudaFuncGetAttributes(&at, kg_tile_count<true, true>); cudaFuncGetAttributes(&at, kg_tile_count<true, false>);
    cudaFuncGetAttributes(&at, kg_tile_count<false, true>); cudaFuncGetAttributes(&at, kg_tile_count<false, false>);
    cudaFuncGetAttributes(&at, kg_tile_pack<true, true>); cudaFuncGetAttributes(&at, kg_tile_pack<true, false>);
    cudaFuncGetAttributes(&at, kg_tile_pack<false, true>); cudaFuncGetAttributes(&at, kg_tile_pack<false, false>);
    cudaFuncGetAttributes(&at, kg_lww_scan); cudaFuncGetAttributes(&at, kg_tile_scan);
    cudaFuncGetAttributes(&at, kg_carry_save); cudaFuncGetAttributes(&at, kg_carry_restore);
    cudaFuncGetAttributes(&at, kg_skm_scatter); cudaFuncGetAttributes(&at, kg_skm_pack_counts); cudaFuncGetAttributes(&at, kg_skm_segments);
    KG_DISPATCH_W(c->W, preload_insert, c->insert_occ);
    cudaGetLastError();
}

// What one rank tells the others about its two batch slots (all-gathered through NCCL inside kg_comm_init).
struct KgPeerHandle {
    uint64_t magic, pid;
    int32_t rank, device;
    uint64_t ptr[2], slab_bytes;
    cudaIpcMemHandle_t ipc[2];
};
#define KG_PEER_MAGIC 0x4B47504545523032ULL   // "KGPEER02"
#define KG_PEER_HANDLE_BYTES 256
static_assert(sizeof(KgPeerHandle) <= KG_PEER_HANDLE_BYTES, "peer handle size");

// Collective.  Besides the NCCL communicator (which only carries the one-word "round" all-reduce) this maps every
// rank's two batch slots into every rank: CUDA IPC handles between processes, plain peer access between the contexts
// of one process.  The exchange itself then needs no NCCL data movement: the copy engines pull the packed words of the
// peers' batches over NVLink and the insert kernel reads the peers' descriptors in place.
extern "C" int kg_comm_init(kg_ctx* c, const void* id, int rank, int world) {
    if (!c || !id) return KG_EBADARG;
    if (rank != c->cfg.rank || world != c->cfg.world || world < 2) { c->err = "kg_comm_init: rank/world differ from kg_config"; return KG_EBADARG; }
    if (c->comm) { c->err = "kg_comm_init: called twice"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    if (!kg_nccl().ok) { c->err = "libnccl.so.2 could not be loaded"; return KG_ENCCL; }
    KG_TRACE(c, "comm_init: ncclCommInitRank");
    KG_NCCL(c, kg_nccl().CommInitRank(&c->comm, world, uid, rank));
    KG_TRACE(c, "comm_init: slots (%zu bytes each)", c->skm_slab_bytes);
    { int rc = alloc_slots(c); if (rc) return rc; }
    KgPeerHandle mine;
    memset(&mine, 0, sizeof(mine));
    mine.magic = KG_PEER_MAGIC;
    mine.pid = (uint64_t)getpid();
    mine.rank = rank;
    mine.device = c->cfg.device;
    mine.slab_bytes = c->skm_slab_bytes;
    for (int b = 0; b < 2; b++) {
        mine.ptr[b] = (uint64_t)(uintptr_t)c->slot[b].slab;
        KG_CUDA(c, cudaIpcGetMemHandle(&mine.ipc[b], c->slot[b].slab));
    }
    char *d_all = nullptr;
    std::vector<char> all((size_t)world * KG_PEER_HANDLE_BYTES, 0);
    KG_CUDA(c, cudaMalloc(&d_all, all.size()));
    struct Guard { char** p; ~Guard() { cudaFree(*p); } } guard{&d_all};
    KG_CUDA(c, cudaMemset(d_all, 0, all.size()));
    KG_CUDA(c, cudaMemcpy(d_all + (size_t)rank * KG_PEER_HANDLE_BYTES, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    KG_TRACE(c, "comm_init: all-gather of the slot handles");
    KG_NCCL(c, kg_nccl().AllGather(d_all + (size_t)rank * KG_PEER_HANDLE_BYTES, d_all, KG_PEER_HANDLE_BYTES, ncclChar, c->comm, c->s_insert));
    KG_CUDA(c, cudaStreamSynchronize(c->s_insert));
    KG_TRACE(c, "comm_init: mapping the peers' slots");
    KG_CUDA(c, cudaMemcpy(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost));
    for (int r = 0; r < world; r++) {
        if (r == rank) continue;
        KgPeerHandle h;
        memcpy(&h, all.data() + (size_t)r * KG_PEER_HANDLE_BYTES, sizeof(h));
        if (h.magic != KG_PEER_MAGIC || h.rank != r || h.slab_bytes != c->skm_slab_bytes) {
            c->err = "kg_comm_init: ranks disagree on the slot layout (k and batch_bytes must be the same on every rank)";
            return KG_EBADARG;
        }
        for (int b = 0; b < 2; b++) {
            SkmSlot& s = c->slot[b];
            if (h.pid == mine.pid) {                        // another context of this process (one host thread per GPU)
                int can = 0;
                KG_CUDA(c, cudaDeviceCanAccessPeer(&can, c->cfg.device, h.device));
                if (!can) { c->err = "kg_comm_init: no peer access between the devices"; return KG_ECUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) KG_CUDA(c, e);
                cudaGetLastError();
                s.peer_slab[r] = (uint8_t*)(uintptr_t)h.ptr[b];
            } else {                                        // another process: map its allocation
                void* p = nullptr;
                KG_CUDA(c, cudaIpcOpenMemHandle(&p, h.ipc[b], cudaIpcMemLazyEnablePeerAccess));
                s.peer_opened[r] = p;
                s.peer_slab[r] = (uint8_t*)p;
            }
            KG_CUDA(c, cudaMalloc(&s.r_buf[r], KG_SKM_META + c->skm_words_bytes));
        }
    }
    { int rc = publish_slot_tables(c); if (rc) return rc; }
    preload_round_kernels(c);
    // warm-up all-reduce: NCCL loads that kernel and sets up its channels here, where every rank is at the same point
    KG_NCCL(c, kg_nccl().AllReduce(c->d_round, c->d_round + 6, 1, ncclUint32, ncclSum, c->comm, c->s_insert));
    KG_CUDA(c, cudaStreamSynchronize(c->s_insert));
    KG_TRACE(c, "comm_init: done");
    return KG_OK;
}

// Decide how the coming pass buckets its batches.  region_bytes = what the inserts of this pass hit at random (count
// table, or the Bloom filter): local partitions are chosen so that one partition's region is ~16-32 MiB, comfortably
// L2-resident next to the streamed descriptors.  Must come out the same on every rank (it fixes the bucket numbering).
static int setup_pass_buckets(kg_ctx* c, size_t region_bytes) {
    c->pass_bucketed = false;
    c->pl = 1;
    c->nb = (u32)c->cfg.world;
    if (!c->bucketed) return KG_OK;
    const u32 world = (u32)c->cfg.world;
    u32 pl = c->cfg.partitions;
    if (pl == 0) {
        // Bloom mode: the filter words of a k-mer are placed by its partition, so the Bloom pass and the count pass must
        // partition alike: size for the larger of the filter and the table the configuration promises (2 x expected
        // distinct k-mers, main.cpp:454), whatever the pass
        if (c->cfg.use_bloom) {
            const size_t est_table = (size_t)(2 * c->cfg.expected_unique / (uint64_t)world) *
                                     kg_slot_stride_words((u32)c->W, c->cfg.table_mode == KG_TABLE_KAARME) * sizeof(u64);
            region_bytes = est_table > c->bloom_bytes ? est_table : c->bloom_bytes;
        }
        pl = 1;
        if (region_bytes > (96u << 20)) while ((size_t)pl * (24u << 20) < region_bytes && pl * 2 * world <= KG_MAX_BUCKETS) pl *= 2;
    }
    if (pl * world > KG_MAX_BUCKETS) pl = KG_MAX_BUCKETS / world;
    if (pl < 1) pl = 1;
    c->pl = pl;
    c->nb = world * pl;
    c->pass_bucketed = world > 1 || pl > 1;
    c->skm_cap = (u32)(c->skm_region_total / ((u64)c->nb * KG_SKM_SUB));   // descriptors per (bucket, sub-region)
    return KG_OK;
}

// part_lo[p] = first slot (or Bloom word) of partition p of `total`
static int upload_partition_bounds(kg_ctx* c, u64* d_dst, u64 total) {
    std::vector<u64> lo(c->pl + 1);
    for (u32 p = 0; p <= c->pl; p++) lo[p] = (u64)(((unsigned __int128)total * p) / c->pl);
    KG_CUDA(c, cudaMemcpyAsync(d_dst, lo.data(), sizeof(u64) * lo.size(), cudaMemcpyHostToDevice, c->s_compute));
    KG_CUDA(c, cudaStreamSynchronize(c->s_compute));
    return KG_OK;
}

extern "C" int kg_pass_begin(kg_ctx* c, int pass) {
    if (!c || (pass != KG_PASS_BLOOM && pass != KG_PASS_COUNT)) return KG_EBADARG;
    if (pass == KG_PASS_BLOOM && !c->cfg.use_bloom) { c->err = "Bloom pass without use_bloom"; return KG_EBADARG; }
    if (pass == KG_PASS_COUNT && c->cfg.use_bloom && !c->bloom_done) { c->err = "count pass before Bloom pass"; return KG_EBADARG; }
    if (c->cfg.world > 1 && !c->comm) { c->err = "world > 1 needs kg_comm_init first"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    KG_TRACE(c, "pass_begin %d", pass);
    c->pass = pass;
    c->stream_open = false;
    c->ev_used = 0;
    c->ins_used = 0;
    c->ins_launches = 0;
    c->raw_bytes_pass = 0;
    c->round = 0;
    KG_CUDA(c, cudaEventRecord(c->ev_pass_begin, c->s_compute));
    KG_CUDA(c, cudaMemsetAsync(c->d_stats, 0, sizeof(KgStats), c->s_compute));
    c->rb_streams = 0;
    if (pass == KG_PASS_BLOOM && c->ref_bloom) {
        KG_CUDA(c, cudaMemsetAsync(c->rb.T1, 0xFF, sizeof(u32) * (c->rb.mask + 1), c->s_compute));
        KG_CUDA(c, cudaMemsetAsync(c->rb.T2, 0xFF, sizeof(u32) * (c->rb.mask + 1), c->s_compute));
        for (auto& b : c->rb_log) { cudaFree(b.words); cudaFree(b.brk); cudaFree(b.st); }
        c->rb_log.clear();
        c->bloom_done = false;
        c->pass_bucketed = false; c->pl = 1; c->nb = 1;
    } else if (pass == KG_PASS_BLOOM) {
        KG_CUDA(c, cudaMemsetAsync(c->bloom.bits, 0, c->bloom_bytes, c->s_compute));
        c->bloom_done = false;
        { int rc = setup_pass_buckets(c, c->bloom_bytes); if (rc) return rc; }
        { int rc = upload_partition_bounds(c, c->d_bpart_lo, c->bloom.nblocks); if (rc) return rc; }
    } else {
        if (c->compacted) {
            cudaFree(c->kaarme.slots); cudaFree(c->kaarme.roots);
            c->kaarme = KgKaarme{nullptr, nullptr, 0, 0};
            c->compacted = false;
        }
        c->counted = false;
        uint64_t want;
        if (c->cfg.use_bloom) want = 2 * c->new_in_second;                      // main.cpp:454
        else want = (c->cfg.min_slots + (uint64_t)c->cfg.world - 1) / (uint64_t)c->cfg.world;
        uint64_t nslots = next_prime3mod4(want);                                 // parallel_parser.hpp:236
        u32 stride = kg_slot_stride_words((u32)c->W, c->cfg.table_mode == KG_TABLE_KAARME);
        // packed 16-byte slots: two key words and >= 26 spare bits for the count (k = 33..51), plain table only;
        // the count field stops at 2^(128-2k) - 2^20 (>= 66 M) instead of wrapping into the key
        u32 packed_tb = 0;
        if (c->W == 2 && c->cfg.table_mode == KG_TABLE_PLAIN && 2 * c->cfg.k - 64 <= 38 && !getenv("KG_NO_PACKED")) {
            packed_tb = 2 * c->cfg.k - 64;
            stride = 2;
        }
        size_t bytes = (size_t)nslots * stride * sizeof(u64);
        // keep the allocation when the new table fits and is not much smaller (Bloom mode sizes the table from
        // new_in_second, which moves a little from run to run: do not pay cudaFree + cudaMalloc of GBs for that)
        if (!c->table.slots || bytes > c->table_bytes || bytes < c->table_bytes / 2) {
            if (c->table.slots) { KG_CUDA(c, cudaFree(c->table.slots)); c->table.slots = nullptr; }
            c->table_bytes = bytes + bytes / 64;
            KG_CUDA(c, cudaMalloc(&c->table.slots, c->table_bytes ? c->table_bytes : 16));
        }
        c->table.nslots = nslots;
        c->table.stride = stride;
        c->table.kaarme = c->cfg.table_mode == KG_TABLE_KAARME;
        c->table.world = 1;            // slots are placed by partition bounds; the hash is not split between shards any more
        c->table.packed_tb = packed_tb;
        c->table.full_flag = &c->d_stats->table_full;
        KG_CUDA(c, cudaMemsetAsync(c->table.slots, 0, bytes, c->s_compute));
        {
            // the partition count must be identical on every rank (it fixes the bucket numbering of the exchange):
            // derive it from configuration only.  After a Bloom pass the shard tables differ a little in size
            // (2 x the LOCAL new_in_second), so use the configured estimate there.
            size_t region = bytes;
            if (c->cfg.world > 1 && c->cfg.use_bloom)
                region = (size_t)(2 * c->cfg.expected_unique / (uint64_t)c->cfg.world) * stride * sizeof(u64);
            int rc = setup_pass_buckets(c, region);
            if (rc) return rc;
        }
        { int rc = upload_partition_bounds(c, c->d_part_lo, nslots); if (rc) return rc; }
        if (c->cfg.use_bloom && !c->ref_bloom) { int rc = upload_partition_bounds(c, c->d_bpart_lo, c->bloom.nblocks); if (rc) return rc; }
    }
    c->bloom.world = 1;
    if (c->pass_bucketed) {   // inserts run on their own stream: order them after the clears above
        KG_CUDA(c, cudaEventRecord(c->ev_pass_ready, c->s_compute));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_insert, c->ev_pass_ready, 0));
    }
    return KG_OK;
}

extern "C" int kg_stream_begin(kg_ctx* c, int starts_in_header) {
    if (!c || !c->pass) return KG_EBADARG;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    if (c->ref_bloom && ++c->rb_streams > 1) { c->err = "KG_CFG_REFERENCE_BLOOM: one stream per pass (window ordinals are positions in it)"; return KG_EBADARG; }
    KgStream s;
    memset(&s, 0, sizeof(s));
    s.in_header = starts_in_header ? 1u : 0u;
    s.pending_break = 1u;  // the first base of a stream starts a run
    // pageable -> async copy is staged by the runtime before returning, so a stack source is fine
    KG_CUDA(c, cudaMemcpyAsync(c->d_stream, &s, sizeof(s), cudaMemcpyHostToDevice, c->s_compute));
    KG_CUDA(c, cudaStreamSynchronize(c->s_compute));
    c->stream_open = true;
    return KG_OK;
}

static cudaEvent_t next_ins_event(kg_ctx* c) {
    if (c->ins_used == c->ins_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        c->ins_pool.push_back(e);
    }
    return c->ins_pool[c->ins_used++];
}

static cudaEvent_t next_event(kg_ctx* c) {
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}

template <int W>
static void launch_count(kg_ctx* c, const KgCountArgs& a, u32 nthreads_words, int sink) {
    const u32 block = 256;
    const u32 grid = (nthreads_words + block - 1) / block;
    if (grid == 0) return;
    switch (sink) {
        case KG_SINK_TABLE: kg_count_kernel<W, KG_SINK_TABLE><<<grid, block, 0, c->s_compute>>>(a); break;
        case KG_SINK_BLOOM1: kg_count_kernel<W, KG_SINK_BLOOM1><<<grid, block, 0, c->s_compute>>>(a); break;
        case KG_SINK_BLOOM2: kg_count_kernel<W, KG_SINK_BLOOM2><<<grid, block, 0, c->s_compute>>>(a); break;
        default: break;
    }
    c->launches++;
}

static int current_sink(const kg_ctx* c) {
    return c->pass == KG_PASS_BLOOM ? KG_SINK_BLOOM1 : (c->cfg.use_bloom ? KG_SINK_BLOOM2 : KG_SINK_TABLE);
}

template <int W, int MINB>
static void launch_skm_insert_b(kg_ctx* c, const KgSkmInsertArgs& a, int sink) {
    // Persistent kernel: exactly as many blocks as can be resident (SMs x occupancy), or fewer per SM when KG_INSERT_GRID
    // says so.  Never more: blocks of this grid that wait for a slot are served before the blocks of the NEXT batch's
    // parse / bucketing kernels on the other stream, which would then never run beside the insert.
    if (!c->insert_grid_auto) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kg_skm_insert<W, KG_SINK_TABLE, MINB>, 256, 0);
        if (per_sm < 1) per_sm = 1;
        c->insert_grid_auto = (u32)per_sm;
    }
    u32 per_sm = c->insert_grid_auto;
    if (c->insert_grid_env && c->insert_grid_env < per_sm) per_sm = c->insert_grid_env;
    const u32 grid = c->sm_count * per_sm;
    switch (sink) {
        case KG_SINK_TABLE: kg_skm_insert<W, KG_SINK_TABLE, MINB><<<grid, 256, 0, c->s_insert>>>(a); break;
        case KG_SINK_BLOOM1: kg_skm_insert<W, KG_SINK_BLOOM1, MINB><<<grid, 256, 0, c->s_insert>>>(a); break;
        case KG_SINK_BLOOM2: kg_skm_insert<W, KG_SINK_BLOOM2, MINB><<<grid, 256, 0, c->s_insert>>>(a); break;
        default: break;
    }
    c->launches++;
}
// two register budgets of the same kernel: as many registers as it wants (MINB = 1), or capped for six resident blocks
// per SM; KG_INSERT_OCC selects (measured in profiles/)
template <int W>
static void launch_skm_insert(kg_ctx* c, const KgSkmInsertArgs& a, int sink) {
    if (c->insert_occ >= 6) launch_skm_insert_b<W, 6>(c, a, sink);
    else if (c->insert_occ == 5) launch_skm_insert_b<W, 5>(c, a, sink);
    else launch_skm_insert_b<W, 1>(c, a, sink);
}

// One round of the minimizer-bucketed path (collective when world > 1; every rank issues the same sequence).
//   s_compute  (have_batch: parse + kg_skm_scatter of this rank's batch were just queued into slot b = round & 1;
//              otherwise the slot's counts are cleared: the rank contributes nothing)            -> ev_ready
//   s_insert   world > 1: one-word all-reduce = "every rank's slot b is ready, and every rank has finished the insert
//              of the previous round" (so the OTHER slot may be overwritten: ev_free), its sum = ranks that had a batch;
//              then the copy engines pull header + counts + packed words of every peer's slot over NVLink
//              kg_skm_segments (where are the descriptors for my partitions, partition-major across senders) and
//              kg_skm_insert, which reads the peers' descriptors in place
// No host synchronisation anywhere: the host only waits when kg_pass_end asks for the all-reduce's sum.
static int skm_round(kg_ctx* c, bool have_batch, bool want_sum) {
    const int b = (int)(c->round & 1);
    SkmSlot& s = c->slot[b];
    const int world = c->cfg.world, me = c->cfg.rank;
    if (world > 1) KG_TRACE(c, "round %llu: %s", (unsigned long long)c->round, have_batch ? "batch" : "no batch");
    if (!have_batch) {
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, s.ev_free, 0));
        KG_CUDA(c, cudaMemsetAsync(s.counts, 0, sizeof(u32) * (c->nb * KG_SKM_SUB + 1), c->s_compute));
    }
    KG_CUDA(c, cudaEventRecord(s.ev_ready, c->s_compute));
    KG_TRACE_SYNC(c, c->s_compute, "parse + scatter");
    KG_CUDA(c, cudaStreamWaitEvent(c->s_insert, s.ev_ready, 0));
    if (world > 1) {
        KG_NCCL(c, kg_nccl().AllReduce(c->d_round + (have_batch ? 1 : 0), c->d_round + 4 + b, 1, ncclUint32, ncclSum, c->comm, c->s_insert));
        KG_TRACE_SYNC(c, c->s_insert, "all-reduce");
        KG_CUDA(c, cudaEventRecord(c->slot[b ^ 1].ev_free, c->s_insert));
        if (want_sum) KG_CUDA(c, cudaMemcpyAsync(c->h_round_sum, c->d_round + 4 + b, sizeof(u32), cudaMemcpyDeviceToHost, c->s_insert));
        for (int r = 0; r < world; r++)
            if (r != me) KG_CUDA(c, cudaMemcpyAsync(s.r_buf[r], s.peer_slab[r], KG_SKM_META + c->skm_words_bytes, cudaMemcpyDefault, c->s_insert));
        KG_TRACE_SYNC(c, c->s_insert, "pulls");
    }
    const u32 nseg = (u32)world * c->pl * KG_SKM_SUB + (u32)world;
    kg_skm_segments<<<1, 1024, 0, c->s_insert>>>(s.d_peers, (u32)world, (u32)me, c->pl, c->nb, c->skm_cap, s.d_seg_start, s.d_seg_ptr);
    c->launches++;
    KG_CUDA(c, cudaMemsetAsync(c->d_work, 0, sizeof(u32), c->s_insert));
    KgSkmInsertArgs a;
    a.seg_start = s.d_seg_start; a.seg_ptr = s.d_seg_ptr; a.nseg = nseg; a.my_rank = (u32)me; a.src = s.d_src;
    // L2 prefetch of the next partition's table region: measured without effect (profiles/r02_insert_variants.txt: the
    // hit rate does not move, DRAM reads grow by what was prefetched), so it is off unless KG_PREFETCH=1
    a.nparts = c->pl; a.segs_per_part = getenv("KG_PREFETCH") ? (u32)world * KG_SKM_SUB : 0u;
    a.part_lo = c->d_part_lo; a.bpart_lo = c->d_bpart_lo; a.table = c->table; a.bloom = c->bloom; a.stats = c->d_stats;
    a.work = c->d_work; a.k = c->cfg.k;
    cudaEvent_t ia = next_ins_event(c), ib = next_ins_event(c);
    if (ia) cudaEventRecord(ia, c->s_insert);
    c->ins_launches++;
    KG_DISPATCH_W(c->W, launch_skm_insert, c, a, current_sink(c));
    if (ib) cudaEventRecord(ib, c->s_insert);
    KG_TRACE_SYNC(c, c->s_insert, "insert");
    if (world == 1) KG_CUDA(c, cudaEventRecord(s.ev_free, c->s_insert));
    c->round++;
    return KG_OK;
}

// bucket the windows of the batch that was just packed into the current slot, then hand the slot to the round
static int bucket_batch(kg_ctx* c, u32 nwords) {
    SkmSlot& s = c->slot[c->round & 1];
    const u32 nregions = c->nb * KG_SKM_SUB;
    KG_CUDA(c, cudaMemsetAsync(c->d_cursors, 0, sizeof(u32) * 8 * (nregions + 1), c->s_compute));
    KgSkmScatterArgs a;
    a.words = s.words; a.brk = c->d_brk; a.st = c->d_stream; a.cursors = c->d_cursors; a.regions = s.desc;
    a.ovf = s.desc + (u64)nregions * c->skm_cap; a.hdr = s.hdr; a.stats = c->d_stats;
    a.k = c->cfg.k; a.m = c->skm_m; a.nb = c->nb; a.pl = c->pl; a.cap = c->skm_cap; a.src = (u32)c->cfg.rank;
    a.ovf_cap = c->skm_ovf_cap; a.nwords = nwords;
    const u32 grid = (nwords + KG_SKM_TPB - 1) / KG_SKM_TPB;
    kg_skm_scatter<<<grid, KG_SKM_TPB, 0, c->s_compute>>>(a);
    kg_skm_pack_counts<<<(nregions + 256) / 256, 256, 0, c->s_compute>>>(c->d_cursors, nregions, c->skm_cap, c->skm_ovf_cap, s.counts);
    c->launches += 2;
    return skm_round(c, true, false);
}

template <int W>
static void launch_rb_sweep(kg_ctx* c, int sweep, const u64* words, const u32* brk, const KgStream* st, u32 nthreads) {
    const u32 grid = (nthreads + 255) / 256;
    if (grid == 0) return;
    switch (sweep) {
        case 1: kg_refbloom_sweep<W, 1><<<grid, 256, 0, c->s_compute>>>(words, brk, st, c->rb, c->d_stats, c->cfg.k); break;
        case 2: kg_refbloom_sweep<W, 2><<<grid, 256, 0, c->s_compute>>>(words, brk, st, c->rb, c->d_stats, c->cfg.k); break;
        case 3: kg_refbloom_sweep<W, 3><<<grid, 256, 0, c->s_compute>>>(words, brk, st, c->rb, c->d_stats, c->cfg.k); break;
        default: break;
    }
    c->launches++;
}
template <int W>
static void launch_rb_count(kg_ctx* c, const KgCountArgs& a, u32 nthreads) {
    const u32 grid = (nthreads + 255) / 256;
    if (grid == 0) return;
    kg_refbloom_count<W><<<grid, 256, 0, c->s_compute>>>(a, c->rb);
    c->launches++;
}

// emulation mode, Bloom pass: sweep 1 over the batch that was just packed, and keep its packed stream (with the stream
// state it was packed under) for sweeps 2 and 3 at kg_pass_end
static int rb_bloom_batch(kg_ctx* c, const u64* words, u32 nthreads) {
    kg_ctx::RbBatch b{nullptr, nullptr, nullptr, nthreads};
    KG_CUDA(c, cudaMalloc(&b.words, sizeof(u64) * nthreads));
    c->rb_log.push_back(b);                                  // owned by the context from here on
    kg_ctx::RbBatch& e = c->rb_log.back();
    KG_CUDA(c, cudaMalloc(&e.brk, sizeof(u32) * nthreads));
    KG_CUDA(c, cudaMalloc(&e.st, sizeof(KgStream)));
    KG_CUDA(c, cudaMemcpyAsync(e.words, words, sizeof(u64) * nthreads, cudaMemcpyDeviceToDevice, c->s_compute));
    KG_CUDA(c, cudaMemcpyAsync(e.brk, c->d_brk, sizeof(u32) * nthreads, cudaMemcpyDeviceToDevice, c->s_compute));
    KG_CUDA(c, cudaMemcpyAsync(e.st, c->d_stream, sizeof(KgStream), cudaMemcpyDeviceToDevice, c->s_compute));
    KG_DISPATCH_W(c->W, launch_rb_sweep, c, 1, e.words, e.brk, e.st, nthreads);
    return KG_OK;
}

// parse + count one device-resident batch (n <= batch_bytes, 16-byte aligned) on the compute stream
static int process_batch(kg_ctx* c, const uint8_t* d_in, size_t n, uint32_t flags) {
    if (n == 0) return KG_OK;
    cudaStream_t s = c->s_compute;
    const u32 ntiles = (u32)((n + KG_TILE - 1) / KG_TILE);
    const bool fasta = c->cfg.input_mode == KG_INPUT_FASTA;
    const size_t nwords = n / 32 + c->carry_max_words + 4;
    // the packed stream of a bucketed batch lives in the batch slot (the insert of this batch, and the peers, read it
    // there): wait until the round that used the slot two batches ago has let go of it
    u64* d_words = c->d_words_direct;
    if (c->bucketed) {
        SkmSlot& slot = c->slot[c->round & 1];
        KG_CUDA(c, cudaStreamWaitEvent(s, slot.ev_free, 0));
        d_words = slot.words;
    }
    cudaEvent_t e0 = next_event(c), e1 = next_event(c), e2 = next_event(c);
    if (e0) cudaEventRecord(e0, s);
    KG_CUDA(c, cudaMemsetAsync(d_words, 0, sizeof(u64) * nwords, s));
    KG_CUDA(c, cudaMemsetAsync(c->d_brk, 0, sizeof(u32) * nwords, s));
    kg_carry_restore<<<1, 32, 0, s>>>(d_words, c->d_brk, c->d_stream, c->d_carry_words, c->d_carry_brk, c->carry_max_words);
    const bool tma = c->parse_tma;   // tiles staged through shared memory by a TMA bulk copy (kg_fetch16<true>)
    if (fasta) {
        if (tma) kg_hdr_summary<true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_eff);
        else kg_hdr_summary<false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_eff);
        kg_lww_scan<<<1, 1024, 0, s>>>(c->d_tile_hdr_eff, c->d_tile_hdr_in, ntiles, &c->d_stream->in_header);
        if (tma) kg_tile_count<true, true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_nbases, c->d_tile_pend_eff);
        else kg_tile_count<true, false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_nbases, c->d_tile_pend_eff);
        c->launches += 3;
    } else {
        if (tma) kg_tile_count<false, true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_nbases, c->d_tile_pend_eff);
        else kg_tile_count<false, false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_nbases, c->d_tile_pend_eff);
        c->launches += 1;
    }
    kg_tile_scan<<<1, 1024, 0, s>>>(c->d_tile_nbases, c->d_tile_off, ntiles, c->d_stream);
    kg_lww_scan<<<1, 1024, 0, s>>>(c->d_tile_pend_eff, c->d_tile_pend_in, ntiles, &c->d_stream->pending_break);
    if (fasta) {
        if (tma) kg_tile_pack<true, true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, d_words, c->d_brk);
        else kg_tile_pack<true, false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, d_words, c->d_brk);
    } else {
        if (tma) kg_tile_pack<false, true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, d_words, c->d_brk);
        else kg_tile_pack<false, false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, d_words, c->d_brk);
    }
    c->launches += 4;
    if (e1) cudaEventRecord(e1, s);
    if (!(flags & KG_FEED_CONTEXT)) {
        const u32 nthreads = (u32)(n / 32 + c->carry_max_words + 2);   // upper bound on packed words
        if (c->ref_bloom) {                                            // bit-exact emulation of the reference's filters
            if (c->pass == KG_PASS_BLOOM) {
                int rc = rb_bloom_batch(c, d_words, nthreads);
                if (rc) return rc;
            } else {
                KgCountArgs a;
                a.words = d_words; a.brk = c->d_brk; a.st = c->d_stream;
                a.table = c->table; a.bloom = c->bloom;
                a.stats = c->d_stats; a.k = c->cfg.k; a.rank = 0; a.world = 1;
                KG_DISPATCH_W(c->W, launch_rb_count, c, a, nthreads);
            }
        } else if (c->pass_bucketed) {
            int rc = bucket_batch(c, nthreads);
            if (rc) return rc;
        } else {
            KgCountArgs a;
            a.words = d_words; a.brk = c->d_brk; a.st = c->d_stream;
            a.table = c->table; a.bloom = c->bloom;
            a.stats = c->d_stats; a.k = c->cfg.k; a.rank = (u32)c->cfg.rank; a.world = (u32)c->cfg.world;
            const int sink = current_sink(c);
            cudaEvent_t ia = next_ins_event(c), ib = next_ins_event(c);
            if (ia) cudaEventRecord(ia, s);
            c->ins_launches++;
            KG_DISPATCH_W(c->W, launch_count, c, a, nthreads, sink);
            if (ib) cudaEventRecord(ib, s);
        }
    }
    kg_carry_save<<<1, 32, 0, s>>>(d_words, c->d_brk, c->d_stream, c->d_carry_words, c->d_carry_brk, c->cfg.k, c->carry_max_words);
    c->launches += 1;
    if (e2) cudaEventRecord(e2, s);
    KG_CUDA(c, cudaGetLastError());
    c->raw_bytes_pass += n;
    return KG_OK;
}

static int check_feed(kg_ctx* c) {
    if (!c || !c->pass) return KG_EBADARG;
    if (c->pass == KG_PASS_COUNT && !c->table.slots) return KG_EBADARG;
    if (!c->stream_open) { c->err = "kg_feed before kg_stream_begin"; return KG_EBADARG; }
    return KG_OK;
}

extern "C" int kg_feed_device(kg_ctx* c, const void* device_bytes, size_t n, uint32_t flags) {
    int rc = check_feed(c);
    if (rc) return rc;
    if (n == 0) return KG_OK;
    if (!device_bytes) return KG_EBADARG;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    const uint8_t* p = (const uint8_t*)device_bytes;
    const bool aligned = ((uintptr_t)p & 15u) == 0;
    for (size_t off = 0; off < n; off += c->batch_bytes) {
        size_t len = n - off < c->batch_bytes ? n - off : c->batch_bytes;
        if (aligned) {
            rc = process_batch(c, p + off, len, flags);
        } else {
            int idx = c->raw_idx; c->raw_idx ^= 1;
            KG_CUDA(c, cudaMemcpyAsync(c->d_raw[idx], p + off, len, cudaMemcpyDeviceToDevice, c->s_compute));
            rc = process_batch(c, c->d_raw[idx], len, flags);
        }
        if (rc) return rc;
    }
    return KG_OK;
}

// Software-pipelined feed of a pinned buffer: the H2D copy of chunk i+1 is issued on the copy stream BEFORE the kernels
// of chunk i are queued, and is gated on the device (cudaStreamWaitEvent on the raw buffer's free event) instead of a
// host wait, so the copies never sit on the compute stream's critical path.  (KG_FEED_PREFETCH=0 selects the simple loop.)
static int feed_pinned_pipelined(kg_ctx* c, const uint8_t* bytes, size_t n, uint32_t flags) {
    const size_t B = c->batch_bytes, nchunks = (n + B - 1) / B;
    const int raw0 = c->raw_idx;
    auto issue = [&](size_t i) -> int {
        const int idx = (int)((raw0 + i) & 1);
        const size_t off = i * B, len = n - off < B ? n - off : B;
        KG_CUDA(c, cudaStreamWaitEvent(c->s_copy, c->ev_raw_free[idx], 0));   // compute is done with this buffer (chunk i-2)
        KG_CUDA(c, cudaMemcpyAsync(c->d_raw[idx], bytes + off, len, cudaMemcpyHostToDevice, c->s_copy));
        KG_CUDA(c, cudaEventRecord(c->ev_copy_done[idx], c->s_copy));
        return KG_OK;
    };
    int rc = issue(0);
    if (rc) return rc;
    for (size_t i = 0; i < nchunks; i++) {
        const int idx = (int)((raw0 + i) & 1);
        const size_t off = i * B, len = n - off < B ? n - off : B;
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_copy_done[idx], 0));
        if (i + 1 < nchunks) { rc = issue(i + 1); if (rc) return rc; }
        rc = process_batch(c, c->d_raw[idx], len, flags);
        if (rc) return rc;
        KG_CUDA(c, cudaEventRecord(c->ev_raw_free[idx], c->s_compute));
    }
    c->raw_idx = (int)((raw0 + nchunks) & 1);
    // the caller owns `bytes` again when we return: the last copies must have left it
    KG_CUDA(c, cudaStreamSynchronize(c->s_copy));
    return KG_OK;
}

extern "C" int kg_feed(kg_ctx* c, const uint8_t* bytes, size_t n, uint32_t flags) {
    int rc = check_feed(c);
    if (rc) return rc;
    if (n == 0) return KG_OK;
    if (!bytes) return KG_EBADARG;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, bytes) == cudaSuccess) {
        if (attr.type == cudaMemoryTypeDevice) return kg_feed_device(c, bytes, n, flags);
        pinned = attr.type == cudaMemoryTypeHost;
    } else {
        cudaGetLastError();
    }
    if (pinned && c->feed_prefetch && n > c->batch_bytes) return feed_pinned_pipelined(c, bytes, n, flags);
    for (size_t off = 0; off < n; off += c->batch_bytes) {
        size_t len = n - off < c->batch_bytes ? n - off : c->batch_bytes;
        const int idx = c->raw_idx; c->raw_idx ^= 1;
        // the compute stream must be done with this raw buffer (two batches ago)
        KG_CUDA(c, cudaEventSynchronize(c->ev_raw_free[idx]));
        const uint8_t* src = bytes + off;
        if (!pinned) {
            if (!c->h_stage[idx]) KG_CUDA(c, cudaHostAlloc((void**)&c->h_stage[idx], c->batch_bytes, cudaHostAllocDefault));
            KG_CUDA(c, cudaEventSynchronize(c->ev_stage_free[idx]));
            memcpy(c->h_stage[idx], src, len);
            src = c->h_stage[idx];
        }
        KG_CUDA(c, cudaMemcpyAsync(c->d_raw[idx], src, len, cudaMemcpyHostToDevice, c->s_copy));
        KG_CUDA(c, cudaEventRecord(c->ev_copy_done[idx], c->s_copy));
        if (!pinned) KG_CUDA(c, cudaEventRecord(c->ev_stage_free[idx], c->s_copy));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_copy_done[idx], 0));
        rc = process_batch(c, c->d_raw[idx], len, flags);
        if (rc) return rc;
        KG_CUDA(c, cudaEventRecord(c->ev_raw_free[idx], c->s_compute));
        // the caller owns `bytes` again when we return: wait for the H2D copy (compute keeps running)
        if (pinned) KG_CUDA(c, cudaEventSynchronize(c->ev_copy_done[idx]));
    }
    return KG_OK;
}

extern "C" int kg_pass_end(kg_ctx* c, kg_pass_stats* out) {
    if (!c || !c->pass) return KG_EBADARG;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    if (c->cfg.world > 1) {
        // keep taking part in rounds until a round in which NO rank had a batch: every rank then has issued the same
        // number of rounds, and entering that last round's all-reduce means every insert of the pass has completed
        for (;;) {
            int rc = skm_round(c, false, true);
            if (rc) return rc;
            KG_CUDA(c, cudaStreamSynchronize(c->s_insert));
            KG_TRACE(c, "pass_end: round %llu had %u ranks with a batch", (unsigned long long)c->round - 1, c->h_round_sum[0]);
            if (c->h_round_sum[0] == 0) break;
        }
    }
    if (c->pass_bucketed) {
        KG_CUDA(c, cudaEventRecord(c->ev_tail, c->s_insert));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_tail, 0));
    }
    if (c->ref_bloom && c->pass == KG_PASS_BLOOM) {
        for (int sweep = 2; sweep <= 3; sweep++)
            for (auto& b : c->rb_log) KG_DISPATCH_W(c->W, launch_rb_sweep, c, sweep, b.words, b.brk, b.st, b.nthreads);
        // window ordinals are 32-bit: the stream must hold fewer than 2^32 - 1 bases
        KgStream last;
        KG_CUDA(c, cudaMemcpyAsync(&last, c->d_stream, sizeof(last), cudaMemcpyDeviceToHost, c->s_compute));
        KG_CUDA(c, cudaStreamSynchronize(c->s_compute));
        if (last.bases_seen + (u64)last.total_bases >= 0xFFFFFFFFull) {
            c->err = "KG_CFG_REFERENCE_BLOOM: input has 2^32 bases or more (window ordinals are 32-bit)";
            c->pass = 0; c->stream_open = false;
            return KG_EBADARG;
        }
        for (auto& b : c->rb_log) { cudaFree(b.words); cudaFree(b.brk); cudaFree(b.st); }
        c->rb_log.clear();
    }
    KG_CUDA(c, cudaEventRecord(c->ev_pass_end, c->s_compute));
    KG_CUDA(c, cudaStreamSynchronize(c->s_compute));
    KgStats st;
    KG_CUDA(c, cudaMemcpy(&st, c->d_stats, sizeof(st), cudaMemcpyDeviceToHost));
    if (c->pass == KG_PASS_BLOOM) {
        c->new_in_second = st.new_in_second;
        c->bloom_done = true;
    }
    if (out) {
        memset(out, 0, sizeof(*out));
        out->input_kmers = st.input_kmers;
        out->inserted_kmers = st.inserted;
        out->distinct = st.distinct;
        out->table_slots = c->pass == KG_PASS_COUNT ? c->table.nslots : 0;
        out->new_in_first = st.new_in_first;
        out->new_in_second = st.new_in_second;
        out->bloom_bits = c->bloom_m;
        out->bloom_hashes = c->bloom.nh;
        out->partitions = c->pass_bucketed ? c->pl : 1;
        out->raw_bytes = c->raw_bytes_pass;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev_pass_begin, c->ev_pass_end);
        out->device_ms = ms;
        for (size_t i = 0; i + 3 <= c->ev_used; i += 3) {
            float a = 0, b = 0;
            if (cudaEventElapsedTime(&a, c->ev_pool[i], c->ev_pool[i + 1]) == cudaSuccess) out->parse_ms += a;
            if (cudaEventElapsedTime(&b, c->ev_pool[i + 1], c->ev_pool[i + 2]) == cudaSuccess) out->count_ms += b;
        }
        for (size_t i = 0; i + 2 <= c->ins_used; i += 2) {
            float a = 0;
            if (cudaEventElapsedTime(&a, c->ins_pool[i], c->ins_pool[i + 1]) == cudaSuccess) out->insert_ms += a;
        }
        out->insert_launches = c->ins_launches;
        cudaGetLastError();
    }
    const int pass = c->pass;
    c->pass = 0;
    c->stream_open = false;
    if (pass == KG_PASS_COUNT) c->counted = true;
    if (st.table_full == 2) { c->err = "internal: descriptor overflow list exhausted"; return KG_ECUDA; }
    if (pass == KG_PASS_COUNT && st.table_full) { c->err = "Hash table is full"; c->pass = 0; return KG_ETABLE_FULL; }
    return KG_OK;
}

template <int W>
static void launch_kaarme_build(kg_ctx* c, const u32* bitmap, const u64* prefix, KgKaarme out, u64* root_counter) {
    const u32 grid = (u32)((c->table.nslots + 255) / 256);
    KgPlacement pm;
    pm.part_lo = c->d_part_lo; pm.pl = c->pass_bucketed ? c->pl : 1; pm.nb = c->pass_bucketed ? c->nb : 1;
    pm.m = c->skm_m; pm.rank = (u32)c->cfg.rank;
    kg_kaarme_build<W><<<grid, 256, 0, c->s_compute>>>(c->table, c->cfg.k, bitmap, prefix, out, root_counter, pm);
    c->launches++;
}
template <int W>
static void launch_kaarme_chain(kg_ctx* c) {
    const u32 grid = (u32)((c->kaarme.n_kmers + 255) / 256);
    if (grid) kg_kaarme_chain_stats<W><<<grid, 256, 0, c->s_compute>>>(c->kaarme, c->cfg.k, c->d_cstats);
    c->launches++;
}
extern "C" int kg_compact(kg_ctx* c, kg_compact_stats* stats) {
    if (!c) return KG_EBADARG;
    if (c->cfg.table_mode != KG_TABLE_KAARME) { c->err = "kg_compact needs table_mode KG_TABLE_KAARME"; return KG_EBADARG; }
    if (!c->counted || !c->table.slots) { c->err = "kg_compact before the count pass"; return KG_EBADARG; }
    if (c->compacted) return KG_OK;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    cudaStream_t s = c->s_compute;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    KG_CUDA(c, cudaEventCreate(&e0));
    KG_CUDA(c, cudaEventCreate(&e1));
    KG_CUDA(c, cudaEventRecord(e0, s));
    const u64 nslots = c->table.nslots;
    const u32 grid = (u32)((nslots + 255) / 256);
    const u64 nwords = (u64)grid * 8;
    const u64 nsb = (nwords + 1023) / 1024;
    u32 *bitmap = nullptr, *wcount = nullptr;
    u64 *prefix = nullptr, *bsum = nullptr, *scalars = nullptr;   // scalars[0] = n_kmers, scalars[1] = root counter
    struct Scratch {   // released on every return path (the KG_CUDA early returns included)
        u32 **a, **b; u64 **p, **q, **r; cudaEvent_t *e0, *e1;
        ~Scratch() {
            cudaFree(*a); cudaFree(*b); cudaFree(*p); cudaFree(*q); cudaFree(*r);
            if (*e0) cudaEventDestroy(*e0);
            if (*e1) cudaEventDestroy(*e1);
        }
    } scratch{&bitmap, &wcount, &prefix, &bsum, &scalars, &e0, &e1};
    KG_CUDA(c, cudaMalloc(&bitmap, sizeof(u32) * nwords));
    KG_CUDA(c, cudaMalloc(&wcount, sizeof(u32) * nwords));
    KG_CUDA(c, cudaMalloc(&prefix, sizeof(u64) * nwords));
    KG_CUDA(c, cudaMalloc(&bsum, sizeof(u64) * nsb));
    KG_CUDA(c, cudaMalloc(&scalars, sizeof(u64) * 2));
    if (!c->d_cstats) KG_CUDA(c, cudaMalloc(&c->d_cstats, sizeof(KgCompactStats)));
    KG_CUDA(c, cudaMemsetAsync(scalars, 0, sizeof(u64) * 2, s));
    KG_CUDA(c, cudaMemsetAsync(c->d_cstats, 0, sizeof(KgCompactStats), s));
    kg_occupancy_bitmap<<<grid, 256, 0, s>>>(c->table, bitmap, wcount);
    kg_scan_blocks<<<(u32)nsb, 1024, 0, s>>>(wcount, bsum, nwords);
    kg_scan_block_sums<<<1, 1024, 0, s>>>(bsum, nsb, scalars);
    kg_scan_finish<<<(u32)nsb, 1024, 0, s>>>(wcount, bsum, prefix, nwords);
    c->launches += 4;
    u64 h[2];
    KG_CUDA(c, cudaMemcpyAsync(h, scalars, sizeof(u64) * 2, cudaMemcpyDeviceToHost, s));
    KG_CUDA(c, cudaStreamSynchronize(s));
    const u64 n_kmers = h[0];
    // sizing pass: how many roots?
    KgKaarme probe{nullptr, nullptr, n_kmers, 0};
    KG_DISPATCH_W(c->W, launch_kaarme_build, c, bitmap, prefix, probe, scalars + 1);
    KG_CUDA(c, cudaMemcpyAsync(h, scalars, sizeof(u64) * 2, cudaMemcpyDeviceToHost, s));
    KG_CUDA(c, cudaStreamSynchronize(s));
    const u64 n_roots = h[1];
    KgKaarme ks{nullptr, nullptr, n_kmers, n_roots};
    cudaFree(c->kaarme.slots); cudaFree(c->kaarme.roots);      // left-overs of an attempt that failed half way
    c->kaarme = KgKaarme{nullptr, nullptr, 0, 0};
    KG_CUDA(c, cudaMalloc(&ks.slots, sizeof(u64) * (n_kmers ? n_kmers : 1)));
    c->kaarme.slots = ks.slots;                                 // owned by the context from here on (freed by kg_destroy)
    KG_CUDA(c, cudaMalloc(&ks.roots, sizeof(u64) * (n_roots ? n_roots : 1) * c->W));
    KG_CUDA(c, cudaMemsetAsync(scalars + 1, 0, sizeof(u64), s));
    KG_DISPATCH_W(c->W, launch_kaarme_build, c, bitmap, prefix, ks, scalars + 1);
    c->kaarme = ks;
    KG_DISPATCH_W(c->W, launch_kaarme_chain, c);
    KG_CUDA(c, cudaEventRecord(e1, s));
    KgCompactStats cs;
    KG_CUDA(c, cudaMemcpyAsync(&cs, c->d_cstats, sizeof(cs), cudaMemcpyDeviceToHost, s));
    KG_CUDA(c, cudaStreamSynchronize(s));
    KG_CUDA(c, cudaGetLastError());
    // the plain table has served its purpose: from here on only the compact structure exists
    cudaFree(c->table.slots);
    c->table.slots = nullptr;
    c->table_bytes = 0;
    c->compacted = true;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (cs.bad) { c->err = "kg_compact: malformed predecessor chain"; return KG_ECUDA; }
    if (stats) {
        stats->kmers = n_kmers;
        stats->roots = n_roots;
        stats->bytes = 8 * n_kmers + 8 * (uint64_t)c->W * n_roots;
        stats->reference_bytes = 8 * nslots + (8 * (uint64_t)c->W + 1) * n_roots;   // kmer.hpp:107, kmer_hash_table.cpp:2144-2145
        stats->max_chain = cs.max_chain;
        stats->device_ms = ms;
    }
    return KG_OK;
}

// -----------------------------------------------------------------------------------------------------------
template <int W>
static void launch_export(kg_ctx* c, u64 b, u64 e, uint64_t min_ab, int count_mode, int buf) {
    const u32 block = 256;
    const u32 grid = (u32)((e - b + block - 1) / block);
    if (c->compacted)
        kg_kaarme_export<W><<<grid, block, 0, c->s_compute>>>(c->kaarme, c->cfg.k, b, e, min_ab, c->d_out_keys[buf],
                                                            c->d_out_counts[buf], c->d_out_n[buf], c->d_cstats);
    else
        kg_export_kernel<W><<<grid, block, 0, c->s_compute>>>(c->table, b, e, min_ab, count_mode, c->cfg.table_mode,
                                                            c->d_out_keys[buf], c->d_out_counts[buf], c->d_out_n[buf]);
    c->launches++;
}

#define KG_TEXT_BUFFER_BYTES (64ull << 20)

// Shared driver of kg_export (records) and kg_export_text (formatted lines).  The table (or the compact structure)
// is scanned in chunks; chunk i is compacted (and formatted) on the compute stream while the host hands chunk i-1
// to the sink.
static int export_impl(kg_ctx* c, uint64_t min_abundance, int count_mode, kg_sink_fn sink, kg_text_sink_fn tsink, void* user) {
    if (!c || (!sink && !tsink) || (!c->table.slots && !c->compacted)) return KG_EBADARG;
    if (min_abundance == 0) return KG_OK;  // parallel_parser.hpp:860-861
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    const int W = c->W;
    const bool text = tsink != nullptr;
    // Slots (= upper bound on records) per chunk: 16 M key words, but never more than the structure holds (a small
    // input must not pay for pinning hundreds of MiB: cudaHostAlloc costs ~1 ms per MiB) and, for text, never more
    // than one text buffer can hold.  Buffers grow on demand and are kept for the next export.
    const u64 nslots = c->compacted ? c->kaarme.n_kmers : c->table.nslots;
    const u32 line_bound = kg_line_bound(c->cfg.k);
    const size_t text_smem = 16 + (size_t)KG_TEXT_TPB * line_bound;
    size_t chunk = (16u << 20) / (size_t)W;
    if (text && KG_TEXT_BUFFER_BYTES / line_bound < chunk) chunk = KG_TEXT_BUFFER_BYTES / line_bound;
    if (nslots < chunk) chunk = (size_t)((nslots + 4095) / 4096 * 4096);
    if (chunk == 0) chunk = 4096;
    if (!c->h_out_n) {
        for (int i = 0; i < 2; i++) KG_CUDA(c, cudaMalloc(&c->d_out_n[i], sizeof(u32)));
        KG_CUDA(c, cudaHostAlloc((void**)&c->h_out_n, 2 * sizeof(u32), cudaHostAllocDefault));
    }
    if (chunk > c->out_cap_dev) {
        for (int i = 0; i < 2; i++) {
            cudaFree(c->d_out_keys[i]); cudaFree(c->d_out_counts[i]);
            c->d_out_keys[i] = nullptr; c->d_out_counts[i] = nullptr;
        }
        c->out_cap_dev = 0;
        for (int i = 0; i < 2; i++) {
            KG_CUDA(c, cudaMalloc(&c->d_out_keys[i], chunk * W * sizeof(u64)));
            KG_CUDA(c, cudaMalloc(&c->d_out_counts[i], chunk * sizeof(u32)));
        }
        c->out_cap_dev = chunk;
    }
    if (!text && chunk > c->out_cap_host) {
        for (int i = 0; i < 2; i++) {
            if (c->h_out_keys[i]) cudaFreeHost(c->h_out_keys[i]);
            if (c->h_out_counts[i]) cudaFreeHost(c->h_out_counts[i]);
            c->h_out_keys[i] = nullptr; c->h_out_counts[i] = nullptr;
        }
        c->out_cap_host = 0;
        for (int i = 0; i < 2; i++) {
            KG_CUDA(c, cudaHostAlloc((void**)&c->h_out_keys[i], chunk * W * sizeof(u64), cudaHostAllocDefault));
            KG_CUDA(c, cudaHostAlloc((void**)&c->h_out_counts[i], chunk * sizeof(u32), cudaHostAllocDefault));
        }
        c->out_cap_host = chunk;
    }
    if (text) {
        const size_t need = ((chunk * (size_t)line_bound) + 4095) / 4096 * 4096;   // every line of a chunk always fits
        if (!c->h_text_cur) {
            for (int i = 0; i < 2; i++) KG_CUDA(c, cudaMalloc(&c->d_text_cur[i], sizeof(u64)));
            KG_CUDA(c, cudaHostAlloc((void**)&c->h_text_cur, 2 * sizeof(u64), cudaHostAllocDefault));
        }
        if (need > c->text_cap) {
            for (int i = 0; i < 2; i++) {
                cudaFree(c->d_text[i]);
                if (c->h_text[i]) cudaFreeHost(c->h_text[i]);
                c->d_text[i] = nullptr; c->h_text[i] = nullptr;
            }
            c->text_cap = 0;
            for (int i = 0; i < 2; i++) {
                KG_CUDA(c, cudaMalloc(&c->d_text[i], need));
                KG_CUDA(c, cudaHostAlloc((void**)&c->h_text[i], need, cudaHostAllocDefault));
            }
            c->text_cap = need;
        }
        if (!c->text_configured) {
            KG_CUDA(c, cudaFuncSetAttribute(kg_format_text, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(16 + (size_t)KG_TEXT_TPB * kg_line_bound(KG_MAX_K))));
            c->text_configured = true;
        }
    }
    const u64 nchunks = (nslots + chunk - 1) / chunk;
    auto finish = [&](u64 i) -> int {
        const int b = (int)(i & 1);
        KG_CUDA(c, cudaEventSynchronize(c->ev_out[b]));
        const u32 n = c->h_out_n[b];
        if (!n) return KG_OK;
        if (text) {
            const u64 bytes = c->h_text_cur[b];
            if (bytes > c->text_cap) { c->err = "kg_export_text: text buffer overrun"; return KG_ECUDA; }
            KG_CUDA(c, cudaMemcpyAsync(c->h_text[b], c->d_text[b], (size_t)bytes, cudaMemcpyDeviceToHost, c->s_copy));
            KG_CUDA(c, cudaStreamSynchronize(c->s_copy));
            if (tsink(user, c->h_text[b], (size_t)bytes, n) != 0) return KG_ESINK;
        } else {
            KG_CUDA(c, cudaMemcpyAsync(c->h_out_keys[b], c->d_out_keys[b], (size_t)n * W * sizeof(u64), cudaMemcpyDeviceToHost, c->s_copy));
            KG_CUDA(c, cudaMemcpyAsync(c->h_out_counts[b], c->d_out_counts[b], (size_t)n * sizeof(u32), cudaMemcpyDeviceToHost, c->s_copy));
            KG_CUDA(c, cudaStreamSynchronize(c->s_copy));
            if (sink(user, (const uint64_t*)c->h_out_keys[b], (const uint32_t*)c->h_out_counts[b], n) != 0) return KG_ESINK;
        }
        return KG_OK;
    };
    for (u64 i = 0; i < nchunks; i++) {
        const int b = (int)(i & 1);
        const u64 sb = i * chunk, se = sb + chunk < nslots ? sb + chunk : nslots;
        KG_CUDA(c, cudaMemsetAsync(c->d_out_n[b], 0, sizeof(u32), c->s_compute));
        switch (W) {
            case 1: launch_export<1>(c, sb, se, min_abundance, count_mode, b); break;
            case 2: launch_export<2>(c, sb, se, min_abundance, count_mode, b); break;
            case 3: launch_export<3>(c, sb, se, min_abundance, count_mode, b); break;
            case 4: launch_export<4>(c, sb, se, min_abundance, count_mode, b); break;
            case 5: launch_export<5>(c, sb, se, min_abundance, count_mode, b); break;
            case 6: launch_export<6>(c, sb, se, min_abundance, count_mode, b); break;
            case 7: launch_export<7>(c, sb, se, min_abundance, count_mode, b); break;
            case 8: launch_export<8>(c, sb, se, min_abundance, count_mode, b); break;
        }
        if (text) {
            KG_CUDA(c, cudaMemsetAsync(c->d_text_cur[b], 0, sizeof(u64), c->s_compute));
            const u32 grid = (u32)((se - sb + KG_TEXT_TPB - 1) / KG_TEXT_TPB);   // upper bound; blocks past *n exit
            kg_format_text<<<grid, KG_TEXT_TPB, text_smem, c->s_compute>>>(c->d_out_keys[b], c->d_out_counts[b], c->d_out_n[b],
                                                                        (u32)W, c->cfg.k, c->d_text[b], c->d_text_cur[b]);
            c->launches++;
            KG_CUDA(c, cudaMemcpyAsync(&c->h_text_cur[b], c->d_text_cur[b], sizeof(u64), cudaMemcpyDeviceToHost, c->s_compute));
        }
        KG_CUDA(c, cudaMemcpyAsync(&c->h_out_n[b], c->d_out_n[b], sizeof(u32), cudaMemcpyDeviceToHost, c->s_compute));
        KG_CUDA(c, cudaEventRecord(c->ev_out[b], c->s_compute));
        if (i > 0) { int rc = finish(i - 1); if (rc) { cudaStreamSynchronize(c->s_compute); return rc; } }
    }
    if (nchunks > 0) { int rc = finish(nchunks - 1); if (rc) return rc; }
    KG_CUDA(c, cudaGetLastError());
    if (c->compacted) {
        KgCompactStats cs;
        KG_CUDA(c, cudaMemcpy(&cs, c->d_cstats, sizeof(cs), cudaMemcpyDeviceToHost));
        if (cs.bad) { c->err = "kg_export: malformed predecessor chain while decoding"; return KG_ECUDA; }
    }
    return KG_OK;
}

extern "C" int kg_export(kg_ctx* c, uint64_t min_abundance, int count_mode, kg_sink_fn sink, void* user) {
    if (!sink) return KG_EBADARG;
    return export_impl(c, min_abundance, count_mode, sink, nullptr, user);
}

extern "C" int kg_export_text(kg_ctx* c, uint64_t min_abundance, int count_mode, kg_text_sink_fn sink, void* user) {
    if (!sink) return KG_EBADARG;
    return export_impl(c, min_abundance, count_mode, nullptr, sink, user);
}

template <int W>
static void launch_checksum(kg_ctx* c, uint64_t min_ab, int count_mode, u64* d_out) {
    if (c->compacted) kg_checksum_kaarme<W><<<148 * 8, 256, 0, c->s_compute>>>(c->kaarme, c->cfg.k, min_ab, d_out, c->d_cstats);
    else kg_checksum_table<W><<<148 * 8, 256, 0, c->s_compute>>>(c->table, min_ab, count_mode, c->cfg.table_mode, d_out);
    c->launches++;
}
extern "C" int kg_checksum(kg_ctx* c, uint64_t min_abundance, int count_mode, uint64_t out[4]) {
    if (!c || !out || (!c->table.slots && !c->compacted)) return KG_EBADARG;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    u64* d_out = nullptr;
    KG_CUDA(c, cudaMalloc(&d_out, sizeof(u64) * 4));
    struct Guard { u64** p; ~Guard() { cudaFree(*p); } } guard{&d_out};
    KG_CUDA(c, cudaMemsetAsync(d_out, 0, sizeof(u64) * 4, c->s_compute));
    if (c->compacted && !c->d_cstats) return KG_EBADARG;
    KG_DISPATCH_W(c->W, launch_checksum, c, min_abundance, c->compacted ? KG_COUNT_REFERENCE : count_mode, d_out);
    KG_CUDA(c, cudaMemcpyAsync(out, d_out, sizeof(u64) * 4, cudaMemcpyDeviceToHost, c->s_compute));
    KG_CUDA(c, cudaStreamSynchronize(c->s_compute));
    KG_CUDA(c, cudaGetLastError());
    return KG_OK;
}

extern "C" int kg_kaarme_download(kg_ctx* c, uint64_t* slots, uint64_t* roots) {
    if (!c || !c->compacted) return KG_EBADARG;
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    if (slots && c->kaarme.n_kmers) KG_CUDA(c, cudaMemcpy(slots, c->kaarme.slots, sizeof(u64) * c->kaarme.n_kmers, cudaMemcpyDeviceToHost));
    if (roots && c->kaarme.n_roots) KG_CUDA(c, cudaMemcpy(roots, c->kaarme.roots, sizeof(u64) * c->kaarme.n_roots * c->W, cudaMemcpyDeviceToHost));
    return KG_OK;
}

extern "C" int kg_kaarme_upload(kg_ctx* c, const uint64_t* slots, uint64_t n_kmers, const uint64_t* roots, uint64_t n_roots) {
    if (!c) return KG_EBADARG;
    if (c->cfg.table_mode != KG_TABLE_KAARME || c->cfg.world != 1) { c->err = "kg_kaarme_upload needs a single-GPU KG_TABLE_KAARME context"; return KG_EBADARG; }
    if (c->pass) { c->err = "kg_kaarme_upload inside a pass"; return KG_EBADARG; }
    if ((n_kmers && !slots) || (n_roots && !roots)) return KG_EBADARG;
    if (n_kmers >> 38 || n_roots >> 38) { c->err = "kg_kaarme_upload: indices are 38 bits (kmer.hpp:108)"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    KG_CUDA(c, cudaStreamSynchronize(c->s_compute));
    cudaFree(c->kaarme.slots); cudaFree(c->kaarme.roots);
    c->kaarme = KgKaarme{nullptr, nullptr, 0, 0};
    c->compacted = false;
    if (c->table.slots) { cudaFree(c->table.slots); c->table.slots = nullptr; c->table_bytes = 0; }
    KgKaarme ks{nullptr, nullptr, n_kmers, n_roots};
    KG_CUDA(c, cudaMalloc(&ks.slots, sizeof(u64) * (n_kmers ? n_kmers : 1)));
    if (cudaMalloc(&ks.roots, sizeof(u64) * (n_roots ? n_roots : 1) * c->W) != cudaSuccess) {
        cudaFree(ks.slots); cudaGetLastError(); c->err = "kg_kaarme_upload: out of device memory"; return KG_ENOMEM;
    }
    c->kaarme = ks;
    if (n_kmers) KG_CUDA(c, cudaMemcpy(ks.slots, slots, sizeof(u64) * n_kmers, cudaMemcpyHostToDevice));
    if (n_roots) KG_CUDA(c, cudaMemcpy(ks.roots, roots, sizeof(u64) * n_roots * c->W, cudaMemcpyHostToDevice));
    if (!c->d_cstats) KG_CUDA(c, cudaMalloc(&c->d_cstats, sizeof(KgCompactStats)));
    KG_CUDA(c, cudaMemset(c->d_cstats, 0, sizeof(KgCompactStats)));
    c->compacted = true;
    c->counted = true;
    return KG_OK;
}

extern "C" int kg_table_info(const kg_ctx* c, uint64_t* slots, uint32_t* slot_bytes, uint32_t* key_words) {
    if (!c) return KG_EBADARG;
    if (slots) *slots = c->table.nslots;
    if (slot_bytes) *slot_bytes = c->table.stride * 8;
    if (key_words) *key_words = (uint32_t)c->W;
    return KG_OK;
}

extern "C" int kg_launch_count(const kg_ctx* c, uint64_t* launches) {
    if (!c || !launches) return KG_EBADARG;
    *launches = c->launches;
    return KG_OK;
}

// ---- random 32-byte-sector atomic ceiling (roofline denominator for the insert kernel) ---------------------
__global__ void kg_atomic_ceiling_kernel(u32* region, u64 nsectors, u64 n_ops, u64 seed) {
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = tid; i < n_ops; i += stride) {
        u64 h = kg_fmix64(i * 0x9E3779B97F4A7C15ULL + seed);
        u64 sector = __umul64hi(h, nsectors);
        atomicAdd(region + sector * 8, 1u);   // one RED per random 32-byte sector
    }
}

extern "C" int kg_atomic_ceiling(int device, uint64_t region_bytes, uint64_t n_ops, int reps, double* sectors_per_s) {
    if (!sectors_per_s || region_bytes < 32 || n_ops == 0) return KG_EBADARG;
    kg_ctx* c = nullptr;
    KG_CUDA(c, cudaSetDevice(device));
    u32* region = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    struct Guard {
        u32** r; cudaEvent_t *a, *b;
        ~Guard() { cudaFree(*r); if (*a) cudaEventDestroy(*a); if (*b) cudaEventDestroy(*b); }
    } guard{&region, &a, &b};
    KG_CUDA(c, cudaMalloc(&region, region_bytes));
    KG_CUDA(c, cudaMemset(region, 0, region_bytes));
    KG_CUDA(c, cudaEventCreate(&a));
    KG_CUDA(c, cudaEventCreate(&b));
    cudaDeviceProp prop;
    KG_CUDA(c, cudaGetDeviceProperties(&prop, device));
    const u32 grid = prop.multiProcessorCount * 8;
    double best = 0;
    for (int r = 0; r < (reps < 1 ? 1 : reps) + 1; r++) {
        KG_CUDA(c, cudaEventRecord(a));
        kg_atomic_ceiling_kernel<<<grid, 256>>>(region, region_bytes / 32, n_ops, (u64)r * 7919u);
        KG_CUDA(c, cudaEventRecord(b));
        KG_CUDA(c, cudaEventSynchronize(b));
        float ms = 0;
        KG_CUDA(c, cudaEventElapsedTime(&ms, a, b));
        if (r > 0 && ms > 0) { double v = (double)n_ops / (ms * 1e-3); if (v > best) best = v; }
    }
    *sectors_per_s = best;
    return KG_OK;
}
