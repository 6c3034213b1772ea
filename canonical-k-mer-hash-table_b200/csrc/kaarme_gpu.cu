// kaarme_gpu.cu -- implementation of the C ABI in include/kaarme_gpu.h (libkaarme_gpu.so, sm_100a only).
// Host-side orchestration: pinned double-buffered H2D on a copy stream, parse + count kernels on a compute
// stream, device-resident stream state (no host round trip per batch), chunked export.
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
#include "kg_exchange_plan.hpp"
#include "kg_kaarme.cuh"
#include "kg_parse.cuh"
#include "kg_refbloom.cuh"
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
    u64* d_words = nullptr;
    u32* d_brk = nullptr;
    size_t words_cap = 0;  // in words
    u64* d_carry_words = nullptr;
    u32* d_carry_brk = nullptr;
    u32 carry_max_words = 0;
    KgStream* d_stream = nullptr;
    KgStats* d_stats = nullptr;
    KgTable table{nullptr, 0, 0, 0, 1, 0};
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
    uint64_t ins_launches = 0, ins_keys_upper = 0;
    uint64_t raw_bytes_pass = 0, bases_pass = 0;
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
    // bucketed path (multi-GPU exchange, or partitions > 1 on one GPU)
    bool bucketed = false;              // this context may bucket (streams/events exist)
    bool pass_bucketed = false;         // the current pass buckets its batches
    u32 nb = 0;                         // buckets of the current pass = world * local partitions
    u32 pl = 1;                         // local partitions of the current pass
    u32 nb_alloc = 0;                   // buckets the scratch below was sized for
    u64 *d_seg[2] = {nullptr, nullptr}, *h_seg[2] = {nullptr, nullptr};   // segment tables (start[nseg+1], src[nseg])
    u32 seg_cap = 0;
    ncclComm_t comm = nullptr;       // key slices (ncclSend/ncclRecv) on s_comm
    ncclComm_t ctl_comm = nullptr;   // per-round count all-gather on s_ctl: its own communicator and stream, so the
                                     // tiny control exchange of round i+1 never queues behind the key transfer of round i
    cudaStream_t s_comm = nullptr, s_insert = nullptr, s_ctl = nullptr;
    u64* d_send[2] = {nullptr, nullptr};
    u64* d_recv[2] = {nullptr, nullptr};
    size_t send_cap = 0, recv_cap = 0;  // in keys
    u32 *d_blk_hist = nullptr, *d_blk_base = nullptr;
    u32 max_blocks = 0;
    u32 *d_bucket_counts = nullptr, *d_bucket_offs = nullptr, *d_matrix = nullptr, *h_matrix = nullptr;
    cudaEvent_t ev_counts = nullptr, ev_scatter = nullptr, ev_matrix = nullptr, ev_pass_ready = nullptr, ev_tail = nullptr;
    cudaEvent_t ev_send_free[2] = {nullptr, nullptr}, ev_recv_free[2] = {nullptr, nullptr}, ev_recv_full[2] = {nullptr, nullptr};
    uint64_t round = 0, subround = 0;
    bool scatter_configured = false;
    u32 reserve_cap = 0;                // keys per bucket region of the one-pass (reserve) scatter
    size_t send_alloc = 0;              // keys each d_send buffer can hold
    // peer exchange (kg_peer_export / kg_peer_connect): scatter kernels store straight into the owners' receive buffers
    bool peer_ready = false;
    u64* peer_recv[2][KG_MAX_WORLD] = {};   // receive buffers of every rank as mapped into this process ([parity][rank])
    void* peer_opened[2][KG_MAX_WORLD] = {};// IPC mappings to close at destroy
    u64** d_peer_ptrs[2] = {nullptr, nullptr};      // device copies of peer_recv[parity][*]
    u64 *d_remote_base[2] = {nullptr, nullptr}, *h_remote_base[2] = {nullptr, nullptr};   // [KG_MAX_BUCKETS] per parity
    u32* d_barrier = nullptr;           // world + 1 words for the "all scatters have landed" collective
    uint64_t peer_rounds = 0, peer_fallback_rounds = 0;
    u32* d_work = nullptr;              // work counter of the persistent insert kernels
    u32 insert_grid = 148 * 8;          // resident blocks of the grid-stride insert kernels (SMs x blocks/SM)
    // bit-exact emulation of the reference's double Bloom filter (kg_refbloom.cuh; opt-in, unvalidated on hardware)
    bool ref_bloom = false;
    KgRefBloom rb{nullptr, nullptr, 0, 0, 0};
    struct RbBatch { u64* words; u32* brk; KgStream* st; u32 nthreads; };
    std::vector<RbBatch> rb_log;        // packed stream of every batch of the Bloom pass, kept for sweeps 2 and 3
    int rb_streams = 0;                 // streams begun in the current pass (the emulation needs exactly one)
    bool feed_prefetch = false;         // KG_FEED_PREFETCH=1: pipelined H2D in kg_feed (unmeasured; opt-in)
    bool parse_tma = false;             // KG_PARSE_TMA=1: parse tiles staged by TMA bulk copies (unmeasured; opt-in)
    // Kaarme representation (after kg_compact)
    KgKaarme kaarme{nullptr, nullptr, 0, 0};
    KgCompactStats* d_cstats = nullptr;
    bool compacted = false, counted = false;
    std::string err;
};

static thread_local std::string g_err;

// NCCL is resolved with dlopen at first use instead of a link-time dependency: inside a Python process torch
// has already loaded its own (newer) libnccl.so.2, and a DT_NEEDED on the system copy would either shadow it
// (breaking `import torch`) or be shadowed by it.  dlopen by soname returns whichever copy is already mapped.
struct KgNccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
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
            KG_SYM(GetUniqueId); KG_SYM(CommInitRank); KG_SYM(CommDestroy); KG_SYM(AllGather); KG_SYM(Send);
            KG_SYM(Recv); KG_SYM(GroupStart); KG_SYM(GroupEnd); KG_SYM(GetErrorString);
#undef KG_SYM
            n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.Send && n.Recv &&
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
static void free_all(kg_ctx* c) {
    cudaSetDevice(c->cfg.device);
    if (c->s_compute) cudaStreamSynchronize(c->s_compute);
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
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
    }
    if (c->h_out_n) cudaFreeHost(c->h_out_n);
    if (c->h_text_cur) cudaFreeHost(c->h_text_cur);
    cudaFree(c->kaarme.slots); cudaFree(c->kaarme.roots); cudaFree(c->d_cstats); cudaFree(c->d_work);
    if (c->s_comm) cudaStreamSynchronize(c->s_comm);
    if (c->s_insert) cudaStreamSynchronize(c->s_insert);
    if (c->s_ctl) cudaStreamSynchronize(c->s_ctl);
    for (int i = 0; i < 2; i++) {
        for (int r = 0; r < KG_MAX_WORLD; r++) if (c->peer_opened[i][r]) cudaIpcCloseMemHandle(c->peer_opened[i][r]);
        cudaFree(c->d_peer_ptrs[i]); cudaFree(c->d_remote_base[i]);
        if (c->h_remote_base[i]) cudaFreeHost(c->h_remote_base[i]);
    }
    cudaFree(c->d_barrier);
    for (auto& b : c->rb_log) { cudaFree(b.words); cudaFree(b.brk); cudaFree(b.st); }
    c->rb_log.clear();
    cudaFree(c->rb.T1); cudaFree(c->rb.T2);
    if (c->comm) { kg_nccl().CommDestroy(c->comm); c->comm = nullptr; }
    if (c->ctl_comm) { kg_nccl().CommDestroy(c->ctl_comm); c->ctl_comm = nullptr; }
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_send[i]); cudaFree(c->d_recv[i]);
        if (c->ev_send_free[i]) cudaEventDestroy(c->ev_send_free[i]);
        if (c->ev_recv_free[i]) cudaEventDestroy(c->ev_recv_free[i]);
        if (c->ev_recv_full[i]) cudaEventDestroy(c->ev_recv_full[i]);
    }
    cudaFree(c->d_blk_hist); cudaFree(c->d_blk_base); cudaFree(c->d_bucket_counts); cudaFree(c->d_bucket_offs); cudaFree(c->d_matrix);
    if (c->h_matrix) cudaFreeHost(c->h_matrix);
    for (int i = 0; i < 2; i++) { cudaFree(c->d_seg[i]); if (c->h_seg[i]) cudaFreeHost(c->h_seg[i]); }
    for (cudaEvent_t e : {c->ev_counts, c->ev_scatter, c->ev_matrix, c->ev_pass_ready, c->ev_tail}) if (e) cudaEventDestroy(e);
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    if (c->s_insert) cudaStreamDestroy(c->s_insert);
    if (c->s_ctl) cudaStreamDestroy(c->s_ctl);
    cudaFree(c->d_tile_hdr_eff); cudaFree(c->d_tile_hdr_in); cudaFree(c->d_tile_nbases);
    cudaFree(c->d_tile_pend_eff); cudaFree(c->d_tile_pend_in); cudaFree(c->d_tile_off);
    cudaFree(c->d_words); cudaFree(c->d_brk); cudaFree(c->d_carry_words); cudaFree(c->d_carry_brk);
    cudaFree(c->d_stream); cudaFree(c->d_stats); cudaFree(c->table.slots); cudaFree(c->bloom.bits);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto e : c->ins_pool) cudaEventDestroy(e);
    if (c->ev_pass_begin) cudaEventDestroy(c->ev_pass_begin);
    if (c->ev_pass_end) cudaEventDestroy(c->ev_pass_end);
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    cudaGetLastError();
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
    c->insert_grid = (u32)prop.multiProcessorCount * 8u;
    if (const char* e = getenv("KG_INSERT_GRID")) c->insert_grid = (u32)atoi(e) * (u32)prop.multiProcessorCount;
    if (const char* e = getenv("KG_PARSE_TMA")) c->parse_tma = atoi(e) != 0;
    if (const char* e = getenv("KG_FEED_PREFETCH")) c->feed_prefetch = atoi(e) != 0;
    c->W = (int)((cfg->k + 31) / 32);
    c->batch_bytes = cfg->batch_bytes ? cfg->batch_bytes : KG_DEFAULT_BATCH;
    if (c->batch_bytes > KG_MAX_BATCH) c->batch_bytes = KG_MAX_BATCH;
    c->batch_bytes = (c->batch_bytes + KG_TILE - 1) / KG_TILE * KG_TILE;
    c->max_tiles = (uint32_t)(c->batch_bytes / KG_TILE);
    c->carry_max_words = (u32)c->W + 2;
    c->words_cap = c->batch_bytes / 32 + c->carry_max_words + 8;

#define KG_TRY(call)                              \
    do {                                          \
        int s_ = [&]() -> int { KG_CUDA(c, call); return KG_OK; }(); \
        if (s_ != KG_OK) { g_err = c->err; free_all(c); delete c; return s_; } \
    } while (0)

    KG_TRY(cudaSetDevice(cfg->device));
    KG_TRY(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
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
    KG_TRY(cudaMalloc(&c->d_words, sizeof(u64) * c->words_cap));
    KG_TRY(cudaMalloc(&c->d_brk, sizeof(u32) * c->words_cap));
    KG_TRY(cudaMalloc(&c->d_carry_words, sizeof(u64) * c->carry_max_words));
    KG_TRY(cudaMalloc(&c->d_carry_brk, sizeof(u32) * c->carry_max_words));
    KG_TRY(cudaMalloc(&c->d_stream, sizeof(KgStream)));
    KG_TRY(cudaMalloc(&c->d_stats, sizeof(KgStats)));
    KG_TRY(cudaMalloc(&c->d_work, sizeof(u32) * 4));
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
        // every shard holds m/world bits per filter (rounded up to whole 64-bit words, at least one)
        uint64_t m_local = (m + (uint64_t)cfg->world - 1) / (uint64_t)cfg->world;
        uint64_t nwords = (m_local + 63) / 64;
        if (nwords == 0) nwords = 1;
        c->bloom_m = m;
        c->bloom.nblocks = nwords;
        c->bloom.nh = nh;
        c->bloom.world = (u32)cfg->world;
        c->bloom_bytes = nwords * 16;                 // [F1 word][F2 word] pairs: 2m bits, as the reference's interleaved array
        KG_TRY(cudaMalloc(&c->bloom.bits, c->bloom_bytes));
    }
    // partitions: 0 = choose per pass from the table / filter size, 1 = never bucket on one GPU, > 1 = as given
    c->bucketed = cfg->world > 1 || (cfg->partitions != 1 && cfg->table_mode != KG_TABLE_KAARME && !c->ref_bloom);
    if (c->bucketed) {
        const size_t max_words = c->batch_bytes / 32 + c->carry_max_words + 2;
        c->max_blocks = (u32)((max_words + 31) / 32);   // smallest block of the hist/scatter pair covers 32 words
        c->send_cap = c->batch_bytes + 64;                       // a batch of n bytes holds < n k-mers
        c->recv_cap = cfg->world > 1 ? 2 * c->send_cap : 0;      // single GPU inserts straight from the send buffer
        KG_TRY(cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
        KG_TRY(cudaStreamCreateWithFlags(&c->s_insert, cudaStreamNonBlocking));
        KG_TRY(cudaStreamCreateWithFlags(&c->s_ctl, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            KG_TRY(cudaEventCreateWithFlags(&c->ev_send_free[i], cudaEventDisableTiming));
            KG_TRY(cudaEventCreateWithFlags(&c->ev_recv_free[i], cudaEventDisableTiming));
            KG_TRY(cudaEventCreateWithFlags(&c->ev_recv_full[i], cudaEventDisableTiming));
        }
        KG_TRY(cudaEventCreateWithFlags(&c->ev_counts, cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_scatter, cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_matrix, cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_pass_ready, cudaEventDisableTiming));
        KG_TRY(cudaEventCreateWithFlags(&c->ev_tail, cudaEventDisableTiming));
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
    static_assert(2 * sizeof(ncclUniqueId) <= KG_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id[2];
    kg_ctx* none = nullptr;
    if (!kg_nccl().ok) { g_err = "libnccl.so.2 could not be loaded"; return KG_ENCCL; }
    KG_NCCL(none, kg_nccl().GetUniqueId(&id[0]));     // key transfers
    KG_NCCL(none, kg_nccl().GetUniqueId(&id[1]));     // control (count all-gather)
    memset(id_out, 0, KG_UNIQUE_ID_BYTES);
    memcpy(id_out, id, sizeof(id));
    return KG_OK;
}

extern "C" int kg_comm_init(kg_ctx* c, const void* id, int rank, int world) {
    if (!c || !id) return KG_EBADARG;
    if (rank != c->cfg.rank || world != c->cfg.world || world < 2) { c->err = "kg_comm_init: rank/world differ from kg_config"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    ncclUniqueId uid[2];
    memcpy(uid, id, sizeof(uid));
    if (!kg_nccl().ok) { c->err = "libnccl.so.2 could not be loaded"; return KG_ENCCL; }
    KG_NCCL(c, kg_nccl().CommInitRank(&c->comm, world, uid[0], rank));
    KG_NCCL(c, kg_nccl().CommInitRank(&c->ctl_comm, world, uid[1], rank));
    return KG_OK;
}

// Decide how the coming pass buckets its batches and size the scratch for it.  region_bytes = what the inserts
// of this pass hit at random (count table, or the Bloom filter): local partitions are chosen so that one
// partition's region is ~16-32 MiB, comfortably L2-resident next to the streaming keys.
static int setup_pass_buckets(kg_ctx* c, size_t region_bytes) {
    c->pass_bucketed = false;
    c->pl = 1;
    c->nb = (u32)c->cfg.world;
    if (!c->bucketed) return KG_OK;
    const u32 world = (u32)c->cfg.world;
    u32 pl = c->cfg.partitions;
    if (c->cfg.table_mode == KG_TABLE_KAARME) pl = 1;          // occurrence records do not travel with bare keys
    else if (pl == 0) {
        pl = 1;
        if (region_bytes > (96u << 20)) while ((size_t)pl * (24u << 20) < region_bytes && pl * 2 * world <= KG_MAX_BUCKETS) pl *= 2;
    }
    if (pl * world > KG_MAX_BUCKETS) pl = KG_MAX_BUCKETS / world;
    if (pl < 1) pl = 1;
    c->pl = pl;
    c->nb = world * pl;
    c->pass_bucketed = world > 1 || pl > 1;
    if (!c->pass_bucketed) return KG_OK;
    // one-pass (reserve) scatter on a single GPU: fixed-capacity bucket regions, 25 % slack + 8192 keys
    size_t need = c->send_cap;
    c->reserve_cap = 0;
    if (world == 1 && c->W <= 4 && !getenv("KG_NO_RESERVE")) {
        c->reserve_cap = (u32)(c->send_cap / pl + c->send_cap / pl / 4 + 8192);
        need = (size_t)c->reserve_cap * pl;
    }
    if (need > c->send_alloc) {
        for (int i = 0; i < 2; i++) { cudaFree(c->d_send[i]); c->d_send[i] = nullptr; }
        c->send_alloc = need;
    }
    for (int i = 0; i < 2; i++) {
        if (!c->d_send[i]) KG_CUDA(c, cudaMalloc(&c->d_send[i], c->send_alloc * c->W * sizeof(u64)));
        if (c->recv_cap && !c->d_recv[i]) KG_CUDA(c, cudaMalloc(&c->d_recv[i], c->recv_cap * c->W * sizeof(u64)));
    }
    if (c->nb > c->nb_alloc) {
        cudaFree(c->d_blk_hist); cudaFree(c->d_blk_base); cudaFree(c->d_bucket_counts); cudaFree(c->d_bucket_offs); cudaFree(c->d_matrix);
        if (c->h_matrix) cudaFreeHost(c->h_matrix);
        c->d_blk_hist = c->d_blk_base = c->d_bucket_counts = c->d_bucket_offs = c->d_matrix = nullptr; c->h_matrix = nullptr;
        c->nb_alloc = 0;   // nothing is allocated until every buffer below exists (an allocation may fail half way)
        KG_CUDA(c, cudaMalloc(&c->d_blk_hist, sizeof(u32) * (size_t)c->max_blocks * c->nb));
        KG_CUDA(c, cudaMalloc(&c->d_blk_base, sizeof(u32) * (size_t)c->max_blocks * c->nb));
        KG_CUDA(c, cudaMalloc(&c->d_bucket_counts, sizeof(u32) * (c->nb + 4)));
        KG_CUDA(c, cudaMalloc(&c->d_bucket_offs, sizeof(u32) * (c->nb + 1)));
        KG_CUDA(c, cudaMalloc(&c->d_matrix, sizeof(u32) * (size_t)(c->nb + 1) * world));
        KG_CUDA(c, cudaHostAlloc((void**)&c->h_matrix, sizeof(u32) * (size_t)(c->nb + 1) * world, cudaHostAllocDefault));
        c->nb_alloc = c->nb;
    }
    if (c->nb + 1 > c->seg_cap) {
        for (int i = 0; i < 2; i++) {
            cudaFree(c->d_seg[i]); if (c->h_seg[i]) cudaFreeHost(c->h_seg[i]);
            c->d_seg[i] = nullptr; c->h_seg[i] = nullptr;
        }
        c->seg_cap = 0;
        for (int i = 0; i < 2; i++) {
            KG_CUDA(c, cudaMalloc(&c->d_seg[i], sizeof(u64) * 2 * (c->nb + 2)));
            KG_CUDA(c, cudaHostAlloc((void**)&c->h_seg[i], sizeof(u64) * 2 * (c->nb + 2), cudaHostAllocDefault));
        }
        c->seg_cap = c->nb + 1;
    }
    return KG_OK;
}

extern "C" int kg_pass_begin(kg_ctx* c, int pass) {
    if (!c || (pass != KG_PASS_BLOOM && pass != KG_PASS_COUNT)) return KG_EBADARG;
    if (pass == KG_PASS_BLOOM && !c->cfg.use_bloom) { c->err = "Bloom pass without use_bloom"; return KG_EBADARG; }
    if (pass == KG_PASS_COUNT && c->cfg.use_bloom && !c->bloom_done) { c->err = "count pass before Bloom pass"; return KG_EBADARG; }
    if (c->cfg.world > 1 && !c->comm) { c->err = "world > 1 needs kg_comm_init first"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    c->pass = pass;
    c->stream_open = false;
    c->ev_used = 0;
    c->ins_used = 0;
    c->ins_launches = 0;
    c->raw_bytes_pass = 0;
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
        // the count field stops at 2^(128-2k) - 2^17 (>= 66.9 M) instead of wrapping into the key
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
        c->table.world = (u32)c->cfg.world;
        c->table.packed_tb = packed_tb;
        KG_CUDA(c, cudaMemsetAsync(c->table.slots, 0, bytes, c->s_compute));
        {
            // the partition count must be identical on every rank (it fixes the bucket layout of the exchange):
            // derive it from configuration only.  After a Bloom pass the shard tables differ a little in size
            // (2 x the LOCAL new_in_second), so use the configured estimate there.
            size_t region = bytes;
            if (c->cfg.world > 1 && c->cfg.use_bloom)
                region = (size_t)(2 * c->cfg.expected_unique / (uint64_t)c->cfg.world) * stride * sizeof(u64);
            int rc = setup_pass_buckets(c, region);
            if (rc) return rc;
        }
    }
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

template <int W>
static void launch_insert_keys(kg_ctx* c, cudaStream_t s, const u64* keys, u64 n_upper, const u32* n_dev, int sink) {
    const u32 block = 256;
    u64 grid = (n_upper + KG_CHUNK - 1) / KG_CHUNK;
    if (grid > c->insert_grid) grid = c->insert_grid;    // persistent kernel: just enough blocks to fill the GPU
    if (grid == 0) return;
    cudaMemsetAsync(c->d_work, 0, sizeof(u32), s);
    switch (sink) {
        case KG_SINK_TABLE: kg_insert_keys_kernel<W, KG_SINK_TABLE><<<(u32)grid, block, 0, s>>>(keys, n_upper, n_dev, c->table, c->bloom, c->d_stats, c->d_work); break;
        case KG_SINK_BLOOM1: kg_insert_keys_kernel<W, KG_SINK_BLOOM1><<<(u32)grid, block, 0, s>>>(keys, n_upper, n_dev, c->table, c->bloom, c->d_stats, c->d_work); break;
        case KG_SINK_BLOOM2: kg_insert_keys_kernel<W, KG_SINK_BLOOM2><<<(u32)grid, block, 0, s>>>(keys, n_upper, n_dev, c->table, c->bloom, c->d_stats, c->d_work); break;
        default: break;
    }
    c->launches++;
}

static void insert_keys(kg_ctx* c, cudaStream_t s, const u64* keys, u64 n_upper, const u32* n_dev) {
    const int sink = current_sink(c);
    cudaEvent_t ia = next_ins_event(c), ib = next_ins_event(c);
    if (ia) cudaEventRecord(ia, s);
    c->ins_launches++;
    switch (c->W) {
        case 1: launch_insert_keys<1>(c, s, keys, n_upper, n_dev, sink); break;
        case 2: launch_insert_keys<2>(c, s, keys, n_upper, n_dev, sink); break;
        case 3: launch_insert_keys<3>(c, s, keys, n_upper, n_dev, sink); break;
        case 4: launch_insert_keys<4>(c, s, keys, n_upper, n_dev, sink); break;
        case 5: launch_insert_keys<5>(c, s, keys, n_upper, n_dev, sink); break;
        case 6: launch_insert_keys<6>(c, s, keys, n_upper, n_dev, sink); break;
        case 7: launch_insert_keys<7>(c, s, keys, n_upper, n_dev, sink); break;
        case 8: launch_insert_keys<8>(c, s, keys, n_upper, n_dev, sink); break;
    }
    if (ib) cudaEventRecord(ib, s);
}

template <int W>
static void launch_bucket(kg_ctx* c, const KgBucketArgs& a, u32 nwords, bool scatter) {
    using G = KgBucketGeom<W>;
    const u32 grid = (nwords + G::WPB - 1) / G::WPB;
    if (scatter) {
        const size_t smem = G::smem_bytes(a.nb);
        if (!c->scatter_configured) {   // per device (one context per device, possibly several in one process)
            cudaFuncSetAttribute(kg_owner_scatter<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes(KG_MAX_BUCKETS));
            c->scatter_configured = true;
        }
        kg_owner_scatter<W><<<grid, G::TPB, smem, c->s_compute>>>(a);
    } else {
        kg_owner_hist<W><<<grid, G::WPB, 0, c->s_compute>>>(a);
    }
    c->launches++;
}
template <int W>
static u32 bucket_grid(u32 nwords) { return (nwords + KgBucketGeom<W>::WPB - 1) / KgBucketGeom<W>::WPB; }
static u32 bucket_blocks(const kg_ctx* c, u32 nwords) {
    switch (c->W) {
        case 1: return bucket_grid<1>(nwords);
        case 2: return bucket_grid<2>(nwords);
        case 3: return bucket_grid<3>(nwords);
        case 4: return bucket_grid<4>(nwords);
        case 5: return bucket_grid<5>(nwords);
        case 6: return bucket_grid<6>(nwords);
        case 7: return bucket_grid<7>(nwords);
        default: return bucket_grid<8>(nwords);
    }
}
static void bucket_kernel(kg_ctx* c, const KgBucketArgs& a, u32 grid, bool scatter) {
    switch (c->W) {
        case 1: launch_bucket<1>(c, a, grid, scatter); break;
        case 2: launch_bucket<2>(c, a, grid, scatter); break;
        case 3: launch_bucket<3>(c, a, grid, scatter); break;
        case 4: launch_bucket<4>(c, a, grid, scatter); break;
        case 5: launch_bucket<5>(c, a, grid, scatter); break;
        case 6: launch_bucket<6>(c, a, grid, scatter); break;
        case 7: launch_bucket<7>(c, a, grid, scatter); break;
        case 8: launch_bucket<8>(c, a, grid, scatter); break;
    }
}

template <int W>
static void launch_insert_segs(kg_ctx* c, cudaStream_t s, const u64* keys, const u64* seg, u32 nseg, u64 n, int sink) {
    u64 grid = (n + KG_CHUNK - 1) / KG_CHUNK;
    if (grid > c->insert_grid) grid = c->insert_grid;
    if (grid == 0) return;
    cudaMemsetAsync(c->d_work, 0, sizeof(u32), s);
    const u64* start = seg;
    const u64* src = seg + (nseg + 1);
    switch (sink) {
        case KG_SINK_TABLE: kg_insert_segs_kernel<W, KG_SINK_TABLE><<<(u32)grid, 256, 0, s>>>(keys, start, src, nseg, c->table, c->bloom, c->d_stats, c->d_work); break;
        case KG_SINK_BLOOM1: kg_insert_segs_kernel<W, KG_SINK_BLOOM1><<<(u32)grid, 256, 0, s>>>(keys, start, src, nseg, c->table, c->bloom, c->d_stats, c->d_work); break;
        case KG_SINK_BLOOM2: kg_insert_segs_kernel<W, KG_SINK_BLOOM2><<<(u32)grid, 256, 0, s>>>(keys, start, src, nseg, c->table, c->bloom, c->d_stats, c->d_work); break;
        default: break;
    }
    c->launches++;
}

template <int W>
static void launch_reserve(kg_ctx* c, const KgReserveArgs& a, u32 nwords, int sink) {
    if constexpr (W <= 4) {
        using G = KgBucketGeom<W>;
        const u32 grid = (nwords + G::WPB - 1) / G::WPB;
        const size_t smem = (size_t)G::KEYS * W * 8 + (size_t)a.nb * 12 + (size_t)G::KEYS * 2 + 64;
        const int max_smem = (int)((size_t)G::KEYS * W * 8 + (size_t)KG_MAX_BUCKETS * 12 + (size_t)G::KEYS * 2 + 64);
        switch (sink) {
            case KG_SINK_TABLE:
                cudaFuncSetAttribute(kg_scatter_reserve<W, KG_SINK_TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
                kg_scatter_reserve<W, KG_SINK_TABLE><<<grid, G::TPB, smem, c->s_compute>>>(a); break;
            case KG_SINK_BLOOM1:
                cudaFuncSetAttribute(kg_scatter_reserve<W, KG_SINK_BLOOM1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
                kg_scatter_reserve<W, KG_SINK_BLOOM1><<<grid, G::TPB, smem, c->s_compute>>>(a); break;
            case KG_SINK_BLOOM2:
                cudaFuncSetAttribute(kg_scatter_reserve<W, KG_SINK_BLOOM2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
                kg_scatter_reserve<W, KG_SINK_BLOOM2><<<grid, G::TPB, smem, c->s_compute>>>(a); break;
            default: break;
        }
        c->launches++;
    }
}

// One exchange round (collective when world > 1).  have_batch: this rank's send buffer (round & 1) was just
// filled by bucket_batch; otherwise the rank contributes nothing and reports "done".  Returns via *all_done
// whether every rank reported done in this round.
//   buckets are owner-major: owner d holds buckets [d*pl, (d+1)*pl) = its local partitions, in table order
//   s_ctl:    all-gather of the per-bucket counts -> host (own communicator, never behind a key transfer)
//   s_comm:   grouped ncclSend/ncclRecv of the key slices
//   s_insert: insert kernel over what arrived (partition-major across senders through a segment table),
//             overlapping the next batch's parse + bucketing on s_compute
static int exchange_round(kg_ctx* c, bool have_batch, bool* all_done, bool counts_gathered = false) {
    const u32 nb = c->nb, pl = c->pl, world = (u32)c->cfg.world, rank = (u32)c->cfg.rank;
    const int sb = (int)(c->round & 1);
    const size_t row = nb + 1;
    if (!counts_gathered) {          // (the peer path gathers the counts itself before it decides to fall back here)
    if (!have_batch) {
        KG_CUDA(c, cudaMemsetAsync(c->d_bucket_counts, 0, sizeof(u32) * nb, c->s_compute));
        KG_CUDA(c, cudaMemsetAsync(c->d_bucket_counts + nb, 1, sizeof(u32), c->s_compute));   // non-zero = done
        KG_CUDA(c, cudaEventRecord(c->ev_counts, c->s_compute));
    }
    KG_CUDA(c, cudaStreamWaitEvent(c->s_ctl, c->ev_counts, 0));
    KG_NCCL(c, kg_nccl().AllGather(c->d_bucket_counts, c->d_matrix, row, ncclUint32, c->ctl_comm, c->s_ctl));
    KG_CUDA(c, cudaMemcpyAsync(c->h_matrix, c->d_matrix, sizeof(u32) * row * world, cudaMemcpyDeviceToHost, c->s_ctl));
    KG_CUDA(c, cudaEventRecord(c->ev_matrix, c->s_ctl));
    KG_CUDA(c, cudaEventSynchronize(c->ev_matrix));
    }
    const u32* M = c->h_matrix;   // M[r*row + b] = keys rank r holds for bucket b; M[r*row + nb] = done flag
    auto to_owner = [&](u32 r, u32 d) { u64 t = 0; for (u32 p = 0; p < pl; p++) t += M[r * row + d * pl + p]; return t; };
    bool done = true;
    u64 max_in = 0, any = 0;
    for (u32 r = 0; r < world; r++) done = done && M[r * row + nb] != 0;
    for (u32 d = 0; d < world; d++) {
        u64 in = 0;
        for (u32 r = 0; r < world; r++) in += to_owner(r, d);
        if (in > max_in) max_in = in;
        any += in;
    }
    *all_done = done;
    if (any) {
        // sub-rounds so that no rank receives more than recv_cap keys at once (identical on every rank)
        const u64 S = (max_in + c->recv_cap - 1) / c->recv_cap;
        std::vector<u64> send_off(world + 1, 0);
        for (u32 d = 0; d < world; d++) send_off[d + 1] = send_off[d] + to_owner(rank, d);
        if (have_batch) KG_CUDA(c, cudaStreamWaitEvent(c->s_comm, c->ev_scatter, 0));
        for (u64 sr = 0; sr < S; sr++) {
            const int rb = (int)(c->subround & 1);
            KG_CUDA(c, cudaStreamWaitEvent(c->s_comm, c->ev_recv_free[rb], 0));
            KG_CUDA(c, cudaEventSynchronize(c->ev_recv_full[rb]));   // the previous upload of h_seg[rb] has been consumed
            std::vector<u64> recv_off(world + 1, 0), rlo_of(world), rhi_of(world);
            KG_NCCL(c, kg_nccl().GroupStart());
            for (u32 peer = 0; peer < world; peer++) {
                const u64 ns = to_owner(rank, peer), lo = ns * sr / S, hi = ns * (sr + 1) / S;      // my slice for peer
                const u64 nr = to_owner(peer, rank), rlo = nr * sr / S, rhi = nr * (sr + 1) / S;    // peer's slice for me
                const u64* src = c->d_send[sb] + (send_off[peer] + lo) * c->W;
                u64* dst = c->d_recv[rb] + recv_off[peer] * c->W;
                if (peer == rank) {
                    if (hi > lo) KG_CUDA(c, cudaMemcpyAsync(dst, src, (hi - lo) * c->W * sizeof(u64), cudaMemcpyDeviceToDevice, c->s_comm));
                } else {
                    if (hi > lo) KG_NCCL(c, kg_nccl().Send(src, (hi - lo) * c->W, ncclUint64, (int)peer, c->comm, c->s_comm));
                    if (rhi > rlo) KG_NCCL(c, kg_nccl().Recv(dst, (rhi - rlo) * c->W, ncclUint64, (int)peer, c->comm, c->s_comm));
                }
                rlo_of[peer] = rlo; rhi_of[peer] = rhi;
                recv_off[peer + 1] = recv_off[peer] + (rhi - rlo);
            }
            KG_NCCL(c, kg_nccl().GroupEnd());
            const u64 n_recv = recv_off[world];
            // segment table, partition-major across senders: sender s delivered keys [rlo, rhi) of its run for me,
            // which is its partitions 0..pl-1 back to back
            u32 nseg = 0;
            u64* hs = c->h_seg[rb];
            const u32 maxseg = c->nb;                      // world * pl segments at most
            u64* h_start = hs;
            u64* h_src = hs + (maxseg + 1);
            u64 acc = 0;
            for (u32 p = 0; p < pl; p++) {
                for (u32 sdr = 0; sdr < world; sdr++) {
                    u64 pbeg = 0;                           // start of partition p inside sender sdr's run for me
                    for (u32 q = 0; q < p; q++) pbeg += M[sdr * row + rank * pl + q];
                    const u64 pend = pbeg + M[sdr * row + rank * pl + p];
                    const u64 a0 = pbeg > rlo_of[sdr] ? pbeg : rlo_of[sdr];
                    const u64 a1 = pend < rhi_of[sdr] ? pend : rhi_of[sdr];
                    if (a1 > a0) {
                        h_start[nseg] = acc;
                        h_src[nseg] = recv_off[sdr] + (a0 - rlo_of[sdr]);
                        acc += a1 - a0;
                        nseg++;
                    }
                }
            }
            h_start[nseg] = acc;
            if (n_recv) {
                // device layout: start[0..nseg], then src[0..nseg-1]
                KG_CUDA(c, cudaMemcpyAsync(c->d_seg[rb], h_start, sizeof(u64) * (nseg + 1), cudaMemcpyHostToDevice, c->s_comm));
                KG_CUDA(c, cudaMemcpyAsync(c->d_seg[rb] + (nseg + 1), h_src, sizeof(u64) * nseg, cudaMemcpyHostToDevice, c->s_comm));
            }
            KG_CUDA(c, cudaEventRecord(c->ev_recv_full[rb], c->s_comm));
            KG_CUDA(c, cudaStreamWaitEvent(c->s_insert, c->ev_recv_full[rb], 0));
            if (n_recv) {
                const int sink = current_sink(c);
                cudaEvent_t ia = next_ins_event(c), ib = next_ins_event(c);
                if (ia) cudaEventRecord(ia, c->s_insert);
                c->ins_launches++;
                KG_DISPATCH_W(c->W, launch_insert_segs, c, c->s_insert, c->d_recv[rb], c->d_seg[rb], nseg, n_recv, sink);
                if (ib) cudaEventRecord(ib, c->s_insert);
            }
            KG_CUDA(c, cudaEventRecord(c->ev_recv_free[rb], c->s_insert));
            c->subround++;
        }
    }
    KG_CUDA(c, cudaEventRecord(c->ev_send_free[sb], c->s_comm));
    c->round++;
    return KG_OK;
}

// ---- peer exchange: the fused bucket -> peer-store path --------------------------------------------------------
// With kg_peer_connect every rank has every other rank's two receive buffers mapped (CUDA IPC between processes, plain
// peer access inside one process).  A round then is:
//   s_compute: hist -> column scan (bucket_batch)                                  counts of this batch
//   s_ctl:     wait "my receive buffer of this parity is free" -> all-gather of the counts -> host
//              (so the gathered matrix also says: EVERY rank's buffer of this parity is free)
//   host:      kg_peer_plan: where each of my (bucket) runs starts in its owner's buffer (partition-major there)
//   s_compute: kg_owner_scatter_peer: the scatter's coalesced 16-byte runs go straight to the owners over NVLink
//   s_comm:    one-word all-gather = "all scatters of this round have completed" (kernel completion makes peer
//              stores visible)
//   s_insert:  kg_insert_keys_kernel over my receive buffer, front to back = partition-major (L2-blocked, section 5)
// No send buffer, no NCCL copy kernels on the SMs, no segment table; NVLink carries each key once.
// A round whose busiest owner would overflow a receive buffer (heavy skew) falls back to the ncclSend/ncclRecv path.
struct KgPeerHandle {               // KG_PEER_HANDLE_BYTES on the wire
    uint64_t magic, pid;
    int32_t rank, device;
    uint64_t ptr[2], cap_keys;
    cudaIpcMemHandle_t ipc[2];
};
static_assert(sizeof(KgPeerHandle) <= KG_PEER_HANDLE_BYTES, "peer handle size");
#define KG_PEER_MAGIC 0x4B47504545523031ULL   // "KGPEER01"

extern "C" int kg_peer_export(kg_ctx* c, void* handle_out) {
    if (!c || !handle_out) return KG_EBADARG;
    if (c->cfg.world < 2 || !c->bucketed) { c->err = "kg_peer_export needs world > 1"; return KG_EBADARG; }
    if (c->pass) { c->err = "kg_peer_export inside a pass"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    for (int i = 0; i < 2; i++)
        if (!c->d_recv[i]) KG_CUDA(c, cudaMalloc(&c->d_recv[i], c->recv_cap * c->W * sizeof(u64)));
    KgPeerHandle h;
    memset(&h, 0, sizeof(h));
    h.magic = KG_PEER_MAGIC;
    h.pid = (uint64_t)getpid();
    h.rank = c->cfg.rank;
    h.device = c->cfg.device;
    h.cap_keys = c->recv_cap;
    for (int i = 0; i < 2; i++) {
        h.ptr[i] = (uint64_t)(uintptr_t)c->d_recv[i];
        KG_CUDA(c, cudaIpcGetMemHandle(&h.ipc[i], c->d_recv[i]));
    }
    memset(handle_out, 0, KG_PEER_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    return KG_OK;
}

extern "C" int kg_peer_connect(kg_ctx* c, const void* handles, int world) {
    if (!c || !handles) return KG_EBADARG;
    if (world != c->cfg.world || world < 2 || world > KG_MAX_WORLD) { c->err = "kg_peer_connect: world differs from kg_config"; return KG_EBADARG; }
    if (c->pass || c->peer_ready) { c->err = "kg_peer_connect: call once, outside a pass"; return KG_EBADARG; }
    if (!c->d_recv[0] || !c->d_recv[1]) { c->err = "kg_peer_connect before kg_peer_export"; return KG_EBADARG; }
    KG_CUDA(c, cudaSetDevice(c->cfg.device));
    const uint64_t me = (uint64_t)getpid();
    for (int r = 0; r < world; r++) {
        KgPeerHandle h;
        memcpy(&h, (const char*)handles + (size_t)r * KG_PEER_HANDLE_BYTES, sizeof(h));
        if (h.magic != KG_PEER_MAGIC || h.rank != r || h.cap_keys != c->recv_cap) { c->err = "kg_peer_connect: bad handle (order must be rank order; same batch size on every rank)"; return KG_EBADARG; }
        for (int i = 0; i < 2; i++) {
            if (r == c->cfg.rank) {
                c->peer_recv[i][r] = c->d_recv[i];
            } else if (h.pid == me) {                       // another context of this process (one host thread per GPU)
                int can = 0;
                KG_CUDA(c, cudaDeviceCanAccessPeer(&can, c->cfg.device, h.device));
                if (!can) { c->err = "kg_peer_connect: no peer access between the devices"; return KG_ECUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) KG_CUDA(c, e);
                cudaGetLastError();
                c->peer_recv[i][r] = (u64*)(uintptr_t)h.ptr[i];
            } else {                                        // another process: map its allocation
                void* p = nullptr;
                KG_CUDA(c, cudaIpcOpenMemHandle(&p, h.ipc[i], cudaIpcMemLazyEnablePeerAccess));
                c->peer_opened[i][r] = p;
                c->peer_recv[i][r] = (u64*)p;
            }
        }
    }
    for (int i = 0; i < 2; i++) {
        KG_CUDA(c, cudaMalloc(&c->d_peer_ptrs[i], sizeof(u64*) * KG_MAX_WORLD));
        KG_CUDA(c, cudaMemcpy(c->d_peer_ptrs[i], c->peer_recv[i], sizeof(u64*) * KG_MAX_WORLD, cudaMemcpyHostToDevice));
        KG_CUDA(c, cudaMalloc(&c->d_remote_base[i], sizeof(u64) * KG_MAX_BUCKETS));
        KG_CUDA(c, cudaHostAlloc((void**)&c->h_remote_base[i], sizeof(u64) * KG_MAX_BUCKETS, cudaHostAllocDefault));
    }
    KG_CUDA(c, cudaMalloc(&c->d_barrier, sizeof(u32) * (KG_MAX_WORLD + 1)));
    KG_CUDA(c, cudaMemset(c->d_barrier, 0, sizeof(u32) * (KG_MAX_WORLD + 1)));
    c->peer_ready = true;
    return KG_OK;
}

extern "C" int kg_peer_stats(const kg_ctx* c, uint64_t* peer_rounds, uint64_t* fallback_rounds) {
    if (!c) return KG_EBADARG;
    if (peer_rounds) *peer_rounds = c->peer_rounds;
    if (fallback_rounds) *fallback_rounds = c->peer_fallback_rounds;
    return KG_OK;
}

template <int W>
static void launch_scatter_peer(kg_ctx* c, const KgBucketArgs& a, const KgPeerArgs& pa, u32 nwords) {
    using G = KgBucketGeom<W>;
    const u32 grid = (nwords + G::WPB - 1) / G::WPB;
    cudaFuncSetAttribute(kg_owner_scatter_peer<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes(KG_MAX_BUCKETS));
    kg_owner_scatter_peer<W><<<grid, G::TPB, G::smem_bytes(a.nb), c->s_compute>>>(a, pa);
    c->launches++;
}

// One round of the peer exchange (collective).  have_batch: hist + column scan of this rank's batch are queued on
// s_compute and ev_counts is recorded (bucket_batch); otherwise the rank contributes nothing and reports "done".
static int peer_round(kg_ctx* c, bool have_batch, const KgBucketArgs* a, u32 nthreads, bool* all_done) {
    const u32 nb = c->nb, pl = c->pl, world = (u32)c->cfg.world, rank = (u32)c->cfg.rank;
    const int rb = (int)(c->round & 1);
    const size_t row = nb + 1;
    if (!have_batch) {
        KG_CUDA(c, cudaMemsetAsync(c->d_bucket_counts, 0, sizeof(u32) * nb, c->s_compute));
        KG_CUDA(c, cudaMemsetAsync(c->d_bucket_counts + nb, 1, sizeof(u32), c->s_compute));   // non-zero = done
        KG_CUDA(c, cudaEventRecord(c->ev_counts, c->s_compute));
    }
    KG_CUDA(c, cudaStreamWaitEvent(c->s_ctl, c->ev_counts, 0));
    KG_CUDA(c, cudaStreamWaitEvent(c->s_ctl, c->ev_recv_free[rb], 0));   // gathered matrix => every rank's buffer rb is free
    KG_NCCL(c, kg_nccl().AllGather(c->d_bucket_counts, c->d_matrix, row, ncclUint32, c->ctl_comm, c->s_ctl));
    KG_CUDA(c, cudaMemcpyAsync(c->h_matrix, c->d_matrix, sizeof(u32) * row * world, cudaMemcpyDeviceToHost, c->s_ctl));
    KG_CUDA(c, cudaEventRecord(c->ev_matrix, c->s_ctl));
    KG_CUDA(c, cudaEventSynchronize(c->ev_matrix));
    const KgPeerPlan plan = kg_peer_plan(c->h_matrix, world, pl, rank);
    *all_done = plan.all_done;
    if (plan.max_in > c->recv_cap) {
        // heavy skew: the busiest owner cannot take its keys in one piece -> this round goes through the send buffer
        // and ncclSend/ncclRecv in sub-rounds (every rank sees the same matrix and takes the same decision)
        c->peer_fallback_rounds++;
        if (have_batch) {
            const int sb = (int)(c->round & 1);
            KgBucketArgs b = *a;
            b.out_keys = c->d_send[sb];
            KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_send_free[sb], 0));
            bucket_kernel(c, b, nthreads, true);
            KG_CUDA(c, cudaEventRecord(c->ev_scatter, c->s_compute));
        }
        return exchange_round(c, have_batch, all_done, /*counts_gathered=*/true);
    }
    c->peer_rounds++;
    if (plan.max_in == 0) { c->round++; return KG_OK; }     // nobody moves anything this round
    if (have_batch) {
        memcpy(c->h_remote_base[rb], plan.remote_base.data(), sizeof(u64) * nb);
        KG_CUDA(c, cudaMemcpyAsync(c->d_remote_base[rb], c->h_remote_base[rb], sizeof(u64) * nb, cudaMemcpyHostToDevice, c->s_compute));
        KgPeerArgs pa;
        pa.peer = c->d_peer_ptrs[rb];
        pa.remote_base = c->d_remote_base[rb];
        pa.pl = pl;
        KG_DISPATCH_W(c->W, launch_scatter_peer, c, *a, pa, nthreads);
        KG_CUDA(c, cudaEventRecord(c->ev_scatter, c->s_compute));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_comm, c->ev_scatter, 0));
    }
    // "every scatter of this round has completed": a collective on s_comm that each rank enters after its own scatter
    KG_NCCL(c, kg_nccl().AllGather(c->d_barrier + KG_MAX_WORLD, c->d_barrier, 1, ncclUint32, c->comm, c->s_comm));
    KG_CUDA(c, cudaEventRecord(c->ev_recv_full[rb], c->s_comm));
    KG_CUDA(c, cudaStreamWaitEvent(c->s_insert, c->ev_recv_full[rb], 0));
    if (plan.my_in) insert_keys(c, c->s_insert, c->d_recv[rb], plan.my_in, nullptr);   // contiguous, partition-major
    KG_CUDA(c, cudaEventRecord(c->ev_recv_free[rb], c->s_insert));
    c->round++;
    return KG_OK;
}

// bucket the k-mers of the batch that was just packed, then hand them on
//   single GPU, W <= 4 : one-pass reserve scatter -> segment table -> insert (kg_scatter_reserve)
//   otherwise          : hist -> scan -> scatter (exact layout, needed for the exchange) -> exchange / insert
static int bucket_batch(kg_ctx* c, u32 nthreads) {
    if (c->cfg.world == 1 && c->reserve_cap) {
        const int b = (int)(c->round & 1);
        const int sink = current_sink(c);
        u32* cursors = c->d_bucket_counts;
        KG_CUDA(c, cudaMemsetAsync(cursors, 0, sizeof(u32) * c->nb, c->s_compute));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_send_free[b], 0));     // insert(round-2) has drained this buffer
        KgReserveArgs r;
        r.words = c->d_words; r.brk = c->d_brk; r.st = c->d_stream; r.cursors = cursors; r.out_keys = c->d_send[b];
        r.stats = c->d_stats; r.table = c->table; r.bloom = c->bloom; r.k = c->cfg.k; r.nb = c->nb; r.cap = c->reserve_cap;
        KG_DISPATCH_W(c->W, launch_reserve, c, r, nthreads, sink);
        kg_seg_from_cursors<<<1, 1024, 0, c->s_compute>>>(cursors, c->nb, c->reserve_cap, c->d_seg[b], c->d_seg[b] + (c->nb + 1));
        c->launches++;
        KG_CUDA(c, cudaEventRecord(c->ev_scatter, c->s_compute));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_insert, c->ev_scatter, 0));
        cudaEvent_t ia = next_ins_event(c), ib = next_ins_event(c);
        if (ia) cudaEventRecord(ia, c->s_insert);
        c->ins_launches++;
        KG_DISPATCH_W(c->W, launch_insert_segs, c, c->s_insert, c->d_send[b], c->d_seg[b], c->nb, (u64)nthreads * 32u, sink);
        if (ib) cudaEventRecord(ib, c->s_insert);
        KG_CUDA(c, cudaEventRecord(c->ev_send_free[b], c->s_insert));
        c->round++;
        return KG_OK;
    }
    const u32 grid = bucket_blocks(c, nthreads);   // blocks of the hist / scatter pair
    const int sb = c->cfg.world > 1 ? (int)(c->round & 1) : 0;
    KgBucketArgs a;
    a.words = c->d_words; a.brk = c->d_brk; a.st = c->d_stream;
    a.blk_hist = c->d_blk_hist; a.blk_base = c->d_blk_base; a.bucket_offs = c->d_bucket_offs; a.out_keys = c->d_send[sb];
    a.stats = c->d_stats; a.k = c->cfg.k; a.nb = c->nb; a.world = (u32)c->cfg.world;
    bucket_kernel(c, a, nthreads, false);
    kg_bucket_colscan<<<c->nb, 1024, 0, c->s_compute>>>(c->d_blk_hist, c->d_blk_base, grid, c->nb, c->d_bucket_counts);
    kg_bucket_offsets<<<1, 1024, 0, c->s_compute>>>(c->d_bucket_counts, c->nb, c->d_bucket_offs);
    c->launches += 2;
    if (c->cfg.world > 1 && c->peer_ready) {   // fused bucket -> peer-store exchange (kg_peer_connect)
        KG_CUDA(c, cudaMemsetAsync(c->d_bucket_counts + c->nb, 0, sizeof(u32), c->s_compute));   // not done
        KG_CUDA(c, cudaEventRecord(c->ev_counts, c->s_compute));
        bool all_done;
        return peer_round(c, true, &a, nthreads, &all_done);
    }
    if (c->cfg.world > 1) {
        KG_CUDA(c, cudaMemsetAsync(c->d_bucket_counts + c->nb, 0, sizeof(u32), c->s_compute));   // not done
        KG_CUDA(c, cudaEventRecord(c->ev_counts, c->s_compute));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_send_free[sb], 0));   // round-2 sends have left this buffer
        bucket_kernel(c, a, nthreads, true);
        KG_CUDA(c, cudaEventRecord(c->ev_scatter, c->s_compute));
        bool all_done;
        return exchange_round(c, true, &all_done);
    }
    // single GPU, partitioned: the send buffer is partition-major; one insert launch walks it in order, so the
    // blocks in flight at any moment hit one or two table regions (L2-resident).  The insert runs on its own
    // stream and the key buffers alternate, so the (ALU-bound) parse + bucketing of the next batch overlaps the
    // (L2-latency-bound) insert of this one.
    {
        const int b = (int)(c->round & 1);
        a.out_keys = c->d_send[b];
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_send_free[b], 0));     // insert(round-2) has drained it
        bucket_kernel(c, a, nthreads, true);
        KG_CUDA(c, cudaMemcpyAsync(c->d_bucket_counts + c->nb + 1 + b, c->d_bucket_offs + c->nb, sizeof(u32),
                                   cudaMemcpyDeviceToDevice, c->s_compute));     // this batch's key total, kept per buffer
        KG_CUDA(c, cudaEventRecord(c->ev_scatter, c->s_compute));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_insert, c->ev_scatter, 0));
        insert_keys(c, c->s_insert, c->d_send[b], (u64)nthreads * 32u, c->d_bucket_counts + c->nb + 1 + b);
        KG_CUDA(c, cudaEventRecord(c->ev_send_free[b], c->s_insert));
        c->round++;
    }
    return KG_OK;
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
static int rb_bloom_batch(kg_ctx* c, u32 nthreads) {
    kg_ctx::RbBatch b{nullptr, nullptr, nullptr, nthreads};
    KG_CUDA(c, cudaMalloc(&b.words, sizeof(u64) * nthreads));
    c->rb_log.push_back(b);                                  // owned by the context from here on
    kg_ctx::RbBatch& e = c->rb_log.back();
    KG_CUDA(c, cudaMalloc(&e.brk, sizeof(u32) * nthreads));
    KG_CUDA(c, cudaMalloc(&e.st, sizeof(KgStream)));
    KG_CUDA(c, cudaMemcpyAsync(e.words, c->d_words, sizeof(u64) * nthreads, cudaMemcpyDeviceToDevice, c->s_compute));
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
    cudaEvent_t e0 = next_event(c), e1 = next_event(c), e2 = next_event(c);
    if (e0) cudaEventRecord(e0, s);
    KG_CUDA(c, cudaMemsetAsync(c->d_words, 0, sizeof(u64) * nwords, s));
    KG_CUDA(c, cudaMemsetAsync(c->d_brk, 0, sizeof(u32) * nwords, s));
    kg_carry_restore<<<1, 32, 0, s>>>(c->d_words, c->d_brk, c->d_stream, c->d_carry_words, c->d_carry_brk, c->carry_max_words);
    const bool tma = c->parse_tma;   // opt-in: tiles staged through shared memory by a TMA bulk copy (kg_fetch16<true>)
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
        if (tma) kg_tile_pack<true, true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, c->d_words, c->d_brk);
        else kg_tile_pack<true, false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, c->d_words, c->d_brk);
    } else {
        if (tma) kg_tile_pack<false, true><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, c->d_words, c->d_brk);
        else kg_tile_pack<false, false><<<ntiles, KG_PT, 0, s>>>(d_in, n, c->d_tile_hdr_in, c->d_tile_off, c->d_tile_pend_in, c->d_words, c->d_brk);
    }
    c->launches += 4;
    if (e1) cudaEventRecord(e1, s);
    if (!(flags & KG_FEED_CONTEXT)) {
        const u32 nthreads = (u32)(n / 32 + c->carry_max_words + 2);   // upper bound on packed words
        if (c->ref_bloom) {                                            // bit-exact emulation of the reference's filters
            if (c->pass == KG_PASS_BLOOM) {
                int rc = rb_bloom_batch(c, nthreads);
                if (rc) return rc;
            } else {
                KgCountArgs a;
                a.words = c->d_words; a.brk = c->d_brk; a.st = c->d_stream;
                a.table = c->table; a.bloom = c->bloom;
                a.stats = c->d_stats; a.k = c->cfg.k; a.rank = 0; a.world = 1;
                KG_DISPATCH_W(c->W, launch_rb_count, c, a, nthreads);
            }
        } else if (c->pass_bucketed) {
            int rc = bucket_batch(c, nthreads);
            if (rc) return rc;
        } else {
            KgCountArgs a;
            a.words = c->d_words; a.brk = c->d_brk; a.st = c->d_stream;
            a.table = c->table; a.bloom = c->bloom;
            a.stats = c->d_stats; a.k = c->cfg.k; a.rank = (u32)c->cfg.rank; a.world = (u32)c->cfg.world;
            const int sink = current_sink(c);
            cudaEvent_t ia = next_ins_event(c), ib = next_ins_event(c);
            if (ia) cudaEventRecord(ia, s);
            c->ins_launches++;
            switch (c->W) {
                case 1: launch_count<1>(c, a, nthreads, sink); break;
                case 2: launch_count<2>(c, a, nthreads, sink); break;
                case 3: launch_count<3>(c, a, nthreads, sink); break;
                case 4: launch_count<4>(c, a, nthreads, sink); break;
                case 5: launch_count<5>(c, a, nthreads, sink); break;
                case 6: launch_count<6>(c, a, nthreads, sink); break;
                case 7: launch_count<7>(c, a, nthreads, sink); break;
                case 8: launch_count<8>(c, a, nthreads, sink); break;
            }
            if (ib) cudaEventRecord(ib, s);
        }
    }
    kg_carry_save<<<1, 32, 0, s>>>(c->d_words, c->d_brk, c->d_stream, c->d_carry_words, c->d_carry_brk, c->cfg.k, c->carry_max_words);
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

// Opt-in (KG_FEED_PREFETCH=1, unmeasured): software-pipelined feed of a pinned buffer.  The H2D copy of chunk i+1 is
// issued on the copy stream BEFORE process_batch(i) -- which, with world > 1, blocks the host on the count all-gather
// of chunk i -- and is gated on the device (cudaStreamWaitEvent on the raw buffer's free event) instead of a host wait,
// so the copies leave the compute stream's critical path.
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
        // keep taking part in exchange rounds until every rank has fed its last batch
        bool all_done = false;
        while (!all_done) {
            int rc = c->peer_ready ? peer_round(c, false, nullptr, 0, &all_done) : exchange_round(c, false, &all_done);
            if (rc) return rc;
        }
        KG_CUDA(c, cudaEventRecord(c->ev_tail, c->s_comm));
        KG_CUDA(c, cudaStreamWaitEvent(c->s_compute, c->ev_tail, 0));
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
    if (pass == KG_PASS_COUNT && st.table_full) { c->err = "Hash table is full"; c->pass = 0; return KG_ETABLE_FULL; }
    return KG_OK;
}

template <int W>
static void launch_kaarme_build(kg_ctx* c, const u32* bitmap, const u64* prefix, KgKaarme out, u64* root_counter) {
    const u32 grid = (u32)((c->table.nslots + 255) / 256);
    kg_kaarme_build<W><<<grid, 256, 0, c->s_compute>>>(c->table, c->cfg.k, bitmap, prefix, out, root_counter);
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
    if (c->cfg.world > 1) { c->err = "kg_compact: occurrence positions are not exchanged between shards yet (single GPU only)"; return KG_EBADARG; }
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
