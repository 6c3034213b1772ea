// mock_abi.cpp -- TEST DOUBLE of libkaarme_gpu.so for CPU tests of the CLI's HOST code (tests/test_cli_mock_cpu.py).
//
// The drop-in executable (canonical-k-mer-hash-table_b200/host/kaarme_main.cpp) is host logic around the C ABI of
// include/kaarme_gpu.h: argument handling, input sharding (make_slice), the reader ring, context feeds, the sinks, the
// parallel writer, the Kaarme file format.  None of that needs a GPU to be WRONG, so this file implements the same ABI
// on the CPU -- with the oracle (oracle/liboracle.so) doing the counting -- and the test suite links a second copy of
// the CLI against it (tests/native/_build/kaarme_mock).  It is test infrastructure only, lives under tests/, is never
// shipped or loaded by the product, and says nothing about the kernels (the -m gpu tests do that).
//
// Model: a context buffers what it is fed (context bytes and counted bytes apart).  At kg_pass_end the k-mers of the
// counted bytes are  count(context + counted) - count(context)  (a window ends either in the context or after it),
// and go into ONE process-wide map shared by all ranks; rank r exports the keys whose hash it owns.  The Bloom pass
// admits everything (so outputs equal the reference's at -a >= 2 only, like any Bloom run).  kg_compact stores every
// k-mer as a root (a legal, if pointless, Kaarme structure).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/kaarme_gpu.h"
#include "../../oracle/kaarme_oracle.h"
#include "../../canonical-k-mer-hash-table_b200/csrc/kg_text.cuh"

namespace {

typedef std::vector<uint64_t> Key;
struct Shared {
    std::mutex m;
    std::map<Key, uint64_t> counts;   // canonical key -> true multiplicity, all ranks
    int users = 0;
};
Shared g_shared;
std::mutex g_oracle;   // one oracle call at a time (keeps the test double's memory use bounded with many ranks)
thread_local std::string g_err;

uint64_t next_prime3mod4(uint64_t c) { return ko_next_prime3mod4(c); }

uint64_t key_hash(const Key& k) {
    uint64_t h = 1469598103934665603ULL;
    for (uint64_t w : k) { h ^= w; h *= 1099511628211ULL; h ^= h >> 29; }
    return h;
}

}  // namespace

struct kg_ctx {
    kg_config cfg;
    uint32_t W = 0;
    int pass = 0;
    bool stream_open = false, in_header = false, bloom_done = false, counted = false, compacted = false;
    std::vector<uint8_t> ctx_bytes, body;
    uint64_t table_slots = 0, distinct_local = 0;
    std::vector<uint64_t> kslots, kroots;   // Kaarme structure
    uint64_t launches = 0;
    std::string err;
};

extern "C" {

int kg_abi_version(void) { return KG_ABI_VERSION; }
const char* kg_strerror(int s) {
    static const char* names[] = {"ok", "bad argument or call order", "CUDA error / no usable sm_100 device", "hash table is full",
                                  "NCCL error", "out of device or pinned memory", "export sink aborted"};
    return s >= 0 && s <= 6 ? names[s] : "unknown status";
}
const char* kg_last_error(const kg_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
int kg_device_count(int* n) {
    if (!n) return KG_EBADARG;
    const char* e = getenv("KG_MOCK_DEVICES");
    *n = e ? atoi(e) : 8;
    return KG_OK;
}
int kg_create(const kg_config* cfg, kg_ctx** out) {
    if (!cfg || !out) return KG_EBADARG;
    *out = nullptr;
    if (cfg->abi_version != KG_ABI_VERSION || cfg->k < 1 || cfg->k > KG_MAX_K) { g_err = "bad config"; return KG_EBADARG; }
    if (cfg->use_bloom ? cfg->expected_unique == 0 : cfg->min_slots == 0) { g_err = "bad sizes"; return KG_EBADARG; }
    kg_ctx* c = new kg_ctx();
    c->cfg = *cfg;
    c->W = (cfg->k + 31) / 32;
    if (getenv("KG_MOCK_SHOW_FLAGS")) fprintf(stderr, "mock: config flags %u\n", cfg->reserved);
    {
        std::lock_guard<std::mutex> g(g_shared.m);
        if (g_shared.users++ == 0) g_shared.counts.clear();
    }
    *out = c;
    return KG_OK;
}
int kg_destroy(kg_ctx* c) {
    if (!c) return KG_OK;
    {
        std::lock_guard<std::mutex> g(g_shared.m);
        g_shared.users--;
    }
    delete c;
    return KG_OK;
}
int kg_host_alloc(size_t bytes, void** out) { *out = malloc(bytes ? bytes : 1); return *out ? KG_OK : KG_ENOMEM; }
int kg_host_free(void* p) { free(p); return KG_OK; }
int kg_comm_unique_id(void* id) { memset(id, 7, KG_UNIQUE_ID_BYTES); return KG_OK; }
int kg_comm_init(kg_ctx* c, const void* id, int rank, int world) {
    if (!c || !id || rank != c->cfg.rank || world != c->cfg.world) return KG_EBADARG;
    return KG_OK;
}


int kg_pass_begin(kg_ctx* c, int pass) {
    if (!c || (pass != KG_PASS_BLOOM && pass != KG_PASS_COUNT)) return KG_EBADARG;
    if (pass == KG_PASS_BLOOM && !c->cfg.use_bloom) return KG_EBADARG;
    if (pass == KG_PASS_COUNT && c->cfg.use_bloom && !c->bloom_done) return KG_EBADARG;
    c->pass = pass;
    c->stream_open = false;
    c->ctx_bytes.clear();
    c->body.clear();
    if (pass == KG_PASS_COUNT) {
        const uint64_t world = (uint64_t)c->cfg.world;
        const uint64_t want = c->cfg.use_bloom ? 2 * c->distinct_local : (c->cfg.min_slots + world - 1) / world;
        c->table_slots = next_prime3mod4(want);
        c->counted = c->compacted = false;
    }
    return KG_OK;
}
int kg_stream_begin(kg_ctx* c, int starts_in_header) {
    if (!c || !c->pass || c->stream_open) return KG_EBADARG;    // the mock models one stream per pass
    c->stream_open = true;
    c->in_header = starts_in_header != 0;
    return KG_OK;
}
int kg_feed(kg_ctx* c, const uint8_t* bytes, size_t n, uint32_t flags) {
    if (!c || !c->pass || !c->stream_open) return KG_EBADARG;
    if (n == 0) return KG_OK;
    if (!bytes) return KG_EBADARG;
    if (flags & KG_FEED_CONTEXT) {
        if (!c->body.empty()) { c->err = "mock: context after counted bytes"; return KG_EBADARG; }
        c->ctx_bytes.insert(c->ctx_bytes.end(), bytes, bytes + n);
    } else {
        c->body.insert(c->body.end(), bytes, bytes + n);
    }
    c->launches++;
    return KG_OK;
}
int kg_feed_device(kg_ctx*, const void*, size_t, uint32_t) { return KG_EBADARG; }

int kg_pass_end(kg_ctx* c, kg_pass_stats* st) {
    if (!c || !c->pass) return KG_EBADARG;
    std::vector<uint8_t> all(c->ctx_bytes);
    all.insert(all.end(), c->body.begin(), c->body.end());
    ko_counts a, b;
    memset(&a, 0, sizeof(a));
    memset(&b, 0, sizeof(b));
    {
        std::lock_guard<std::mutex> g(g_oracle);
        if (ko_count(all.data(), all.size(), c->cfg.k, c->cfg.input_mode, c->in_header, &a) != 0) return KG_ENOMEM;
        if (ko_count(c->ctx_bytes.data(), c->ctx_bytes.size(), c->cfg.k, c->cfg.input_mode, c->in_header, &b) != 0) return KG_ENOMEM;
    }
    std::map<Key, uint64_t> local;
    for (uint64_t i = 0; i < a.n; i++) local[Key(a.keys + i * a.W, a.keys + (i + 1) * a.W)] += a.counts[i];
    for (uint64_t i = 0; i < b.n; i++) {
        auto it = local.find(Key(b.keys + i * b.W, b.keys + (i + 1) * b.W));
        if (it == local.end() || it->second < b.counts[i]) { c->err = "mock: context k-mers are not a subset"; return KG_ECUDA; }
        it->second -= b.counts[i];
        if (it->second == 0) local.erase(it);
    }
    const uint64_t windows = a.total_windows - b.total_windows;
    ko_counts_free(&a);
    ko_counts_free(&b);
    if (st) {
        memset(st, 0, sizeof(*st));
        st->input_kmers = windows;
        st->raw_bytes = all.size();
        st->partitions = 1;
    }
    const int pass = c->pass;
    c->pass = 0;
    c->stream_open = false;
    if (pass == KG_PASS_BLOOM) {
        c->bloom_done = true;
        c->distinct_local = local.size();
        if (st) { st->new_in_first = local.size(); st->new_in_second = local.size(); st->bloom_bits = 64; st->bloom_hashes = 1; }
        return KG_OK;
    }
    {
        std::lock_guard<std::mutex> g(g_shared.m);
        for (auto& kv : local) g_shared.counts[kv.first] += kv.second;
    }
    c->counted = true;
    c->distinct_local = local.size();
    if (st) {
        st->distinct = local.size();
        st->inserted_kmers = windows;
        st->table_slots = c->table_slots;
    }
    if (local.size() > c->table_slots) { c->err = "Hash table is full"; return KG_ETABLE_FULL; }
    return KG_OK;
}

// keys owned by this rank, with their true counts (call after every rank has finished its count pass)
static std::vector<std::pair<Key, uint64_t>> owned(kg_ctx* c) {
    std::vector<std::pair<Key, uint64_t>> v;
    std::lock_guard<std::mutex> g(g_shared.m);
    for (auto& kv : g_shared.counts)
        if (key_hash(kv.first) % (uint64_t)c->cfg.world == (uint64_t)c->cfg.rank) v.push_back(kv);
    return v;
}

int kg_compact(kg_ctx* c, kg_compact_stats* st) {
    if (!c || c->cfg.table_mode != KG_TABLE_KAARME || !c->counted) return KG_EBADARG;
    if (!c->compacted) {
        const auto v = owned(c);
        const uint32_t k = c->cfg.k;
        c->kslots.clear();
        c->kroots.clear();
        for (size_t i = 0; i < v.size(); i++) {
            const Key& key = v[i].first;
            const uint64_t cnt = std::min<uint64_t>(v[i].second, 16383);
            const uint32_t top_chars = k - 32 * (c->W - 1);
            const uint64_t left = (key[0] >> (2 * (top_chars - 1))) & 3, right = key[c->W - 1] & 3;
            c->kslots.push_back(((uint64_t)i << 26) | (cnt << 12) | (left << 10) | (right << 8) | (1u << 5) | (1u << 4) | 1u);
            c->kroots.insert(c->kroots.end(), key.begin(), key.end());
        }
        c->compacted = true;
    }
    if (st) {
        memset(st, 0, sizeof(*st));
        st->kmers = c->kslots.size();
        st->roots = c->kslots.size();
        st->bytes = 8 * st->kmers + 8 * (uint64_t)c->W * st->roots;
        st->reference_bytes = 8 * c->table_slots + (8 * (uint64_t)c->W + 1) * st->roots;
    }
    return KG_OK;
}

int kg_kaarme_download(kg_ctx* c, uint64_t* slots, uint64_t* roots) {
    if (!c || !c->compacted) return KG_EBADARG;
    if (slots) memcpy(slots, c->kslots.data(), 8 * c->kslots.size());
    if (roots) memcpy(roots, c->kroots.data(), 8 * c->kroots.size());
    return KG_OK;
}
int kg_kaarme_upload(kg_ctx* c, const uint64_t* slots, uint64_t n_kmers, const uint64_t* roots, uint64_t n_roots) {
    if (!c || c->cfg.table_mode != KG_TABLE_KAARME || c->cfg.world != 1 || c->pass) return KG_EBADARG;
    c->kslots.assign(slots, slots + n_kmers);
    c->kroots.assign(roots, roots + n_roots * c->W);
    c->compacted = c->counted = true;
    return KG_OK;
}

// records selected by (min_abundance, count_mode), in key order
static int select_records(kg_ctx* c, uint64_t min_ab, int count_mode, std::vector<uint64_t>& keys, std::vector<uint32_t>& counts) {
    const uint32_t k = c->cfg.k, W = c->W;
    if (c->compacted) {
        std::vector<uint8_t> codes(k);
        const uint64_t n_roots = c->kroots.size() / W;
        // bound every root index a walk could reach before handing the structure to the decoder
        for (uint64_t s : c->kslots) if (!((s >> 1) & 1) && (s >> 26) >= n_roots) { c->err = "malformed structure"; return KG_ECUDA; }
        for (uint64_t i = 0; i < c->kslots.size(); i++) {
            const uint64_t d = c->kslots[i];
            const uint32_t cnt = (uint32_t)(d >> 12) & 16383u;
            if (cnt < min_ab) continue;
            if (ko_kaarme_decode(c->kslots.data(), c->kslots.size(), c->kroots.data(), k, i, codes.data()) < 0) { c->err = "malformed chain"; return KG_ECUDA; }
            Key key(W, 0);
            for (uint32_t j = 0; j < k; j++) {
                const uint32_t pos = k - 1 - j;
                key[W - 1 - pos / 32] |= (uint64_t)codes[j] << (2 * (pos % 32));
            }
            keys.insert(keys.end(), key.begin(), key.end());
            counts.push_back(cnt);
        }
        return KG_OK;
    }
    if (!c->counted) return KG_EBADARG;
    for (auto& kv : owned(c)) {
        const uint64_t rep = count_mode == KG_COUNT_EXACT ? kv.second : ko_reported_count(kv.second, c->cfg.table_mode);
        if (rep == 0 || rep < min_ab) continue;
        keys.insert(keys.end(), kv.first.begin(), kv.first.end());
        counts.push_back((uint32_t)rep);
    }
    return KG_OK;
}

#define MOCK_SINK_RECORDS 1000   // several sink calls even for small inputs

int kg_export(kg_ctx* c, uint64_t min_ab, int count_mode, kg_sink_fn sink, void* user) {
    if (!c || !sink) return KG_EBADARG;
    if (min_ab == 0) return KG_OK;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> counts;
    int rc = select_records(c, min_ab, count_mode, keys, counts);
    if (rc) return rc;
    for (size_t i = 0; i < counts.size(); i += MOCK_SINK_RECORDS) {
        const size_t n = std::min<size_t>(MOCK_SINK_RECORDS, counts.size() - i);
        if (sink(user, keys.data() + i * c->W, counts.data() + i, n) != 0) return KG_ESINK;
    }
    return KG_OK;
}
int kg_export_text(kg_ctx* c, uint64_t min_ab, int count_mode, kg_text_sink_fn sink, void* user) {
    if (!c || !sink) return KG_EBADARG;
    if (min_ab == 0) return KG_OK;
    std::vector<uint64_t> keys;
    std::vector<uint32_t> counts;
    int rc = select_records(c, min_ab, count_mode, keys, counts);
    if (rc) return rc;
    std::vector<char> text;
    for (size_t i = 0; i < counts.size(); i += MOCK_SINK_RECORDS) {
        const size_t n = std::min<size_t>(MOCK_SINK_RECORDS, counts.size() - i);
        text.resize(n * kg_line_bound(c->cfg.k));
        size_t bytes = 0;
        for (size_t j = 0; j < n; j++)   // the product's own per-record formatter (host compilation of csrc/kg_text.cuh)
            bytes += kg_format_line((const unsigned long long*)keys.data() + (i + j) * c->W, c->W, c->cfg.k, counts[i + j], text.data() + bytes);
        if (sink(user, text.data(), bytes, n) != 0) return KG_ESINK;
    }
    return KG_OK;
}

int kg_table_info(const kg_ctx* c, uint64_t* slots, uint32_t* slot_bytes, uint32_t* key_words) {
    if (!c) return KG_EBADARG;
    if (slots) *slots = c->table_slots;
    if (slot_bytes) *slot_bytes = 16;
    if (key_words) *key_words = c->W;
    return KG_OK;
}
int kg_atomic_ceiling(int, uint64_t, uint64_t, int, double*) { return KG_ECUDA; }
int kg_checksum(kg_ctx* c, uint64_t min_ab, int, uint64_t out[4]) {      // the CLI does not call it; present for the symbol set
    if (!c || !out) return KG_EBADARG;
    out[0] = out[1] = out[2] = out[3] = 0;
    for (auto& kv : owned(c)) if (min_ab && kv.second >= min_ab) { out[0]++; out[1] += kv.second; }
    return KG_OK;
}
int kg_launch_count(const kg_ctx* c, uint64_t* n) { if (!c || !n) return KG_EBADARG; *n = c->launches; return KG_OK; }

}  // extern "C"
