// kaarme_main.cpp -- the drop-in `kaarme` executable: the reference's command line (main.cpp:134-151,
// README.md:31-52), file-format sniffing (main.cpp:27-68), log lines and output format, driving the sm_100a
// kernels of libkaarme_gpu.so through the C ABI (include/kaarme_gpu.h).  No CPU counting path exists here:
// the host only reads bytes, feeds them, and formats what the GPU exports.
//
//   kaarme [OPTIONS] INPUT KLEN
//     -m,--hash-table-type {0,2}   -a,--min-k-abu N   -t,--threads N   -o,--output-file PATH
//     -b,--use-bfilter  -f,--bfilter-fpr F   exactly one of  -s,--hash-tab-size N | -u,--unq-kmers N
//   GPU-side extras (do not collide with the reference's flags):
//     --device N   --gpus N   --batch-mb N   --exact-counts   --stats-json PATH   --host-format
//     --dump-kaarme PATH   (-m 2: save the compact structure)      --from-kaarme   (INPUT is such a file: decode it)
//
// Exit codes follow the reference: 0 ok; CLI11's 105 (validation), 106 (required), 107 (requires),
// 109 (unexpected argument); 1 for an ill-formed input file (main.cpp:168-171) or a full table
// (kmer_hash_table.cpp:2552-2556; the plain table's silent truncation, parallel_parser.hpp:742-746, is
// deliberately NOT reproduced).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <filesystem>
#include <fstream>
#include <condition_variable>
#include <iostream>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "kaarme_gpu.h"
#define KG_READER_WITH_ZLIB 1
#include "kg_reader.hpp"
#include "kg_writer.hpp"

namespace {

struct Args {
    std::string input, output, stats_json;
    long long k = 0;
    int mode = 2;                 // main.cpp:137
    uint64_t min_abundance = 2;   // main.cpp:138
    int threads = 0;              // main.cpp:139 has no usable default; 0 = pick hardware_concurrency
    bool bloom = false;
    double fpr = 0.01;            // main.cpp:142
    bool has_s = false, has_u = false, has_f = false, has_t = false;
    uint64_t slots = 0, unique = 0;
    int device = 0;
    int gpus = 1;
    uint64_t batch_mb = 0;
    bool exact_counts = false;
    bool print_slices = false;    // --print-slices: show how --gpus N would shard the input, then exit (no GPU needed)
    bool host_format = false;     // --host-format: export (key, count) records and format the lines on host threads
    std::string dump_kaarme;      // --dump-kaarme PATH: write the compact structure (-m 2, one GPU) as a binary file
    bool reference_bloom = false; // --reference-bloom: -b builds the reference's own filter bit for bit (experimental)
    bool from_kaarme = false;     // --from-kaarme: INPUT is a file written by --dump-kaarme; decode it on the GPU
};

[[noreturn]] void cli_fail(int code, const std::string& msg) {
    std::cerr << msg << "\nRun with --help for more information.\n";
    std::exit(code);
}

void print_help(const char* argv0) {
    std::cout << "Space-efficient k-mer counter (B200 GPU build)\n"
                 "Usage: " << argv0 << " [OPTIONS] INPUT KLEN\n\n"
                 "Positionals:\n"
                 "  INPUT TEXT REQUIRED        Input file (automatic format detection)\n"
                 "  KLEN INT REQUIRED          k-mer length\n\n"
                 "Options:\n"
                 "  -h,--help                  Print this help message and exit\n"
                 "  -m,--hash-table-type INT   Hash table type: 0 for plain and 2 for kaarme (def. 2)\n"
                 "  -a,--min-k-abu UINT        Minimum abundance threshold for the output k-mers (def. 2)\n"
                 "  -t,--threads UINT          Number of host threads formatting the output (def. all cores)\n"
                 "  -o,--output-file TEXT      Output file where the k-mer counts will be stored\n"
                 "  -b,--use-bfilter           Use bloom filters to discard unique k-mers\n"
                 "  -f,--bfilter-fpr FLOAT     Bloom filter false positive rate (def. 0.01)\n"
                 "  --device INT               CUDA device ordinal (def. 0; with --gpus N the first of N consecutive devices)\n"
                 "  --gpus INT                 Hash-shard the k-mers over N GPUs of this box (NCCL exchange; def. 1)\n"
                 "  --reference-bloom          With -b on one GPU: reproduce the reference's double Bloom filter bit for bit\n"
                 "                             (its hash functions, counters, table size, false positives; experimental)\n"
                 "  --batch-mb UINT            Raw bytes per device batch in MiB (def. 128)\n"
                 "  --exact-counts             Report true 32-bit counts instead of emulating the reference's\n"
                 "                             16-bit wrap (-m 0) / 14-bit saturation (-m 2)\n"
                 "  --stats-json TEXT          Write pass statistics and device timings as JSON\n"
                 "  --host-format              Format the output lines on host threads instead of on the GPU\n"
                 "  --dump-kaarme TEXT         -m 2, one GPU: also save the compact Kaarme structure (8-byte slots + roots)\n"
                 "  --from-kaarme              INPUT is a file saved with --dump-kaarme: decode it on the GPU and write\n"
                 "                             its k-mers with count >= -a (no -s/-u needed)\n\n"
                 "[Exactly 1 of the following options is required]\n"
                 "Mandatory params:\n"
                 "  -s,--hash-tab-size UINT    Hash table size\n"
                 "  -u,--unq-kmers UINT        Estimated number of unique k-mers\n";
}

bool parse_u64(const std::string& s, uint64_t& v) {
    if (s.empty() || s[0] == '-') return false;
    char* end = nullptr;
    errno = 0;
    unsigned long long x = strtoull(s.c_str(), &end, 10);
    if (errno || *end) return false;
    v = x;
    return true;
}
bool parse_i64(const std::string& s, long long& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    errno = 0;
    long long x = strtoll(s.c_str(), &end, 10);
    if (errno || *end) return false;
    v = x;
    return true;
}
bool parse_f64(const std::string& s, double& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    errno = 0;
    double x = strtod(s.c_str(), &end);
    if (errno || *end) return false;
    v = x;
    return true;
}

Args parse_args(int argc, char** argv) {
    Args a;
    std::vector<std::string> pos;
    auto need_value = [&](int& i, const std::string& name) -> std::string {
        if (i + 1 >= argc) cli_fail(114, name + ": 1 required");
        return argv[++i];
    };
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        std::string val;
        bool has_eq = false;
        if (s.rfind("--", 0) == 0) {
            size_t eq = s.find('=');
            if (eq != std::string::npos) { val = s.substr(eq + 1); s = s.substr(0, eq); has_eq = true; }
        }
        auto value = [&](const std::string& name) { return has_eq ? val : need_value(i, name); };
        if (s == "-h" || s == "--help") { print_help(argv[0]); std::exit(0); }
        else if (s == "-m" || s == "--hash-table-type") {
            long long v;
            std::string t = value("--hash-table-type");
            if (!parse_i64(t, v) || v < 0 || v > 2) cli_fail(105, "--hash-table-type: Value " + t + " not in range 0 to 2");
            a.mode = (int)v;
        } else if (s == "-a" || s == "--min-k-abu") {
            std::string t = value("--min-k-abu");
            if (!parse_u64(t, a.min_abundance)) cli_fail(104, "Could not convert: --min-k-abu = " + t);
        } else if (s == "-t" || s == "--threads") {
            long long v;
            std::string t = value("--threads");
            if (!parse_i64(t, v) || v < 3 || v > 64) cli_fail(105, "--threads: Value " + t + " not in range 3 to 64");
            a.threads = (int)v; a.has_t = true;
        } else if (s == "-o" || s == "--output-file") a.output = value("--output-file");
        else if (s == "-b" || s == "--use-bfilter") a.bloom = true;
        else if (s == "-f" || s == "--bfilter-fpr") {
            std::string t = value("--bfilter-fpr");
            if (!parse_f64(t, a.fpr) || a.fpr < 0.001 || a.fpr > 0.999)
                cli_fail(105, "--bfilter-fpr: Value " + t + " not in range 0.001000 to 0.999000");
            a.has_f = true;
        } else if (s == "-s" || s == "--hash-tab-size") {
            std::string t = value("--hash-tab-size");
            if (!parse_u64(t, a.slots)) cli_fail(104, "Could not convert: --hash-tab-size = " + t);
            a.has_s = true;
        } else if (s == "-u" || s == "--unq-kmers") {
            std::string t = value("--unq-kmers");
            if (!parse_u64(t, a.unique)) cli_fail(104, "Could not convert: --unq-kmers = " + t);
            a.has_u = true;
        } else if (s == "--device") {
            long long v; std::string t = value("--device");
            if (!parse_i64(t, v) || v < 0) cli_fail(105, "--device: Value " + t + " not a device ordinal");
            a.device = (int)v;
        } else if (s == "--gpus") {
            long long v; std::string t = value("--gpus");
            if (!parse_i64(t, v) || v < 1 || v > 64) cli_fail(105, "--gpus: Value " + t + " not in range 1 to 64");
            a.gpus = (int)v;
        } else if (s == "--batch-mb") {
            std::string t = value("--batch-mb");
            if (!parse_u64(t, a.batch_mb) || a.batch_mb == 0 || a.batch_mb > 1024) cli_fail(105, "--batch-mb: Value " + t + " not in range 1 to 1024");
        } else if (s == "--exact-counts") a.exact_counts = true;
        else if (s == "--print-slices") a.print_slices = true;
        else if (s == "--host-format") a.host_format = true;
        else if (s == "--from-kaarme") a.from_kaarme = true;
        else if (s == "--reference-bloom") a.reference_bloom = true;
        else if (s == "--dump-kaarme") a.dump_kaarme = value("--dump-kaarme");
        else if (s == "--stats-json") a.stats_json = value("--stats-json");
        else if (s.size() > 1 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) cli_fail(109, "The following argument was not expected: " + s);
        else pos.push_back(s);
    }
    if (pos.size() > 2) cli_fail(109, "The following argument was not expected: " + pos[2]);
    if (pos.empty()) cli_fail(106, "INPUT is required");
    a.input = pos[0];
    {
        struct stat st{};
        if (stat(a.input.c_str(), &st) != 0) cli_fail(105, "INPUT: File does not exist: " + a.input);
        if (S_ISDIR(st.st_mode)) cli_fail(105, "INPUT: File is actually a directory: " + a.input);
    }
    if (pos.size() < 2) cli_fail(106, "KLEN is required");
    if (!parse_i64(pos[1], a.k) || a.k <= 0) cli_fail(105, "KLEN: Value " + pos[1] + " not in range 0 to inf (positive number required)");
    int given = (a.has_s ? 1 : 0) + (a.has_u ? 1 : 0);
    if (a.from_kaarme && given == 0 && !a.bloom && !a.has_f) return a;   // decoding a saved structure sizes nothing
    if (given == 0) cli_fail(106, "Exactly 1 option from [-s,--hash-tab-size,-u,--unq-kmers] is required");
    if (given == 2) cli_fail(106, "Exactly 1 option from [-s,--hash-tab-size,-u,--unq-kmers] is required and 2 were given");
    if (a.has_u && !a.bloom) cli_fail(107, "--unq-kmers requires --use-bfilter");
    if (a.bloom && !a.has_u) cli_fail(107, "--use-bfilter requires --unq-kmers");
    if (a.has_f && !a.bloom) cli_fail(107, "--bfilter-fpr requires --use-bfilter");
    return a;
}

// main.cpp:19-68 (is_gz, file_format): format by extension + first byte
struct Format { char header_symbol; bool ill_formed; bool gz; };
Format file_format(const std::string& path) {
    Format f{0, false, false};
    unsigned char b[2] = {0, 0};
    {
        std::ifstream ifs(path, std::ios::binary);
        ifs.read((char*)b, 2);
        f.gz = ifs.gcount() == 2 && b[0] == 0x1f && b[1] == 0x8b;
    }
    std::filesystem::path pt(path);
    std::string ext = pt.extension().string();
    char sym = (char)b[0];
    if (f.gz) {   // main.cpp:35-43: the format is that of the compressed content
        while (ext == ".gz") { pt.replace_extension(); ext = pt.extension().string(); }
        sym = 0;
        gzFile g = gzopen(path.c_str(), "rb");
        if (g) { if (gzread(g, &sym, 1) != 1) sym = 0; gzclose(g); }
    }
    if (ext == ".fasta" || ext == ".fa") { f.header_symbol = '>'; f.ill_formed = sym != '>'; }
    else if (ext == ".fastq" || ext == ".fq") { f.header_symbol = '@'; f.ill_formed = sym != '@'; }
    else { f.header_symbol = 0; f.ill_formed = std::string("actgACGT").find(sym) == std::string::npos || sym == 0; }
    return f;
}

#define KG_CHECK(call)                                                                              \
    do {                                                                                            \
        int rc_ = (call);                                                                           \
        if (rc_ == KG_ETABLE_FULL) { std::cout << "Hash table is full... Cannot handle this yet" << std::endl; std::_Exit(1); } \
        if (rc_ != KG_OK) {                                                                         \
            std::cerr << "kaarme: " #call " failed: " << kg_strerror(rc_) << " (" << kg_last_error(ctx) << ")" << std::endl; \
            std::_Exit(2);   /* worker threads may be running: no static destructors */            \
        }                                                                                           \
    } while (0)

// Host-side sharding of one input across ranks (the role of text_reader.h:141-184): a rank that owns bytes
// [lo, hi) must also see the k-1 bases before lo (fed with KG_FEED_CONTEXT, not counted) and must know whether its
// first byte lies inside a FASTA header.
struct Slice { off_t ctx_lo, lo, hi; bool in_header; };
// pread that loops over short reads; any failure is fatal (a slice computed from a partial read would silently drop the
// k-mers that straddle a rank boundary, or misparse a slice that starts inside a header)
static void pread_exact(int fd, char* dst, size_t n, off_t at) {
    size_t got = 0;
    while (got < n) {
        const ssize_t r = pread(fd, dst + got, n - got, at + (off_t)got);
        if (r <= 0) { std::cerr << "kaarme: IO error while reading the input file\n"; std::_Exit(1); }
        got += (size_t)r;
    }
}
Slice make_slice(int fd, off_t file_size, int rank, int world, uint32_t k, bool fasta) {
    if (fd < 0) { std::cerr << "kaarme: cannot open the input file\n"; std::_Exit(1); }
    Slice s;
    s.lo = file_size * rank / world;
    s.hi = file_size * (rank + 1) / world;
    s.in_header = false;
    // walk back over k-1 non-newline bytes
    off_t need = (off_t)k - 1, i = s.lo;
    std::vector<char> buf(1 << 16);
    while (i > 0 && need > 0) {
        off_t n = std::min<off_t>((off_t)buf.size(), i);
        pread_exact(fd, buf.data(), (size_t)n, i - n);
        for (off_t j = n - 1; j >= 0 && need > 0; j--) { i--; if (buf[j] != '\n') need--; }
    }
    s.ctx_lo = i;
    if (fasta) {   // a '>' since the last newline before ctx_lo => the context starts inside a header
        off_t j = s.ctx_lo;
        bool decided = false;
        while (j > 0 && !decided) {
            off_t n = std::min<off_t>((off_t)buf.size(), j);
            pread_exact(fd, buf.data(), (size_t)n, j - n);
            for (off_t q = n - 1; q >= 0; q--) {
                if (buf[q] == '\n') { decided = true; break; }
                if (buf[q] == '>') { s.in_header = true; decided = true; break; }
            }
            j -= n;
        }
    }
    return s;
}

// one pass over this rank's slice: the reader ring (kg_reader.hpp) fills pinned buffers with concurrent pread(2)s
// (or, for a .gz input, inflates the stream in order) while this thread feeds the GPU; a buffer returns to the ring
// as soon as kg_feed has copied it to the device
template <class Reader>
void drain_reader(kg_ctx* ctx, Reader& reader) {
    kg::ReadChunk c;
    while (reader.next(c)) {
        KG_CHECK(kg_feed(ctx, c.data, c.len, c.context ? KG_FEED_CONTEXT : 0));   // returns once the H2D copy is done
        reader.release(c.buf);
    }
    if (reader.failed()) { std::cerr << "kaarme: " << reader.error() << "\n"; std::_Exit(1); }
}
void feed_file(kg_ctx* ctx, const std::string& path, bool gz, uint8_t* const* bufs, int nbufs, size_t buf_bytes, const Slice& sl,
               int io_threads) {
    KG_CHECK(kg_stream_begin(ctx, sl.in_header ? 1 : 0));
    if (gz) {
        kg::GzReader reader(path, bufs, nbufs, buf_bytes);
        drain_reader(ctx, reader);
    } else {
        kg::SliceReader reader(path, sl.ctx_lo, sl.lo, sl.hi, bufs, nbufs, buf_bytes, io_threads);
        drain_reader(ctx, reader);
    }
}

// reusable barrier for the per-GPU host threads
class Barrier {
  public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const int gen = gen_;
        if (++count_ == n_) { gen_++; count_ = 0; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen != gen_; });
    }
  private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

struct Writer {
    int fd = -1;
    kg::ParallelWriter out;       // appends whole buffers with concurrent pwrite(2)s (host/kg_writer.hpp)
    uint32_t k = 0, W = 0;
    int threads = 1;
    uint64_t written = 0;
    std::vector<std::vector<char>> bufs;
};

bool write_all(int fd, const char* p, size_t n) {
    while (n) {
        const ssize_t r = write(fd, p, n);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}

// kmer_hash_table.cpp:2022-2043: k characters, a space, the decimal count, newline  (--host-format path; the default
// path receives these lines ready-made from the GPU, csrc/kg_text.cuh)
size_t format_range(const Writer& w, const uint64_t* keys, const uint32_t* counts, size_t b, size_t e, std::vector<char>& out) {
    const uint32_t k = w.k, W = w.W;
    out.resize((e - b) * (k + 12));
    char* p = out.data();
    for (size_t i = b; i < e; i++) {
        const uint64_t* key = keys + i * W;
        for (uint32_t j = 0; j < k; j++) {
            uint32_t pos = k - 1 - j;
            *p++ = "ACGT"[(key[W - 1 - pos / 32] >> (2 * (pos % 32))) & 3];
        }
        *p++ = ' ';
        char num[12];
        int nd = 0;
        uint32_t v = counts[i];
        do { num[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (nd) *p++ = num[--nd];
        *p++ = '\n';
    }
    return (size_t)(p - out.data());
}

int sink(void* user, const uint64_t* keys, const uint32_t* counts, size_t n) {
    Writer& w = *(Writer*)user;
    const int T = (int)std::min<size_t>((size_t)w.threads, std::max<size_t>(1, n / 65536));
    w.bufs.resize(T);
    std::vector<size_t> lens(T);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) {
        size_t b = n * t / T, e = n * (t + 1) / T;
        th.emplace_back([&, t, b, e]() { lens[t] = format_range(w, keys, counts, b, e, w.bufs[t]); });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < T; t++)
        if (lens[t] && !w.out.append(w.bufs[t].data(), lens[t])) return 1;
    w.written += n;
    return 0;
}

// default: the lines were formatted by kg_format_text on the GPU; the host only writes them
int text_sink(void* user, const char* text, size_t bytes, size_t records) {
    Writer& w = *(Writer*)user;
    if (!w.out.append(text, bytes)) return 1;
    w.written += records;
    return 0;
}

// ---- saved Kaarme structure (--dump-kaarme / --from-kaarme) ----------------------------------------------------
// little-endian: magic "KAARMEG1", u32 version, u32 k, u32 key words W, u32 flags, u64 n_kmers, u64 n_roots,
// u64 reserved[3]  (64 bytes), then n_kmers slot words in the bit layout of kmer.hpp:108-123 (pointers are indices
// into this same array / into the roots), then n_roots * W root words (secondary array, kmer_hash_table.cpp:2144-2145)
struct KaarmeFileHeader {
    char magic[8];
    uint32_t version, k, W, flags;
    uint64_t n_kmers, n_roots, reserved[3];
};
static_assert(sizeof(KaarmeFileHeader) == 64, "header is 64 bytes");
const char KAARME_MAGIC[8] = {'K', 'A', 'A', 'R', 'M', 'E', 'G', '1'};

bool save_kaarme(const std::string& path, uint32_t k, uint32_t W, const std::vector<uint64_t>& slots, const std::vector<uint64_t>& roots,
                 uint64_t n_kmers, uint64_t n_roots) {
    int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    KaarmeFileHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, KAARME_MAGIC, 8);
    h.version = 1; h.k = k; h.W = W; h.n_kmers = n_kmers; h.n_roots = n_roots;
    bool ok = write_all(fd, (const char*)&h, sizeof(h)) && write_all(fd, (const char*)slots.data(), n_kmers * 8) &&
              write_all(fd, (const char*)roots.data(), n_roots * W * 8);
    return close(fd) == 0 && ok;
}

bool load_kaarme(const std::string& path, KaarmeFileHeader& h, std::vector<uint64_t>& slots, std::vector<uint64_t>& roots, std::string& why) {
    std::ifstream f(path, std::ios::binary);
    if (!f.read((char*)&h, sizeof(h)) || memcmp(h.magic, KAARME_MAGIC, 8) != 0 || h.version != 1) { why = "not a Kaarme structure file"; return false; }
    if (h.k < 1 || h.k > KG_MAX_K || h.W != (h.k + 31) / 32 || (h.n_kmers >> 38) || (h.n_roots >> 38)) { why = "bad header"; return false; }
    struct stat st{};
    if (stat(path.c_str(), &st) != 0 || (uint64_t)st.st_size != sizeof(h) + 8 * h.n_kmers + 8 * (uint64_t)h.W * h.n_roots) { why = "file size does not match its header"; return false; }
    slots.resize(h.n_kmers);
    roots.resize(h.n_roots * h.W);
    if ((h.n_kmers && !f.read((char*)slots.data(), (std::streamsize)(8 * h.n_kmers))) ||
        (!roots.empty() && !f.read((char*)roots.data(), (std::streamsize)(8 * roots.size())))) { why = "short read"; return false; }
    return true;
}

void json_pass(std::ostream& o, const char* name, const kg_pass_stats& s) {
    o << "\"" << name << "\": {\"input_kmers\": " << s.input_kmers << ", \"inserted_kmers\": " << s.inserted_kmers
      << ", \"distinct\": " << s.distinct << ", \"table_slots\": " << s.table_slots << ", \"new_in_first\": " << s.new_in_first
      << ", \"new_in_second\": " << s.new_in_second << ", \"bloom_bits\": " << s.bloom_bits << ", \"bloom_hashes\": " << s.bloom_hashes
      << ", \"raw_bytes\": " << s.raw_bytes << ", \"device_ms\": " << s.device_ms << ", \"parse_ms\": " << s.parse_ms
      << ", \"count_ms\": " << s.count_ms << "}";
}

// the product path needs `need` usable sm_100 devices; say so before any file is created
// KAARME_TIMING=1: wall clock of the start-up stages on stderr (where does a small input spend its second?)
static void stage_time(const char* what) {
    static const bool on = getenv("KAARME_TIMING") != nullptr;
    static const auto t0 = std::chrono::steady_clock::now();
    static auto last = t0;
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[timing] %-28s +%8.1f ms  (%8.1f ms)\n", what, std::chrono::duration<double, std::milli>(now - last).count(),
            std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
}

void require_devices(int need) {
    int have = 0;
    const int rc = kg_device_count(&have);
    if (rc != KG_OK || have < need) {
        std::cerr << "kaarme: cannot initialise the GPU path: " << (rc != KG_OK ? kg_strerror(rc) : "not enough compute capability 10.x devices")
                  << " (" << have << " usable, " << need << " needed). This build has no CPU fallback.\n";
        std::exit(2);
    }
}

// --from-kaarme: INPUT is a saved compact structure; the GPU walks the predecessor chains (reconstruct_kmer_in_slot,
// kmer_hash_table.cpp:3848-4058) and formats every k-mer with count >= -a
int decode_kaarme_file(const Args& args) {
    KaarmeFileHeader h;
    std::vector<uint64_t> slots, roots;
    std::string why;
    if (!load_kaarme(args.input, h, slots, roots, why)) { std::cerr << "Input file " << args.input << " is ill-formed (" << why << ")" << std::endl; return 1; }
    if ((long long)h.k != args.k) { std::cerr << "kaarme: " << args.input << " holds " << h.k << "-mers, KLEN says " << args.k << "\n"; return 1; }
    require_devices(args.device + 1);
    std::string output = args.output.empty() ? std::filesystem::path(args.input).replace_extension().filename().string() + ".kaarme_counts" : args.output;
    kg_ctx* ctx = nullptr;
    kg_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = KG_ABI_VERSION; cfg.k = h.k; cfg.table_mode = KG_TABLE_KAARME; cfg.input_mode = KG_INPUT_FASTA;
    cfg.min_slots = 1; cfg.device = args.device; cfg.world = 1; cfg.partitions = 1; cfg.batch_bytes = 1 << 20;
    if (kg_create(&cfg, &ctx) != KG_OK) { std::cerr << "kaarme: cannot initialise the GPU path (" << kg_last_error(nullptr) << "). This build has no CPU fallback.\n"; return 2; }
    KG_CHECK(kg_kaarme_upload(ctx, slots.data(), h.n_kmers, roots.data(), h.n_roots));
    Writer w;
    const unsigned hc = std::thread::hardware_concurrency();
    w.k = h.k; w.W = h.W; w.threads = std::max(1, (args.has_t ? args.threads : (int)std::min(64u, hc ? hc : 3u)) - 2);
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (args.min_abundance > 0) {
        w.fd = open(output.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (w.fd < 0) { std::cerr << "kaarme: cannot open output file " << output << "\n"; return 1; }
        w.out.reset(w.fd, std::max(1, std::min(8, w.threads)));
        // counts in the structure are already 14-bit saturated (kmer.cpp:699-714): nothing left to emulate
        if (args.host_format) KG_CHECK(kg_export(ctx, args.min_abundance, KG_COUNT_EXACT, sink, &w));
        else KG_CHECK(kg_export_text(ctx, args.min_abundance, KG_COUNT_EXACT, text_sink, &w));
        if (close(w.fd) != 0) { std::cerr << "kaarme: error closing " << output << "\n"; return 1; }
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    kg_destroy(ctx);
    std::cout << "Kaarme structure: " << h.n_kmers << " k-mers, " << h.n_roots << " roots, k = " << h.k << "\n";
    std::cout << "Time used to write k-mers in a file: " << std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() << " microseconds\n";
    std::cout << "Written k-mers: " << w.written << "\n";                        // kmer_hash_table.cpp:4522-4523
    std::cout << "Skipped k-mers: " << (h.n_kmers - w.written) << "\n";
    return 0;
}

}  // namespace

// CUDA initialises every visible device at the first runtime call (about a second on an 8-GPU box, most of a small
// run's wall clock).  Unless the user has narrowed the set already, make only the GPUs this run uses visible; they are
// then devices 0 .. gpus-1.
static void narrow_visible_devices(Args& args) {
    // several GPUs driven by threads of this one process: no lazy kernel loading while NCCL kernels may be spinning
    // (the library also preloads what a round launches, kg_comm_init; this covers everything else)
    if (args.gpus > 1) setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    if (getenv("CUDA_VISIBLE_DEVICES")) return;
    std::string v;
    for (int i = 0; i < std::max(1, args.gpus); i++) v += (i ? "," : "") + std::to_string(args.device + i);
    setenv("CUDA_VISIBLE_DEVICES", v.c_str(), 1);
    args.device = 0;
}

int main(int argc, char** argv) {
    Args args = parse_args(argc, argv);
    narrow_visible_devices(args);
    if (args.from_kaarme) return decode_kaarme_file(args);
    Format fmt = file_format(args.input);
    if (fmt.ill_formed) {
        std::cerr << "Input file " << args.input << " is ill-formed" << std::endl;
        return 1;  // main.cpp:168-171
    }
    std::string fmt_name;
    int input_mode;
    if (fmt.header_symbol == '>') { fmt_name = "FASTA"; input_mode = KG_INPUT_FASTA; }
    else if (fmt.header_symbol == '@') { fmt_name = "FASTQ"; input_mode = 1; }
    else { fmt_name = "ONE-STR-PER-LINE"; input_mode = KG_INPUT_PLAIN; }
    if (args.output.empty())
        args.output = std::filesystem::path(args.input).replace_extension().filename().string() + ".kaarme_counts";  // main.cpp:189-191
    if (!args.has_t) {
        unsigned hc = std::thread::hardware_concurrency();
        args.threads = (int)std::max(3u, std::min(64u, hc ? hc : 3u));
    }
    // main.cpp:193-208
    std::cout << "Running settings: " << std::endl;
    std::cout << "  input file:               " << std::filesystem::path(args.input).filename().string() << std::endl;
    std::cout << "  input format:             " << fmt_name << std::endl;
    std::cout << "  gzip compressed:          " << (fmt.gz ? "yes" : "no") << std::endl;
    std::cout << "  k-mer length:             " << args.k << std::endl;
    std::cout << "  min. abundance threshold: " << args.min_abundance << std::endl;
    std::cout << "  hash table type:          " << (args.mode == 0 ? "plain" : "kaarme") << std::endl;
    std::cout << "  using bloom filers:       " << (args.bloom ? "yes" : "no") << std::endl;
    if (args.bloom) {
        std::cout << "    est. unique k-mers:     " << args.unique << std::endl;
        std::cout << "    false positive rate:    " << args.fpr << std::endl;
    } else {
        std::cout << "    est. hash table size:   " << args.slots << std::endl;
    }
    std::cout << "  working threads:          " << args.threads << std::endl;
    std::cout << "  output file:              " << args.output << std::endl;

    if (fmt.gz && args.gpus > 1) { std::cerr << "kaarme: a gzip stream cannot be cut into byte ranges: use --gpus 1 for .gz input\n"; return 1; }
    if (input_mode == 1) { std::cout << "Not implemented yet" << std::endl; return 0; }  // parallel_parser.hpp:797-800
    if (args.mode == 1) { std::cout << "Chosen mode not recognized\n"; return 0; }       // -m 1 (superseded variant) is out of scope
    if (args.k > 256) { std::cerr << "kaarme: k > 256 is not supported by the GPU build\n"; return 1; }

    if (args.print_slices) {   // host-side sharding only: byte range, context start and header state of every rank
        int fd = open(args.input.c_str(), O_RDONLY);
        struct stat st{};
        fstat(fd, &st);
        for (int r = 0; r < args.gpus; r++) {
            const Slice sl = make_slice(fd, st.st_size, r, args.gpus, (uint32_t)args.k, input_mode == KG_INPUT_FASTA);
            std::cout << "slice " << r << " " << sl.ctx_lo << " " << sl.lo << " " << sl.hi << " " << (sl.in_header ? 1 : 0) << "\n";
        }
        close(fd);
        return 0;
    }
    if (args.gpus > 1 && args.mode == KG_TABLE_KAARME)
        std::cout << "note: with --gpus > 1 the Kaarme compaction is skipped (k-mers are exported from the sharded plain tables)\n";

    const int world = args.gpus;
    struct stat fst{};
    stat(args.input.c_str(), &fst);
    // batch = ring buffer = device batch: --batch-mb (def. 128 MiB), but no larger than a rank's share of the file
    // (small inputs should not pay for pinning and allocating hundreds of MiB)
    size_t buf_bytes = (args.batch_mb ? args.batch_mb : 128) << 20;
    {
        const size_t share = (size_t)fst.st_size / (size_t)world + (size_t)args.k + 4096;
        const size_t rounded = std::max<size_t>(1u << 20, (share + (1u << 20) - 1) >> 20 << 20);
        if (!args.batch_mb && !fmt.gz && rounded < buf_bytes) buf_bytes = rounded;   // (a .gz size says nothing about its content)
    }
    char nccl_id[KG_UNIQUE_ID_BYTES];
    if (world > 1) {
        kg_ctx* ctx = nullptr;
        KG_CHECK(kg_comm_unique_id(nccl_id));
    }
    // shared across the per-GPU host threads
    std::vector<kg_pass_stats> bloom_stats(world), count_stats(world);
    std::vector<kg_compact_stats> compact_stats(world);
    std::vector<uint64_t> table_slots(world, 0);
    for (int r = 0; r < world; r++) { memset(&bloom_stats[r], 0, sizeof(kg_pass_stats)); memset(&count_stats[r], 0, sizeof(kg_pass_stats)); memset(&compact_stats[r], 0, sizeof(kg_compact_stats)); }
    stage_time("arguments parsed");
    require_devices(args.device + world);   // before anything is created on disk
    stage_time("device count");
    if (!args.dump_kaarme.empty() && (args.mode != KG_TABLE_KAARME || world != 1)) {
        std::cerr << "kaarme: --dump-kaarme needs -m 2 on one GPU\n";
        return 1;
    }
    Writer w;
    w.k = (uint32_t)args.k; w.W = ((uint32_t)args.k + 31) / 32; w.threads = std::max(1, (args.threads - 2) / world);
    std::mutex out_mutex;
    if (args.min_abundance > 0) {
        w.fd = open(args.output.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (w.fd < 0) { std::cerr << "kaarme: cannot open output file " << args.output << "\n"; return 1; }
        w.out.reset(w.fd, std::max(1, std::min(8, args.threads - 2)));
    }
    const int io_threads = std::max(1, std::min(8, (args.threads - 2) / world));
    const int nbufs = 3;
    Barrier barrier(world);
    std::chrono::high_resolution_clock::time_point t_bloom0, t_bloom1, t_build0, t_build1, t_write1;

    // one host thread per GPU: its own context (= hash shard), its own byte range of the input
    auto rank_main = [&](int rank) {
        kg_ctx* ctx = nullptr;
        kg_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.abi_version = KG_ABI_VERSION;
        cfg.k = (uint32_t)args.k;
        cfg.table_mode = args.mode;
        cfg.input_mode = input_mode;
        cfg.min_slots = args.slots;
        cfg.use_bloom = args.bloom ? 1 : 0;
        cfg.device = args.device + rank;
        cfg.fpr = args.fpr;
        cfg.expected_unique = args.unique;
        cfg.batch_bytes = buf_bytes;
        cfg.rank = rank;
        cfg.world = world;
        cfg.reserved = args.reference_bloom ? KG_CFG_REFERENCE_BLOOM : 0;
        {
            int rc = kg_create(&cfg, &ctx);
            if (rc != KG_OK) {
                std::cerr << "kaarme: cannot initialise the GPU path on device " << cfg.device << ": " << kg_strerror(rc) << " ("
                          << kg_last_error(nullptr) << "). This build has no CPU fallback.\n";
                std::exit(2);
            }
        }
        if (rank == 0) stage_time("kg_create");
        if (world > 1) KG_CHECK(kg_comm_init(ctx, nccl_id, rank, world));
        if (rank == 0) stage_time("kg_comm_init");
        uint8_t* bufs[nbufs] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < nbufs; i++) KG_CHECK(kg_host_alloc(buf_bytes, (void**)&bufs[i]));
        if (rank == 0) stage_time("pinned ring buffers");
        int fd = open(args.input.c_str(), O_RDONLY);
        const Slice sl = fmt.gz ? Slice{0, 0, 0, false} : make_slice(fd, fst.st_size, rank, world, (uint32_t)args.k, input_mode == KG_INPUT_FASTA);
        close(fd);

        if (args.bloom) {
            barrier.wait();
            if (rank == 0) { std::cout << "Starting parallel bloom filtering\n"; t_bloom0 = std::chrono::high_resolution_clock::now(); }  // parallel_parser.hpp:2689
            KG_CHECK(kg_pass_begin(ctx, KG_PASS_BLOOM));
            barrier.wait();   // (see the count pass below)
            feed_file(ctx, args.input, fmt.gz, bufs, nbufs, buf_bytes, sl, io_threads);
            KG_CHECK(kg_pass_end(ctx, &bloom_stats[rank]));
            barrier.wait();
            if (rank == 0) t_bloom1 = std::chrono::high_resolution_clock::now();
        }
        if (args.bloom && rank == 0 && world == 1) {
            std::cout << "New k-mers in first bloom filter " << bloom_stats[0].new_in_first << "\n";   // parallel_parser.hpp:2900-2901
            std::cout << "New k-mers in second bloom filter " << bloom_stats[0].new_in_second << "\n";
            std::cout << "Time used to bloom filter k-mers: " << std::chrono::duration_cast<std::chrono::microseconds>(t_bloom1 - t_bloom0).count() << " microseconds\n";
        }
        barrier.wait();
        if (rank == 0) t_build0 = std::chrono::high_resolution_clock::now();
        KG_CHECK(kg_pass_begin(ctx, KG_PASS_COUNT));
        // All contexts of this process allocate (the table!) before any of them feeds: with peer access enabled between
        // the GPUs a cudaMalloc has to update the peers' mappings, and would wait for ever on a peer whose NCCL kernel
        // is already spinning for this thread's batch.
        barrier.wait();
        if (rank == 0) stage_time("count pass begin (table)");
        kg_table_info(ctx, &table_slots[rank], nullptr, nullptr);
        if (rank == 0) {
            std::cout << (args.mode == 0 ? "Starting atomic flag basic hash table\n" : "Starting atomic variable pointer hash table\n");
            if (world == 1) std::cout << "Hash table size is: " << table_slots[0] << "\n";  // functions_math.cpp:90
        }
        feed_file(ctx, args.input, fmt.gz, bufs, nbufs, buf_bytes, sl, io_threads);
        KG_CHECK(kg_pass_end(ctx, &count_stats[rank]));
        if (rank == 0) stage_time("count pass fed + ended");
        barrier.wait();                            // every shard's table is complete
        if (args.mode == KG_TABLE_KAARME) {        // every shard builds its own self-contained structure
            KG_CHECK(kg_compact(ctx, &compact_stats[rank]));
            if (!args.dump_kaarme.empty()) {
                const kg_compact_stats& cs = compact_stats[rank];
                std::vector<uint64_t> slots(cs.kmers), roots(cs.roots * w.W);
                KG_CHECK(kg_kaarme_download(ctx, slots.data(), roots.data()));
                if (!save_kaarme(args.dump_kaarme, w.k, w.W, slots, roots, cs.kmers, cs.roots)) {
                    std::cerr << "kaarme: cannot write " << args.dump_kaarme << "\n";
                    std::_Exit(1);
                }
            }
        }
        barrier.wait();
        if (rank == 0) t_build1 = std::chrono::high_resolution_clock::now();
        if (args.min_abundance > 0) {
            // shards own disjoint k-mers and the output order is unspecified: the ranks take turns per chunk
            struct Locked { Writer* w; std::mutex* m; } lk{&w, &out_mutex};
            const int count_mode = args.exact_counts ? KG_COUNT_EXACT : KG_COUNT_REFERENCE;
            if (args.host_format) {
                auto locked_sink = [](void* user, const uint64_t* keys, const uint32_t* counts, size_t n) -> int {
                    auto* l = static_cast<Locked*>(user);
                    std::lock_guard<std::mutex> g(*l->m);
                    return sink(l->w, keys, counts, n);
                };
                KG_CHECK(kg_export(ctx, args.min_abundance, count_mode, locked_sink, &lk));
            } else {
                auto locked_text = [](void* user, const char* text, size_t bytes, size_t records) -> int {
                    auto* l = static_cast<Locked*>(user);
                    std::lock_guard<std::mutex> g(*l->m);
                    return text_sink(l->w, text, bytes, records);
                };
                KG_CHECK(kg_export_text(ctx, args.min_abundance, count_mode, locked_text, &lk));
            }
        }
        barrier.wait();
        if (rank == 0) { t_write1 = std::chrono::high_resolution_clock::now(); stage_time("export + write"); }
        for (int i = 0; i < nbufs; i++) kg_host_free(bufs[i]);
        kg_destroy(ctx);
        if (rank == 0) stage_time("destroy");
    };
    if (world == 1) rank_main(0);
    else {
        std::vector<std::thread> th;
        for (int r = 0; r < world; r++) th.emplace_back(rank_main, r);
        for (auto& t : th) t.join();
    }
    if (w.fd >= 0 && close(w.fd) != 0) { std::cerr << "kaarme: error closing " << args.output << "\n"; return 1; }

    // logs, in the reference's order (sums over the shards)
    kg_pass_stats bsum = bloom_stats[0], csum = count_stats[0];
    uint64_t slots_sum = table_slots[0];
    for (int r = 1; r < world; r++) {
        bsum.new_in_first += bloom_stats[r].new_in_first; bsum.new_in_second += bloom_stats[r].new_in_second;
        bsum.device_ms = std::max(bsum.device_ms, bloom_stats[r].device_ms);
        csum.input_kmers += count_stats[r].input_kmers; csum.inserted_kmers += count_stats[r].inserted_kmers;
        csum.distinct += count_stats[r].distinct; csum.table_slots += count_stats[r].table_slots; csum.raw_bytes += count_stats[r].raw_bytes;
        csum.device_ms = std::max(csum.device_ms, count_stats[r].device_ms);
        slots_sum += table_slots[r];
    }
    using us = std::chrono::microseconds;
    if (args.bloom && world > 1) {
        std::cout << "New k-mers in first bloom filter " << bsum.new_in_first << "\n";   // parallel_parser.hpp:2900-2901
        std::cout << "New k-mers in second bloom filter " << bsum.new_in_second << "\n";
        std::cout << "Time used to bloom filter k-mers: " << std::chrono::duration_cast<us>(t_bloom1 - t_bloom0).count() << " microseconds\n";
    }
    if (world > 1) std::cout << "Hash table size is: " << slots_sum << " (" << world << " shards)\n";
    std::cout << "Time used to build hash table: " << std::chrono::duration_cast<us>(t_build1 - t_build0).count() << " microseconds\n";
    std::cout << "Time used to write k-mers in a file: " << std::chrono::duration_cast<us>(t_write1 - t_build1).count() << " microseconds\n";
    kg_compact_stats cs = compact_stats[0];
    for (int r = 1; r < world; r++) {
        cs.kmers += compact_stats[r].kmers; cs.roots += compact_stats[r].roots; cs.bytes += compact_stats[r].bytes;
        cs.reference_bytes += compact_stats[r].reference_bytes;
        cs.max_chain = std::max(cs.max_chain, compact_stats[r].max_chain);
    }
    if (args.mode == KG_TABLE_KAARME) {
        std::cout << "Written k-mers: " << w.written << "\n";                         // kmer_hash_table.cpp:4522-4523
        std::cout << "Skipped k-mers: " << (csum.distinct - w.written) << "\n";
        std::cout << "Main array slots used " << csum.distinct << " / " << csum.table_slots << "\n";  // parallel_parser.hpp:1560-1561
        std::cout << "Max secondary array slots used " << cs.roots << "\n";
        std::cout << "Kaarme bytes: " << cs.bytes << " (" << (cs.kmers ? (double)cs.bytes / cs.kmers : 0.0)
                  << " B/k-mer; reference layout would hold " << cs.reference_bytes << " B)\n";
        if (world > 1)
            for (int r = 0; r < world; r++)
                std::cout << "  shard " << r << ": " << compact_stats[r].kmers << " k-mers, " << compact_stats[r].roots << " roots, "
                          << (compact_stats[r].kmers ? (double)compact_stats[r].bytes / compact_stats[r].kmers : 0.0) << " B/k-mer\n";
    }
    const double dev_ms = csum.device_ms + bsum.device_ms;
    std::cout << "GPU x" << world << ": input k-mers " << csum.input_kmers << ", distinct " << csum.distinct << ", device time " << dev_ms
              << " ms (" << (dev_ms > 0 ? csum.input_kmers / (dev_ms * 1e3) : 0.0) << " M k-mers/s)\n";
    if (!args.stats_json.empty()) {
        std::ofstream o(args.stats_json);
        o << "{\"gpus\": " << world << ", ";
        if (args.bloom) { json_pass(o, "bloom", bsum); o << ", "; }
        json_pass(o, "count", csum);
        o << ", \"written\": " << w.written << ", \"kaarme\": {\"kmers\": " << cs.kmers << ", \"roots\": " << cs.roots
          << ", \"bytes\": " << cs.bytes << ", \"reference_bytes\": " << cs.reference_bytes << ", \"max_chain\": "
          << cs.max_chain << "}}\n";
    }
    return 0;
}
