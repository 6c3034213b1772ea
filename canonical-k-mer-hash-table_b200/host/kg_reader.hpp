// kg_reader.hpp -- host reader pipeline of the drop-in CLI.
//
// Replaces text_reader.h:91-226 (read_chunk_from_file) + the io_worker / ts_queue hand-off of
// parallel_parser.hpp:275-338: there, ONE thread read(2)s 10 MiB chunks and spin-waits on a queue while the
// workers parse.  Here the "workers" are a GPU that consumes ~30 GB/s of text, so the host side has to deliver
// bytes at memory speed and never sit between a read and a copy:
//   * one producer thread walks the rank's byte range in file order and fills a ring of caller-provided (pinned)
//     buffers; a large chunk is split over `io_threads` concurrent pread(2)s (page-cache reads are memcpy-bound
//     per thread, NVMe reads are queue-depth-bound: both scale with concurrency);
//   * the consumer (the thread that owns the kg_ctx) takes chunks in order, kg_feed()s them (H2D copy on the copy
//     stream) and hands the buffer back, so reading chunk i+1.. overlaps the copy and the kernels of chunk i.
// Context bytes (the k-1 bases before the rank's range, fed with KG_FEED_CONTEXT) and counted bytes never share a
// chunk -- the same contract as the overlap of text_reader.h:141-151,210.
// No CUDA here: the class only needs memory to read into, so tests/native/reader_host.cpp checks it on a CPU.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <mutex>
#include <string>
#include <sys/types.h>
#include <thread>
#include <unistd.h>
#include <vector>
#ifdef KG_READER_WITH_ZLIB
#include <zlib.h>
#endif

namespace kg {

struct ReadChunk {
    uint8_t* data = nullptr;
    size_t len = 0;
    bool context = false;   // bytes before the rank's range: warm the window only
    int buf = -1;           // ring slot to give back with release()
};

class SliceReader {
  public:
    // [ctx_lo, lo) is context, [lo, hi) is counted.  bufs: nbufs buffers of buf_bytes each (owned by the caller).
    SliceReader(const std::string& path, off_t ctx_lo, off_t lo, off_t hi, uint8_t* const* bufs, int nbufs,
                size_t buf_bytes, int io_threads)
        : path_(path), ctx_lo_(ctx_lo), lo_(lo), hi_(hi), bufs_(bufs, bufs + nbufs), buf_bytes_(buf_bytes),
          io_threads_(std::max(1, io_threads)) {
        for (int i = 0; i < nbufs; i++) free_.push_back(i);
        producer_ = std::thread([this] { produce(); });
    }
    ~SliceReader() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        if (producer_.joinable()) producer_.join();
    }
    SliceReader(const SliceReader&) = delete;
    SliceReader& operator=(const SliceReader&) = delete;

    // next chunk in file order; false at the end of the range or after an IO error (see failed())
    bool next(ReadChunk& out) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !ready_.empty() || done_; });
        if (ready_.empty()) return false;
        out = ready_.front();
        ready_.pop_front();
        return true;
    }
    void release(int buf) {
        {
            std::lock_guard<std::mutex> g(m_);
            free_.push_back(buf);
        }
        cv_.notify_all();
    }
    bool failed() const { return failed_; }
    const std::string& error() const { return error_; }

  private:
    static bool pread_full(int fd, uint8_t* dst, size_t want, off_t pos, size_t& got) {
        got = 0;
        while (got < want) {
            const ssize_t r = pread(fd, dst + got, want - got, pos + (off_t)got);
            if (r < 0) return false;
            if (r == 0) break;   // end of file (the file shrank): deliver what there is
            got += (size_t)r;
        }
        return true;
    }

    // fill dst with file bytes [pos, pos+want) using up to io_threads_ concurrent preads
    bool fill(int fd, uint8_t* dst, size_t want, off_t pos, size_t& got) {
        const size_t min_piece = 4u << 20;
        const int pieces = (int)std::min<size_t>((size_t)io_threads_, std::max<size_t>(1, want / min_piece));
        if (pieces <= 1) return pread_full(fd, dst, want, pos, got);
        std::vector<size_t> g(pieces, 0);
        std::vector<char> ok(pieces, 1);
        std::vector<std::thread> th;
        auto piece_lo = [&](int p) { return (want * (size_t)p / (size_t)pieces) & ~(size_t)4095; };
        for (int p = 0; p < pieces; p++) {
            const size_t a = piece_lo(p), b = p + 1 == pieces ? want : piece_lo(p + 1);
            th.emplace_back([&, p, a, b] { ok[p] = pread_full(fd, dst + a, b - a, pos + (off_t)a, g[p]) ? 1 : 0; });
        }
        for (auto& t : th) t.join();
        got = 0;
        for (int p = 0; p < pieces; p++) {
            if (!ok[p]) return false;
            const size_t a = piece_lo(p), b = p + 1 == pieces ? want : piece_lo(p + 1);
            got += g[p];
            if (g[p] < b - a) break;   // short piece: everything after it is past the end of the file
        }
        return true;
    }

    void produce() {
        const int fd = open(path_.c_str(), O_RDONLY);
        if (fd < 0) { fail("cannot open " + path_); return; }
#ifdef __linux__
        posix_fadvise(fd, ctx_lo_, hi_ - ctx_lo_, POSIX_FADV_SEQUENTIAL);   // parallel_parser.hpp:280
#endif
        off_t pos = ctx_lo_;
        while (pos < hi_) {
            const bool context = pos < lo_;
            const off_t end = context ? lo_ : hi_;
            const size_t want = (size_t)std::min<off_t>((off_t)buf_bytes_, end - pos);
            int b;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return !free_.empty() || stop_; });
                if (stop_) break;
                b = free_.front();
                free_.pop_front();
            }
            size_t got = 0;
            if (!fill(fd, bufs_[b], want, pos, got)) { fail("read error on " + path_); break; }
            if (got == 0) break;
            ReadChunk c;
            c.data = bufs_[b]; c.len = got; c.context = context; c.buf = b;
            {
                std::lock_guard<std::mutex> g(m_);
                ready_.push_back(c);
            }
            cv_.notify_all();
            pos += (off_t)got;
            if (got < want) break;   // short read: end of file
        }
        close(fd);
        {
            std::lock_guard<std::mutex> g(m_);
            done_ = true;
        }
        cv_.notify_all();
    }
    void fail(const std::string& what) {
        std::lock_guard<std::mutex> g(m_);
        failed_ = true;
        error_ = what;
        done_ = true;
        cv_.notify_all();
    }

    std::string path_;
    off_t ctx_lo_, lo_, hi_;
    std::vector<uint8_t*> bufs_;
    size_t buf_bytes_;
    int io_threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<int> free_;
    std::deque<ReadChunk> ready_;
    bool done_ = false, stop_ = false, failed_ = false;
    std::string error_;
    std::thread producer_;
};

// gzip input (a genuine extension: the reference's zlib path, text_reader.h:38-89, seeks backwards in the compressed
// stream by the COMPRESSED size and does not work).  Same ring, same interface; one producer thread inflates the
// stream in order with zlib (concatenated members included), so a .gz input is bound by inflate (~0.3-0.5 GB/s), not
// by the GPU.  A compressed stream cannot be cut into byte ranges: one rank only (the CLI refuses --gpus > 1).
#ifdef KG_READER_WITH_ZLIB
class GzReader {
  public:
    GzReader(const std::string& path, uint8_t* const* bufs, int nbufs, size_t buf_bytes)
        : path_(path), bufs_(bufs, bufs + nbufs), buf_bytes_(buf_bytes) {
        for (int i = 0; i < nbufs; i++) free_.push_back(i);
        producer_ = std::thread([this] { produce(); });
    }
    ~GzReader() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        if (producer_.joinable()) producer_.join();
    }
    GzReader(const GzReader&) = delete;
    GzReader& operator=(const GzReader&) = delete;
    bool next(ReadChunk& out) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !ready_.empty() || done_; });
        if (ready_.empty()) return false;
        out = ready_.front();
        ready_.pop_front();
        return true;
    }
    void release(int buf) {
        {
            std::lock_guard<std::mutex> g(m_);
            free_.push_back(buf);
        }
        cv_.notify_all();
    }
    bool failed() const { return failed_; }
    const std::string& error() const { return error_; }

  private:
    void produce() {
        gzFile f = gzopen(path_.c_str(), "rb");
        if (!f) { finish("cannot open " + path_); return; }
        gzbuffer(f, 1u << 20);
        std::string err;
        for (;;) {
            int b;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return !free_.empty() || stop_; });
                if (stop_) break;
                b = free_.front();
                free_.pop_front();
            }
            size_t got = 0;
            while (got < buf_bytes_) {
                const unsigned want = (unsigned)std::min<size_t>(buf_bytes_ - got, 1u << 30);
                const int r = gzread(f, bufs_[b] + got, want);
                if (r < 0) { int e = 0; const char* msg = gzerror(f, &e); err = std::string("gzip error in ") + path_ + ": " + (msg ? msg : "?"); break; }
                if (r == 0) break;
                got += (size_t)r;
            }
            if (!err.empty()) break;
            if (got == 0) {   // clean end of stream -- unless zlib saw a truncated member
                int e = 0;
                const char* msg = gzerror(f, &e);
                if (e != Z_OK && e != Z_STREAM_END) err = std::string("gzip error in ") + path_ + ": " + (msg ? msg : "?");
                break;
            }
            ReadChunk c;
            c.data = bufs_[b]; c.len = got; c.context = false; c.buf = b;
            {
                std::lock_guard<std::mutex> g(m_);
                ready_.push_back(c);
            }
            cv_.notify_all();
        }
        if (err.empty() && !stop_) {   // gzclose reports a stream that ended inside a member (Z_BUF_ERROR)
            const int rc = gzclose_r(f);
            if (rc != Z_OK) err = "gzip error in " + path_ + ": unexpected end of file";
        } else {
            gzclose_r(f);
        }
        finish(err);
    }
    void finish(const std::string& err) {
        std::lock_guard<std::mutex> g(m_);
        if (!err.empty()) { failed_ = true; error_ = err; }
        done_ = true;
        cv_.notify_all();
    }
    std::string path_;
    std::vector<uint8_t*> bufs_;
    size_t buf_bytes_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<int> free_;
    std::deque<ReadChunk> ready_;
    bool done_ = false, stop_ = false, failed_ = false;
    std::string error_;
    std::thread producer_;
};
#endif

}  // namespace kg
