// kg_writer.hpp -- output side of the drop-in CLI (host half).
//
// The reference writes its result with one thread and `ofstream <<` per character (kmer_hash_table.cpp:2013-2050,
// :4318-4524).  Here the lines arrive from the GPU as finished 64 MiB text buffers (csrc/kg_text.cuh), so what is
// left on the host is moving bytes into the page cache -- a memcpy that one thread does at a few GB/s.  Output
// order is unspecified, but a buffer is appended as a whole, so the file is simply the buffers back to back:
// ParallelWriter reserves [end, end+n) and lets `threads` workers pwrite(2) disjoint pieces of it concurrently.
// Descriptors that cannot seek (pipes, /dev/stdout) fall back to sequential write(2).
// No CUDA here: tests/native/writer_host.cpp checks the class on a CPU.
#pragma once
#include <algorithm>
#include <cerrno>
#include <cstddef>
#include <cstdint>
#include <sys/types.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace kg {

class ParallelWriter {
  public:
    ParallelWriter() = default;
    ParallelWriter(int fd, int threads) { reset(fd, threads); }
    void reset(int fd, int threads) {
        fd_ = fd;
        threads_ = std::max(1, threads);
        const off_t cur = fd >= 0 ? lseek(fd, 0, SEEK_CUR) : (off_t)-1;
        seekable_ = cur != (off_t)-1;
        end_ = seekable_ ? (uint64_t)cur : 0;
        written_ = 0;
    }
    int fd() const { return fd_; }
    uint64_t bytes() const { return written_; }

    // append n bytes; returns when they have all been handed to the kernel (the caller may reuse p)
    bool append(const char* p, size_t n) {
        if (n == 0) return true;
        if (fd_ < 0) return false;
        bool ok;
        if (!seekable_) {
            ok = write_seq(p, n);
        } else {
            const uint64_t at = end_;
            end_ += n;
            const size_t min_piece = 4u << 20;
            const int pieces = (int)std::min<size_t>((size_t)threads_, std::max<size_t>(1, n / min_piece));
            if (pieces <= 1) {
                ok = pwrite_all(p, n, (off_t)at);
            } else {
                std::vector<char> res(pieces, 1);
                std::vector<std::thread> th;
                auto cut = [&](int i) { return i == pieces ? n : (n * (size_t)i / (size_t)pieces) & ~(size_t)4095; };
                for (int i = 0; i < pieces; i++) {
                    const size_t a = cut(i), b = cut(i + 1);
                    th.emplace_back([&, i, a, b] { res[i] = pwrite_all(p + a, b - a, (off_t)(at + a)) ? 1 : 0; });
                }
                for (auto& t : th) t.join();
                ok = std::all_of(res.begin(), res.end(), [](char c) { return c != 0; });
            }
        }
        if (ok) written_ += n;
        if (ok && seekable_) lseek(fd_, (off_t)end_, SEEK_SET);   // keep the descriptor's own offset at the end: a plain
                                                                   // write(2) by anyone sharing it lands after the data
        return ok;
    }

  private:
    bool pwrite_all(const char* p, size_t n, off_t at) const {
        while (n) {
            const ssize_t r = pwrite(fd_, p, n, at);
            if (r < 0) { if (errno == EINTR) continue; return false; }
            p += r; n -= (size_t)r; at += r;
        }
        return true;
    }
    bool write_seq(const char* p, size_t n) const {
        while (n) {
            const ssize_t r = write(fd_, p, n);
            if (r < 0) { if (errno == EINTR) continue; return false; }
            p += r; n -= (size_t)r;
        }
        return true;
    }
    int fd_ = -1;
    int threads_ = 1;
    bool seekable_ = false;
    uint64_t end_ = 0, written_ = 0;
};

}  // namespace kg
