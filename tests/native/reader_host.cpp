// CPU check of the CLI's reader pipeline (canonical-k-mer-hash-table_b200/host/kg_reader.hpp): drives a SliceReader
// over plain malloc'ed buffers and writes what a kg_ctx would have been fed.
// usage: reader_host PATH CTX_LO LO HI BUF_BYTES NBUFS IO_THREADS
// stdout: u64 context_bytes, u64 counted_bytes, u64 chunks, then the context bytes followed by the counted bytes
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/host/kg_reader.hpp"

int main(int argc, char** argv) {
    if (argc != 8) return 2;
    const std::string path = argv[1];
    const off_t ctx_lo = atoll(argv[2]), lo = atoll(argv[3]), hi = atoll(argv[4]);
    const size_t buf_bytes = (size_t)atoll(argv[5]);
    const int nbufs = atoi(argv[6]), io_threads = atoi(argv[7]);
    std::vector<std::vector<uint8_t>> store(nbufs, std::vector<uint8_t>(buf_bytes));
    std::vector<uint8_t*> bufs;
    for (auto& s : store) bufs.push_back(s.data());
    std::vector<uint8_t> ctx, body;
    uint64_t chunks = 0;
    bool seen_body = false;
    {
        kg::SliceReader rd(path, ctx_lo, lo, hi, bufs.data(), nbufs, buf_bytes, io_threads);
        kg::ReadChunk c;
        while (rd.next(c)) {
            if (c.len == 0 || c.len > buf_bytes) return 3;
            if (c.context && seen_body) return 4;            // context always precedes the counted bytes
            if (!c.context) seen_body = true;
            auto& dst = c.context ? ctx : body;
            dst.insert(dst.end(), c.data, c.data + c.len);
            chunks++;
            rd.release(c.buf);
        }
        if (rd.failed()) { fprintf(stderr, "%s\n", rd.error().c_str()); return 5; }
    }
    const uint64_t hdr[3] = {ctx.size(), body.size(), chunks};
    fwrite(hdr, 8, 3, stdout);
    fwrite(ctx.data(), 1, ctx.size(), stdout);
    fwrite(body.data(), 1, body.size(), stdout);
    return 0;
}
