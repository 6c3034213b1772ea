// CPU check of the CLI's output writer (canonical-k-mer-hash-table_b200/host/kg_writer.hpp).
// usage: writer_host IN OUT|- THREADS CHUNK_BYTES [PREFIX [SUFFIX]]
// Appends IN to OUT in CHUNK_BYTES pieces through kg::ParallelWriter ("-" = stdout, e.g. a pipe: not seekable).
// PREFIX (optional) is written with plain write(2) first: the writer must continue at the current file offset;
// SUFFIX (optional) with plain write(2) afterwards: it must land after everything the writer appended.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <string>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/host/kg_writer.hpp"

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<char> data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const bool to_stdout = strcmp(argv[2], "-") == 0;
    const int fd = to_stdout ? 1 : open(argv[2], O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return 3;
    if (argc > 5 && write(fd, argv[5], strlen(argv[5])) != (ssize_t)strlen(argv[5])) return 4;
    kg::ParallelWriter w(fd, atoi(argv[3]));
    const size_t chunk = (size_t)atoll(argv[4]);
    for (size_t off = 0; off < data.size(); off += chunk)
        if (!w.append(data.data() + off, std::min(chunk, data.size() - off))) return 5;
    if (!w.append(nullptr, 0)) return 6;
    if (w.bytes() != data.size()) return 7;
    if (argc > 6 && write(fd, argv[6], strlen(argv[6])) != (ssize_t)strlen(argv[6])) return 9;   // SUFFIX via plain write(2)
    if (!to_stdout && close(fd) != 0) return 8;
    return 0;
}
