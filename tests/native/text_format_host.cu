// Host-side check of the per-record formatter that kg_format_text runs on the GPU (csrc/kg_text.cuh):
// kg_format_line / kg_ndigits are __host__ __device__, so the very same code is exercised here on a CPU.
// stdin : u32 k, u32 n, n*W u64 key words, n u32 counts (W = ceil(k/32))     stdout : the n formatted lines
// Built by tests/test_text_format_cpu.py with `nvcc` (host compilation only; no CUDA call is made).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/csrc/kg_text.cuh"

int main() {
    uint32_t hdr[2];
    if (fread(hdr, 4, 2, stdin) != 2) return 2;
    const uint32_t k = hdr[0], n = hdr[1], W = (k + 31) / 32;
    std::vector<unsigned long long> keys((size_t)n * W);
    std::vector<uint32_t> counts(n);
    if (n && (fread(keys.data(), 8, keys.size(), stdin) != keys.size() || fread(counts.data(), 4, n, stdin) != n)) return 2;
    std::vector<char> line(kg_line_bound(k));
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t len = kg_format_line(keys.data() + (size_t)i * W, W, k, counts[i], line.data());
        if (len > kg_line_bound(k) || len != k + 2 + kg_ndigits(counts[i])) return 3;
        fwrite(line.data(), 1, len, stdout);
    }
    return 0;
}
