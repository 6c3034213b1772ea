// CPU check of the peer-exchange layout arithmetic (canonical-k-mer-hash-table_b200/csrc/kg_exchange_plan.hpp).
// stdin: u32 world, u32 pl, then the gathered count matrix M[world][world*pl + 1] (u32, last column = done flag)
// stdout (text): one line per rank: "rank my_in max_in all_done | in_keys[world] | remote_base[nb]"
#include <cstdio>
#include <vector>
#include "../../canonical-k-mer-hash-table_b200/csrc/kg_exchange_plan.hpp"

int main() {
    uint32_t hdr[2];
    if (fread(hdr, 4, 2, stdin) != 2) return 2;
    const uint32_t world = hdr[0], pl = hdr[1], nb = world * pl;
    std::vector<uint32_t> M((size_t)world * (nb + 1));
    if (fread(M.data(), 4, M.size(), stdin) != M.size()) return 2;
    for (uint32_t r = 0; r < world; r++) {
        const KgPeerPlan p = kg_peer_plan(M.data(), world, pl, r);
        printf("%u %llu %llu %d |", r, (unsigned long long)p.my_in, (unsigned long long)p.max_in, p.all_done ? 1 : 0);
        for (uint32_t d = 0; d < world; d++) printf(" %llu", (unsigned long long)p.in_keys[d]);
        printf(" |");
        for (uint32_t b = 0; b < nb; b++) printf(" %llu", (unsigned long long)p.remote_base[b]);
        printf("\n");
    }
    return 0;
}
