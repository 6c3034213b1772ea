// kg_exchange_plan.hpp -- host arithmetic of the multi-GPU key exchange (no CUDA: tests/native/exchange_plan_host.cpp
// checks it on a CPU against a brute-force layout).
//
// Every round each rank all-gathers its per-bucket key counts.  M[r*row + b] (row = nb + 1, nb = world * pl) = keys
// rank r holds for bucket b; buckets are owner-major: bucket b = d * pl + p is local partition p of owner d.
//
// Peer exchange (kg_peer_connect): a sender's scatter kernel stores its keys STRAIGHT into the owner's receive buffer
// over NVLink, so every (sender, bucket) run needs its place there before the scatter starts.  The owner's buffer is
// laid out partition-major, senders in rank order inside a partition:
//     [p=0: from rank 0 | from rank 1 | ...][p=1: from rank 0 | ...] ...
// which is exactly the order the L2-blocked insert wants: it walks the buffer front to back and
// the live table region moves with it.  No segment table, no send buffer, no copy kernel.
#pragma once
#include <cstdint>
#include <vector>

struct KgPeerPlan {
    std::vector<uint64_t> remote_base;   // [nb] first key index, in the OWNER's receive buffer, of my run for bucket b
    std::vector<uint64_t> in_keys;       // [world] keys every owner receives this round (all senders, itself included)
    uint64_t my_in = 0;                  // in_keys[rank]
    uint64_t max_in = 0;                 // max over owners: must fit a receive buffer, else the round falls back
    bool all_done = true;                // every rank raised its done flag (M[r*row + nb] != 0)
};

inline KgPeerPlan kg_peer_plan(const uint32_t* M, uint32_t world, uint32_t pl, uint32_t rank) {
    const uint32_t nb = world * pl;
    const size_t row = (size_t)nb + 1;
    KgPeerPlan plan;
    plan.remote_base.assign(nb, 0);
    plan.in_keys.assign(world, 0);
    for (uint32_t d = 0; d < world; d++) {
        uint64_t at = 0;                                   // running offset in owner d's receive buffer
        for (uint32_t p = 0; p < pl; p++) {
            const uint32_t b = d * pl + p;
            for (uint32_t s = 0; s < world; s++) {
                if (s == rank) plan.remote_base[b] = at;
                at += M[(size_t)s * row + b];
            }
        }
        plan.in_keys[d] = at;
        if (at > plan.max_in) plan.max_in = at;
    }
    plan.my_in = plan.in_keys[rank];
    for (uint32_t r = 0; r < world; r++) plan.all_done = plan.all_done && M[(size_t)r * row + nb] != 0;
    return plan;
}
