// bsi_target_main executed on the host: the kernel text of bsi_template.cuh under cuda_shim.h with
// the round logic of bsi_target_run in csrc/bsi.cpp (horizon rounds with a cycle-closure test when
// max_t lies beyond the horizon).  Test infrastructure (tests/test_kernel_emul.py).
//
// usage: target_emul idx_begin idx_end max_t        (max_t 0 = no cap)
// prints "hit <index> <t>" per hit in index order, then "totals <hits> <closed without a hit> <steps>".
#include "cuda_shim.h"

#define BSI_HOST_EMUL 1
#include BSI_HEADER
#undef BSI_UNIT
#define BSI_UNIT 0
#include BSI_TEMPLATE

#include <algorithm>
#include <functional>
#include <utility>
#include <vector>

unsigned char bsi_smem[16];

static void launch(unsigned grid, const std::function<void()> &kernel) {
    gridDim.x = grid;
    for (unsigned blk = 0; blk < grid; ++blk) {
        blockIdx.x = blk;
        bsi_emul_block(32, kernel);
    }
}

int main(int argc, char **argv) {
    static_assert(BSI_BLOCK == 32, "generate the header with block_threads = 32");
    if (argc < 4) return 2;
    const u64 INF = ~0ull, kMaxH = 1ull << 30;
    const u64 b = strtoull(argv[1], nullptr, 10), e = strtoull(argv[2], nullptr, 10);
    u64 max_t = strtoull(argv[3], nullptr, 10);
    if (max_t == 0) max_t = INF;
    const u64 n_lanes = e - b;
    BsiCounters *ctr = (BsiCounters *)calloc(1, sizeof(BsiCounters));
    std::vector<u64> hit_idx(n_lanes + 1), hit_t(n_lanes + 1), closed(n_lanes + 1);
    std::vector<u64> retry[2] = {std::vector<u64>(n_lanes + 1), std::vector<u64>(n_lanes + 1)};
    std::vector<std::pair<u64, u64>> hits;
    u64 T = max_t == INF ? std::max<u64>(1024, 2 * (u64)BSI_LP_MAX) : max_t;
    u64 n_in = n_lanes, n_without = 0;
    int which = 0;
    bool dense = true;
    while (n_lanes) {
        const int check_closure = (max_t == INF || T < max_t) ? 1 : 0;
        memset(ctr, 0, 56);                                 // kPerRoundBytes of csrc/bsi.cpp
        ctr->n_hits = 0;
        const u64 *in_q = dense ? nullptr : retry[which ^ 1].data();
        u64 *retry_out = check_closure ? retry[which].data() : nullptr;
        u64 *closed_out = check_closure ? closed.data() : nullptr;
        const u64 hit_cap = n_lanes, n_words = (n_in + 31) >> 5;
        launch((unsigned)((n_words + 31) / 32), [&] {
            bsi_target_main(b, in_q, n_in, T, check_closure, hit_idx.data(), hit_t.data(), hit_cap, retry_out,
                            closed_out, ctr);
        });
        for (u64 i = 0; i < ctr->n_hits; ++i) hits.push_back({hit_idx[i], hit_t[i]});
        n_without += ctr->n_found;
        if (!check_closure || ctr->n_retry == 0) break;
        if (T >= kMaxH) return 4;
        n_in = ctr->n_retry;
        dense = false;
        which ^= 1;
        T = std::min(T * 4, kMaxH);
        if (max_t != INF && T > max_t) T = max_t;
    }
    std::sort(hits.begin(), hits.end());
    for (auto &h : hits) printf("hit %llu %llu\n", (unsigned long long)h.first, (unsigned long long)h.second);
    printf("totals %zu %llu %llu\n", hits.size(), (unsigned long long)(n_without + ctr->n_without),
           (unsigned long long)ctr->steps);
    return 0;
}
