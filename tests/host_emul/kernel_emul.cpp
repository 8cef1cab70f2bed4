// The attract kernels of bsi_template.cuh -- the real template text, not a restatement -- executed
// on the host: cuda_shim.h supplies a 32-thread warp of host threads, this file the launch
// sequence of csrc/bsi.cpp (attract_enqueue: detection rounds, the dense / queue trajectory
// kernel, then the four resolution kernels, the device choosing between them) with one block of
// BSI_BLOCK threads (1 to 8 independent warps) as the whole grid.  Test infrastructure: tests/test_kernel_emul.py compares the
// attractor table and the counters it prints with the C oracle.
//
// usage: kernel_emul idx_begin idx_end max_t max_attractor_l first_horizon few_words no_early table_log2
//        (max_t / max_attractor_l: 0 = no cap)
#include "cuda_shim.h"

#define BSI_HOST_EMUL 1
#include BSI_HEADER
#undef BSI_UNIT
#define BSI_UNIT 0
#include BSI_TEMPLATE

#include <algorithm>
#include <functional>
#include <vector>

extern "C" int bsi_attract_schedule(uint64_t horizon, uint32_t *ck_time, uint32_t *ck_slot, uint32_t cap,
                                    uint32_t *n_ck, uint32_t *end, uint32_t *first_final);

unsigned char bsi_smem[3 * BSI_N * BSI_BLOCK * 4 + 16];

static void launch(const std::function<void()> &kernel) { bsi_emul_block(BSI_BLOCK, kernel); }

template <typename T>
static T *device_alloc(size_t n) { return (T *)calloc(n ? n : 1, sizeof(T)); }

int main(int argc, char **argv) {
    if (argc < 9) return 2;
    static_assert(BSI_BLOCK % 32 == 0 && BSI_BLOCK <= 32 * BSI_EMUL_MAX_WARPS, "generate the header with block_threads = 32 .. 256");
    const u64 idx_begin = strtoull(argv[1], nullptr, 10), idx_end = strtoull(argv[2], nullptr, 10);
    const u64 INF = ~0ull;
    u64 max_t = strtoull(argv[3], nullptr, 10), max_l = strtoull(argv[4], nullptr, 10);
    if (max_t == 0) max_t = INF;
    if (max_l == 0) max_l = INF;
    u64 first_horizon = strtoull(argv[5], nullptr, 10);
    u64 few = strtoull(argv[6], nullptr, 10);
    if (few == 0) few = 32768;
    const int no_early = atoi(argv[7]);
    u32 table_log2 = (u32)atoi(argv[8]);

    const u64 lp = BSI_LP_ORIGIN;
    const u64 n_problems = idx_end - idx_begin;
    BsiCounters *ctr = device_alloc<BsiCounters>(1);
    BsiSlot *table = device_alloc<BsiSlot>((size_t)1 << table_log2);
    if (n_problems == 0 || max_t <= lp) {
        printf("counters %llu 0 %llu 0 0\n", (unsigned long long)(n_problems * lp), (unsigned long long)n_problems);
        return 0;
    }
    const u64 T = max_t == INF ? INF : max_t - lp;
    const u64 kMaxH = 1ull << 30;
    u32 *scratch = device_alloc<u32>((size_t)BSI_BLOCK * (BSI_SCRATCH_WORDS + 3 * BSI_N));
    const u64 b = idx_begin, e = idx_end;
    const u64 n_lanes = e - b;
    BsiFound *found_q = device_alloc<BsiFound>(n_lanes + 64 + 1024);
    u64 *retry_q[2] = {device_alloc<u64>(n_lanes + 64), device_alloc<u64>(n_lanes + 64)};
    const u64 word0 = b >> 5, n_words_dense = ((e + 31) >> 5) - word0;
    void *early_q = calloc(((n_words_dense + 31) / 32) * BSI_EQ_BLOCK_BYTES + 64, 1);

    // cycle dictionary (argv[9]: 1 = off; log2 of the orbit storage in argv[10], default 12)
    const int no_dict = argc > 9 ? atoi(argv[9]) : 0;
    const u32 dict_log2_states = argc > 10 ? (u32)atoi(argv[10]) : 12u;
    BsiDict dict;
    memset(&dict, 0, sizeof dict);
    if (!no_dict) {
        dict.slots = device_alloc<BsiDictSlot>((size_t)2 << dict_log2_states);
        dict.orbit = device_alloc<u64>(((size_t)1 << dict_log2_states) * BSI_KW);
        dict.n_states = device_alloc<u64>(2);
        dict.n_attr = dict.n_states + 1;
        dict.cap_attr = 64;
        dict.dir = device_alloc<u32>(2 * dict.cap_attr);
        dict.log2_slots = dict_log2_states + 1;
        dict.cap_states = 1u << dict_log2_states;
    }

    u64 H = first_horizon ? first_horizon : 1024;
    H = std::min(H, kMaxH);
    int which = 0, rounds = 0;
    bool dense = true;
    u64 n_in = n_lanes;
    for (;;) {
        if (H >= T) H = T;
        const int final_round = H == T ? 1 : 0;
        BsiSched sched;
        memset(&sched, 0, sizeof sched);
        if (bsi_attract_schedule(H, sched.ck_time, sched.ck_slot, BSI_MAX_CK, &sched.n_ck, &sched.end,
                                 &sched.first_final)) return 3;
        memset(ctr, 0, 56);                                  // kPerRoundBytes of csrc/bsi.cpp
        u64 stop_t = max_t;
        u64 *retry_out = final_round ? nullptr : retry_q[which];
        u64 early_cap = 0;
        if (dense) {
            early_cap = no_early ? 0 : n_words_dense;
            launch([&] { bsi_attract_dense(word0, n_words_dense, b, e, sched, final_round, stop_t, found_q,
                                           retry_out, scratch, early_cap ? early_q : nullptr, early_cap, ctr); });
        } else {
            const u64 *in_q = retry_q[which ^ 1];
            launch([&] { bsi_attract_queue(in_q, n_in, sched, final_round, stop_t, found_q, retry_out, scratch, ctr); });
        }
        const u64 *n_found_p = &ctr->n_found;
        void *early_p = early_cap ? early_q : nullptr;
        if (dense && !no_early) {
            const u64 *n_early_p = &ctr->n_early;
            launch([&] { bsi_resolve_early(early_p, early_cap, n_early_p, n_found_p, few, T, max_l, max_t, table,
                                           table_log2, ctr, dict); });
        }
        launch([&] { bsi_resolve_mu(found_q, n_found_p, few, T, max_l, max_t, table, table_log2, ctr, dict); });
        launch([&] { bsi_resolve(found_q, n_found_p, few, early_p, early_cap, T, max_l, max_t, table, table_log2, ctr, dict); });
        launch([&] { bsi_resolve_finish(found_q, n_found_p, few, T, max_l, max_t, table, table_log2, ctr); });
        ++rounds;
        if (final_round) break;
        if (ctr->n_retry == 0) break;
        if (H >= kMaxH) return 4;
        const bool mostly_open = ctr->n_retry * 2 > n_in;
        n_in = ctr->n_retry;
        dense = false;
        which ^= 1;
        if (mostly_open) H = T != INF ? std::min(T, kMaxH) : std::min(H * 16, kMaxH);
        else H = std::min(H * 4, kMaxH);
    }
    printf("counters %llu %llu %llu %llu %llu rounds %d dict %llu %llu dbg %llu %llu %llu %llu\n", (unsigned long long)ctr->steps,
           (unsigned long long)ctr->n_accepted, (unsigned long long)ctr->n_without,
           (unsigned long long)ctr->table_full, (unsigned long long)ctr->internal_error, rounds,
           (unsigned long long)(no_dict ? 0 : *dict.n_attr), (unsigned long long)(no_dict ? 0 : *dict.n_states),
           (unsigned long long)ctr->dbg[0], (unsigned long long)ctr->dbg[1], (unsigned long long)ctr->dbg[2],
           (unsigned long long)ctr->dbg[3]);
    // every stored orbit must be a cycle of the right length that starts at its minimum state
    if (!no_dict) {
        for (u64 a = 0; a < *dict.n_attr && a < dict.cap_attr; ++a) {
            const u32 base = dict.dir[2 * a], lam = dict.dir[2 * a + 1];
            for (u32 ph = 1; ph < lam; ++ph) {
                bool less = false, decided = false;
                for (int k = BSI_KW - 1; k >= 0 && !decided; --k) {
                    const u64 x = dict.orbit[(u64)(base + ph) * BSI_KW + k], y = dict.orbit[(u64)base * BSI_KW + k];
                    if (x != y) { less = x < y; decided = true; }
                }
                if (!decided || less) { printf("BAD ORBIT attractor %llu phase %u\n", (unsigned long long)a, ph); return 5; }
            }
        }
    }
    for (u64 i = 0; i < (1ull << table_log2); ++i) {
        const BsiSlot &s = table[i];
        if (s.flag != 2u) continue;
        printf("rec");
        for (int k = 0; k < BSI_KW; ++k) printf(" %llu", (unsigned long long)s.key[k]);
        printf(" | %llu %llu %llu %u\n", (unsigned long long)s.n, (unsigned long long)s.s1,
               (unsigned long long)s.s2, s.lam);
    }
    return 0;
}
