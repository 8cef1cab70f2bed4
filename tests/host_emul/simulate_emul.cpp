// bsi_trajectories (simulate mode; also the orbit / hit-trajectory generator) executed on the host:
// the kernel text of bsi_template.cuh under cuda_shim.h, launched block by block as
// run_trajectories in csrc/bsi.cpp does.  Test infrastructure (tests/test_kernel_emul.py).
//
// usage: simulate_emul idx_begin idx_end n_steps out_file      (explicit lists: see --list below)
//        simulate_emul --list n_steps out_file idx idx idx ...
// out_file receives u64 [n][n_steps + 1][BSI_KW], little endian.
#include "cuda_shim.h"

#define BSI_HOST_EMUL 1
#include BSI_HEADER
#undef BSI_UNIT
#define BSI_UNIT 0
#include BSI_TEMPLATE

#include <functional>
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
    if (argc < 5) return 2;
    std::vector<u64> indices;
    u64 idx_begin = 0, n = 0, n_steps = 0;
    const char *out_path = nullptr;
    if (!strcmp(argv[1], "--list")) {
        n_steps = strtoull(argv[2], nullptr, 10);
        out_path = argv[3];
        for (int i = 4; i < argc; ++i) indices.push_back(strtoull(argv[i], nullptr, 10));
        n = indices.size();
    } else {
        idx_begin = strtoull(argv[1], nullptr, 10);
        n = strtoull(argv[2], nullptr, 10) - idx_begin;
        n_steps = strtoull(argv[3], nullptr, 10);
        out_path = argv[4];
    }
    std::vector<u64> out((size_t)(n ? n : 1) * (n_steps + 1) * BSI_KW, 0ull);
    const u64 *idx_p = indices.empty() ? nullptr : indices.data();
    const u64 n_words = (n + 31) >> 5;
    if (n) launch((unsigned)((n_words + 31) / 32), [&] { bsi_trajectories(idx_begin, idx_p, nullptr, n, n_steps, out.data()); });
    FILE *f = fopen(out_path, "wb");
    if (!f) return 3;
    fwrite(out.data(), 8, (size_t)n * (n_steps + 1) * BSI_KW, f);
    fclose(f);
    return 0;
}
