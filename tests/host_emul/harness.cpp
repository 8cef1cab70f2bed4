// Host emulation of the generated macro header (test infrastructure).  The header libbsi prepends
// to the kernel template is plain preprocessor text over 32-bit variables, so a C++ program that
// supplies bsi_lop3 / bsi_transpose32 can execute BSI_UPDATE, BSI_PERTURB, BSI_INIT_DENSE,
// BSI_LANE_SETUP, BSI_PACK / BSI_UNPACK and the filter compares on the CPU, word for word what
// the GPU executes per thread.  tests/test_codegen_host.py feeds it and checks the answers
// against the truth tables / the problem decoder.
//
// stdin:  commands, one per line
//   U  a_0 .. a_{N-1}                 -> BSI_UPDATE(b, a): prints b_0 .. b_{N-1}
//   Q  a_0 .. a_{N-1}                 -> BSI_UPDATE_QUIET(b, a) (fixed nodes folded into their readers)
//   P  t  a_0 .. a_{N-1}              -> BSI_PERTURB(a, t): prints a_0 .. a_{N-1}
//   D  W                              -> BSI_INIT_DENSE(a, W): prints a_0 .. a_{N-1}
//   L  idx                            -> BSI_LANE_SETUP(0, idx, c): prints c[0..KW)
//   T  a_0 .. a_{N-1}                 -> BSI_UNPACK then BSI_PACK: prints the 32 lane codes (KW words
//                                        each), then the re-packed words
//   F  alive a_0 .. a_{N-1} k_0 .. k_{N-1}  -> the hot-loop filter of a against checkpoint k:
//                                        prints p (lanes that cannot match), the reference value
//                                        computed term by term, the full-compare mask, and p
//                                        again through the single-chain hot-loop macro
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

typedef unsigned char u8;
typedef unsigned int u32;
typedef unsigned long long u64;

template <int LUT>
static inline u32 bsi_lop3(u32 a, u32 b, u32 c) {
    u32 r = 0;
    for (int i = 0; i < 8; ++i)
        if ((LUT >> i) & 1) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
    return r;
}

// the contract of the device function (B[r] bit c = A[31-c] bit 31-r), bit by bit
#define BSI_TRANSPOSE32(A) bsi_transpose32(A)
static void bsi_transpose32(unsigned int (&A)[32]) {
    unsigned int B[32];
    for (int r = 0; r < 32; ++r) {
        B[r] = 0;
        for (int c = 0; c < 32; ++c) B[r] |= ((A[31 - c] >> (31 - r)) & 1u) << c;
    }
    for (int r = 0; r < 32; ++r) A[r] = B[r];
}

#include BSI_HEADER

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        char cmd = 0;
        in >> cmd;
#define DECL(i) u32 a##i = 0, b##i = 0, k##i = 0;
        BSI_FOR_NODES(DECL)
#undef DECL
#define READ_A(i) in >> a##i;
#define READ_K(i) in >> k##i;
#define PRINT_A(i) std::printf("%u ", a##i);
#define PRINT_B(i) std::printf("%u ", b##i);
        if (cmd == 'U') {
            BSI_FOR_NODES(READ_A)
            BSI_UPDATE(b, a)
            BSI_FOR_NODES(PRINT_B)
        } else if (cmd == 'Q') {
            BSI_FOR_NODES(READ_A)
            BSI_UPDATE_QUIET(b, a)
            BSI_FOR_NODES(PRINT_B)
        } else if (cmd == 'P') {
            u64 t = 0;
            in >> t;
            BSI_FOR_NODES(READ_A)
            BSI_PERTURB(a, t)
            BSI_FOR_NODES(PRINT_A)
        } else if (cmd == 'D') {
            u64 W = 0;
            in >> W;
            BSI_INIT_DENSE(a, W)
            BSI_FOR_NODES(PRINT_A)
        } else if (cmd == 'L') {
            u64 idx = 0, c[BSI_KW];
            in >> idx;
            for (int i = 0; i < BSI_KW; ++i) c[i] = 0;
            u64 lp_lane_ = 0;
            BSI_LANE_SETUP(0, idx, c)
            (void)lp_lane_;
            for (int i = 0; i < BSI_KW; ++i) std::printf("%llu ", c[i]);
        } else if (cmd == 'T') {
            u64 cw[32][BSI_KW];
            BSI_FOR_NODES(READ_A)
            BSI_UNPACK(cw, a)
            for (int l = 0; l < 32; ++l)
                for (int i = 0; i < BSI_KW; ++i) std::printf("%llu ", cw[l][i]);
            BSI_PACK(b, cw)
            BSI_FOR_NODES(PRINT_B)
        } else if (cmd == 'F') {
            u32 alive = 0;
            in >> alive;
            BSI_FOR_NODES(READ_A)
            BSI_FOR_NODES(READ_K)
            u32 p = ~alive, q = 0u;
            BSI_CMP_FILTER(p, q, a, k)
            p |= q;
            u32 full = p;
            BSI_CMP_REST(full, a, k)
            u32 ref = ~alive;
#define REF(i) ref |= a##i ^ k##i;
            BSI_FOR_FILTER(REF)
            const u32 ref_filter = ref;
            BSI_FOR_REST(REF)
#undef REF
            u32 p1 = 0x12345678u;                       // (whatever it held before is overwritten)
            BSI_CMP_FILTER1(p1, alive, a, k)
            std::printf("%u %u %u %u %u", p, ref_filter, full, ref, p1);
        }
        std::printf("\n");
    }
    return 0;
}
