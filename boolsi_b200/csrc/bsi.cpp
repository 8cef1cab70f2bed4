// libbsi -- host side of the C ABI declared in include/bsi.h.
//
// A plan = one network + problem space + mode, specialised through NVRTC into the hand-written
// kernel template (bsi_template.cuh) for sm_100a, loaded with the CUDA driver API (libcuda is
// opened lazily with dlopen so that the library itself loads -- and can specialise kernels --
// on a machine without a GPU; every run entry point fails with BSI_E_CUDA there: there is no
// CPU fallback).
//
// What the run functions replace in the reference: the batch/problem loops of
// mpi.py:131-190, 498-544 and the per-mode solvers of attract.py:262-455, target.py:109-133,
// simulate.py:97-178 (see include/bsi.h for the per-function citations).
#include <cuda.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <future>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "bsi_codegen.hpp"
#include "bsi_template_inc.h"   // generated from bsi_template.cuh: bsi_template_src[]

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// BSI_DEBUG_TIMING=1: host-side stage times on stderr
struct StageTimer {
    bool on = getenv("BSI_DEBUG_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[bsi] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---- CUDA driver API, bound at first use -------------------------------------------------------
#define BSI_STR2(x) #x
#define BSI_STR(x) BSI_STR2(x)
#define BSI_CU_FUNCS(X)                                                                          \
    X(cuInit) X(cuDeviceGetCount) X(cuDeviceGet) X(cuDeviceGetAttribute) X(cuDeviceGetName)       \
    X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease) X(cuCtxPushCurrent)                  \
    X(cuCtxPopCurrent) X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction)               \
    X(cuFuncSetAttribute) X(cuFuncGetAttribute) X(cuOccupancyMaxActiveBlocksPerMultiprocessor)    \
    X(cuLaunchKernel) X(cuMemAlloc) X(cuMemFree) X(cuMemsetD8Async) X(cuMemcpyDtoHAsync)          \
    X(cuMemcpyHtoDAsync) X(cuMemcpyDtoDAsync) X(cuStreamSynchronize) X(cuStreamCreate)            \
    X(cuStreamDestroy) X(cuEventCreate) X(cuEventRecord) X(cuEventSynchronize)                    \
    X(cuEventElapsedTime) X(cuEventDestroy) X(cuGetErrorString) X(cuMemAllocHost)                 \
    X(cuMemFreeHost) X(cuDeviceCanAccessPeer) X(cuCtxEnablePeerAccess) X(cuMemcpyPeerAsync)       \
    X(cuStreamWaitEvent)

struct CuApi {
#define X(name) decltype(&name) p_##name = nullptr;
    BSI_CU_FUNCS(X)
#undef X
    bool ok = false;
    std::string why;
};

CuApi &cu() {
    static CuApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.why = "libcuda.so.1 not found (no NVIDIA driver): libbsi has no CPU fallback"; return; }
        // cuda.h maps the plain names to their current versioned symbols (e.g. cuMemAlloc_v2)
#define X(name)                                                                                   \
        api.p_##name = (decltype(&name))dlsym(h, BSI_STR(name));                                  \
        if (!api.p_##name) { api.why = std::string("missing driver symbol ") + BSI_STR(name); return; }
        BSI_CU_FUNCS(X)
#undef X
        CUresult r = api.p_cuInit(0);
        if (r != CUDA_SUCCESS) { api.why = "cuInit failed (" + std::to_string((int)r) + ")"; return; }
        api.ok = true;
    });
    return api;
}

const char *cu_errstr(CUresult r) {
    const char *s = nullptr;
    if (cu().p_cuGetErrorString && cu().p_cuGetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
    return "unknown CUDA error";
}

#define CU_CHECK(call)                                                                            \
    do {                                                                                          \
        CUresult r_ = cu().p_##call;                                                              \
        if (r_ != CUDA_SUCCESS)                                                                   \
            return fail(BSI_E_CUDA, "%s failed: %s (%s:%d)", #call, cu_errstr(r_), __FILE__, __LINE__); \
    } while (0)

// mirrors of the device structs in bsi_template.cuh
struct Counters {
    uint64_t next_word, next_word2, next_word3, n_found, n_retry, n_early, next_word4;
    uint64_t n_without, n_accepted, steps, device_steps, table_full, n_hits, internal_error;
    uint64_t pad[2];
    uint64_t dbg[4];
};
constexpr size_t kPerRoundBytes = 7 * sizeof(uint64_t);
constexpr int kMaxCk = 72;
// (bsi_template.cuh: BSI_SCRATCH_WORDS, BSI_EQ_ENTRY_BYTES)
constexpr size_t kScratchWords = 160;
constexpr size_t early_entry_bytes(int n_nodes) { return 8 + 4 + 4 + 32 * 4 + 32 * 4 + 4 * (size_t)n_nodes; }
struct Sched {
    uint32_t n_ck, end, first_final, pad;
    uint32_t ck_time[kMaxCk];
    uint32_t ck_slot[kMaxCk];
};

// cycle dictionary (see bsi_template.cuh): kernel parameter, passed by value
struct DictParam {
    CUdeviceptr slots, orbit, n_states, n_attr, dir;
    uint32_t log2_slots, cap_states, cap_attr, pad;
};
constexpr uint32_t kDictLog2States = 18;          // orbit storage: 2^18 states, hash slots: twice that
constexpr uint32_t kDictCapAttr = 1u << 14;

// Checkpoint schedule for horizon H >= 1 (see the header comment of bsi_template.cuh).
Sched make_sched(uint64_t H) {
    Sched s{};
    const uint64_t c1 = (H + 2) / 3, c2 = (2 * H + 2) / 3, c3 = H;
    auto push = [&](uint64_t t, uint32_t slot) {
        s.ck_time[s.n_ck] = (uint32_t)t;
        s.ck_slot[s.n_ck] = slot;
        ++s.n_ck;
    };
    // early detection: 0, 1, 2, 3, 4, 6, 8, 12, 16, 24, ... while below the first final checkpoint
    if (c1 > 0) push(0, 0);
    for (uint64_t p = 1; p < c1; p *= 2) {
        push(p, 0);
        if (p >= 2 && p + p / 2 < c1) push(p + p / 2, 0);
    }
    uint64_t finals[3] = {c1, c2, c3};
    uint64_t prev = 0, max_gap = 0;
    uint32_t slot = 0;
    bool first = true;
    for (uint64_t c : finals) {
        if (!first && c == prev) continue;
        // a lane with prev < mu <= c (first: 0 <= mu <= c) has lambda <= H - mu
        uint64_t gap = first ? c : c - prev - 1;
        max_gap = std::max(max_gap, gap);
        push(c, slot++);
        prev = c;
        first = false;
    }
    s.end = (uint32_t)(H + max_gap);
    s.first_final = (uint32_t)c1;
    return s;
}

struct DevBuf {
    CUdeviceptr p = 0;
    size_t bytes = 0;
};

// The primary context of a device is retained once and kept for the life of the process (as the
// CUDA runtime does): releasing the last reference would tear the context down, and the next plan
// would pay hundreds of ms to bring it back.
std::mutex g_dev_mu;
std::map<int, CUcontext> g_ctx;

// Freed device buffers are parked per device and reused by later plans (the found queue of a
// 2^26-state chunk is 3 GB; cuMemAlloc/cuMemFree cost from a fraction of a ms to tens of ms).
struct Parked { CUdeviceptr p; size_t bytes; };
std::map<int, std::vector<Parked>> g_pool;
size_t g_pool_bytes = 0;
constexpr size_t kPoolMinBytes = 256, kPoolMaxBytes = (size_t)24 << 30;

}  // namespace

struct bsi_plan {
    bsi::Model m;
    bsi::GenInfo info;
    bsi_plan_opts opts{};
    std::string source, log, header;       // header: the generated macro header (identifies the plan's units)
    std::shared_ptr<const std::vector<char>> cubins[7];     // per BSI_UNIT (1 dense, 2 queue, 3 + 5 + 6 resolve, 4 the rest)
    int block = 64;
    int device = -1;           // -1: compile-only plan
    CUdevice dev = 0;
    CUcontext ctx = nullptr;
    CUmodule mods[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool mod_cached[7] = {false, false, false, false, false, false, false};   // owned by the module cache
    std::string unit_key[7];         // identifies a unit's source (module cache key)
    CUstream own_stream = nullptr, stream = nullptr;
    CUfunction f_dense = nullptr, f_queue = nullptr, f_mu = nullptr, f_early = nullptr,
               f_resolve = nullptr, f_finish = nullptr,
               f_export = nullptr, f_merge = nullptr, f_merge_payloads = nullptr, f_traj = nullptr, f_counts = nullptr, f_learn_keys = nullptr,
               f_target = nullptr;
    int n_sms = 0;
    int main_blocks_per_sm = 0, key_blocks_per_sm = 0, mu_blocks_per_sm = 0, early_blocks_per_sm = 0, resolve_blocks_per_sm = 0;
    size_t main_smem = 0;
    int smem_slots = 3;
    uint32_t table_log2 = 20;
    DevBuf ctr, table, found_q, retry_q[2], scratch, early_q, misc, misc2, misc3, misc4, misc5;
    DevBuf dict_slots, dict_orbit, dict_dir, dict_ctr;     // cycle dictionary (attract)
    DictParam dict{};
    bsi_progress_fn progress = nullptr;                    // reports finished problems of long runs
    void *progress_user = nullptr;
    std::vector<CUevent> events;
    size_t events_used = 0;
    uint64_t launches = 0;
    // a run enqueued by bsi_attract_run_async and not yet collected by bsi_attract_run_finish
    bool pending = false, pending_trivial = false;
    bsi_run_stats pending_stats{};
    uint64_t pending_launches0 = 0;
    CUevent pending_e0 = nullptr, pending_e1 = nullptr;
    std::vector<std::pair<CUevent, CUevent>> pending_main, pending_resolve;
};

namespace {

struct CtxGuard {
    bool pushed = false;
    explicit CtxGuard(bsi_plan *p) {
        if (p && p->ctx && cu().ok) pushed = cu().p_cuCtxPushCurrent(p->ctx) == CUDA_SUCCESS;
    }
    ~CtxGuard() {
        if (pushed) { CUcontext c; cu().p_cuCtxPopCurrent(&c); }
    }
};

int device_context(int device, CUdevice *dev, CUcontext *ctx) {
    CU_CHECK(cuDeviceGet(dev, device));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    auto it = g_ctx.find(device);
    if (it == g_ctx.end()) {
        CUcontext c = nullptr;
        CU_CHECK(cuDevicePrimaryCtxRetain(&c, *dev));
        it = g_ctx.emplace(device, c).first;
    }
    *ctx = it->second;
    return BSI_OK;
}

void give_back(int device, DevBuf &b) {
    if (!b.p) return;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        if (b.bytes >= kPoolMinBytes && g_pool_bytes + b.bytes <= kPoolMaxBytes) {
            g_pool[device].push_back({b.p, b.bytes});
            g_pool_bytes += b.bytes;
            b.p = 0;
            b.bytes = 0;
            return;
        }
    }
    cu().p_cuMemFree(b.p);
    b.p = 0;
    b.bytes = 0;
}

int ensure(bsi_plan *p, DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return BSI_OK;
    give_back(p->device, b);
    size_t want = std::max<size_t>(bytes, 256);
    if (want < ((size_t)1 << 20)) {          // small buffers: power-of-two sizes, so that they get reused
        size_t r = 256;
        while (r < want) r <<= 1;
        want = r;
    }
    if (want >= kPoolMinBytes) {
        // best fit among the parked buffers of this device, wasting at most 2x
        std::lock_guard<std::mutex> lk(g_dev_mu);
        auto &v = g_pool[p->device];
        int best = -1;
        for (int i = 0; i < (int)v.size(); ++i)
            if (v[i].bytes >= want && v[i].bytes <= 2 * want && (best < 0 || v[i].bytes < v[best].bytes)) best = i;
        if (best >= 0) {
            b.p = v[best].p;
            b.bytes = v[best].bytes;
            g_pool_bytes -= v[best].bytes;
            v.erase(v.begin() + best);
            return BSI_OK;
        }
    }
    CUresult r = cu().p_cuMemAlloc(&b.p, want);
    if (r == CUDA_ERROR_OUT_OF_MEMORY) {
        // drop everything parked on this device and retry once
        std::vector<Parked> drop;
        {
            std::lock_guard<std::mutex> lk(g_dev_mu);
            drop.swap(g_pool[p->device]);
            for (auto &q : drop) g_pool_bytes -= q.bytes;
        }
        for (auto &q : drop) cu().p_cuMemFree(q.p);
        r = cu().p_cuMemAlloc(&b.p, want);
    }
    if (r != CUDA_SUCCESS) {
        b.p = 0;
        return fail(r == CUDA_ERROR_OUT_OF_MEMORY ? BSI_E_NOMEM : BSI_E_CUDA,
                    "cuMemAlloc(%zu bytes) failed: %s", want, cu_errstr(r));
    }
    b.bytes = want;
    (void)p;
    return BSI_OK;
}

int get_event(bsi_plan *p, CUevent *out) {
    if (p->events_used == p->events.size()) {
        CUevent e;
        CU_CHECK(cuEventCreate(&e, CU_EVENT_DEFAULT));
        p->events.push_back(e);
    }
    *out = p->events[p->events_used++];
    return BSI_OK;
}

int launch(bsi_plan *p, CUfunction f, unsigned grid, unsigned block, size_t smem, void **args) {
    if (grid == 0) return BSI_OK;
    CU_CHECK(cuLaunchKernel(f, grid, 1, 1, block, 1, 1, (unsigned)smem, p->stream, args, nullptr));
    ++p->launches;
    return BSI_OK;
}

// ---- filter-node choice: a small host-side activity probe of the network ---------------------
// (tuning only: the filter never decides a result, every filter pass is followed by a full
//  compare.)  Nodes that keep toggling late in random trajectories discriminate states best.
std::vector<int> choose_filter(const bsi::Model &m, int want, std::vector<int> *second, int want2) {
    const int n = m.n;
    want = std::min(want, n);
    std::vector<uint8_t> x(n), y(n);
    std::vector<long> toggles(n, 0);
    std::mt19937_64 rng(0xB51u);
    bool too_wide = false;
    for (int i = 0; i < n; ++i) if (m.fn[i].k() > 20) too_wide = true;
    const int n_traj = too_wide ? 0 : 48, n_steps = 192;
    for (int tr = 0; tr < n_traj; ++tr) {
        for (int i = 0; i < n; ++i) x[i] = (uint8_t)(rng() & 1);
        for (int t = 0; t < n_steps; ++t) {
            for (int i = 0; i < n; ++i) {
                auto fx = m.fixed.find(i);
                if (fx != m.fixed.end()) { y[i] = (uint8_t)fx->second; continue; }
                size_t j = 0;
                for (int v : m.fn[i].vars) j = (j << 1) | x[v];
                y[i] = m.fn[i].tt[j];
            }
            if (t >= n_steps / 2)
                for (int i = 0; i < n; ++i) toggles[i] += x[i] != y[i];
            x.swap(y);
        }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return toggles[a] > toggles[b]; });
    std::vector<int> f(order.begin(), order.begin() + want);
    std::sort(f.begin(), f.end());
    if (second) {
        second->assign(order.begin() + want, order.begin() + std::min(n, want + want2));
        std::sort(second->begin(), second->end());
    }
    return f;
}

struct UnitResult {
    int status = BSI_OK;
    std::string error, log, key;
    std::vector<char> cubin;
};

// One NVRTC program -> sm_100a cubin.
UnitResult compile_unit(std::string src) {
    UnitResult out;
    nvrtcProgram prog;
    nvrtcResult r = nvrtcCreateProgram(&prog, src.c_str(), "bsi_kernels.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) {
        out.status = BSI_E_NVRTC;
        out.error = std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r);
        return out;
    }
    // (#550 "set but never used": generated per-node temporaries of rules that do not use them)
    const char *opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", "-default-device",
                          "--diag-suppress=550"};
    r = nvrtcCompileProgram(prog, 5, opts);
    size_t log_size = 0;
    nvrtcGetProgramLogSize(prog, &log_size);
    out.log.assign(log_size ? log_size : 1, '\0');
    if (log_size) nvrtcGetProgramLog(prog, &out.log[0]);
    while (!out.log.empty() && out.log.back() == '\0') out.log.pop_back();
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        out.status = BSI_E_NVRTC;
        out.error = std::string("kernel specialisation failed (") + nvrtcGetErrorString(r) + "):\n" + out.log.substr(0, 3500);
        return out;
    }
    size_t cubin_size = 0;
    r = nvrtcGetCUBINSize(prog, &cubin_size);
    if (r != NVRTC_SUCCESS || cubin_size == 0) {
        nvrtcDestroyProgram(&prog);
        out.status = BSI_E_NVRTC;
        out.error = std::string("nvrtcGetCUBINSize: ") + nvrtcGetErrorString(r);
        return out;
    }
    out.cubin.resize(cubin_size);
    nvrtcGetCUBIN(prog, out.cubin.data());
    nvrtcDestroyProgram(&prog);
    return out;
}

// In-process cache of compiled units.  The template text is a constant of the library, so a unit is
// identified by the generated header (network + space + mode + tuning) and its BSI_UNIT number.
struct CachedUnit {
    std::shared_ptr<const std::vector<char>> cubin;
    std::string log, key;
};
std::mutex g_unit_mu;
std::map<std::pair<std::string, int>, CachedUnit> g_units;

// Second level on disk, for the NEXT process that runs the same network (NVRTC takes a few seconds
// per network, a 2^30-state cambium run 50 ms): on by default in the user's cache directory.  One file per unit, named by a hash of (header, template, NVRTC version, unit); the
// file repeats the header, which is compared on load, so a hash collision cannot hand out the
// wrong kernels.  Written to a temporary name and renamed: readers never see a partial file.
uint64_t fnv1a(const std::string &text, uint64_t h = 0xCBF29CE484222325ull) {
    for (unsigned char c : text) { h ^= c; h *= 0x100000001B3ull; }
    return h;
}
// The cache directory: BSI_CACHE_DIR if set ("" or "off" switch the disk level off), else
// $XDG_CACHE_HOME/boolsi_b200, else ~/.cache/boolsi_b200 (created on first use; an unwritable or
// missing home simply means no disk level).
std::string disk_cache_dir() {
    const char *dir = getenv("BSI_CACHE_DIR");
    if (dir) return (!*dir || !strcmp(dir, "off")) ? std::string() : std::string(dir);
    std::string base;
    const char *xdg = getenv("XDG_CACHE_HOME"), *home = getenv("HOME");
    if (xdg && *xdg) base = xdg;
    else if (home && *home) base = std::string(home) + "/.cache";
    else return std::string();
    mkdir(base.c_str(), 0755);
    base += "/boolsi_b200";
    mkdir(base.c_str(), 0755);
    return base;
}
std::string disk_cache_path(const std::string &header, int unit) {
    const std::string dir_s = disk_cache_dir();
    const char *dir = dir_s.c_str();
    if (dir_s.empty()) return std::string();
    int major = 0, minor = 0;
    nvrtcVersion(&major, &minor);
    uint64_t h = fnv1a(header);
    h = fnv1a(bsi_template_src, h);
    h = fnv1a("nvrtc" + std::to_string(major) + "." + std::to_string(minor) + "/unit" + std::to_string(unit), h);
    char name[64];
    snprintf(name, sizeof name, "/bsi-%016llx-u%d.cubin", (unsigned long long)h, unit);
    return std::string(dir) + name;
}
const char kDiskMagic[8] = {'B', 'S', 'I', 'C', 'U', 'B', '1', '\n'};
bool disk_cache_load(const std::string &path, const std::string &header, CachedUnit *out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = false;
    char magic[8];
    uint64_t n[3];
    std::string h, log;
    std::vector<char> cubin;
    if (fread(magic, 1, 8, f) == 8 && memcmp(magic, kDiskMagic, 8) == 0 && fread(n, 8, 3, f) == 3 &&
        n[0] == header.size() && n[1] < (1u << 24) && n[2] > 0 && n[2] < (1u << 28)) {
        h.resize(n[0]);
        log.resize(n[1]);
        cubin.resize(n[2]);
        ok = fread(&h[0], 1, n[0], f) == n[0] && (n[1] == 0 || fread(&log[0], 1, n[1], f) == n[1]) &&
             fread(cubin.data(), 1, n[2], f) == n[2] && h == header;
    }
    fclose(f);
    if (!ok) return false;
    out->cubin = std::make_shared<const std::vector<char>>(std::move(cubin));
    out->log = log;
    return true;
}
void disk_cache_store(const std::string &path, const std::string &header, const CachedUnit &c) {
    const std::string tmp = path + ".tmp" + std::to_string((unsigned long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                   // an unwritable cache directory is not an error
    const uint64_t n[3] = {header.size(), c.log.size(), c.cubin->size()};
    bool ok = fwrite(kDiskMagic, 1, 8, f) == 8 && fwrite(n, 8, 3, f) == 3 &&
              fwrite(header.data(), 1, n[0], f) == n[0] &&
              (n[1] == 0 || fwrite(c.log.data(), 1, n[1], f) == n[1]) &&
              fwrite(c.cubin->data(), 1, n[2], f) == n[2];
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

// every queue of every run on this plan is short: see bsi_latency_mode in bsi_template.cuh
bool small_space(const bsi_plan *p) {
    const uint64_t few = p->opts.resolve_few_words ? p->opts.resolve_few_words : 32768;
    return p->m.n_problems / 32 + 2 <= few;
}

// Units already compiled in this process come from the cache without starting a thread.
int compile_units(bsi_plan *p, const std::vector<int> &units) {
    const std::string &header = p->header;
    // the source of a unit = its number + header + template (p->source may already carry the
    // "#define BSI_UNIT 0" line of bsi_plan_source: the units are built from the parts)
    const std::string body = header + "\n#line 1 \"bsi_template.cuh\"\n" + bsi_template_src;
    std::vector<std::pair<int, std::future<UnitResult>>> jobs;
    {
        std::lock_guard<std::mutex> lk(g_unit_mu);
        for (int u : units) {
            if (p->cubins[u] && !p->cubins[u]->empty()) continue;
            auto it = g_units.find({header, u});
            if (it != g_units.end()) {
                p->cubins[u] = it->second.cubin;
                p->unit_key[u] = it->second.key;
                if (!it->second.log.empty()) p->log += it->second.log + "\n";
                continue;
            }
            const std::string unit_key = std::to_string(header.size()) + ":" +
                                         std::to_string(std::hash<std::string>()(header)) + ":" + std::to_string(u);
            const std::string path = disk_cache_path(header, u);
            CachedUnit from_disk;
            if (!path.empty() && disk_cache_load(path, header, &from_disk)) {
                from_disk.key = unit_key;
                p->cubins[u] = from_disk.cubin;
                p->unit_key[u] = unit_key;
                if (!from_disk.log.empty()) p->log += from_disk.log + "\n";
                g_units[{header, u}] = std::move(from_disk);
                continue;
            }
            std::string src = "#define BSI_UNIT " + std::to_string(u) + "\n" + body;
            jobs.emplace_back(u, std::async(std::launch::async, compile_unit, std::move(src)));
        }
    }
    int rc = BSI_OK;
    for (auto &job : jobs) {
        UnitResult r = job.second.get();
        if (r.status != BSI_OK) {
            if (rc == BSI_OK) { rc = r.status; g_err = r.error; }
            continue;
        }
        CachedUnit c;
        c.cubin = std::make_shared<const std::vector<char>>(std::move(r.cubin));
        c.log = r.log;
        c.key = std::to_string(header.size()) + ":" + std::to_string(std::hash<std::string>()(header)) + ":" + std::to_string(job.first);
        if (!c.log.empty()) p->log += c.log + "\n";
        p->cubins[job.first] = c.cubin;
        p->unit_key[job.first] = c.key;
        const std::string path = disk_cache_path(header, job.first);
        if (!path.empty()) disk_cache_store(path, header, c);
        std::lock_guard<std::mutex> lk(g_unit_mu);
        if (g_units.size() >= 1024) g_units.clear();
        g_units[{header, job.first}] = std::move(c);
    }
    return rc;
}

int compile_plan(bsi_plan *p) {
    StageTimer tm;
    const int f_nodes = p->opts.filter_nodes ? (int)p->opts.filter_nodes : 12;
    std::vector<int> filter2, filter;
    {
        // the activity probe only depends on the rules and the origin fixed nodes: remember it
        static std::mutex mu;
        static std::map<std::string, std::pair<std::vector<int>, std::vector<int>>> probes;
        std::string key = std::to_string(f_nodes) + "|";
        for (int i = 0; i < p->m.n; ++i) key += p->m.fn[i].key() + ";";
        for (auto &kv : p->m.fixed) key += "f" + std::to_string(kv.first) + "=" + std::to_string(kv.second);
        std::lock_guard<std::mutex> lk(mu);
        auto it = probes.find(key);
        if (it == probes.end()) {
            filter = choose_filter(p->m, f_nodes, &filter2, 12);
            if (probes.size() < 256) probes[key] = {filter, filter2};
        } else {
            filter = it->second.first;
            filter2 = it->second.second;
        }
    }
    tm.lap("choose_filter");
    p->block = p->opts.block_threads ? (int)p->opts.block_threads : 64;
    if (p->block % 32 != 0 || p->block > 1024) return fail(BSI_E_ARG, "block_threads must be a multiple of 32");
    // Occupancy of the trajectory kernels.  A B200 scheduler owns 16 K registers: two warps fit at up
    // to 255 registers per thread, a third needs <= 168 (min_blocks = 6) AND six 64-thread blocks of
    // shared memory per SM (<= 37 KB each: smem_slots = 2 at N = 64, the third checkpoint slot in
    // the L2-resident scratch).  Measured at N = 64, K = 3: 2.2 % faster when almost no lane closes
    // (seed 0; the cap only spills slot-0 filter words in the three-slots-armed loop, 13 LDL per
    // step) but 7 % slower when 13 % of the lanes close (seed 2: the match handlers spill).  Not the
    // default; both knobs stay for tuning.
    const int min_blocks = p->opts.min_blocks ? (int)p->opts.min_blocks : 1;
    // Cycles up to this length are walked by the resolution kernels, longer ones looked up in the cycle
    // dictionary: the key walk + advance of a word cost about (2 U + 4 N) LOP3 per step of its longest
    // cycle, 32 lookups + gathers + the 32x32 transposes about 2800 instructions.
    const int walk_per_step = 2 * std::max(1, p->m.n) + 4 * p->m.n;      // (U <= N for K <= 3; a bound is enough)
    const int dict_min_lam = std::max(8, 2800 / walk_per_step);
    // Early words get keys AND P(0) out of ONE walk of max lambda - 1 steps, each an update, the
    // bit-sliced minimum and its select (about 4 N + 15 instructions, a long dependent chain), while
    // their way through the dictionary -- transposes to state codes, per-lane lookups and gathers,
    // transposes back -- costs about 2600 + 55 N instructions of memory latency that overlaps.
    // Measured on both sides of the line: cambium2 (N = 30, threshold 31) walks its 18-cycle -- the
    // dictionary cost the 6 % of its words that hold one 12 % of the kernel -- and the seed-1 network
    // (N = 64, threshold 22) looks its 29- and 58-cycles up (walking them: 8.6 instead of 5.5 ms).
    const int early_dict_min_lam = std::max(8, (2600 + 55 * p->m.n) / (4 * p->m.n + 15));
    p->smem_slots = p->opts.smem_slots ? (int)p->opts.smem_slots : 3;
    if (p->smem_slots < 1 || p->smem_slots > 3) return fail(BSI_E_ARG, "smem_slots must be 1, 2 or 3");
    const std::string header = bsi::generate_header(p->m, filter, filter2, p->block, min_blocks, (int)p->opts.debug_stats, p->info) +
                               "#define BSI_SMEM_SLOTS " + std::to_string(p->smem_slots) + "\n" +
                               "#define BSI_DICT_MIN_LAM " + std::to_string(dict_min_lam) + "u\n" +
                               "#define BSI_EARLY_DICT_MIN_LAM " + std::to_string(early_dict_min_lam) + "u\n" +
                               "#define BSI_TUNE " + std::to_string(p->opts.tune) + "\n";
    p->source = header + "\n#line 1 \"bsi_template.cuh\"\n" + bsi_template_src;
    tm.lap("generate source");
    if (p->opts.source_only) {                             // inspection / host emulation: no NVRTC
        p->source = "#define BSI_UNIT 0\n" + p->source;
        return BSI_OK;
    }

    p->header = header;
    // The translation unit is compiled once per BSI_UNIT, in parallel host threads (NVRTC is
    // thread-safe across programs): the two trajectory kernels dominate the compile time.  A SMALL
    // problem space (every queue of every run is short, so the device always picks the latency-
    // oriented resolution kernels) only needs the dense trajectory kernel, those kernels and the
    // rest; the queue kernel and the two throughput-mode resolution kernels are compiled when a run
    // first needs them (a retry round; ensure_unit).  For a three-node example that halves the cold
    // start, which is all a job of eight trajectories consists of.
    std::vector<int> units;
    if (p->m.mode != BSI_MODE_ATTRACT) units = {4};
    else if (small_space(p)) units = {1, 5, 4};
    else units = {1, 2, 3, 5, 6, 4};
    p->log.clear();
    int rc = compile_units(p, units);
    tm.lap("nvrtc (parallel units)");
    p->source = "#define BSI_UNIT 0\n" + p->source;       // what bsi_plan_source shows: the whole TU
    return rc;
}

// loaded modules are kept per (device, unit source) for the life of the process: a later plan
// for the same network skips cuModuleLoadData / cuModuleUnload (several ms)
int load_modules(bsi_plan *p) {
    static std::map<std::string, CUmodule> loaded;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    for (int u = 1; u <= 6; ++u) {
        if (p->mods[u] || !p->cubins[u] || p->cubins[u]->empty()) continue;
        const std::string key = std::to_string(p->device) + "/" + p->unit_key[u];
        auto it = loaded.find(key);
        if (it != loaded.end()) {
            p->mods[u] = it->second;
            p->mod_cached[u] = true;
            continue;
        }
        CU_CHECK(cuModuleLoadData(&p->mods[u], p->cubins[u]->data()));
        if (loaded.size() < 128) {
            loaded[key] = p->mods[u];
            p->mod_cached[u] = true;
        }
    }
    return BSI_OK;
}

// the kernels of the units that are compiled on demand (see compile_plan)
int bind_lazy_unit(bsi_plan *p, int u) {
    int a = 0;
    if (u == 2 && !p->f_queue) {
        CU_CHECK(cuModuleGetFunction(&p->f_queue, p->mods[2], "bsi_attract_queue"));
        if (p->main_smem > 48 * 1024)
            CU_CHECK(cuFuncSetAttribute(p->f_queue, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)p->main_smem));
    } else if (u == 3 && !p->f_mu) {
        CU_CHECK(cuModuleGetFunction(&p->f_mu, p->mods[3], "bsi_resolve_mu"));
        CU_CHECK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&a, p->f_mu, p->block, 0));
        p->mu_blocks_per_sm = std::max(1, a);
    } else if (u == 6 && !p->f_early) {
        CU_CHECK(cuModuleGetFunction(&p->f_early, p->mods[6], "bsi_resolve_early"));
        CU_CHECK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&a, p->f_early, p->block, 0));
        p->early_blocks_per_sm = std::max(1, a);
    }
    return BSI_OK;
}
// compile (if this plan skipped it), load and bind unit u: called with the plan's context current
int ensure_unit(bsi_plan *p, int u) {
    if (p->mods[u]) return BSI_OK;
    int rc = compile_units(p, {u});
    if (rc == BSI_OK) rc = load_modules(p);
    if (rc == BSI_OK) rc = bind_lazy_unit(p, u);
    return rc;
}

int load_plan(bsi_plan *p, int device) {
    if (!cu().ok) return fail(BSI_E_CUDA, "%s", cu().why.c_str());
    int count = 0;
    CU_CHECK(cuDeviceGetCount(&count));
    if (device < 0 || device >= count) return fail(BSI_E_CUDA, "device %d out of range (%d devices)", device, count);
    StageTimer tm;
    p->device = device;
    if (int rc0 = device_context(device, &p->dev, &p->ctx)) return rc0;
    CtxGuard g(p);
    tm.lap("context");
    int major = 0, minor = 0;
    CU_CHECK(cuDeviceGetAttribute(&major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, p->dev));
    CU_CHECK(cuDeviceGetAttribute(&minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, p->dev));
    if (major != 10) return fail(BSI_E_CUDA, "device %d is sm_%d%d; libbsi kernels are built for sm_100a (B200) only", device, major, minor);
    CU_CHECK(cuDeviceGetAttribute(&p->n_sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, p->dev));
    if (int rc1 = load_modules(p)) return rc1;
    tm.lap("cuModuleLoadData");
    CU_CHECK(cuStreamCreate(&p->own_stream, CU_STREAM_NON_BLOCKING));
    p->stream = p->own_stream;
    CU_CHECK(cuModuleGetFunction(&p->f_traj, p->mods[4], "bsi_trajectories"));
    CU_CHECK(cuModuleGetFunction(&p->f_counts, p->mods[4], "bsi_orbit_counts"));
    if (p->m.mode == BSI_MODE_ATTRACT) {
        CU_CHECK(cuModuleGetFunction(&p->f_dense, p->mods[1], "bsi_attract_dense"));
        CU_CHECK(cuModuleGetFunction(&p->f_resolve, p->mods[5], "bsi_resolve"));
        CU_CHECK(cuModuleGetFunction(&p->f_learn_keys, p->mods[5], "bsi_dict_learn_keys"));
        CU_CHECK(cuModuleGetFunction(&p->f_finish, p->mods[5], "bsi_resolve_finish"));
        CU_CHECK(cuModuleGetFunction(&p->f_export, p->mods[4], "bsi_table_export"));
        CU_CHECK(cuModuleGetFunction(&p->f_merge, p->mods[4], "bsi_table_merge"));
        CU_CHECK(cuModuleGetFunction(&p->f_merge_payloads, p->mods[4], "bsi_table_merge_payloads"));
        p->main_smem = (size_t)p->smem_slots * p->m.n * p->block * 4;      // the checkpoint slots held on chip
        if (p->main_smem > 48 * 1024)
            CU_CHECK(cuFuncSetAttribute(p->f_dense, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)p->main_smem));
        for (int u : {2, 3, 6})
            if (p->mods[u])
                if (int rcu = bind_lazy_unit(p, u)) return rcu;
        // (the persistent grid of both trajectory kernels and the scratch behind it follow the dense
        //  kernel; the queue kernel takes its work from a counter, so a grid it cannot hold at once
        //  still finishes)
        int a = 0;
        CU_CHECK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&a, p->f_dense, p->block, p->main_smem));
        p->main_blocks_per_sm = std::max(1, a);
        CU_CHECK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&a, p->f_finish, p->block, 0));
        p->key_blocks_per_sm = std::max(1, std::min(a, 8));
        CU_CHECK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&a, p->f_resolve, p->block, 0));
        p->resolve_blocks_per_sm = std::max(1, a);
        p->table_log2 = p->opts.table_log2_slots ? p->opts.table_log2_slots : 20;
        if (p->table_log2 < 4 || p->table_log2 > 28) return fail(BSI_E_ARG, "table_log2_slots out of range");
        const size_t slot_bytes = (size_t)8 * p->m.kw + 32;
        int rc = ensure(p, p->table, slot_bytes << p->table_log2);
        if (rc) return rc;
        CU_CHECK(cuMemsetD8Async(p->table.p, 0, p->table.bytes, p->stream));
        if (!p->opts.no_dict) {
            const size_t dslot_bytes = (16 + 8 * (size_t)p->m.kw + 15) / 16 * 16;
            const uint32_t log2_states = kDictLog2States;
            if ((rc = ensure(p, p->dict_slots, dslot_bytes << (log2_states + 1))) ||
                (rc = ensure(p, p->dict_orbit, ((size_t)8 * p->m.kw) << log2_states)) ||
                (rc = ensure(p, p->dict_dir, (size_t)8 * kDictCapAttr)) || (rc = ensure(p, p->dict_ctr, 256)))
                return rc;
            p->dict = DictParam{p->dict_slots.p, p->dict_orbit.p, p->dict_ctr.p, p->dict_ctr.p + 8, p->dict_dir.p,
                                log2_states + 1, 1u << log2_states, kDictCapAttr, 0};
            CU_CHECK(cuMemsetD8Async(p->dict_slots.p, 0, dslot_bytes << (log2_states + 1), p->stream));
            CU_CHECK(cuMemsetD8Async(p->dict_ctr.p, 0, 256, p->stream));
        }
    }
    if (p->m.mode == BSI_MODE_TARGET) CU_CHECK(cuModuleGetFunction(&p->f_target, p->mods[4], "bsi_target_main"));
    int rc = ensure(p, p->ctr, sizeof(Counters));
    if (rc) return rc;
    CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, sizeof(Counters), p->stream));
    CU_CHECK(cuStreamSynchronize(p->stream));
    tm.lap("functions + table alloc");
    return BSI_OK;
}

int create_plan(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                const bsi_target_desc *target, const bsi_plan_opts *opts, int device, bsi_plan **out) {
    if (!out) return fail(BSI_E_ARG, "null output pointer");
    *out = nullptr;
    if (mode < BSI_MODE_SIMULATE || mode > BSI_MODE_TARGET) return fail(BSI_E_ARG, "bad mode %d", mode);
    bsi_plan *p = new bsi_plan();
    if (opts) p->opts = *opts;
    std::string err = bsi::load_model(network, space, mode, target, p->m);
    if (!err.empty()) {
        delete p;
        bool unsup = err.find("not supported") != std::string::npos || err.find("unsupported") != std::string::npos;
        return fail(unsup ? BSI_E_UNSUPPORTED : BSI_E_ARG, "%s", err.c_str());
    }
    if (p->opts.source_only && device >= 0) {
        delete p;
        return fail(BSI_E_ARG, "source_only is for offline plans (bsi_plan_create_offline)");
    }
    int rc = compile_plan(p);
    if (rc == BSI_OK && device >= 0) rc = load_plan(p, device);
    if (rc != BSI_OK) {
        std::string keep = g_err;
        bsi_plan_destroy(p);
        g_err = keep;
        return rc;
    }
    *out = p;
    return BSI_OK;
}

int need_device(bsi_plan *p, int mode) {
    if (!p) return fail(BSI_E_ARG, "null plan");
    if (p->device < 0 || !p->mods[4]) return fail(BSI_E_CUDA, "this plan was specialised without a device and cannot run (no CPU fallback)");
    if (mode >= 0 && p->m.mode != mode) return fail(BSI_E_ARG, "plan was created for another mode");
    return BSI_OK;
}

int read_counters(bsi_plan *p, Counters *c) {
    CU_CHECK(cuMemcpyDtoHAsync(c, p->ctr.p, sizeof(Counters), p->stream));
    CU_CHECK(cuStreamSynchronize(p->stream));
    return BSI_OK;
}

struct Timer {
    std::vector<std::pair<CUevent, CUevent>> main, resolve;
    CUevent e0 = nullptr, e1 = nullptr;
};

int elapsed(CUevent a, CUevent b, double *ms) {
    float f = 0;
    CU_CHECK(cuEventElapsedTime(&f, a, b));
    *ms += f;
    return BSI_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int bsi_abi_version(void) { return BSI_ABI_VERSION; }
const char *bsi_last_error(void) { return g_err.c_str(); }

int bsi_device_count(int *count) {
    if (!count) return fail(BSI_E_ARG, "null pointer");
    *count = 0;
    if (!cu().ok) return fail(BSI_E_CUDA, "%s", cu().why.c_str());
    CU_CHECK(cuDeviceGetCount(count));
    return BSI_OK;
}

int bsi_plan_create(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                    const bsi_target_desc *target, const bsi_plan_opts *opts, int device,
                    bsi_plan **out) {
    if (device < 0) return fail(BSI_E_ARG, "device must be >= 0 (use bsi_plan_create_offline to specialise without a GPU)");
    return create_plan(network, space, mode, target, opts, device, out);
}

int bsi_plan_create_offline(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                            const bsi_target_desc *target, const bsi_plan_opts *opts, bsi_plan **out) {
    return create_plan(network, space, mode, target, opts, -1, out);
}

void bsi_plan_destroy(bsi_plan *p) {
    if (!p) return;
    StageTimer tm;
    if (p->ctx && cu().ok) {
        {
            CtxGuard g(p);
            if (p->stream) cu().p_cuStreamSynchronize(p->stream);
            for (DevBuf *b : {&p->ctr, &p->table, &p->found_q, &p->retry_q[0], &p->retry_q[1], &p->scratch,
                              &p->early_q, &p->misc, &p->misc2, &p->misc3, &p->misc4, &p->misc5, &p->dict_slots, &p->dict_orbit,
                              &p->dict_dir, &p->dict_ctr})
                give_back(p->device, *b);
            tm.lap("destroy: buffers");
            for (CUevent e : p->events) cu().p_cuEventDestroy(e);
            if (p->own_stream) cu().p_cuStreamDestroy(p->own_stream);
            for (int u = 0; u < 7; ++u) if (p->mods[u] && !p->mod_cached[u]) cu().p_cuModuleUnload(p->mods[u]);
            tm.lap("destroy: module/stream");
        }
    }
    delete p;
}

int bsi_attract_schedule(uint64_t horizon, uint32_t *ck_time, uint32_t *ck_slot, uint32_t cap,
                         uint32_t *n_ck, uint32_t *end, uint32_t *first_final) {
    if (horizon == 0 || horizon > ((uint64_t)1 << 30)) return fail(BSI_E_ARG, "horizon out of range");
    Sched s = make_sched(horizon);
    if (n_ck) *n_ck = s.n_ck;
    if (end) *end = s.end;
    if (first_final) *first_final = s.first_final;
    if (s.n_ck > cap) return fail(BSI_E_CAPACITY, "schedule has %u checkpoints", s.n_ck);
    for (uint32_t i = 0; i < s.n_ck; ++i) {
        if (ck_time) ck_time[i] = s.ck_time[i];
        if (ck_slot) ck_slot[i] = s.ck_slot[i];
    }
    return BSI_OK;
}

int bsi_release_cached_memory(void) {
    if (!cu().ok) return BSI_OK;
    std::map<int, std::vector<Parked>> drop;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        drop.swap(g_pool);
        g_pool_bytes = 0;
    }
    for (auto &kv : drop) {
        auto it = g_ctx.find(kv.first);
        if (it == g_ctx.end()) continue;
        if (cu().p_cuCtxPushCurrent(it->second) != CUDA_SUCCESS) continue;
        for (auto &q : kv.second) cu().p_cuMemFree(q.p);
        CUcontext c;
        cu().p_cuCtxPopCurrent(&c);
    }
    return BSI_OK;
}

int bsi_plan_n_problems(const bsi_plan *p, uint64_t *n) {
    if (!p || !n) return fail(BSI_E_ARG, "null pointer");
    *n = p->m.n_problems;
    return BSI_OK;
}

int bsi_plan_key_words(const bsi_plan *p, uint32_t *kw) {
    if (!p || !kw) return fail(BSI_E_ARG, "null pointer");
    *kw = (uint32_t)p->m.kw;
    return BSI_OK;
}

int bsi_plan_set_progress(bsi_plan *p, bsi_progress_fn fn, void *user) {
    if (!p) return fail(BSI_E_ARG, "null plan");
    p->progress = fn;
    p->progress_user = user;
    return BSI_OK;
}

int bsi_plan_set_stream(bsi_plan *p, void *cuda_stream) {
    if (!p) return fail(BSI_E_ARG, "null plan");
    // NULL is the CUDA default (legacy) stream -- what torch.cuda.current_stream().cuda_stream is
    // when no stream context is active; ((void *)-1) goes back to the plan's own stream
    if (cuda_stream == (void *)-1) p->stream = p->own_stream;
    else p->stream = cuda_stream ? (CUstream)cuda_stream : CU_STREAM_LEGACY;
    return BSI_OK;
}

const char *bsi_plan_source(const bsi_plan *p) { return p ? p->source.c_str() : ""; }
const char *bsi_plan_build_log(const bsi_plan *p) { return p ? p->log.c_str() : ""; }

int bsi_plan_cubin(const bsi_plan *p, const void **data, uint64_t *bytes) {
    if (!p || !data || !bytes) return fail(BSI_E_ARG, "null pointer");
    static const std::vector<char> none;
    const auto &cp = p->cubins[p->m.mode == BSI_MODE_ATTRACT ? 1 : 4];
    const std::vector<char> &c = cp ? *cp : none;
    *data = c.data();
    *bytes = c.size();
    return BSI_OK;
}

int bsi_plan_op_counts(const bsi_plan *p, uint32_t *update_lop3, uint32_t *compare_lop3,
                       uint32_t *filter_lop3) {
    if (!p) return fail(BSI_E_ARG, "null plan");
    if (update_lop3) *update_lop3 = (uint32_t)p->info.update_lop3;
    if (compare_lop3) *compare_lop3 = (uint32_t)p->info.compare_lop3;
    if (filter_lop3) *filter_lop3 = (uint32_t)p->info.filter_lop3;
    return BSI_OK;
}

int bsi_plan_update_quiet_lop3(const bsi_plan *p) {
    if (!p) return fail(BSI_E_ARG, "null plan");
    return p->info.update_quiet_lop3;
}

int bsi_plan_launch_config(const bsi_plan *p, uint32_t *n_sms, uint32_t *block_threads,
                           uint32_t *blocks_per_sm) {
    if (!p) return fail(BSI_E_ARG, "null plan");
    if (n_sms) *n_sms = (uint32_t)p->n_sms;
    if (block_threads) *block_threads = (uint32_t)p->block;
    if (blocks_per_sm) *blocks_per_sm = (uint32_t)p->main_blocks_per_sm;
    return BSI_OK;
}

int bsi_compile_only(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                     const bsi_target_desc *target, const bsi_plan_opts *opts,
                     uint64_t *cubin_bytes, char *log, uint64_t log_cap) {
    bsi_plan *p = nullptr;
    int rc = create_plan(network, space, mode, target, opts, -1, &p);
    if (log && log_cap) {
        const std::string &src = rc == BSI_OK ? p->log : g_err;
        snprintf(log, (size_t)log_cap, "%s", src.c_str());
    }
    if (rc == BSI_OK) {
        if (cubin_bytes) { *cubin_bytes = 0; for (auto &c : p->cubins) if (c) *cubin_bytes += c->size(); }
        bsi_plan_destroy(p);
    }
    return rc;
}

// ---- attract ------------------------------------------------------------------------------------
static int attract_collect(bsi_plan *p, bsi_run_stats *stats);

static int attract_enqueue(bsi_plan *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                           uint64_t max_attractor_l, bsi_run_stats *stats, bool async) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (p->pending) return fail(BSI_E_ARG, "a run enqueued by bsi_attract_run_async has not been collected (bsi_attract_run_finish)");
    if (idx_begin > idx_end || idx_end > p->m.n_problems) return fail(BSI_E_ARG, "bad index range");
    bsi_run_stats st{};
    st.n_problems = idx_end - idx_begin;
    const uint64_t lp = p->m.lp_origin;
    if (st.n_problems == 0 || max_t <= lp) {
        // nothing to launch.  max_t <= lp: model.py:201, the loop body never runs; every problem
        // stops at t = lp with no attractor.  The device counters still get the totals (the
        // exchange payload header is written from them).
        st.n_without = st.n_problems;
        st.steps = st.n_problems * lp;
        CtxGuard g0(p);
        Counters h{};
        h.n_without = st.n_without;
        h.steps = st.steps;
        CU_CHECK(cuMemcpyHtoDAsync(p->ctr.p, &h, sizeof h, p->stream));
        CU_CHECK(cuStreamSynchronize(p->stream));
        if (async) {
            p->pending = true;
            p->pending_trivial = true;
            p->pending_stats = st;
        }
        if (stats) *stats = st;
        return BSI_OK;
    }
    const uint64_t T = max_t == BSI_INF ? BSI_INF : max_t - lp;     // horizon relative to lp
    const uint64_t kMaxH = (uint64_t)1 << 30;
    CtxGuard g(p);
    p->events_used = 0;
    const uint64_t launches0 = p->launches;
    Timer tm;
    if ((rc = get_event(p, &tm.e0)) || (rc = get_event(p, &tm.e1))) return rc;

    const uint32_t chunk_log2 = p->opts.chunk_log2 ? p->opts.chunk_log2 : 26;
    if (chunk_log2 < 5 || chunk_log2 > 34) return fail(BSI_E_ARG, "chunk_log2 out of range");
    const uint64_t chunk = (uint64_t)1 << chunk_log2;
    const size_t found_bytes = 24 + 24 * (size_t)p->m.kw;
    const unsigned main_grid = (unsigned)(p->n_sms * p->main_blocks_per_sm);
    // per thread: detection records (3 x 32 words), early events (2 x 32 words) + the checkpoint slots that are not in shared memory
    const size_t scratch_bytes = (size_t)main_grid * p->block * (kScratchWords + (size_t)(3 - p->smem_slots) * p->m.n) * sizeof(uint32_t);
    if ((rc = ensure(p, p->scratch, scratch_bytes))) return rc;

    CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, sizeof(Counters), p->stream));
    CU_CHECK(cuEventRecord(tm.e0, p->stream));

    // the found queue is sized once for the largest chunk: growing it between chunks would hand a
    // buffer that earlier kernels may still be writing to back to the pool
    if ((rc = ensure(p, p->found_q, found_bytes * (std::min(chunk, idx_end - idx_begin) + 64)))) return rc;
    // chunks are aligned to word boundaries (32 problems)
    uint64_t b = idx_begin;
    while (b < idx_end) {
        uint64_t e = std::min(idx_end, ((b >> 5) << 5) + chunk);
        uint64_t n_lanes = e - b;
        uint64_t H = p->opts.first_horizon ? p->opts.first_horizon : 1024;
        H = std::min(H, kMaxH);
        int which = 0;
        bool dense = true;
        uint64_t n_in = n_lanes;
        for (;;) {
            if (H >= T) H = T;
            const int final_round = H == T ? 1 : 0;
            if (!final_round)
                if ((rc = ensure(p, p->retry_q[which], sizeof(uint64_t) * (n_in + 32)))) return rc;
            Sched sched = make_sched(H);
            CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, kPerRoundBytes, p->stream));
            CUevent m0, m1, r0, r1;
            if ((rc = get_event(p, &m0)) || (rc = get_event(p, &m1)) || (rc = get_event(p, &r0)) ||
                (rc = get_event(p, &r1)))
                return rc;
            CU_CHECK(cuEventRecord(m0, p->stream));
            uint64_t stop_t = max_t;      // only used in the final round
            CUdeviceptr retry_out = final_round ? 0 : p->retry_q[which].p;
            if (dense) {
                uint64_t word0 = b >> 5, n_words = ((e + 31) >> 5) - word0;
                // early-word queue: one entry per word at most
                uint64_t early_cap = p->opts.no_early_words ? 0 : n_words;
                if (early_cap &&
                    (rc = ensure(p, p->early_q, ((early_cap + 31) / 32) * 32 * early_entry_bytes(p->m.n))))   // whole blocks of 32 entries
                    return rc;
                CUdeviceptr early_p = early_cap ? p->early_q.p : 0;
                void *args[] = {&word0, &n_words, &b, &e, &sched, (void *)&final_round, &stop_t,
                                &p->found_q.p, &retry_out, &p->scratch.p, &early_p, &early_cap, &p->ctr.p};
                unsigned grid = (unsigned)std::min<uint64_t>(main_grid, (n_words + p->block - 1) / p->block);
                if ((rc = launch(p, p->f_dense, grid, p->block, p->main_smem, args))) return rc;
            } else {
                CUdeviceptr in_q = p->retry_q[which ^ 1].p;
                uint64_t n_words = (n_in + 31) >> 5;
                void *args[] = {&in_q, &n_in, &sched, (void *)&final_round, &stop_t, &p->found_q.p,
                                &retry_out, &p->scratch.p, &p->ctr.p};
                unsigned grid = (unsigned)std::min<uint64_t>(main_grid, (n_words + p->block - 1) / p->block);
                if ((rc = ensure_unit(p, 2))) return rc;
                if ((rc = launch(p, p->f_queue, grid, p->block, p->main_smem, args))) return rc;
            }
            CU_CHECK(cuEventRecord(m1, p->stream));
            tm.main.push_back({m0, m1});
            // resolution of the matched lanes: the kernels read the queue length on the device
            CU_CHECK(cuEventRecord(r0, p->stream));
            {
                CUdeviceptr n_found_p = p->ctr.p + offsetof(Counters, n_found);
                // all three resolution kernels are launched; the device picks the ones that fit
                // the queue length (see bsi_template.cuh), the others return at once -- unless the
                // round is so small that both queues MUST be short: then the throughput-mode
                // kernels are not even launched (nor compiled, if the plan has not needed them yet)
                uint64_t few = p->opts.resolve_few_words ? p->opts.resolve_few_words : 32768;
                const bool surely_short = n_in / 32 + 2 <= few;
                if (!surely_short && ((rc = ensure_unit(p, 3)) || (rc = ensure_unit(p, 6)))) return rc;
                uint64_t t_rel = T, cap_l = max_attractor_l, stop_cap = max_t;
                void *args_m[] = {&p->found_q.p, &n_found_p, &few, &t_rel, &cap_l, &stop_cap, &p->table.p,
                                  &p->table_log2, &p->ctr.p, &p->dict};
                unsigned grid_k = (unsigned)(p->n_sms * p->key_blocks_per_sm);
                unsigned grid_m = (unsigned)(p->n_sms * p->mu_blocks_per_sm);
                unsigned grid_r = (unsigned)(p->n_sms * p->resolve_blocks_per_sm);
                uint64_t early_cap = 0;
                CUdeviceptr early_p = 0;
                if (dense && !p->opts.no_early_words) {
                    early_cap = ((e + 31) >> 5) - (b >> 5);
                    early_p = p->early_q.p;
                    CUdeviceptr n_early_p = p->ctr.p + offsetof(Counters, n_early);
                    void *args_e[] = {&early_p, &early_cap, &n_early_p, &n_found_p, &few, &t_rel, &cap_l,
                                      &stop_cap, &p->table.p, &p->table_log2, &p->ctr.p, &p->dict};
                    if (!surely_short &&
                        (rc = launch(p, p->f_early, (unsigned)(p->n_sms * p->early_blocks_per_sm), p->block, 0, args_e)))
                        return rc;
                }
                void *args_r[] = {&p->found_q.p, &n_found_p, &few, &early_p, &early_cap, &t_rel, &cap_l,
                                  &stop_cap, &p->table.p, &p->table_log2, &p->ctr.p, &p->dict};
                if (!surely_short && (rc = launch(p, p->f_mu, grid_m, p->block, 0, args_m))) return rc;
                if ((rc = launch(p, p->f_resolve, grid_r, p->block, 0, args_r))) return rc;
                if ((rc = launch(p, p->f_finish, grid_k, p->block, 0, args_m))) return rc;
            }
            CU_CHECK(cuEventRecord(r1, p->stream));
            tm.resolve.push_back({r0, r1});
            ++st.rounds;
            if (final_round) break;
            Counters c;
            if ((rc = read_counters(p, &c))) return rc;
            if (c.n_retry == 0) break;
            if (H >= kMaxH) return fail(BSI_E_UNSUPPORTED, "a trajectory did not close within 2^30 steps");
            // next horizon: x4 -- or, when most lanes are still open (chaotic regime), straight to
            // the cap (finite max_t) / x16 (no cap): every round replays its lanes from t = 0
            const bool mostly_open = c.n_retry * 2 > n_in;
            n_in = c.n_retry;
            dense = false;
            which ^= 1;
            if (mostly_open) H = T != BSI_INF ? std::min(T, kMaxH) : std::min(H * 16, kMaxH);
            else H = std::min(H * 4, kMaxH);
        }
        b = e;
        // (with several rounds per chunk the host has waited for this chunk: it is finished)
        if (p->progress) p->progress(b - idx_begin, idx_end - idx_begin, p->progress_user);
    }
    CU_CHECK(cuEventRecord(tm.e1, p->stream));
    p->pending = true;
    p->pending_stats = st;
    p->pending_launches0 = launches0;
    p->pending_e0 = tm.e0;
    p->pending_e1 = tm.e1;
    p->pending_main = tm.main;
    p->pending_resolve = tm.resolve;
    if (async) return BSI_OK;
    return attract_collect(p, stats);
}

// waits for an enqueued run, reads its counters and timings, checks the invariants
static int attract_collect(bsi_plan *p, bsi_run_stats *stats) {
    int rc;
    p->pending = false;
    bsi_run_stats st = p->pending_stats;
    if (p->pending_trivial) {
        p->pending_trivial = false;
        if (stats) *stats = st;
        return BSI_OK;
    }
    Counters c;
    if ((rc = read_counters(p, &c))) return rc;
    CU_CHECK(cuEventSynchronize(p->pending_e1));
    if (c.table_full) return fail(BSI_E_CAPACITY, "attractor table overflow (raise table_log2_slots)");
    if (c.internal_error) return fail(BSI_E_CUDA, "internal invariant failed in bsi_resolve_mu");
    st.n_found = c.n_accepted;
    st.n_without = c.n_without;
    st.steps = c.steps;
    st.device_steps = c.device_steps;
    st.launches = p->launches - p->pending_launches0;
    for (int k = 0; k < 4; ++k) st.reserved[k] = c.dbg[k];      // debug_stats counters
    if ((rc = elapsed(p->pending_e0, p->pending_e1, &st.ms_total))) return rc;
    for (auto &pr : p->pending_main) if ((rc = elapsed(pr.first, pr.second, &st.ms_main))) return rc;
    for (auto &pr : p->pending_resolve) if ((rc = elapsed(pr.first, pr.second, &st.ms_resolve))) return rc;
    if (st.n_found + st.n_without != st.n_problems)
        return fail(BSI_E_CUDA, "internal: %llu found + %llu without != %llu problems",
                    (unsigned long long)st.n_found, (unsigned long long)st.n_without,
                    (unsigned long long)st.n_problems);
    if (stats) *stats = st;
    return BSI_OK;
}

int bsi_attract_run(bsi_plan *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                    uint64_t max_attractor_l, bsi_run_stats *stats) {
    return attract_enqueue(p, idx_begin, idx_end, max_t, max_attractor_l, stats, false);
}

int bsi_attract_run_async(bsi_plan *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                          uint64_t max_attractor_l) {
    return attract_enqueue(p, idx_begin, idx_end, max_t, max_attractor_l, nullptr, true);
}

int bsi_attract_run_finish(bsi_plan *p, bsi_run_stats *stats) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (!p->pending) return fail(BSI_E_ARG, "no run is pending");
    CtxGuard g(p);
    return attract_collect(p, stats);
}

int bsi_attract_table_clear(bsi_plan *p) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    CtxGuard g(p);
    // asynchronous on the plan's stream (every later call is ordered after it)
    CU_CHECK(cuMemsetD8Async(p->table.p, 0, p->table.bytes, p->stream));
    CU_CHECK(cuMemsetD8Async(p->ctr.p + offsetof(Counters, table_full), 0, 8, p->stream));
    // the cycle dictionary goes with the table: a cleared plan has learnt nothing (keep_dict: it stays,
    // the dictionary only holds facts about the network's dynamics)
    if (p->dict.slots && !p->opts.keep_dict) {
        CU_CHECK(cuMemsetD8Async(p->dict_slots.p, 0, p->dict_slots.bytes, p->stream));
        CU_CHECK(cuMemsetD8Async(p->dict_ctr.p, 0, 256, p->stream));
    }
    return BSI_OK;
}

static int export_table(bsi_plan *p, CUdeviceptr out, uint64_t cap, uint64_t *n, CUdeviceptr header,
                        bool sync, bool zero_fill = true) {
    int rc = ensure(p, p->misc, 256);
    if (rc) return rc;
    CU_CHECK(cuMemsetD8Async(p->misc.p, 0, 16, p->stream));     // [0] record count, [1] blocks done
    if (out && cap && zero_fill) CU_CHECK(cuMemsetD8Async(out, 0, cap * sizeof(bsi_attractor_rec), p->stream));
    CUdeviceptr done = p->misc.p + 8;
    void *args[] = {&p->table.p, &p->table_log2, &out, &cap, &p->misc.p, &header, &p->ctr.p, &done};
    if ((rc = launch(p, p->f_export, (unsigned)(p->n_sms * 4), 256, 0, args))) return rc;
    if (!sync) return BSI_OK;
    uint64_t count = 0, full = 0;
    CU_CHECK(cuMemcpyDtoHAsync(&count, p->misc.p, 8, p->stream));
    CU_CHECK(cuMemcpyDtoHAsync(&full, p->ctr.p + offsetof(Counters, table_full), 8, p->stream));
    CU_CHECK(cuStreamSynchronize(p->stream));
    if (n) *n = count;
    // (also reports an overflow of an earlier, unsynchronised bsi_attract_table_merge_device)
    if (full) return fail(BSI_E_CAPACITY, "attractor table overflow (raise table_log2_slots)");
    return BSI_OK;
}

int bsi_attract_table_count(bsi_plan *p, uint64_t *n) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    CtxGuard g(p);
    return export_table(p, 0, 0, n, 0, true);
}

int bsi_attract_table_read(bsi_plan *p, bsi_attractor_rec *out, uint64_t cap, uint64_t *n) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    CtxGuard g(p);
    if ((rc = ensure(p, p->misc2, std::max<uint64_t>(cap, 1) * sizeof(bsi_attractor_rec)))) return rc;
    uint64_t count = 0;
    if ((rc = export_table(p, p->misc2.p, cap, &count, 0, true))) return rc;
    if (n) *n = count;
    if (count > cap) return fail(BSI_E_CAPACITY, "attractor table holds %llu records, buffer %llu",
                                 (unsigned long long)count, (unsigned long long)cap);
    if (count && out) {
        CU_CHECK(cuMemcpyDtoHAsync(out, p->misc2.p, count * sizeof(bsi_attractor_rec), p->stream));
        CU_CHECK(cuStreamSynchronize(p->stream));
    }
    return BSI_OK;
}

int bsi_attract_table_export_device(bsi_plan *p, void *dev_out, uint64_t cap, uint64_t *n) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (!dev_out) return fail(BSI_E_ARG, "null device buffer");
    CtxGuard g(p);
    uint64_t count = 0;
    if ((rc = export_table(p, (CUdeviceptr)dev_out, cap, &count, 0, true))) return rc;
    if (n) *n = count;
    if (count > cap) return fail(BSI_E_CAPACITY, "attractor table holds %llu records, buffer %llu",
                                 (unsigned long long)count, (unsigned long long)cap);
    return BSI_OK;
}

int bsi_attract_table_export_payload(bsi_plan *p, void *dev_out, uint64_t cap) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (!dev_out || cap < 1) return fail(BSI_E_ARG, "need a device buffer of at least one record");
    CtxGuard g(p);
    CUdeviceptr base = (CUdeviceptr)dev_out;
    return export_table(p, base + sizeof(bsi_attractor_rec), cap - 1, nullptr, base, false);
}

int bsi_attract_table_merge_device(bsi_plan *p, const void *dev_records, uint64_t n) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (n == 0) return BSI_OK;
    if (!dev_records) return fail(BSI_E_ARG, "null device buffer");
    CtxGuard g(p);
    CUdeviceptr recs = (CUdeviceptr)dev_records;
    void *args[] = {&p->table.p, &p->table_log2, &recs, &n, &p->ctr.p};
    unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)p->n_sms * 4);
    // asynchronous on the plan's stream: a table overflow is reported by the next
    // bsi_attract_table_count / _read / _export_device
    return launch(p, p->f_merge, grid, 256, 0, args);
}

// ---- trajectories (simulate, orbits) --------------------------------------------------------------
static int run_trajectories(bsi_plan *p, uint64_t idx_begin, const uint64_t *indices,
                            const uint64_t *codes, uint64_t n, uint64_t n_steps, uint64_t *out,
                            bsi_run_stats *stats,
                            const std::function<int(uint64_t, uint64_t)> &after_chunk = nullptr) {
    int rc;
    const uint64_t kw = (uint64_t)p->m.kw;
    if (n_steps >= ((uint64_t)1 << 40)) return fail(BSI_E_ARG, "max_t too large for an explicit trajectory");
    const uint64_t per_lane = (n_steps + 1) * kw * sizeof(uint64_t);
    uint64_t lanes_per_call = std::max<uint64_t>(32, (((uint64_t)1 << 30) / per_lane) & ~(uint64_t)31);
    bsi_run_stats st{};
    st.n_problems = n;
    p->events_used = 0;
    const uint64_t launches0 = p->launches;
    CUevent e0, e1;
    if ((rc = get_event(p, &e0)) || (rc = get_event(p, &e1))) return rc;
    CU_CHECK(cuEventRecord(e0, p->stream));
    for (uint64_t done = 0; done < n;) {
        uint64_t m = std::min(lanes_per_call, n - done);
        if ((rc = ensure(p, p->misc, m * per_lane))) return rc;
        CUdeviceptr d_idx = 0, d_codes = 0;
        if (indices) {
            if ((rc = ensure(p, p->misc2, m * 8))) return rc;
            CU_CHECK(cuMemcpyHtoDAsync(p->misc2.p, indices + done, m * 8, p->stream));
            d_idx = p->misc2.p;
        }
        if (codes) {
            if ((rc = ensure(p, p->misc3, m * 8 * kw))) return rc;
            CU_CHECK(cuMemcpyHtoDAsync(p->misc3.p, codes + done * kw, m * 8 * kw, p->stream));
            d_codes = p->misc3.p;
        }
        uint64_t first = idx_begin + done;
        void *args[] = {&first, &d_idx, &d_codes, &m, &n_steps, &p->misc.p};
        uint64_t n_words = (m + 31) >> 5;
        if ((rc = launch(p, p->f_traj, (unsigned)((n_words + p->block - 1) / p->block), p->block, 0, args))) return rc;
        // (after_chunk: the chunk's states [m][n_steps + 1][kw] are in p->misc, on the device)
        if (after_chunk && (rc = after_chunk(done, m))) return rc;
        if (out) CU_CHECK(cuMemcpyDtoHAsync(out + done * (n_steps + 1) * kw, p->misc.p, m * per_lane, p->stream));
        CU_CHECK(cuStreamSynchronize(p->stream));
        done += m;
    }
    CU_CHECK(cuEventRecord(e1, p->stream));
    CU_CHECK(cuEventSynchronize(e1));
    if ((rc = elapsed(e0, e1, &st.ms_total))) return rc;
    st.ms_main = st.ms_total;
    st.steps = n * n_steps;
    st.device_steps = ((n + 31) >> 5) * n_steps;
    st.launches = p->launches - launches0;
    st.n_found = n;
    if (stats) *stats = st;
    return BSI_OK;
}

int bsi_simulate_run(bsi_plan *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                     uint64_t *states_out, bsi_run_stats *stats) {
    int rc = need_device(p, -1);
    if (rc) return rc;
    if (idx_begin > idx_end || idx_end > p->m.n_problems) return fail(BSI_E_ARG, "bad index range");
    if (!states_out && idx_end > idx_begin) return fail(BSI_E_ARG, "null output buffer");
    if (max_t < p->m.lp_max) return fail(BSI_E_ARG, "simulate: max_t is smaller than the last perturbation time");
    CtxGuard g(p);
    return run_trajectories(p, idx_begin, nullptr, nullptr, idx_end - idx_begin, max_t, states_out, stats);
}

int bsi_simulate_list(bsi_plan *p, const uint64_t *indices, uint64_t n, uint64_t max_t,
                      uint64_t *states_out, bsi_run_stats *stats) {
    int rc = need_device(p, -1);
    if (rc) return rc;
    if (n && (!indices || !states_out)) return fail(BSI_E_ARG, "null buffer");
    for (uint64_t i = 0; i < n; ++i)
        if (indices[i] >= p->m.n_problems) return fail(BSI_E_ARG, "problem index out of range");
    CtxGuard g(p);
    return run_trajectories(p, 0, indices, nullptr, n, max_t, states_out, stats);
}

int bsi_attractor_states(bsi_plan *p, const uint64_t *key, uint64_t lambda, uint64_t *out) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (!key || !out || lambda == 0) return fail(BSI_E_ARG, "bad arguments");
    CtxGuard g(p);
    return run_trajectories(p, 0, nullptr, key, 1, lambda - 1, out, nullptr);
}

// Orbits the resolution kernels stored in the cycle dictionary, copied to the host: for each
// stored attractor (key state = first state of its orbit) the offset and length of its orbit.
static int read_dictionary(bsi_plan *p, std::vector<uint64_t> *orbit, std::map<std::vector<uint64_t>, std::pair<uint64_t, uint64_t>> *by_key) {
    orbit->clear();
    by_key->clear();
    if (!p->dict.slots) return BSI_OK;
    uint64_t cnt[2] = {0, 0};
    CU_CHECK(cuMemcpyDtoHAsync(cnt, p->dict_ctr.p, 16, p->stream));
    CU_CHECK(cuStreamSynchronize(p->stream));
    const uint64_t n_attr = std::min<uint64_t>(cnt[1], p->dict.cap_attr);
    if (n_attr == 0) return BSI_OK;
    const uint64_t kw = (uint64_t)p->m.kw;
    std::vector<uint32_t> dir(2 * n_attr);
    CU_CHECK(cuMemcpyDtoHAsync(dir.data(), p->dict_dir.p, 8 * n_attr, p->stream));
    CU_CHECK(cuStreamSynchronize(p->stream));
    uint64_t hi = 0;
    for (uint64_t a = 0; a < n_attr; ++a) hi = std::max<uint64_t>(hi, (uint64_t)dir[2 * a] + dir[2 * a + 1]);
    hi = std::min<uint64_t>(hi, p->dict.cap_states);
    orbit->resize(hi * kw);
    CU_CHECK(cuMemcpyDtoHAsync(orbit->data(), p->dict_orbit.p, 8 * hi * kw, p->stream));
    CU_CHECK(cuStreamSynchronize(p->stream));
    for (uint64_t a = 0; a < n_attr; ++a) {
        const uint64_t base = dir[2 * a], lam = dir[2 * a + 1];
        if (base + lam > hi) continue;
        std::vector<uint64_t> key(orbit->begin() + base * kw, orbit->begin() + (base + 1) * kw);
        (*by_key)[key] = {base, lam};
    }
    return BSI_OK;
}

int bsi_attractor_states_batch(bsi_plan *p, const uint64_t *keys, uint64_t n, uint64_t n_states,
                               uint64_t *out) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (n == 0) return BSI_OK;
    if (!keys || !out || n_states == 0) return fail(BSI_E_ARG, "bad arguments");
    CtxGuard g(p);
    // orbits the run already stored (cycle dictionary) are copied; only the others are regenerated
    const uint64_t kw = (uint64_t)p->m.kw;
    std::vector<uint64_t> orbit;
    std::map<std::vector<uint64_t>, std::pair<uint64_t, uint64_t>> by_key;
    if ((rc = read_dictionary(p, &orbit, &by_key))) return rc;
    std::vector<uint64_t> rest_keys, rest_pos;
    for (uint64_t i = 0; i < n; ++i) {
        auto it = by_key.empty() ? by_key.end()
                                 : by_key.find(std::vector<uint64_t>(keys + i * kw, keys + (i + 1) * kw));
        if (it == by_key.end()) {
            rest_pos.push_back(i);
            rest_keys.insert(rest_keys.end(), keys + i * kw, keys + (i + 1) * kw);
            continue;
        }
        const uint64_t base = it->second.first, lam = it->second.second;
        for (uint64_t t = 0; t < n_states; ++t)
            memcpy(out + (i * n_states + t) * kw, orbit.data() + (base + t % lam) * kw, 8 * kw);
    }
    if (rest_pos.empty()) return BSI_OK;
    if (rest_pos.size() == n) return run_trajectories(p, 0, nullptr, keys, n, n_states - 1, out, nullptr);
    std::vector<uint64_t> tmp(rest_pos.size() * n_states * kw);
    if ((rc = run_trajectories(p, 0, nullptr, rest_keys.data(), rest_pos.size(), n_states - 1, tmp.data(), nullptr)))
        return rc;
    for (size_t r = 0; r < rest_pos.size(); ++r)
        memcpy(out + rest_pos[r] * n_states * kw, tmp.data() + r * n_states * kw, 8 * n_states * kw);
    return BSI_OK;
}

int bsi_attractor_orbits(bsi_plan *p, const uint64_t *keys, const uint64_t *lams, uint64_t n,
                         uint64_t *states, uint64_t *counts) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (n == 0) return BSI_OK;
    if (!keys || !lams || !states) return fail(BSI_E_ARG, "null pointer");
    CtxGuard g(p);
    const uint64_t kw = (uint64_t)p->m.kw, nn = (uint64_t)p->m.n;
    std::vector<uint64_t> offset(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (lams[i] == 0) return fail(BSI_E_ARG, "attractor length 0");
        offset[i + 1] = offset[i] + lams[i];
    }
    std::vector<uint64_t> bases(n, ~0ull), host_counts;
    if (p->dict.slots) {
        // one pass on the device: find or store every orbit (bases), count node activity where the orbits
        // lie -- two launches, no host round trip between them
        if ((rc = ensure(p, p->misc4, 8 * n * (kw + 2))) || (counts && (rc = ensure(p, p->misc5, 8 * n * nn)))) return rc;
        CUdeviceptr d_keys = p->misc4.p, d_lams = d_keys + 8 * n * kw, d_bases = d_lams + 8 * n;
        CU_CHECK(cuMemcpyHtoDAsync(d_keys, keys, 8 * n * kw, p->stream));
        CU_CHECK(cuMemcpyHtoDAsync(d_lams, lams, 8 * n, p->stream));
        uint64_t nk = n;
        void *args[] = {&d_keys, &d_lams, &nk, &d_bases, &p->dict};
        // (one thread per block: the threads of a warp would walk their orbits one after the other)
        if ((rc = launch(p, p->f_learn_keys, (unsigned)n, 32, 0, args))) return rc;
        if (counts) {
            void *cargs[] = {&p->dict_orbit.p, &d_bases, &d_lams, &p->misc5.p};
            if ((rc = launch(p, p->f_counts, (unsigned)n, 128, 0, cargs))) return rc;
            host_counts.resize(n * nn);
            CU_CHECK(cuMemcpyDtoHAsync(host_counts.data(), p->misc5.p, 8 * n * nn, p->stream));
        }
        CU_CHECK(cuMemcpyDtoHAsync(bases.data(), d_bases, 8 * n, p->stream));
        CU_CHECK(cuStreamSynchronize(p->stream));
        // the stored orbits: one copy of the range that holds them
        uint64_t lo = ~0ull, hi = 0;
        for (uint64_t i = 0; i < n; ++i)
            if (bases[i] != ~0ull) { lo = std::min(lo, bases[i]); hi = std::max(hi, bases[i] + lams[i]); }
        if (hi > lo) {
            std::vector<uint64_t> orbit((hi - lo) * kw);
            CU_CHECK(cuMemcpyDtoHAsync(orbit.data(), p->dict_orbit.p + 8 * lo * kw, 8 * (hi - lo) * kw, p->stream));
            CU_CHECK(cuStreamSynchronize(p->stream));
            for (uint64_t i = 0; i < n; ++i)
                if (bases[i] != ~0ull)
                    memcpy(states + offset[i] * kw, orbit.data() + (bases[i] - lo) * kw, 8 * kw * lams[i]);
        }
    }
    // fixed points are their own orbit; what the dictionary could not take (storage exhausted, dictionary
    // off) is regenerated by the trajectory kernel and counted with bsi_attractor_activity's fallback
    std::vector<uint64_t> rest;
    for (uint64_t i = 0; i < n; ++i) {
        if (bases[i] != ~0ull) {
            if (counts) memcpy(counts + i * nn, host_counts.data() + i * nn, 8 * nn);
        } else if (lams[i] == 1) {
            memcpy(states + offset[i] * kw, keys + i * kw, 8 * kw);
            if (counts) for (uint64_t v = 0; v < nn; ++v) counts[i * nn + v] = (keys[i * kw + (v >> 6)] >> (v & 63)) & 1;
        } else {
            rest.push_back(i);
        }
    }
    for (uint64_t i : rest) {
        if ((rc = run_trajectories(p, 0, nullptr, keys + i * kw, 1, lams[i] - 1, states + offset[i] * kw, nullptr)))
            return rc;
        if (counts && (rc = bsi_attractor_activity(p, keys + i * kw, lams + i, 1, counts + i * nn))) return rc;
    }
    return BSI_OK;
}

int bsi_attractor_activity(bsi_plan *p, const uint64_t *keys, const uint64_t *lams, uint64_t n,
                           uint64_t *counts) {
    int rc = need_device(p, BSI_MODE_ATTRACT);
    if (rc) return rc;
    if (n == 0) return BSI_OK;
    if (!keys || !lams || !counts) return fail(BSI_E_ARG, "null pointer");
    CtxGuard g(p);
    const uint64_t kw = (uint64_t)p->m.kw, nn = (uint64_t)p->m.n;
    std::vector<uint64_t> orbit;
    std::map<std::vector<uint64_t>, std::pair<uint64_t, uint64_t>> by_key;
    if ((rc = read_dictionary(p, &orbit, &by_key))) return rc;
    // one launch over the orbits stored by the run; the others are regenerated and counted chunk by chunk
    std::vector<uint64_t> start, lam, pos, rest_keys, rest_pos;
    for (uint64_t i = 0; i < n; ++i) {
        if (lams[i] == 0) return fail(BSI_E_ARG, "attractor length 0");
        auto it = by_key.empty() ? by_key.end()
                                 : by_key.find(std::vector<uint64_t>(keys + i * kw, keys + (i + 1) * kw));
        if (it != by_key.end() && it->second.second == lams[i]) {
            start.push_back(it->second.first);
            lam.push_back(lams[i]);
            pos.push_back(i);
        } else {
            rest_pos.push_back(i);
            rest_keys.insert(rest_keys.end(), keys + i * kw, keys + (i + 1) * kw);
        }
    }
    auto count_on_device = [&](CUdeviceptr states, const std::vector<uint64_t> &st, const std::vector<uint64_t> &lm,
                               const uint64_t *where) -> int {
        const uint64_t m = st.size();
        int r;
        if ((r = ensure(p, p->misc4, 16 * m)) || (r = ensure(p, p->misc5, 8 * m * nn))) return r;
        CU_CHECK(cuMemcpyHtoDAsync(p->misc4.p, st.data(), 8 * m, p->stream));
        CU_CHECK(cuMemcpyHtoDAsync(p->misc4.p + 8 * m, lm.data(), 8 * m, p->stream));
        CUdeviceptr d_start = p->misc4.p, d_lam = p->misc4.p + 8 * m;
        void *args[] = {&states, &d_start, &d_lam, &p->misc5.p};
        if ((r = launch(p, p->f_counts, (unsigned)m, 128, 0, args))) return r;
        std::vector<uint64_t> host(m * nn);
        CU_CHECK(cuMemcpyDtoHAsync(host.data(), p->misc5.p, 8 * m * nn, p->stream));
        CU_CHECK(cuStreamSynchronize(p->stream));
        for (uint64_t j = 0; j < m; ++j) memcpy(counts + where[j] * nn, host.data() + j * nn, 8 * nn);
        return BSI_OK;
    };
    if (!pos.empty() && (rc = count_on_device(p->dict_orbit.p, start, lam, pos.data()))) return rc;
    if (rest_pos.empty()) return BSI_OK;
    // regenerate the rest, one launch per power-of-two class of lengths (a long cycle must not inflate
    // the buffers of many short ones)
    std::map<int, std::vector<uint64_t>> classes;
    for (uint64_t r = 0; r < rest_pos.size(); ++r) {
        int cls = 0;
        while ((1ull << cls) < lams[rest_pos[r]]) ++cls;
        classes[cls].push_back(r);
    }
    for (auto &kv : classes) {
        const std::vector<uint64_t> &members = kv.second;
        uint64_t n_states = 0;
        std::vector<uint64_t> ck;
        for (uint64_t r : members) {
            n_states = std::max(n_states, lams[rest_pos[r]]);
            ck.insert(ck.end(), rest_keys.begin() + r * kw, rest_keys.begin() + (r + 1) * kw);
        }
        rc = run_trajectories(p, 0, nullptr, ck.data(), members.size(), n_states - 1, nullptr, nullptr,
                              [&](uint64_t done, uint64_t m) -> int {
                                  std::vector<uint64_t> st(m), lm(m), where(m);
                                  for (uint64_t j = 0; j < m; ++j) {
                                      st[j] = j * n_states;
                                      lm[j] = lams[rest_pos[members[done + j]]];
                                      where[j] = rest_pos[members[done + j]];
                                  }
                                  return count_on_device(p->misc.p, st, lm, where.data());
                              });
        if (rc) return rc;
    }
    return BSI_OK;
}

// ---- target ---------------------------------------------------------------------------------------
int bsi_target_run(bsi_plan *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                   uint64_t *hit_index, uint64_t *hit_t, uint64_t cap, uint64_t *n_hits,
                   bsi_run_stats *stats) {
    int rc = need_device(p, BSI_MODE_TARGET);
    if (rc) return rc;
    if (idx_begin > idx_end || idx_end > p->m.n_problems) return fail(BSI_E_ARG, "bad index range");
    if (cap && (!hit_index || !hit_t)) return fail(BSI_E_ARG, "null hit buffers");
    // (hit times and last-perturbation times travel as 32-bit words on the device)
    if (max_t != BSI_INF && max_t >= ((uint64_t)1 << 32)) return fail(BSI_E_ARG, "target: a finite max_t must be below 2^32");
    if (p->m.lp_max >= ((uint64_t)1 << 32)) return fail(BSI_E_UNSUPPORTED, "target: perturbation times of 2^32 and up are not supported");
    CtxGuard g(p);
    bsi_run_stats st{};
    st.n_problems = idx_end - idx_begin;
    p->events_used = 0;
    const uint64_t launches0 = p->launches;
    CUevent e0, e1;
    if ((rc = get_event(p, &e0)) || (rc = get_event(p, &e1))) return rc;
    CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, sizeof(Counters), p->stream));
    CU_CHECK(cuEventRecord(e0, p->stream));
    const uint32_t chunk_log2 = p->opts.chunk_log2 ? p->opts.chunk_log2 : 26;
    const uint64_t chunk = (uint64_t)1 << chunk_log2;
    const uint64_t kMaxH = (uint64_t)1 << 30;
    uint64_t written = 0, total_hits = 0;
    std::vector<std::pair<uint64_t, uint64_t>> hits;
    std::vector<uint64_t> hi, ht;
    auto fetch_hits = [&](uint64_t n) -> int {
        hi.resize(n);
        ht.resize(n);
        CU_CHECK(cuMemcpyDtoHAsync(hi.data(), p->misc.p, 8 * n, p->stream));
        CU_CHECK(cuMemcpyDtoHAsync(ht.data(), p->misc2.p, 8 * n, p->stream));
        CU_CHECK(cuStreamSynchronize(p->stream));
        for (uint64_t i = 0; i < n; ++i) hits.push_back({hi[i], ht[i]});
        return BSI_OK;
    };
    // hits of one list in index order (the enumeration order the reference reports in, simulate.py:149-178).
    // Only the first `cap - written` are handed out: when that is a small part of the list they are
    // SELECTED (partial sort, linear in the list), otherwise the list is radix-sorted on the index
    // (11-bit digits over the bits the indices of this run can have) -- a comparison sort of some
    // 10^5 hits costs more host time than the kernels of a 2^26-state chunk take.
    auto emit_sorted = [&]() {
        total_hits += hits.size();
        const uint64_t want = std::min<uint64_t>(cap - std::min(written, cap), hits.size());
        if (want == 0) return;
        if (want < hits.size() / 8) {
            std::partial_sort(hits.begin(), hits.begin() + want, hits.end());
        } else if (hits.size() < 4096) {
            std::sort(hits.begin(), hits.end());
        } else {
            int bits = 1;
            while (bits < 64 && (idx_end >> bits)) ++bits;
            std::vector<std::pair<uint64_t, uint64_t>> tmp(hits.size());
            for (int shift = 0; shift < bits; shift += 11) {
                size_t count[2049] = {0};
                for (auto &h : hits) ++count[((h.first >> shift) & 2047u) + 1];
                for (int d = 0; d < 2048; ++d) count[d + 1] += count[d];
                for (auto &h : hits) tmp[count[(h.first >> shift) & 2047u]++] = h;
                hits.swap(tmp);
            }
        }
        for (uint64_t i = 0; i < want; ++i) {
            hit_index[written] = hits[i].first;
            hit_t[written] = hits[i].second;
            ++written;
        }
    };
    // A finite max_t needs one round per chunk and no decision of the host between them: all chunks are
    // enqueued back to back, their hits appended to ONE list (the kernel's counter runs on), and the
    // host reads it once -- the GPU does not wait for the host between chunks.  The list holds 2^24
    // hits; a job with more falls back to the chunk-by-chunk form below.
    bool done = false;
    const uint64_t kOnePassHits = p->opts.target_hit_list ? p->opts.target_hit_list : (uint64_t)1 << 24;
    if (max_t != BSI_INF && p->opts.target_hit_list != 1) {
        const uint64_t total = idx_end - idx_begin;
        uint64_t hit_cap = std::min(total, kOnePassHits);
        if ((rc = ensure(p, p->misc, 8 * std::max<uint64_t>(hit_cap, 1))) ||
            (rc = ensure(p, p->misc2, 8 * std::max<uint64_t>(hit_cap, 1)))) return rc;
        uint64_t T = max_t;
        int check_closure = 0, n_chunks = 0;
        CUdeviceptr none = 0;
        // the list's length after every chunk stays on the device (misc3[k], copied there in stream order):
        // chunks run one after the other, so the hits of chunk k lie before those of chunk k + 1 in the
        // list, and the first `cap` hits in index order are all within the prefix that ends with the
        // first chunk whose running count reaches `cap` -- only that prefix is fetched and ordered
        const uint64_t total_chunks = (total + chunk - 1) / chunk + 1;
        if ((rc = ensure(p, p->misc3, 8 * total_chunks))) return rc;
        for (uint64_t b = idx_begin; b < idx_end;) {
            uint64_t e = std::min(idx_end, b + chunk);
            uint64_t n_in = e - b;
            CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, kPerRoundBytes, p->stream));
            void *args[] = {&b, &none, &n_in, &T, &check_closure, &p->misc.p, &p->misc2.p, &hit_cap,
                            &none, &none, &p->ctr.p};
            uint64_t n_words = (n_in + 31) >> 5;
            if ((rc = launch(p, p->f_target, (unsigned)((n_words + p->block - 1) / p->block), p->block, 0, args))) return rc;
            CU_CHECK(cuMemcpyDtoDAsync(p->misc3.p + 8 * (uint64_t)st.rounds, p->ctr.p + offsetof(Counters, n_hits), 8, p->stream));
            ++st.rounds;
            b = e;
            // (a progress hook wants to hear from a long run: wait for the device every 16 chunks)
            if (p->progress && (++n_chunks % 16 == 0 || b == idx_end)) {
                CU_CHECK(cuStreamSynchronize(p->stream));
                p->progress(b - idx_begin, idx_end - idx_begin, p->progress_user);
            }
        }
        CU_CHECK(cuEventRecord(e1, p->stream));
        Counters c;
        if ((rc = read_counters(p, &c))) return rc;
        if (c.n_hits <= hit_cap) {
            uint64_t fetch = c.n_hits;
            if (cap < c.n_hits && st.rounds > 1) {
                std::vector<uint64_t> after(st.rounds);
                CU_CHECK(cuMemcpyDtoHAsync(after.data(), p->misc3.p, 8 * st.rounds, p->stream));
                CU_CHECK(cuStreamSynchronize(p->stream));
                for (uint64_t k = 0; k < st.rounds; ++k)
                    if (after[k] >= cap) { fetch = std::min(fetch, after[k]); break; }
            }
            if (fetch && (rc = fetch_hits(fetch))) return rc;
            emit_sorted();
            total_hits = c.n_hits;              // (the list was read in part only)
            done = true;
        } else {
            // (start over: the counters of the attempt are dropped)
            CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, sizeof(Counters), p->stream));
            CU_CHECK(cuEventRecord(e0, p->stream));
            st.rounds = 0;
        }
    }
    for (uint64_t b = idx_begin; !done && b < idx_end;) {
        uint64_t e = std::min(idx_end, b + chunk);
        uint64_t n_lanes = e - b;
        if ((rc = ensure(p, p->misc, 8 * n_lanes)) || (rc = ensure(p, p->misc2, 8 * n_lanes))) return rc;
        hits.clear();
        uint64_t T = max_t == BSI_INF ? std::max<uint64_t>(1024, 2 * p->m.lp_max) : max_t;
        uint64_t n_in = n_lanes;
        int which = 0;
        bool dense = true;
        for (;;) {
            int check_closure = (max_t == BSI_INF || T < max_t) ? 1 : 0;
            if (check_closure) {
                if ((rc = ensure(p, p->retry_q[which], 8 * n_in)) || (rc = ensure(p, p->found_q, 8 * n_in))) return rc;
            }
            // per round: hits of this round start at 0
            CU_CHECK(cuMemsetD8Async(p->ctr.p, 0, kPerRoundBytes, p->stream));
            CU_CHECK(cuMemsetD8Async(p->ctr.p + offsetof(Counters, n_hits), 0, 8, p->stream));
            CUdeviceptr in_q = dense ? 0 : p->retry_q[which ^ 1].p;
            CUdeviceptr retry_out = check_closure ? p->retry_q[which].p : 0;
            CUdeviceptr closed_out = check_closure ? p->found_q.p : 0;
            uint64_t hit_cap = n_lanes;
            void *args[] = {&b, &in_q, &n_in, &T, &check_closure, &p->misc.p, &p->misc2.p, &hit_cap,
                            &retry_out, &closed_out, &p->ctr.p};
            uint64_t n_words = (n_in + 31) >> 5;
            if ((rc = launch(p, p->f_target, (unsigned)((n_words + p->block - 1) / p->block), p->block, 0, args))) return rc;
            ++st.rounds;
            Counters c;
            if ((rc = read_counters(p, &c))) return rc;
            if (c.n_hits && (rc = fetch_hits(c.n_hits))) return rc;
            st.n_without += c.n_found;     // closed without a hit
            if (!check_closure || c.n_retry == 0) break;
            if (T >= kMaxH) return fail(BSI_E_UNSUPPORTED, "a trajectory did not close within 2^30 steps");
            n_in = c.n_retry;
            dense = false;
            which ^= 1;
            T = std::min(T * 4, kMaxH);
            if (max_t != BSI_INF && T > max_t) T = max_t;
        }
        emit_sorted();
        b = e;
        if (p->progress) p->progress(b - idx_begin, idx_end - idx_begin, p->progress_user);
    }
    if (!done) CU_CHECK(cuEventRecord(e1, p->stream));
    Counters c;
    if ((rc = read_counters(p, &c))) return rc;
    CU_CHECK(cuEventSynchronize(e1));
    if ((rc = elapsed(e0, e1, &st.ms_total))) return rc;
    st.ms_main = st.ms_total;
    st.n_found = total_hits;
    st.n_without += c.n_without;
    st.steps = c.steps;
    st.device_steps = c.device_steps;
    st.launches = p->launches - launches0;
    if (n_hits) *n_hits = written;
    if (stats) *stats = st;
    return BSI_OK;
}

// ---- groups: one plan per device, one host process ------------------------------------------------
// The one-process replacement of the reference's MPI task farm (mpi.py:193-340: master + workers, one
// process per core): every device of the group takes a contiguous, word-aligned slice of the index
// range (one host thread per device enqueues it), and the collection stage (mpi.py:296-332) is the
// peers' export kernels writing their compacted tables straight into device 0's memory over NVLink
// (peer access; a peer copy where that is not available), followed by ONE merge launch on device 0.
}  // extern "C"

struct bsi_group {
    std::vector<bsi_plan *> plans;
    std::vector<char> direct;           // device d writes into device 0's memory directly (peer access)
    std::vector<DevBuf> staging;        // per device, on that device: payload when there is no peer access
    std::vector<CUevent> done;          // per device: payload has reached device 0
    DevBuf gather;                      // on device 0: headers [n - 1] + records [n - 1][cap]
    uint64_t cap = 1024;                // records per payload
    int mode = BSI_MODE_ATTRACT;
};

namespace {

void shard(uint64_t b, uint64_t e, int d, int n, uint64_t *lo, uint64_t *hi) {
    // contiguous slices of [b, e), cut at word boundaries (multiples of 32 problems)
    const uint64_t w0 = b >> 5, w1 = (e + 31) >> 5, words = w1 - w0;
    uint64_t a = (w0 + words * (uint64_t)d / (uint64_t)n) << 5, z = (w0 + words * (uint64_t)(d + 1) / (uint64_t)n) << 5;
    *lo = std::min(std::max(a, b), e);
    *hi = d == n - 1 ? e : std::min(std::max(z, b), e);
}

// runs fn(d) for every device of the group in its own host thread; first error wins
template <typename F>
int for_each_device(bsi_group *g, F fn) {
    const int n = (int)g->plans.size();
    std::vector<int> rcs(n, BSI_OK);
    std::vector<std::string> errs(n);
    auto body = [&](int d) {
        rcs[d] = fn(d);
        if (rcs[d] != BSI_OK) errs[d] = g_err;
    };
    if (n == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int d = 0; d < n; ++d) th.emplace_back(body, d);
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < n; ++d)
        if (rcs[d] != BSI_OK) { g_err = "device " + std::to_string(g->plans[d]->device) + ": " + errs[d]; return rcs[d]; }
    return BSI_OK;
}

int group_gather_buffers(bsi_group *g) {
    const int n = (int)g->plans.size();
    if (n < 2) return BSI_OK;
    bsi_plan *p0 = g->plans[0];
    const size_t rec = sizeof(bsi_attractor_rec);
    int rc;
    {
        CtxGuard c0(p0);
        if ((rc = ensure(p0, g->gather, (size_t)(n - 1) * (1 + g->cap) * rec))) return rc;
    }
    for (int d = 1; d < n; ++d) {
        if (g->direct[d]) continue;
        CtxGuard cd(g->plans[d]);
        if ((rc = ensure(g->plans[d], g->staging[d], (1 + g->cap) * rec))) return rc;
    }
    return BSI_OK;
}

int regrow_table(bsi_plan *p, uint32_t log2) {
    CtxGuard c(p);
    CU_CHECK(cuStreamSynchronize(p->stream));
    p->table_log2 = log2;
    const size_t slot_bytes = (size_t)8 * p->m.kw + 32;
    int rc = ensure(p, p->table, slot_bytes << p->table_log2);
    if (rc) return rc;
    CU_CHECK(cuMemsetD8Async(p->table.p, 0, p->table.bytes, p->stream));
    return BSI_OK;
}

}  // namespace

extern "C" {

int bsi_group_create(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                     const bsi_target_desc *target, const bsi_plan_opts *opts, const int *devices,
                     int n_devices, bsi_group **out) {
    if (!out) return fail(BSI_E_ARG, "null output pointer");
    *out = nullptr;
    if (!cu().ok) return fail(BSI_E_CUDA, "%s", cu().why.c_str());
    int count = 0;
    CU_CHECK(cuDeviceGetCount(&count));
    std::vector<int> devs;
    if (n_devices <= 0) {
        for (int d = 0; d < count; ++d) devs.push_back(d);
    } else {
        for (int d = 0; d < n_devices; ++d) devs.push_back(devices ? devices[d] : d);
    }
    if (devs.empty()) return fail(BSI_E_CUDA, "no CUDA device");
    for (size_t a = 0; a < devs.size(); ++a)
        for (size_t b = a + 1; b < devs.size(); ++b)
            if (devs[a] == devs[b]) return fail(BSI_E_ARG, "device %d listed twice", devs[a]);
    bsi_group *g = new bsi_group();
    g->mode = mode;
    const int n = (int)devs.size();
    g->plans.assign(n, nullptr);
    g->direct.assign(n, 0);
    g->staging.resize(n);
    g->done.assign(n, nullptr);
    // the first plan specialises the kernels (NVRTC; cached per process), the others only load them
    int rc = create_plan(network, space, mode, target, opts, devs[0], &g->plans[0]);
    if (rc == BSI_OK && n > 1) {
        std::vector<int> rcs(n, BSI_OK);
        std::vector<std::string> errs(n);
        std::vector<std::thread> th;
        for (int d = 1; d < n; ++d)
            th.emplace_back([&, d] {
                rcs[d] = create_plan(network, space, mode, target, opts, devs[d], &g->plans[d]);
                if (rcs[d] != BSI_OK) errs[d] = g_err;
            });
        for (auto &t : th) t.join();
        for (int d = 1; d < n && rc == BSI_OK; ++d)
            if (rcs[d] != BSI_OK) { rc = rcs[d]; g_err = errs[d]; }
    }
    if (rc == BSI_OK && n > 1 && mode == BSI_MODE_ATTRACT) {
        bsi_plan *p0 = g->plans[0];
        for (int d = 1; d < n; ++d) {
            bsi_plan *pd = g->plans[d];
            CtxGuard cd(pd);
            int can = 0;
            if (cu().p_cuDeviceCanAccessPeer(&can, pd->dev, p0->dev) == CUDA_SUCCESS && can) {
                CUresult r = cu().p_cuCtxEnablePeerAccess(p0->ctx, 0);
                g->direct[d] = (r == CUDA_SUCCESS || r == CUDA_ERROR_PEER_ACCESS_ALREADY_ENABLED) ? 1 : 0;
            }
            if (cu().p_cuEventCreate(&g->done[d], CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) {
                rc = fail(BSI_E_CUDA, "cuEventCreate failed");
                break;
            }
        }
    }
    if (rc != BSI_OK) {
        std::string keep = g_err;
        bsi_group_destroy(g);
        g_err = keep;
        return rc;
    }
    *out = g;
    return BSI_OK;
}

void bsi_group_destroy(bsi_group *g) {
    if (!g) return;
    for (size_t d = 0; d < g->plans.size(); ++d) {
        bsi_plan *p = g->plans[d];
        if (!p) continue;
        {
            CtxGuard c(p);
            if (p->stream) cu().p_cuStreamSynchronize(p->stream);
            if (g->done[d]) cu().p_cuEventDestroy(g->done[d]);
            give_back(p->device, g->staging[d]);
            if (d == 0) give_back(p->device, g->gather);
        }
        bsi_plan_destroy(p);
    }
    delete g;
}

int bsi_group_size(const bsi_group *g, int *n_devices) {
    if (!g || !n_devices) return fail(BSI_E_ARG, "null pointer");
    *n_devices = (int)g->plans.size();
    return BSI_OK;
}

int bsi_group_plan(bsi_group *g, int i, bsi_plan **plan) {
    if (!g || !plan) return fail(BSI_E_ARG, "null pointer");
    if (i < 0 || i >= (int)g->plans.size()) return fail(BSI_E_ARG, "no such member");
    *plan = g->plans[i];
    return BSI_OK;
}

int bsi_group_attract_run(bsi_group *g, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                          uint64_t max_attractor_l, bsi_run_stats *total, bsi_run_stats *per_device) {
    if (!g) return fail(BSI_E_ARG, "null group");
    if (g->mode != BSI_MODE_ATTRACT) return fail(BSI_E_ARG, "group was created for another mode");
    const int n = (int)g->plans.size();
    bsi_plan *p0 = g->plans[0];
    if (idx_begin > idx_end || idx_end > p0->m.n_problems) return fail(BSI_E_ARG, "bad index range");
    const size_t rec = sizeof(bsi_attractor_rec);
    std::vector<bsi_run_stats> st(n);
    for (int attempt = 0;; ++attempt) {
        int rc = group_gather_buffers(g);
        if (rc) return rc;
        // every device: clear, run its slice, hand its table to device 0 -- all enqueued, no host wait
        // (a run of several detection rounds reads queue lengths between rounds, in its own thread)
        rc = for_each_device(g, [&](int d) -> int {
            bsi_plan *p = g->plans[d];
            uint64_t lo, hi;
            shard(idx_begin, idx_end, d, n, &lo, &hi);
            int r = bsi_attract_table_clear(p);
            if (r) return r;
            if ((r = attract_enqueue(p, lo, hi, max_t, max_attractor_l, nullptr, true))) return r;
            if (d == 0) return BSI_OK;
            CtxGuard c(p);
            const CUdeviceptr hdr0 = g->gather.p + (size_t)(d - 1) * rec;
            const CUdeviceptr recs0 = g->gather.p + (size_t)(n - 1) * rec + (size_t)(d - 1) * g->cap * rec;
            if (g->direct[d]) {
                if ((r = export_table(p, recs0, g->cap, nullptr, hdr0, false, false))) return r;
            } else {
                const CUdeviceptr mine = g->staging[d].p;
                if ((r = export_table(p, mine + rec, g->cap, nullptr, mine, false, false))) return r;
                CU_CHECK(cuMemcpyPeerAsync(hdr0, p0->ctx, mine, p->ctx, rec, p->stream));
                CU_CHECK(cuMemcpyPeerAsync(recs0, p0->ctx, mine + rec, p->ctx, g->cap * rec, p->stream));
            }
            CU_CHECK(cuEventRecord(g->done[d], p->stream));
            return BSI_OK;
        });
        bool failed = rc != BSI_OK;
        std::string err = g_err;
        if (!failed && n > 1) {
            CtxGuard c0(p0);
            for (int d = 1; d < n; ++d) CU_CHECK(cuStreamWaitEvent(p0->stream, g->done[d], 0));
            CUdeviceptr hdrs = g->gather.p, recs = g->gather.p + (size_t)(n - 1) * rec;
            uint32_t n_pay = (uint32_t)(n - 1);
            void *args[] = {&p0->table.p, &p0->table_log2, &hdrs, &recs, &n_pay, &g->cap, &p0->ctr.p};
            if ((rc = launch(p0, p0->f_merge_payloads, (unsigned)std::min<uint64_t>(p0->n_sms, 8 * n_pay), 256, 0, args))) {
                failed = true;
                err = g_err;
            }
        }
        // collect every device's run (also those that were enqueued before another one failed)
        bool overflow = false;
        for (int d = 0; d < n; ++d) {
            bsi_plan *p = g->plans[d];
            if (!p->pending) continue;
            CtxGuard c(p);
            int r = attract_collect(p, &st[d]);
            if (r == BSI_E_CAPACITY) overflow = true;
            else if (r != BSI_OK && !failed) { failed = true; rc = r; err = g_err; }
        }
        if (failed) { g_err = err; return rc; }
        uint64_t need = 0;
        if (n > 1 && !overflow) {
            // did every payload fit, and did the merged table hold?
            CtxGuard c0(p0);
            std::vector<bsi_attractor_rec> hdr(n - 1);
            uint64_t full = 0;
            CU_CHECK(cuMemcpyDtoHAsync(hdr.data(), g->gather.p, (size_t)(n - 1) * rec, p0->stream));
            CU_CHECK(cuMemcpyDtoHAsync(&full, p0->ctr.p + offsetof(Counters, table_full), 8, p0->stream));
            CU_CHECK(cuStreamSynchronize(p0->stream));
            for (auto &h : hdr) need = std::max<uint64_t>(need, h.key[0]);
            if (full) overflow = true;
        }
        if (overflow) {
            // more attractors than the table holds: the reference's store simply grows (attract.py:405-455),
            // so grow the tables of all members and run the job again
            uint32_t log2 = 0;
            for (bsi_plan *p : g->plans) log2 = std::max(log2, p->table_log2);
            if (log2 >= 28 || attempt >= 6) return fail(BSI_E_CAPACITY, "attractor table overflow at 2^%u slots", log2);
            log2 = std::min<uint32_t>(28, log2 + 2);
            for (bsi_plan *p : g->plans) if ((rc = regrow_table(p, log2))) return rc;
            continue;
        }
        if (need > g->cap) {                 // a payload was larger than the gather buffer: run again with room
            while (g->cap < need) g->cap *= 2;
            if (attempt >= 6) return fail(BSI_E_CAPACITY, "table exchange did not converge");
            continue;
        }
        break;
    }
    if (per_device) for (int d = 0; d < n; ++d) per_device[d] = st[d];
    if (total) {
        bsi_run_stats t{};
        for (int d = 0; d < n; ++d) {
            t.n_problems += st[d].n_problems; t.n_found += st[d].n_found; t.n_without += st[d].n_without;
            t.steps += st[d].steps; t.device_steps += st[d].device_steps; t.launches += st[d].launches;
            t.rounds = std::max(t.rounds, st[d].rounds);
            t.ms_total = std::max(t.ms_total, st[d].ms_total);
            t.ms_main = std::max(t.ms_main, st[d].ms_main);
            t.ms_resolve = std::max(t.ms_resolve, st[d].ms_resolve);
        }
        t.launches += n > 1 ? (uint64_t)n : 0;        // the peers' exports + the merge
        *total = t;
    }
    return BSI_OK;
}

int bsi_group_target_run(bsi_group *g, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                         uint64_t *hit_index, uint64_t *hit_t, uint64_t cap, uint64_t *n_hits,
                         bsi_run_stats *total) {
    if (!g) return fail(BSI_E_ARG, "null group");
    if (g->mode != BSI_MODE_TARGET) return fail(BSI_E_ARG, "group was created for another mode");
    const int n = (int)g->plans.size();
    if (idx_begin > idx_end || idx_end > g->plans[0]->m.n_problems) return fail(BSI_E_ARG, "bad index range");
    if (cap && (!hit_index || !hit_t)) return fail(BSI_E_ARG, "null hit buffers");
    // slices are ascending and each device reports its hits in index order, so concatenation in device
    // order IS the enumeration order the reference reports in (simulate.py:149-178); the first `cap`
    // overall are among the first `cap` of each slice
    std::vector<std::vector<uint64_t>> hi(n), ht(n);
    std::vector<uint64_t> got(n, 0);
    std::vector<bsi_run_stats> st(n);
    int rc = for_each_device(g, [&](int d) -> int {
        uint64_t lo, up;
        shard(idx_begin, idx_end, d, n, &lo, &up);
        const uint64_t c = std::min<uint64_t>(cap, up - lo);
        hi[d].resize(c + 1);
        ht[d].resize(c + 1);
        return bsi_target_run(g->plans[d], lo, up, max_t, hi[d].data(), ht[d].data(), c, &got[d], &st[d]);
    });
    if (rc) return rc;
    uint64_t written = 0;
    bsi_run_stats t{};
    for (int d = 0; d < n; ++d) {
        for (uint64_t i = 0; i < got[d] && written < cap; ++i, ++written) {
            hit_index[written] = hi[d][i];
            hit_t[written] = ht[d][i];
        }
        t.n_problems += st[d].n_problems; t.n_found += st[d].n_found; t.n_without += st[d].n_without;
        t.steps += st[d].steps; t.device_steps += st[d].device_steps; t.launches += st[d].launches;
        t.rounds = std::max(t.rounds, st[d].rounds);
        t.ms_total = std::max(t.ms_total, st[d].ms_total);
        t.ms_main = std::max(t.ms_main, st[d].ms_main);
    }
    if (n_hits) *n_hits = written;
    if (total) *total = t;
    return BSI_OK;
}

int bsi_group_simulate_run(bsi_group *g, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                           uint64_t *states_out, bsi_run_stats *total) {
    if (!g) return fail(BSI_E_ARG, "null group");
    const int n = (int)g->plans.size();
    if (idx_begin > idx_end || idx_end > g->plans[0]->m.n_problems) return fail(BSI_E_ARG, "bad index range");
    std::vector<bsi_run_stats> st(n);
    const uint64_t per_problem = (max_t + 1) * (uint64_t)g->plans[0]->m.kw;
    int rc = for_each_device(g, [&](int d) -> int {
        uint64_t lo, up;
        shard(idx_begin, idx_end, d, n, &lo, &up);
        if (lo == up) return BSI_OK;
        return bsi_simulate_run(g->plans[d], lo, up, max_t, states_out + (lo - idx_begin) * per_problem, &st[d]);
    });
    if (rc) return rc;
    if (total) {
        bsi_run_stats t{};
        for (int d = 0; d < n; ++d) {
            t.n_problems += st[d].n_problems; t.n_found += st[d].n_found; t.steps += st[d].steps;
            t.device_steps += st[d].device_steps; t.launches += st[d].launches;
            t.ms_total = std::max(t.ms_total, st[d].ms_total);
            t.ms_main = std::max(t.ms_main, st[d].ms_main);
        }
        *total = t;
    }
    return BSI_OK;
}

// ---- the bare update loop of this plan's network --------------------------------------------------
int bsi_measure_update_rate(bsi_plan *p, double *word_steps_per_s, double *lop3_per_s) {
    int rc = need_device(p, -1);
    if (rc) return rc;
    CtxGuard g(p);
    CUfunction f = nullptr;
    CU_CHECK(cuModuleGetFunction(&f, p->mods[4], "bsi_update_rate"));
    if ((rc = ensure(p, p->misc, 256))) return rc;
    int per_sm = 0;
    CU_CHECK(cuOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, p->block, 0));
    per_sm = std::max(1, std::min(per_sm, 8));
    unsigned grid = (unsigned)(p->n_sms * per_sm);
    uint32_t pairs = 20000, seed = 777;
    void *args[] = {&pairs, &seed, &p->misc.p};
    p->events_used = 0;
    CUevent e0, e1;
    if ((rc = get_event(p, &e0)) || (rc = get_event(p, &e1))) return rc;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CU_CHECK(cuEventRecord(e0, p->stream));
        if ((rc = launch(p, f, grid, p->block, 0, args))) return rc;
        CU_CHECK(cuEventRecord(e1, p->stream));
        CU_CHECK(cuEventSynchronize(e1));
        double ms = 0;
        if ((rc = elapsed(e0, e1, &ms))) return rc;
        if (rep > 0) best = std::max(best, (double)grid * p->block * pairs * 2.0 / (ms * 1e-3));
    }
    if (word_steps_per_s) *word_steps_per_s = best;
    if (lop3_per_s) *lop3_per_s = best * p->info.update_lop3;
    return BSI_OK;
}

// ---- integer-logic roofline microbenchmark --------------------------------------------------------
static const char *kPeakSrc = R"BSIRAW(
typedef unsigned int u32;
// 16 independent chains per thread, each link one lop3.b32 (LUT 0x96 = a ^ b ^ c).
extern "C" __global__ void __launch_bounds__(256) bsi_lop3_peak(u32 *out, u32 iters, u32 seed) {
    u32 r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = seed * (i + 1) + threadIdx.x;
    u32 k1 = seed ^ 0x9E3779B9u, k2 = seed + blockIdx.x;
    for (u32 it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(k1), "r"(k2));
        }
    }
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
    if (acc == 0x12345678u) out[0] = acc;
}
)BSIRAW";

int bsi_measure_lop3_peak(int device, double *lop3_per_s, double *sm_clock_mhz, int *n_sms) {
    if (!cu().ok) return fail(BSI_E_CUDA, "%s", cu().why.c_str());
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, kPeakSrc, "bsi_peak.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(BSI_E_NVRTC, "nvrtcCreateProgram failed");
    const char *opts[] = {"--gpu-architecture=sm_100a", "-lineinfo"};
    if (nvrtcCompileProgram(prog, 2, opts) != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(BSI_E_NVRTC, "peak kernel failed to compile");
    }
    size_t sz = 0;
    nvrtcGetCUBINSize(prog, &sz);
    std::vector<char> cubin(sz);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    bsi_plan tmp;
    int count = 0;
    CU_CHECK(cuDeviceGetCount(&count));
    if (device < 0 || device >= count) return fail(BSI_E_CUDA, "device out of range");
    int rc = device_context(device, &tmp.dev, &tmp.ctx);
    if (rc) return rc;
    {
        CtxGuard g(&tmp);
        CUmodule mod = nullptr;
        CUfunction f = nullptr;
        CUstream s = nullptr;
        CUevent e0 = nullptr, e1 = nullptr;
        CUdeviceptr out = 0;
        int sms = 0, clk_khz = 0;
        auto body = [&]() -> int {
            CU_CHECK(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, tmp.dev));
            CU_CHECK(cuDeviceGetAttribute(&clk_khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE, tmp.dev));
            CU_CHECK(cuModuleLoadData(&mod, cubin.data()));
            CU_CHECK(cuModuleGetFunction(&f, mod, "bsi_lop3_peak"));
            CU_CHECK(cuStreamCreate(&s, CU_STREAM_NON_BLOCKING));
            CU_CHECK(cuEventCreate(&e0, CU_EVENT_DEFAULT));
            CU_CHECK(cuEventCreate(&e1, CU_EVENT_DEFAULT));
            CU_CHECK(cuMemAlloc(&out, 256));
            uint32_t iters = 20000, seed = 12345;
            unsigned grid = (unsigned)sms * 8;
            void *args[] = {&out, &iters, &seed};
            double best = 0;
            for (int rep = 0; rep < 5; ++rep) {
                CU_CHECK(cuEventRecord(e0, s));
                CU_CHECK(cuLaunchKernel(f, grid, 1, 1, 256, 1, 1, 0, s, args, nullptr));
                CU_CHECK(cuEventRecord(e1, s));
                CU_CHECK(cuEventSynchronize(e1));
                float ms = 0;
                CU_CHECK(cuEventElapsedTime(&ms, e0, e1));
                double ops = (double)grid * 256.0 * iters * 8.0 * 16.0;
                if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
            }
            if (lop3_per_s) *lop3_per_s = best;
            if (n_sms) *n_sms = sms;
            // clock implied by 64 lop3 / clk / SM (alu pipe: 16 lanes x 4 sub-partitions)
            if (sm_clock_mhz) *sm_clock_mhz = best / (64.0 * sms) / 1e6;
            return BSI_OK;
        };
        rc = body();
        if (out) cu().p_cuMemFree(out);
        if (e0) cu().p_cuEventDestroy(e0);
        if (e1) cu().p_cuEventDestroy(e1);
        if (s) cu().p_cuStreamDestroy(s);
        if (mod) cu().p_cuModuleUnload(mod);
    }
    tmp.ctx = nullptr;
    return rc;
}

}  // extern "C"
