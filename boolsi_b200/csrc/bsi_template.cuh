// Hand-written sm_100a kernel template of libbsi.  Compiled at plan-creation time by NVRTC after
// the per-network macro header (bsi_codegen.hpp) has been prepended.
//
// Layout: trajectories are bit-sliced, 32 per 32-bit word ("lanes"), one register word per
// network node; one thread owns one word at a time.  A synchronous step (reference
// model.py:16-73) is BSI_UPDATE: one lop3.b32 per node with <= 3 inputs, a Shannon tree of lop3
// for wider rules, constants and AND/OR masks for fixed nodes and perturbations.  State never
// leaves registers; HBM sees only result records.
//
// attract (reference model.py:152-236 + attract.py:262-302, default "storing all states" mode).
// The reference stops a trajectory at the first repeated state, t = mu + lambda (relative to the
// last perturbation time lp), provided that is <= T = max_t - lp.  Here, store-free:
//
//   bsi_attract_dense / bsi_attract_queue   one forward pass per word (dense: a contiguous index
//       range; queue: the lanes a previous round left open) with up to three live CHECKPOINTS.  A
//       lane whose state equals a checkpoint taken at time c is on its cycle, c >= mu, and because
//       every later state is compared the first such match is at time c + lambda: lambda is exact.
//       Checkpoint times for a horizon H: 0, 1, 2, 3, 4, 6, 8, 12, 16, 24, ... while they are below
//       ceil(H/3) (early detection, Brent style, all in slot 0), then ceil(H/3), ceil(2H/3), H in
//       slots 0, 1, 2.  A lane with c_(i-1) < mu <= c_i and mu + lambda <= H is matched by time
//       c_i + lambda < H + gap, so running to H + max gap - 1 PROVES "mu + lambda > H" for every
//       lane still unmatched (tests/test_schedule.py checks the obligation by brute force).
//       The per-step compare is a FILTER over a dozen nodes held in registers; only when some lane
//       passes the filter (and a second filter over a dozen more words) is the full state
//       compared against the checkpoint, which is kept in shared memory.  Lanes matched against
//       an early checkpoint leave bit-sliced through the early-word queue, the others as
//       48-byte records in the found queue, open lanes as indices in the retry queue.
//   bsi_resolve_mu, bsi_resolve_early (long queues), bsi_resolve + bsi_resolve_finish (short ones)
//       canonical key = minimum state code over the cycle (model.py:140-149, attract.py:296) by a
//       bit-sliced lexicographic-minimum walk of the cycle; exact mu by walking x_j (from the
//       previous checkpoint, whose time is provably < mu, or from the problem index) in lock-step
//       with the periodic extension of the cycle; then acceptance (mu + lambda <= T, lambda <=
//       max_attractor_l, attract.py:294) and aggregation into the device attractor table.

typedef unsigned char u8;
typedef unsigned int u32;
typedef unsigned long long u64;

template <int LUT>
__device__ __forceinline__ u32 bsi_lop3(u32 a, u32 b, u32 c) {
#ifdef BSI_HOST_EMUL        // tests/host_emul: this very template, executed by 32 host threads per warp
    return bsi_host_lop3(LUT, a, b, c);
#else
    u32 d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
#endif
}

#define BSI_MAX_CK 72
#ifndef BSI_DICT_MIN_LAM
#define BSI_DICT_MIN_LAM 8u          // cycles up to this length are walked, longer ones looked up
#endif

// 32x32 bit-matrix transpose in place (Hacker's Delight 7-3): B[r] bit c = A[31-c] bit 31-r.
// Turns 32 node words (bit = lane) into 32 lane words (bit = node) and back in 5 x 16 butterfly
// steps instead of 32 x 32 single-bit moves (BSI_PACK / BSI_UNPACK in the generated header).
#ifndef BSI_TUNE
#define BSI_TUNE 0
#endif
__device__ __forceinline__ void bsi_transpose32(unsigned int (&A)[32]) {
    unsigned int m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const unsigned int t = (A[k] ^ (A[k + j] >> j)) & m;
            A[k] ^= t;
            A[k + j] ^= (t << j);
        }
    }
}

// (BSI_TUNE bit 0: in the resolution kernels the transpose is a function call on a local-memory array,
//  not ~400 inlined instructions per use: those kernels hold six to ten uses, most of their code)
__device__ __noinline__ void bsi_transpose32_call(unsigned int (&A)[32]) { bsi_transpose32(A); }
#define BSI_TRANSPOSE32(A) bsi_transpose32(A)

struct BsiCounters {
    // reset before every round
    u64 next_word;      // work counter of the persistent trajectory kernel
    u64 next_word2;     // work counter of bsi_resolve_key
    u64 next_word3;     // work counter of bsi_resolve_mu
    u64 n_found;        // entries in the found queue (target: closed queue)
    u64 n_retry;        // entries in the retry queue
    u64 n_early;        // entries in the early-word queue
    u64 next_word4;     // work counter of bsi_resolve_early
    // accumulated over a run
    u64 n_without;      // lanes proven to have no attractor within the caps
    u64 n_accepted;     // lanes whose attractor went into the table
    u64 steps;          // sum of reference stop times (model.py:235) of finished lanes
    u64 device_steps;   // word-steps executed by the trajectory kernels
    u64 table_full;     // set when the attractor table overflowed
    u64 n_hits;         // target: number of hits
    u64 internal_error; // set when an invariant of the resolution kernels failed
    u64 pad[2];
    u64 dbg[4];         // debug_stats counters
};

struct BsiSched {       // checkpoint schedule of one detection round (times relative to lp)
    u32 n_ck;
    u32 end;            // last compared time
    u32 first_final;    // time of the first final checkpoint (earlier ones are overwritten)
    u32 pad;
    u32 ck_time[BSI_MAX_CK];
    u32 ck_slot[BSI_MAX_CK];
};

struct BsiFound {       // one lane that matched a checkpoint
    u64 idx;            // problem index
    u64 y[BSI_KW];      // the matched state (on the cycle), state code words
    u64 key[BSI_KW];    // canonical key, filled by bsi_resolve_key
    u64 z[BSI_KW];      // state at time c0 (valid when c0 > 0)
    u32 lam;            // cycle length
    u32 c;              // time of the matched checkpoint (>= mu), relative to lp
    u32 c0;             // a time < mu: the previous final checkpoint, or 0 (then x_0 comes from idx)
    u32 pad;
};

// Early-word queue (dense rounds only): a word whose lanes matched EARLY checkpoints hands them
// over bit-sliced -- no per-lane records, no transposes.  What it hands over is the word's state at
// the time j_end its loop ended: every matched lane has stayed on its cycle since it matched, so
// that ONE state (all lanes at the same time) is an on-cycle state of each of them -- no latch of the
// matched checkpoints is needed -- plus the list of its match EVENTS (lanes, lambda): lanes that
// matched the same checkpoint at the same time share their cycle length, so a word of 32 lanes has
// a handful of events and the resolution works per event, not per lane.  Entries come in blocks of
// 32 (the 32 threads of a warp reserve consecutive entries and later read consecutive entries), and a
// block is one contiguous piece of memory, structure of arrays inside:
//   u64 word[32]; u32 mask[32]; u32 info[32] (n_events | j_end << 8); u32 evm[32 events][32];
//   u32 evl[32 events][32]; u32 y[N][32]
// so that a warp's loads per entry hit a few adjacent DRAM pages instead of planes that lie `cap`
// entries apart.
#define BSI_EQ_ENTRY_BYTES (8ull + 4ull + 4ull + 32ull * 4ull + 32ull * 4ull + 4ull * BSI_N)
#define BSI_EQ_BLOCK_BYTES (32ull * BSI_EQ_ENTRY_BYTES)
struct BsiEarlyQ {
    unsigned char *base;
    u64 cap;
};
struct BsiEarlyEntry {
    u64 *word;          // [0]
    u32 *mask;          // [0]  the lanes handed over
    u32 *info;          // [0]  n_events | j_end << 8
    u32 *evm;           // [event * 32]  lanes of the event
    u32 *evl;           // [event * 32]  their cycle length
    u32 *y;             // [node * 32]   x at time j_end
};
__device__ __forceinline__ BsiEarlyQ bsi_early_view(void *base, u64 cap) {
    BsiEarlyQ q;
    q.base = (unsigned char *)base;
    q.cap = cap;
    return q;
}
__device__ __forceinline__ BsiEarlyEntry bsi_early_entry(const BsiEarlyQ &q, u64 e) {
    unsigned char *blk = q.base + (e >> 5) * BSI_EQ_BLOCK_BYTES;
    const u32 k = (u32)e & 31u;
    BsiEarlyEntry r;
    r.word = (u64 *)blk + k;
    r.mask = (u32 *)(blk + 32ull * 8ull) + k;
    r.info = r.mask + 32;
    r.evm = r.info + 32;
    r.evl = r.evm + 32 * 32;
    r.y = r.evl + 32 * 32;
    return r;
}

struct BsiSlot {
    u64 key[BSI_KW];
    u64 n, s1, s2;
    u32 lam;
    u32 flag;           // 0 empty, 1 being written, 2 ready
};

struct BsiRec {         // == bsi_attractor_rec (include/bsi.h)
    u64 key[4];
    u64 n, s1, s2, lam;
};

#define BSI_DECL_ONE(P, i) u32 P##i;
#define BSI_DECL_A(i) BSI_DECL_ONE(a, i)
#define BSI_DECL_B(i) BSI_DECL_ONE(b, i)
#define BSI_DECL_C(i) BSI_DECL_ONE(c, i)
#define BSI_DECL_K0(i) BSI_DECL_ONE(k0_, i)
#define BSI_DECL_K1(i) BSI_DECL_ONE(k1_, i)
#define BSI_DECL_K2(i) BSI_DECL_ONE(k2_, i)
#define BSI_ZERO_A(i) a##i = 0u;
#define BSI_COPY_AB(i) a##i = b##i;
#define BSI_COPY_CA(i) c##i = a##i;
#define BSI_COPY_CB_(i) c##i = b##i;
#define BSI_SAVE_K0(i) k0_##i = a##i;
#define BSI_SAVE_K1(i) k1_##i = a##i;
#define BSI_SAVE_K2(i) k2_##i = a##i;
#define BSI_ZERO_K(i) k0_##i = 0u; k1_##i = 0u; k2_##i = 0u;

// lane masks of per-lane fixed nodes / perturbation variations (simulate and target only)
#define BSI_DECL_FIXVAR(i) u32 fm##i = 0u, fv##i = 0u;
#define BSI_DECL_PVAR(q) u32 pm##q = 0u, pv##q = 0u;

__device__ __forceinline__ u64 bsi_hash_key(const u64 *k) {
    u64 h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < BSI_KW; ++i) {
        h ^= k[i];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 29;
    }
    // (a product's low bits only see the low bits of its input: without this round, states that
    //  differ in nodes 48 and up only -- half of the states of a typical cycle pair up that way --
    //  shared a slot index, and every second lookup took the probe sequence instead of one load)
    h ^= h >> 32;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 31;
    return h;
}

// flag of a table slot, with acquire semantics: what its writer stored before the release (key,
// lambda) is visible after it -- no fence on the reader's side (a membar.gl per table update was a
// quarter of the early-word kernel's stall samples)
__device__ __forceinline__ u32 bsi_ld_acquire_u32(const u32 *p) {
#ifdef BSI_HOST_EMUL
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}

// Adds (n, s1, s2) for `key` to the open-addressing table (the replacement for the ZODB store and
// the Welford/Chan merges of reference attract.py:35-45, 405-455: integer sums are exact and
// order-independent).
__device__ __noinline__ void bsi_table_add(BsiSlot *table, u32 log2_slots, const u64 *key, u64 n,
                                           u64 s1, u64 s2, u32 lam, BsiCounters *ctr) {
    const u64 mask = ((u64)1 << log2_slots) - 1;
    u64 pos = bsi_hash_key(key) & mask;
    // (a bounded probe sequence: a table this full is reported as overflow, not searched on)
    for (u64 probe = 0; probe <= mask && probe < 4096; ++probe, pos = (pos + 1) & mask) {
        BsiSlot *s = table + pos;
        u32 f = bsi_ld_acquire_u32(&s->flag);
        if (f == 0u) {
            u32 old = atomicCAS(&s->flag, 0u, 1u);
            if (old == 0u) {
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) s->key[i] = key[i];
                s->lam = lam;
                __threadfence();
                atomicExch(&s->flag, 2u);
                f = 2u;
            } else {
                f = old;
            }
        }
        while (f != 2u) f = bsi_ld_acquire_u32(&s->flag);
        bool same = true;
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) same = same && (((volatile u64 *)s->key)[i] == key[i]);
        if (same) {
            atomicAdd(&s->n, n);
            atomicAdd(&s->s1, s1);
            atomicAdd(&s->s2, s2);
            return;
        }
    }
    atomicExch(&ctr->table_full, 1ull);
}

// ---------------------------------------------------------------------------------------------
// cycle dictionary: on-cycle state -> (attractor, phase)
// ---------------------------------------------------------------------------------------------
// The number of DIFFERENT attractors of a network is tiny next to the number of trajectories that
// fall into them (reference attract.py:374-402 aggregates millions of problems into a handful of
// records), so the canonical key of an attractor -- the minimum state code over the cycle,
// model.py:140-149 / attract.py:296 -- is worked out by a walk ONCE per attractor and then looked
// up: the first thread that resolves a lane of a new attractor stores the whole orbit, key state
// first, in `orbit` and enters every state into an open-addressing table.  A later lane that
// matched state y finds (base, phase) with one probe; its key is orbit[base] and the periodic
// extension of its cycle at any time is orbit[base + (phase - dt) mod lambda] -- no key walk and no
// per-lane advance walk.  The dictionary is a CACHE of facts about the network's dynamics: a
// miss, a full table, or a racing writer only mean that the lane takes the walks; results never
// depend on it.  Slots go 0 (empty) -> 1 (being written, a few instructions) -> 2 (ready), 3
// (tombstone: orbit storage exhausted) or 4 (the key state of an attractor whose orbit is still
// being stored: reserved, not yet usable); a slot's fields are written once.  NO state of an
// attractor is published before its whole orbit is in memory: a reader that finds a state may
// read any phase of the orbit.
struct __align__(16) BsiDictSlot {
    u32 flag, idx, base, lam;       // idx = base + phase
    u64 st[BSI_KW];
};
struct BsiDict {
    BsiDictSlot *slots;             // nullptr: dictionary off
    u64 *orbit;                     // [cap_states][KW]
    u64 *n_states;                  // fill counter of `orbit` (may run past cap_states)
    u64 *n_attr;                    // number of stored attractors
    u32 *dir;                       // [cap_attr][2]: (base, lambda) of every stored attractor
    u32 log2_slots, cap_states, cap_attr, pad;
};

__device__ __forceinline__ void bsi_dict_head(const BsiDictSlot *s, u32 &f, u32 &idx, u32 &base, u32 &lam) {
#ifdef BSI_HOST_EMUL
    f = __atomic_load_n(&s->flag, __ATOMIC_SEQ_CST);
    idx = s->idx; base = s->base; lam = s->lam;
#else
    // one 16-byte transaction: the four fields are seen together
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(f), "=r"(idx), "=r"(base), "=r"(lam) : "l"(s));
#endif
}
__device__ __forceinline__ u64 bsi_ld_u64(const u64 *p) { return *((const volatile u64 *)p); }
// The same through the L2 only (ld.global.cg), NOT volatile: independent loads of a batch may be
// issued back to back, so that a word's 32 lookups cost a few L2 round trips and not 32.
__device__ __forceinline__ void bsi_dict_head_cg(const BsiDictSlot *s, u32 &f, u32 &idx, u32 &base, u32 &lam) {
#ifdef BSI_HOST_EMUL
    bsi_dict_head(s, f, idx, base, lam);
#else
    asm("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(f), "=r"(idx), "=r"(base), "=r"(lam) : "l"(s));
#endif
}
// Read-once data (the early-word queue: 1.1 GB per 2^26 states, every entry read by exactly one
// thread): streaming loads (ld.global.cs, evict-first), so that the queue does not push the
// resolution kernels' local-memory arrays out of the L2 on its way through.  (BSI_TUNE 128: plain loads.)
__device__ __forceinline__ u32 bsi_ld_stream(const u32 *p) {
#if defined(BSI_HOST_EMUL) || (BSI_TUNE & 128)
    return *p;
#else
    return __ldcs(p);
#endif
}
__device__ __forceinline__ u64 bsi_ld_stream(const u64 *p) {
#if defined(BSI_HOST_EMUL) || (BSI_TUNE & 128)
    return *p;
#else
    return (u64)__ldcs((const unsigned long long *)p);
#endif
}
// a found-queue record, read once (throughput-mode resolution): streaming loads as well
__device__ __forceinline__ BsiFound bsi_found_stream(const BsiFound *p) {
    union { BsiFound r; u64 w[sizeof(BsiFound) / 8]; } u;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(BsiFound) / 8); ++i) u.w[i] = bsi_ld_stream((const u64 *)p + i);
    return u.r;
}
__device__ __forceinline__ u64 bsi_ldcg_u64(const u64 *p) {
#ifdef BSI_HOST_EMUL
    return *((const volatile u64 *)p);
#else
    u64 v;
    asm("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#endif
}

// state -> (idx, base, lambda); false on a miss (or on a slot that is still being written)
__device__ __noinline__ bool bsi_dict_find(const BsiDict &d, const u64 *st, u32 &idx, u32 &base, u32 &lam) {
    const u32 mask = (1u << d.log2_slots) - 1u;
    u32 pos = (u32)bsi_hash_key(st) & mask;
#pragma unroll 1
    for (u32 probe = 0; probe < 32u; ++probe, pos = (pos + 1u) & mask) {
        const BsiDictSlot *s = d.slots + pos;
        u32 f;
        bsi_dict_head(s, f, idx, base, lam);
        if (f == 2u || f == 4u) {
            // (a state word that lags behind the flag reads as zero or as itself: at worst a miss)
            bool same = true;
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) same = same && (bsi_ld_u64(&s->st[i]) == st[i]);
            if (same) return f == 2u;
        } else if (f != 3u) {
            return false;
        }
    }
    return false;
}

// enter state `st` = orbit[idx] of the attractor stored at `base` (the caller has written the orbit entry)
__device__ __noinline__ void bsi_dict_put(const BsiDict &d, const u64 *st, u32 idx, u32 base, u32 lam) {
    const u32 mask = (1u << d.log2_slots) - 1u;
    u32 pos = (u32)bsi_hash_key(st) & mask;
    for (u32 probe = 0; probe < 32u; ++probe, pos = (pos + 1u) & mask) {
        BsiDictSlot *s = d.slots + pos;
        u32 f = *((volatile u32 *)&s->flag);
        if (f == 0u) {
            const u32 old = atomicCAS(&s->flag, 0u, 1u);
            if (old == 0u) {
                s->idx = idx; s->base = base; s->lam = lam;
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) s->st[i] = st[i];
                __threadfence();
                atomicExch(&s->flag, 2u);
                return;
            }
            f = old;
        }
        while (f == 1u) f = *((volatile u32 *)&s->flag);      // (a writer never waits inside this section)
        if (f == 2u || f == 4u) {
            __threadfence();
            bool same = true;
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) same = same && (bsi_ld_u64(&s->st[i]) == st[i]);
            if (same) return;
        }
    }
}

// Claim the attractor with key state `key` and length `lam`: true if THIS thread reserved the key
// (slot state 4; it then owns orbit[base .. base + lam), must store the orbit, enter the other states
// and finally publish the key with bsi_dict_publish), false if the key is already there or reserved,
// the table is crowded, or the orbit storage is exhausted.
__device__ __noinline__ bool bsi_dict_claim(const BsiDict &d, const u64 *key, u32 lam, u32 &base, u32 &slot) {
    if (bsi_ld_u64(d.n_states) + lam > (u64)d.cap_states || bsi_ld_u64(d.n_attr) >= (u64)d.cap_attr) return false;
    const u32 mask = (1u << d.log2_slots) - 1u;
    u32 pos = (u32)bsi_hash_key(key) & mask;
    for (u32 probe = 0; probe < 32u; ++probe, pos = (pos + 1u) & mask) {
        BsiDictSlot *s = d.slots + pos;
        u32 f = *((volatile u32 *)&s->flag);
        if (f == 0u) {
            const u32 old = atomicCAS(&s->flag, 0u, 1u);
            if (old == 0u) {
                const u64 b = atomicAdd(d.n_states, (u64)lam);
                const u64 a = b + lam <= (u64)d.cap_states ? atomicAdd(d.n_attr, 1ull) : (u64)d.cap_attr;
                if (a >= (u64)d.cap_attr) {
                    atomicExch(&s->flag, 3u);
                    return false;
                }
                base = (u32)b;
                slot = pos;
                d.dir[2 * a] = base;
                d.dir[2 * a + 1] = lam;
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) { d.orbit[b * BSI_KW + i] = key[i]; s->st[i] = key[i]; }
                s->idx = base; s->base = base; s->lam = lam;
                __threadfence();
                atomicExch(&s->flag, 4u);
                return true;
            }
            f = old;
        }
        while (f == 1u) f = *((volatile u32 *)&s->flag);
        if (f == 2u || f == 4u) {
            __threadfence();
            bool same = true;
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) same = same && (bsi_ld_u64(&s->st[i]) == key[i]);
            if (same) return false;
        }
    }
    return false;
}
__device__ __forceinline__ void bsi_dict_publish(const BsiDict &d, u32 slot) {
    __threadfence();
    atomicExch(&d.slots[slot].flag, 2u);
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------

// lanes of word W (problems 32W .. 32W+31) that fall into [idx_begin, idx_end)
__device__ __forceinline__ u32 bsi_valid_mask(u64 W, u64 idx_begin, u64 idx_end) {
    u64 base = W << 5;
    u32 v = 0xFFFFFFFFu;
    if (base < idx_begin) {
        u64 skip = idx_begin - base;
        v = skip >= 32 ? 0u : (0xFFFFFFFFu << (u32)skip);
    }
    if (base + 32 > idx_end) {
        u64 keep = idx_end > base ? idx_end - base : 0;
        v &= keep >= 32 ? 0xFFFFFFFFu : ((1u << (u32)keep) - 1u);
    }
    return v;
}

// each warp takes 32 consecutive words from a global counter; returns this thread's word
__device__ __forceinline__ u64 bsi_grab_word(u64 *counter) {
    u64 base = 0;
    if ((threadIdx.x & 31u) == 0u) base = atomicAdd(counter, 32ull);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    return base + (threadIdx.x & 31u);
}

// warp-aggregated reservation of `n` queue entries per thread; returns this thread's first index
__device__ __forceinline__ u64 bsi_reserve(u64 *counter, u32 n) {
    const u32 lane = threadIdx.x & 31u;
    if (!__any_sync(0xFFFFFFFFu, n != 0u)) return 0ull;      // (the usual case for two of the three queues)
    u32 incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    const u32 tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    u64 base = 0;
    if (lane == 31u && tot) base = atomicAdd(counter, (u64)tot);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    return base + (incl - n);
}

// Warp-aggregated table update, called by ALL 32 threads of a warp (n == 0: nothing to add).
// In the ordered regime millions of lanes fall into a handful of attractors: one atomicAdd
// triple per THREAD then serialises on a few L2 addresses (measured: 2.3 ns per word, the whole
// resolution time).  Threads holding the same key as the first pending thread are summed over
// the warp (redux on 16-bit pieces: exact, no carries lost) and that thread takes the total; after
// a few rounds (many distinct keys in the warp: nothing to gain) the others keep their own.  The
// table is only touched AFTER the rounds, by all the threads that hold a total at once, so the
// probe latencies overlap instead of adding up round by round.
__device__ __forceinline__ u64 bsi_warp_sum(u64 x) {
    const u32 p0 = __reduce_add_sync(0xFFFFFFFFu, (u32)x & 0xFFFFu);
    const u32 p1 = __reduce_add_sync(0xFFFFFFFFu, (u32)(x >> 16) & 0xFFFFu);
    const u32 p2 = __reduce_add_sync(0xFFFFFFFFu, (u32)(x >> 32) & 0xFFFFu);
    const u32 p3 = __reduce_add_sync(0xFFFFFFFFu, (u32)(x >> 48));
    return (u64)p0 + ((u64)p1 << 16) + ((u64)p2 << 32) + ((u64)p3 << 48);
}
__device__ __forceinline__ void bsi_warp_table_add(BsiSlot *table, u32 log2_slots, const u64 *key, u64 n,
                                                   u64 s1, u64 s2, u32 lam, BsiCounters *ctr) {
    const u32 lane = threadIdx.x & 31u;
    u32 pending = __ballot_sync(0xFFFFFFFFu, n != 0ull);
#pragma unroll 1
    for (int round = 0; pending && round < 3; ++round) {
        const u32 leader = __ffs(pending) - 1;
        bool same = n != 0ull;
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) {
            const u64 k = __shfl_sync(0xFFFFFFFFu, key[i], leader);     // (every thread takes part:
            same = same && (k == key[i]);                               //  no short-circuit around it)
        }
        const u32 grp = __ballot_sync(0xFFFFFFFFu, same);
        if (grp != (1u << leader)) {                                    // (uniform over the warp)
            const u64 a = bsi_warp_sum(same ? n : 0ull);
            const u64 b = bsi_warp_sum(same ? s1 : 0ull);
            const u64 c = bsi_warp_sum(same ? s2 : 0ull);
            if (lane == leader) { n = a; s1 = b; s2 = c; }
            else if (same) n = 0ull;                                    // handed over to the leader
        }
        pending &= ~grp;
    }
    if (n != 0ull) bsi_table_add(table, log2_slots, key, n, s1, s2, lam, ctr);
}

// perturbation prefix: LP steps from t = 0, perturbations overriding at their time
// (reference model.py:76-128).  Ends with the state at time LP in a*.
#define BSI_PREFIX(LP)                                                              \
    {                                                                               \
        u64 t_ = 1;                                                                 \
        for (; t_ + 1 <= (LP); t_ += 2) {                                           \
            BSI_UPDATE(b, a) BSI_PERTURB(b, t_)                                     \
            BSI_UPDATE(a, b) BSI_PERTURB(a, (t_ + 1))                               \
        }                                                                           \
        if (t_ <= (LP)) { BSI_UPDATE(b, a) BSI_PERTURB(b, t_) BSI_FOR_NODES(BSI_COPY_AB) } \
    }

// ---------------------------------------------------------------------------------------------
// attract
// ---------------------------------------------------------------------------------------------
#if BSI_MODE == 1

// dynamic shared memory of the main kernel: the three checkpoint slots in full,
// u32 ck[3][N][BLOCK] (48 KB per 64-thread block at N = 64; four blocks per SM fit).  The
// per-thread scratch in global memory (sized by the persistent grid) only holds the detection
// records of the thread's 32 lanes: u32 lam[32], u32 c[32], u32 c0[32] for matches against final
// checkpoints (the matched states themselves stay in shared memory until the end of the word) and
// the event list of the early matches, u32 evm[32], u32 evl[32] (BSI_SCRATCH_WORDS in all).
#define BSI_SCRATCH_WORDS 160ull
#ifdef BSI_HOST_EMUL
extern unsigned char bsi_smem[];
#else
extern __shared__ __align__(16) unsigned char bsi_smem[];
#endif

// full compare of a state held in local memory with a checkpoint in shared memory: lanes of
// `cand` whose N words all agree
__device__ __forceinline__ u32 bsi_slow_match(const u32 *xs, const u32 *ck, u64 cstride, u32 cand) {
    // (a lane that passed both filters almost always matches, so there is no early exit)
    u32 acc = ~cand;
#pragma unroll 8
    for (int i = 0; i < BSI_N; ++i) acc |= xs[i] ^ ck[i * cstride];
    return ~acc;
}

// detection record of the lanes in m: cycle length, checkpoint time and the matched state code
// A match against a FINAL checkpoint (never overwritten) only notes the slot: the matched state
// (= the checkpoint itself) and the previous checkpoint are extracted from shared memory at the end
// of the word, when all threads of the warp do so together instead of one diverged thread now.
#define BSI_DEFERRED 0x80000000u
// A match against an EARLY checkpoint (slot 0 before its final time; it will be overwritten) is one
// EVENT: its lanes and their common cycle length, two words, whatever the number of lanes.  Nothing
// of the matched state is kept: the lanes go on being updated with the rest of the word, so the
// word's state when its loop ends is an on-cycle state of every one of them (see the early-word
// queue above).  A word has at most 32 events (every event retires at least one lane).
__device__ __forceinline__ void bsi_record(u32 m, u32 lam, u32 c, u32 *lam_s, u32 *c_s, u64 sstride, u32 slot) {
    u32 *c0_s = c_s + 32ull * sstride;
    while (m) {
        const u32 l = __ffs(m) - 1;
        m &= m - 1;
        lam_s[l * sstride] = lam;
        c_s[l * sstride] = c;
        c0_s[l * sstride] = BSI_DEFERRED | slot;
    }
}

// DENSE: words of the contiguous index range [idx_begin, idx_end), word0 = idx_begin >> 5.
// otherwise: word w = lanes 32w .. 32w+31 of the retry queue `in_queue`.
template <bool DENSE>
__device__ __forceinline__ void bsi_main_body(u64 word0, u64 n_words, u64 idx_begin, u64 idx_end,
                                              const u64 *__restrict__ in_queue, u64 n_in,
                                              const BsiSched &sched, int final_round, u64 stop_t,
                                              BsiFound *found_q, u64 *retry_q, u32 *scratch,
                                              void *early_base, u64 early_cap, BsiCounters *ctr) {
    const u32 tid = threadIdx.x;
    u32 *const ck_s = (u32 *)bsi_smem + tid;                          // checkpoints [slot][node][BLOCK]
    const u64 sstride = (u64)gridDim.x * BSI_BLOCK;
    const u64 gtid = (u64)blockIdx.x * BSI_BLOCK + tid;
    // detection records of this thread's 32 lanes (rarely written), in global memory
    u32 *const lam_s = scratch + gtid;                                // [32][sstride]
    u32 *const c_s = lam_s + 32ull * sstride;                         // [32][sstride], then c0 [32]
    // checkpoint slot S: the first BSI_SMEM_SLOTS slots live in shared memory, the later ones
    // (armed for the shortest time) in this thread's scratch, [slot][node][sstride], L2-resident
    u32 *const evm_s = scratch + 96ull * sstride + gtid;              // [32][sstride] early events: lanes
    u32 *const evl_s = evm_s + 32ull * sstride;                       // [32][sstride]               lambda
    u32 *const ck_g = scratch + BSI_SCRATCH_WORDS * sstride + gtid;
    (void)ck_g;
#define BSI_CK_ONCHIP(S) (BSI_SMEM_SLOTS == 3 || (S) < BSI_SMEM_SLOTS)
#define BSI_CK_PTR(S) (BSI_CK_ONCHIP(S) ? ck_s + (S) * BSI_N * BSI_BLOCK \
                                            : ck_g + (u64)((S) - BSI_SMEM_SLOTS) * BSI_N * sstride)
#define BSI_CK_STRIDE(S) (BSI_CK_ONCHIP(S) ? (u64)BSI_BLOCK : sstride)
    u64 dev_steps = 0, n_without = 0, ref_steps = 0;

    for (;;) {
        const u64 w = bsi_grab_word(&ctr->next_word);
        if (__all_sync(0xFFFFFFFFu, w >= n_words)) break;
        const bool active = w < n_words;
        BSI_FOR_NODES(BSI_DECL_A)
        BSI_FOR_NODES(BSI_DECL_B)
        BSI_FOR_FILTER(BSI_DECL_K0)
        BSI_FOR_FILTER(BSI_DECL_K1)
        BSI_FOR_FILTER(BSI_DECL_K2)
        u32 valid = 0u, F = 0u, alive = 0u, early = 0u;
        u32 n_ev = 0u, j_end = 0u;               // early events so far; the time the loop ended at
        u32 st0 = 0, st1 = 0, st2 = 0;           // checkpoint time held by each slot

        if (active) {
            // ---- x_0 .. x_lp ----------------------------------------------------------------
            if (DENSE) {
                const u64 W = word0 + w;
                valid = bsi_valid_mask(W, idx_begin, idx_end);
                BSI_INIT_DENSE(a, W)
            } else {
                u64 cw_[32][BSI_KW];
#pragma unroll 1
                for (u32 l = 0; l < 32; ++l) {
                    u64 q = (w << 5) + l;
                    u64 c_[BSI_KW];
#pragma unroll
                    for (int i = 0; i < BSI_KW; ++i) c_[i] = 0ull;
                    if (q < n_in) {
                        u64 lp_lane_ = 0;
                        BSI_LANE_SETUP(l, in_queue[q], c_)
                        (void)lp_lane_;
                        valid |= 1u << l;
                    }
#pragma unroll
                    for (int i = 0; i < BSI_KW; ++i) cw_[l][i] = c_[i];
                }
                BSI_PACK(a, cw_)
            }
            BSI_PREFIX(BSI_LP_ORIGIN)
            BSI_FOR_FILTER(BSI_ZERO_K)

            alive = valid;
            u32 j = 0, ci = 0, armed = 0;

            // Cold path, taken when some lane passes a register filter.  Stage 2 refines the
            // candidate lanes against a dozen more words of the checkpoint in shared memory
            // (inline, a handful of LDS); only if a lane survives that too is the state dumped to
            // local memory for the full compare against the shared-memory checkpoint (rolled
            // loops: the hot loop's register allocation only sees the state words, the filter
            // words and a few scalars).
#define BSI_FULL(XS, J, S, ST, ACC)                                                  \
    {                                                                                \
        u32 cand_ = ~(ACC) & alive;                                                  \
        if (cand_) {                                                                 \
            BSI_COUNT(1)                                                             \
            u32 m_ = bsi_slow_match(XS, BSI_CK_PTR(S), BSI_CK_STRIDE(S), cand_);       \
            if (m_) { BSI_MATCHED(m_, J, S, ST) }                                    \
        }                                                                            \
    }
            // lanes M matched the checkpoint of slot S (taken at time ST) at time J
#define BSI_MATCHED(M, J, S, ST)                                                     \
    {                                                                                \
        F |= (M);                                                                    \
        alive &= ~(M);                                                               \
        if ((S) > 0 || (ST) >= sched.first_final) {                                  \
            bsi_record((M), (J) - (ST), (ST), lam_s, c_s, sstride, (S));             \
        } else {                                                                     \
            early |= (M);                                                            \
            evm_s[n_ev * sstride] = (M);                                             \
            evl_s[n_ev * sstride] = (J) - (ST);                                      \
            ++n_ev;                                                                  \
        }                                                                            \
    }
#if BSI_STATS
#define BSI_COUNT(k) atomicAdd(&ctr->dbg[k], 1ull);
#else
#define BSI_COUNT(k)
#endif
#define BSI_STAGE2(X, S, ACC)                                                        \
    if (ACC != 0xFFFFFFFFu) {                                                        \
        BSI_COUNT(0)                                                                 \
        const u32 *p2_ = BSI_CK_PTR(S);                                              \
        BSI_CMP_FILTER2A_MEM(ACC, X, p2_, BSI_CK_STRIDE(S))                          \
        if (ACC != 0xFFFFFFFFu) {                                                    \
            BSI_CMP_FILTER2B_MEM(ACC, X, p2_, BSI_CK_STRIDE(S))                      \
        }                                                                            \
    }
            // register filters of state X against the armed slots: p_s = lanes that CANNOT match
#define BSI_FILT0(X) {}
#define BSI_FILT1(X)                                                                 \
    {                                                                                \
        BSI_CMP_FILTER1(p0, alive, X, k0_)                                           \
    }
#define BSI_FILT2(X)                                                                 \
    {                                                                                \
        BSI_CMP_FILTER1(p0, alive, X, k0_)                                           \
        BSI_CMP_FILTER1(p1, alive, X, k1_)                                           \
    }
#define BSI_FILT3(X)                                                                 \
    {                                                                                \
        BSI_CMP_FILTER1(p0, alive, X, k0_)                                           \
        BSI_CMP_FILTER1(p1, alive, X, k1_)                                           \
        BSI_CMP_FILTER1(p2, alive, X, k2_)                                           \
    }
#define BSI_ANY0 false
#define BSI_ANY1 (p0 != 0xFFFFFFFFu)
#define BSI_ANY2 ((p0 & p1) != 0xFFFFFFFFu)
#define BSI_ANY3 (bsi_lop3<0x80>(p0, p1, p2) != 0xFFFFFFFFu)
            // slow handlers: stage 2, then the full compare of the surviving candidates
#define BSI_SLOW0(X, J) {}
            // (one slot armed -- the early phase, where whole words close lane by lane: the full
            //  compare runs straight from the state registers, one LDS + one lop3 per node, no dump)
#define BSI_SLOW1(X, J)                                                              \
    {                                                                                \
        BSI_STAGE2(X, 0, p0)                                                         \
        if (BSI_ANY1) {                                                              \
            BSI_COUNT(1)                                                             \
            u32 acc_ = p0 | ~alive;                                                  \
            BSI_CMP_ALL_MEM(acc_, X, BSI_CK_PTR(0), BSI_CK_STRIDE(0))                \
            const u32 m_ = ~acc_;                                                    \
            if (m_) { BSI_MATCHED(m_, J, 0u, st0) }                                  \
        }                                                                            \
    }
#define BSI_SLOW2(X, J)                                                              \
    {                                                                                \
        BSI_STAGE2(X, 0, p0)                                                         \
        BSI_STAGE2(X, 1, p1)                                                         \
        if (BSI_ANY2) {                                                              \
            u32 xs_[BSI_N];                                                          \
            BSI_DUMP(X, xs_)                                                         \
            BSI_FULL(xs_, J, 0, st0, p0)                                             \
            BSI_FULL(xs_, J, 1, st1, p1)                                             \
        }                                                                            \
    }
#define BSI_SLOW3(X, J)                                                              \
    {                                                                                \
        BSI_STAGE2(X, 0, p0)                                                         \
        BSI_STAGE2(X, 1, p1)                                                         \
        BSI_STAGE2(X, 2, p2)                                                         \
        if (BSI_ANY3) {                                                              \
            u32 xs_[BSI_N];                                                          \
            BSI_DUMP(X, xs_)                                                         \
            BSI_FULL(xs_, J, 0, st0, p0)                                             \
            BSI_FULL(xs_, J, 1, st1, p1)                                             \
            BSI_FULL(xs_, J, 2, st2, p2)                                             \
        }                                                                            \
    }
            // n_ steps from time j, comparing every new state with the armed slots.  The inner
            // loop is the hot path: two updates, their register filters and ONE taken branch;
            // a filter pass leaves the loop for the handler below and re-enters afterwards.
#define BSI_SEGMENT(A)                                                               \
    {                                                                                \
        u32 n_ = seg_end - j;                                                        \
        u32 p0 = 0xFFFFFFFFu, p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;                    \
        for (; n_ >= 2u && alive; n_ -= 2u, j += 2u) {                               \
            BSI_UPDATE(b, a)                                                         \
            BSI_FILT##A(b)                                                           \
            if (__builtin_expect(BSI_ANY##A, 0)) goto slow_b##A;                     \
          back_b##A:                                                                 \
            BSI_UPDATE(a, b)                                                         \
            BSI_FILT##A(a)                                                           \
            if (__builtin_expect(BSI_ANY##A, 0)) goto slow_a##A;                     \
          back_a##A:;                                                                \
        }                                                                            \
        if (n_ == 1u && alive) {                                                     \
            BSI_UPDATE(b, a)                                                         \
            BSI_FILT##A(b)                                                           \
            if (BSI_ANY##A) { BSI_SLOW##A(b, j + 1u) }                               \
            BSI_FOR_NODES(BSI_COPY_AB)                                               \
            j += 1u;                                                                 \
            n_ = 0u;                                                                 \
        }                                                                            \
        if (false) {                                                                 \
          slow_b##A:                                                                 \
            BSI_SLOW##A(b, j + 1u)                                                   \
            goto back_b##A;                                                          \
          slow_a##A:                                                                 \
            BSI_SLOW##A(a, j + 2u)                                                   \
            goto back_a##A;                                                          \
        }                                                                            \
        (void)p0; (void)p1; (void)p2;                                                \
        dev_steps += (seg_end - j0_) - n_;                                           \
    }
            while (alive) {
                // checkpoint due at time j: state a (already compared) goes into its slot
                if (ci < sched.n_ck && sched.ck_time[ci] == j) {
                    const u32 s = sched.ck_slot[ci];
                    if (BSI_CK_ONCHIP(s)) {
                        u32 *p_ = ck_s + s * BSI_N * BSI_BLOCK;
                        BSI_STORE_ALL(a, p_, BSI_BLOCK)
                    } else {
                        u32 *p_ = ck_g + (u64)(s - BSI_SMEM_SLOTS) * BSI_N * sstride;
                        BSI_STORE_ALL(a, p_, sstride)
                    }
                    if (s == 0u) { BSI_FOR_FILTER(BSI_SAVE_K0) st0 = j; }
                    else if (s == 1u) { BSI_FOR_FILTER(BSI_SAVE_K1) st1 = j; }
                    else { BSI_FOR_FILTER(BSI_SAVE_K2) st2 = j; }
                    if (armed < s + 1u) armed = s + 1u;
                    ++ci;
                    continue;
                }
                if (j >= sched.end) break;
                const u32 seg_end = ci < sched.n_ck ? sched.ck_time[ci] : sched.end;
                const u32 j0_ = j;
                if (armed == 0u) BSI_SEGMENT(0)
                else if (armed == 1u) BSI_SEGMENT(1)
                else if (armed == 2u) BSI_SEGMENT(2)
                else BSI_SEGMENT(3)
            }
#undef BSI_SEGMENT
#undef BSI_FULL
#undef BSI_STAGE2
#undef BSI_MATCHED
            j_end = j;
        }

        // ---- results ------------------------------------------------------------------------------
        const u32 matched = F;
        // early matches of a dense word -> early-word queue: the word's state now (time j_end) and
        // the event list
        if (DENSE && early_cap) {
            const u64 e = bsi_reserve(&ctr->n_early, early ? 1u : 0u);
            if (early) {
                const BsiEarlyEntry q = bsi_early_entry(bsi_early_view(early_base, early_cap), e);
                q.word[0] = word0 + w;
                q.mask[0] = early;
                q.info[0] = n_ev | (j_end << 8);
                // (unrolled with predication: the event loads are L2 round trips, in flight together)
#pragma unroll 8
                for (u32 k = 0; k < 32; ++k)
                    if (k < n_ev) {
                        q.evm[k * 32u] = evm_s[k * sstride];
                        q.evl[k * 32u] = evl_s[k * sstride];
                    }
                BSI_STORE_ALL(a, q.y, 32)
                F &= ~early;
                early = 0u;
            }
        }
        // other matched lanes -> found queue, open lanes -> retry queue / "no attractor"
        const u32 n_f = __popc(F);
        const u32 U = valid & ~matched;
        const u32 n_u = __popc(U);
        u64 base_f = bsi_reserve(&ctr->n_found, n_f);
        if (n_f) {
            // early matches without an early-word queue (retry rounds): per-lane records with the
            // lane's state now as the on-cycle state (y = x_c with c = j_end >= mu, c0 = 0) -- one
            // transpose pass over the word's state for the state codes
            u64 cw_[32][BSI_KW];
            if (early) {
                BSI_UNPACK(cw_, a)
#pragma unroll 1
                for (u32 k = 0; k < n_ev; ++k) {
                    u32 m = evm_s[k * sstride];
                    const u32 lam = evl_s[k * sstride];
                    while (m) {
                        const u32 l = __ffs(m) - 1;
                        m &= m - 1;
                        lam_s[l * sstride] = lam;
                        c_s[l * sstride] = j_end;
                        c_s[(32 + l) * sstride] = 0u;
                    }
                }
            }
            // Lanes that matched a FINAL checkpoint take y (that checkpoint) and z (the one before it) out
            // of the bit-sliced checkpoint words.  One or two lanes: bit by bit (64 single-bit moves per
            // state).  More -- the closing-heavy chaotic regime, a handful of lanes in every word -- : the
            // checkpoints these lanes need go through 32x32 transposes once, for all 32 lanes together
            // (ncu, seed-2 network: the bit-by-bit form was 12 % of the kernel's instructions).
            u32 defer = 0u, need = 0u;
            {
                u32 m = F;
                while (m) {
                    const u32 l = __ffs(m) - 1;
                    m &= m - 1;
                    const u32 c0 = c_s[(32 + l) * sstride];
                    if (c0 & BSI_DEFERRED) {
                        const u32 slot = c0 & 3u;
                        defer |= 1u << l;
                        need |= (1u << slot) | (slot ? 1u << (slot - 1u) : 0u);
                    }
                }
            }
            // (BSI_TUNE 32: from the first lane on, 64: never -- the tests run both forms on the same inputs)
            const bool by_transpose = (BSI_TUNE & 64) ? false : (u32)__popc(defer) >= ((BSI_TUNE & 32) ? 1u : 3u);
            u64 ckw_[3][32][BSI_KW];
            if (by_transpose) {
#pragma unroll 1
                for (u32 s = 0; s < 3u; ++s) {
                    if (!((need >> s) & 1u)) continue;
                    const u32 *p = BSI_CK_PTR(s);
                    const u64 st = BSI_CK_STRIDE(s);
#pragma unroll 1
                    for (u32 b = 0; b < (u32)((BSI_N + 31) / 32); ++b) {
                        unsigned int A_[32];
#pragma unroll 8
                        for (u32 k = 0; k < 32; ++k) {
                            const u32 node = 32u * b + 31u - k;       // (the transpose reverses both indices)
                            A_[k] = node < (u32)BSI_N ? p[node * st] : 0u;
                        }
                        bsi_transpose32_call(A_);
#pragma unroll 8
                        for (u32 k = 0; k < 32; ++k) {
                            if (b & 1u) ckw_[s][31u - k][b >> 1] |= (u64)A_[k] << 32;
                            else ckw_[s][31u - k][b >> 1] = (u64)A_[k];
                        }
                    }
                }
            }
            u32 m = F;
            while (m) {
                u32 l = __ffs(m) - 1;
                m &= m - 1;
                BsiFound r;
                r.idx = DENSE ? (((word0 + w) << 5) + l) : in_queue[(w << 5) + l];
                r.lam = lam_s[l * sstride];
                r.c = c_s[l * sstride];
                r.c0 = c_s[(32 + l) * sstride];
                r.pad = 0u;
                if (r.c0 & BSI_DEFERRED) {
                    // y = the matched final checkpoint, z = the final checkpoint before it
                    const u32 slot = r.c0 & 3u;
                    const u32 *py = BSI_CK_PTR(slot);
                    const u64 sy = BSI_CK_STRIDE(slot);
                    const u32 *pz = BSI_CK_PTR(slot ? slot - 1u : 0u);
                    const u64 sz = BSI_CK_STRIDE(slot ? slot - 1u : 0u);
                    r.c0 = slot == 0u ? 0u : (slot == 1u ? st0 : st1);
                    if (by_transpose) {
#pragma unroll
                        for (int k = 0; k < BSI_KW; ++k) {
                            const u64 y = ckw_[slot][l][k];
                            r.y[k] = y; r.key[k] = y; r.z[k] = slot ? ckw_[slot - 1u][l][k] : 0ull;
                        }
                    } else
#pragma unroll
                    for (int k = 0; k < BSI_KW; ++k) {
                        u64 y = 0ull, z = 0ull;
#pragma unroll 8
                        for (int i = 64 * k; i < BSI_N && i < 64 * k + 64; ++i)
                            y |= (u64)((py[i * sy] >> l) & 1u) << (i & 63);
                        if (slot) {
#pragma unroll 8
                            for (int i = 64 * k; i < BSI_N && i < 64 * k + 64; ++i)
                                z |= (u64)((pz[i * sz] >> l) & 1u) << (i & 63);
                        }
                        r.y[k] = y; r.key[k] = y; r.z[k] = z;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < BSI_KW; ++k) { r.y[k] = cw_[l][k]; r.key[k] = cw_[l][k]; r.z[k] = 0ull; }
                }
                found_q[base_f++] = r;
            }
        }
        if (final_round) {
            n_without += n_u;
            ref_steps += (u64)n_u * stop_t;      // model.py:201: the loop ran out at t = max_t
        } else {
            u64 base_u = bsi_reserve(&ctr->n_retry, n_u);
            u32 m = U;
            while (m) {
                u32 l = __ffs(m) - 1;
                m &= m - 1;
                retry_q[base_u++] = DENSE ? (((word0 + w) << 5) + l) : in_queue[(w << 5) + l];
            }
        }
    }
    if (dev_steps) atomicAdd(&ctr->device_steps, dev_steps);
    if (n_without) atomicAdd(&ctr->n_without, n_without);
    if (ref_steps) atomicAdd(&ctr->steps, ref_steps);
}

// The translation unit is compiled once per BSI_UNIT (in parallel, see bsi.cpp): 1 = dense
// trajectory kernel, 2 = queue trajectory kernel, 3 = resolution kernels, 4 = everything else;
// 5 = the latency-oriented resolution kernels, 6 = the early-word resolution kernel; 0 = all of it (the ahead-of-time nvcc build check).
#if BSI_UNIT == 0 || BSI_UNIT == 1
extern "C" __global__ void __launch_bounds__(BSI_BLOCK, BSI_MIN_BLOCKS)
bsi_attract_dense(u64 word0, u64 n_words, u64 idx_begin, u64 idx_end, const BsiSched sched,
                  int final_round, u64 stop_t, BsiFound *found_q, u64 *retry_q, u32 *scratch,
                  void *early_q, u64 early_cap, BsiCounters *ctr) {
    bsi_main_body<true>(word0, n_words, idx_begin, idx_end, nullptr, 0, sched, final_round, stop_t,
                        found_q, retry_q, scratch, early_q, early_cap, ctr);
}
#endif

#if BSI_UNIT == 0 || BSI_UNIT == 2
extern "C" __global__ void __launch_bounds__(BSI_BLOCK, BSI_MIN_BLOCKS)
bsi_attract_queue(const u64 *in_queue, u64 n_in, const BsiSched sched, int final_round, u64 stop_t,
                  BsiFound *found_q, u64 *retry_q, u32 *scratch, BsiCounters *ctr) {
    bsi_main_body<false>(0, (n_in + 31) >> 5, 0, 0, in_queue, n_in, sched, final_round, stop_t,
                         found_q, retry_q, scratch, nullptr, 0, ctr);
}
#endif

// ---------------------------------------------------------------------------------------------
// attract: exact resolution of the matched lanes
// ---------------------------------------------------------------------------------------------
#if BSI_TUNE & 1
#undef BSI_TRANSPOSE32
#define BSI_TRANSPOSE32(A) bsi_transpose32_call(A)
#endif
#if BSI_UNIT == 0 || BSI_UNIT == 3 || BSI_UNIT == 5 || BSI_UNIT == 6

// Which records of the found queue form "word" w of the resolution kernels: words come in groups
// of 32 (one warp), a group owns 1024 consecutive records, and lane l of the group's t-th word is
// record 1024 g + 32 l + t -- so that, for a fixed lane, the 32 threads of a warp read 32
// CONSECUTIVE records (coalesced 48-byte structs) instead of records 1.5 KB apart.
#define BSI_Q(w, l) ((((w) & ~31ull) << 5) + ((u64)(l) << 5) + ((w) & 31ull))
// number of words to hand out for n records (whole groups)
#define BSI_QWORDS(n) ((((n) + 1023ull) >> 10) << 5)

// Canonical key = lexicographic minimum over the cycle with node N-1 most significant, i.e. the
// minimum state code (reference model.py:140-149 / attract.py:296).  Lane l of thread w walks
// the cycle of found_q[32w + l] from y; walking more than lambda steps only revisits cycle
// states, so all lanes of a word share the step count max lambda - 1.  lambda == 1: key = y.
// latency mode = both queues are short (see the comment above bsi_resolve_mu)
__device__ __forceinline__ bool bsi_latency_mode(u64 n_found, const BsiCounters *ctr, u64 few_words) {
    return ((n_found + 31) >> 5) <= few_words && ctr->n_early <= few_words;
}

// in: cw_[32][KW] = on-cycle state codes of the valid lanes; out: their canonical keys, in place
__device__ __forceinline__ void bsi_key_core(u64 (*cw_)[BSI_KW], u32 valid, u32 lam_max) {
    if (lam_max <= 1u) return;                    // fixed points: the state is its own key
    BSI_FOR_NODES(BSI_DECL_A)
    BSI_FOR_NODES(BSI_DECL_B)
    BSI_FOR_NODES(BSI_DECL_C)
    {
        BSI_PACK(a, cw_)
        BSI_FOR_NODES(BSI_COPY_CA)
        // lt_: lanes where X < c (decided at the most significant differing node)
#define BSI_LEX_NODE(i)                                                           \
    lt_ |= bsi_lop3<0x08>(X_(i), c##i, eq_);                                      \
    eq_ = bsi_lop3<0x90>(eq_, X_(i), c##i);
#define BSI_SEL_NODE(i) c##i = bsi_lop3<0xCA>(lt_, X_(i), c##i);
        // 0x08: ~x & c & eq;  0x90: eq & ~(x ^ c);  0xCA: lt ? x : c
#define BSI_LEX_STEP                                                              \
    {                                                                             \
        u32 lt_ = 0u, eq_ = valid;                                                \
        BSI_LEX_CHUNKS                                                            \
        if (lt_) { BSI_FOR_NODES(BSI_SEL_NODE) }                                  \
    }
        for (u32 k = 1; k < lam_max; k += 2) {
            BSI_UPDATE(b, a)
#define X_(i) b##i
            BSI_LEX_STEP
#undef X_
            BSI_UPDATE(a, b)
#define X_(i) a##i
            BSI_LEX_STEP
#undef X_
        }
#undef BSI_LEX_STEP
#undef BSI_SEL_NODE
#undef BSI_LEX_NODE
        BSI_UNPACK(cw_, c)
    }
}

// Store the orbit of `key` (length lam >= 2, reserved at `base` / `slot` by bsi_dict_claim), enter
// its states into the dictionary and publish it.  The thread's 32 lanes walk 32 different phases of
// the SAME orbit: a first pass brings lane j to phase j * s (s = ceil(lam / 32); lanes drop out of the
// advancing set one by one), a second pass of s steps then yields 32 states per step, so the 32x32
// transposes that turn bit-sliced words into state codes are paid s times and not lam times.  Only
// when the whole orbit is in memory are its states entered (read back from there), the key last.
__device__ __noinline__ void bsi_dict_store_orbit(const BsiDict &d, const u64 *key, u32 lam, u32 base) {
    BSI_FOR_NODES(BSI_DECL_A)
    BSI_FOR_NODES(BSI_DECL_B)
#define BSI_BCAST_NODE(i) a##i = 0u - (u32)((key[(i) >> 6] >> ((i) & 63)) & 1ull);
    BSI_FOR_NODES(BSI_BCAST_NODE)
#undef BSI_BCAST_NODE
    const u32 s = (lam + 31u) / 32u;
#define BSI_HOLD_AB(i) a##i = bsi_lop3<0xCA>(go_, b##i, a##i);
    for (u32 k = 0; k < 31u * s; ++k) {
        const u32 go_ = ~((2u << (k / s)) - 1u);          // lanes j with j * s > k
        BSI_UPDATE(b, a)
        BSI_FOR_NODES(BSI_HOLD_AB)
    }
#undef BSI_HOLD_AB
    u64 cw_[32][BSI_KW];
    for (u32 q = 0; q < s; ++q) {
        BSI_UNPACK(cw_, a)
#pragma unroll 1
        for (u32 j = 0; j < 32u; ++j) {
            const u32 ph = j * s + q;
            if (ph >= 1u && ph < lam) {
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) d.orbit[(u64)(base + ph) * BSI_KW + i] = cw_[j][i];
            }
        }
        BSI_UPDATE(b, a)
        BSI_FOR_NODES(BSI_COPY_AB)
    }
    __threadfence();
}
// enter the states of a stored orbit, phases first, first + stride, ... (the key state, phase 0, is there)
__device__ __forceinline__ void bsi_dict_put_orbit(const BsiDict &d, u32 lam, u32 base, u32 first, u32 stride) {
#pragma unroll 1
    for (u32 ph = first; ph < lam; ph += stride) {
        if (ph == 0u) continue;
        u64 st[BSI_KW];
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) st[i] = bsi_ld_u64(&d.orbit[(u64)(base + ph) * BSI_KW + i]);
        bsi_dict_put(d, st, base + ph, base, lam);
    }
}
__device__ __noinline__ void bsi_dict_learn(const BsiDict &d, const u64 *key, u32 lam, u32 base, u32 slot) {
    bsi_dict_store_orbit(d, key, lam, base);
    bsi_dict_put_orbit(d, lam, base, 1u, 1u);
    bsi_dict_publish(d, slot);
}

__device__ __forceinline__ void bsi_key_word(BsiFound *found_q, u64 n_found, u64 w, const BsiDict &dict) {
    u32 valid = 0u, lam_max = 0u, miss = 0u;
    u64 cw_[32][BSI_KW];
    u32 lam_[32];
#pragma unroll 1
    for (u32 l = 0; l < 32; ++l) {
        const u64 q = BSI_Q(w, l);
        const bool in = q < n_found;
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) cw_[l][i] = in ? found_q[q].y[i] : 0ull;
        lam_[l] = in ? found_q[q].lam : 0u;
        if (in) {
            valid |= 1u << l;
            lam_max = max(lam_max, lam_[l]);
        }
    }
    if (lam_max <= 1u) return;
    // lanes whose cycle is already in the dictionary take their key from there
    if (dict.slots) {
#pragma unroll 4
        for (u32 l = 0; l < 32; ++l) {
            if (lam_[l] <= 1u) continue;
            u32 idx, base, lamd;
            if (bsi_dict_find(dict, cw_[l], idx, base, lamd) && lamd == lam_[l]) {
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) found_q[BSI_Q(w, l)].key[i] = bsi_ld_u64(&dict.orbit[(u64)base * BSI_KW + i]);
            } else {
                miss |= 1u << l;
            }
        }
        if (!miss) return;
    } else {
        miss = valid;
    }
    bsi_key_core(cw_, valid, lam_max);
#pragma unroll 1
    for (u32 l = 0; l < 32; ++l) {
        if ((miss >> l) & 1u) {
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) found_q[BSI_Q(w, l)].key[i] = cw_[l][i];
            // (no learning here: this is the latency-oriented path, where every word of the queue is in
            //  flight at once -- nobody would profit before the kernel ends, and the walk of a long
            //  orbit would sit on the critical path of the step)
        }
    }
}

// The dictionary's part of resolving one word, OUT OF LINE (the resolution kernels are big; what a
// word of short cycles never executes must not sit in the instruction cache's way): look the matched
// states ycw[l] of the lanes on cycles up -- eight independent probes in flight at a time -- and, if
// the dictionary knows them ALL, fetch P(c0) = y advanced by adv_s[l * adv_stride] steps and, with
// want_keys, the canonical keys (keys[l]) from the stored orbits; pb receives P(c0) bit-sliced, one
// word per node, dbase_[l] where the lane's orbit is stored (one number per attractor).  Returns
// the lanes it does not know (0: pb / keys / dbase_ are filled).
struct BsiKeep {        // what bsi_finish_word needs of the 32 records, when it runs in the same kernel
    u64 key[32][BSI_KW];
    u32 lam[32], c0[32];
};
__device__ __noinline__ u32 bsi_dict_word(const BsiDict &dict, u64 (*ycw)[BSI_KW], const u32 *lam, const u32 *adv_s,
                                          u32 adv_stride, bool want_keys, u64 (*keys)[BSI_KW], u32 *pb,
                                          u32 *dbase_) {
    const u32 mask_ = (1u << dict.log2_slots) - 1u;
    u32 dphase_[32];
    u32 miss = 0u;
#pragma unroll 1
    for (u32 l0 = 0; l0 < 32; l0 += 8) {
        u32 f_[8], idx_[8], base_[8], lamd_[8];
        u64 st_[8][BSI_KW];
#pragma unroll
        for (u32 k = 0; k < 8; ++k) {
            const BsiDictSlot *s_ = dict.slots + ((u32)bsi_hash_key(ycw[l0 + k]) & mask_);
            bsi_dict_head_cg(s_, f_[k], idx_[k], base_[k], lamd_[k]);
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) st_[k][i] = bsi_ldcg_u64(&s_->st[i]);
        }
#pragma unroll
        for (u32 k = 0; k < 8; ++k) {
            const u32 l = l0 + k;
            if (lam[l] <= 1u) continue;
            bool hit = f_[k] == 2u;
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) hit = hit && st_[k][i] == ycw[l][i];
            // (not in its first slot: the full probe sequence)
            if (!hit) hit = bsi_dict_find(dict, ycw[l], idx_[k], base_[k], lamd_[k]);
            if (hit && lamd_[k] == lam[l]) {
                dbase_[l] = base_[k];
                dphase_[l] = idx_[k] - base_[k];
            } else {
                miss |= 1u << l;
            }
        }
    }
    if (miss) return miss;
    u64 cp_[32][BSI_KW];
#pragma unroll 1
    for (u32 l0 = 0; l0 < 32; l0 += 8) {
#pragma unroll
        for (u32 k = 0; k < 8; ++k) {
            const u32 l = l0 + k;
            const u32 lm = lam[l];
            if (lm > 1u) {
                u32 ph = dphase_[l] + adv_s[l * adv_stride];
                if (ph >= lm) ph -= lm;
                const u64 *po = dict.orbit + (u64)(dbase_[l] + ph) * BSI_KW;
                const u64 *pk = dict.orbit + (u64)dbase_[l] * BSI_KW;
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) {
                    cp_[l][i] = bsi_ldcg_u64(po + i);
                    if (want_keys) keys[l][i] = bsi_ldcg_u64(pk + i);
                }
            } else {
                // (fixed points: P is the state itself, and so is the key)
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) cp_[l][i] = lm ? ycw[l][i] : 0ull;
            }
        }
    }
    {
        BSI_FOR_NODES(BSI_DECL_B)
        BSI_PACK(b, cp_)
#define BSI_STORE_PB(i) pb[i] = b##i;
        BSI_FOR_NODES(BSI_STORE_PB)
#undef BSI_STORE_PB
    }
    return 0u;
}

// Exact mu, then acceptance and aggregation.  With y = x_c on the cycle, let P be the periodic
// extension of the cycle to all times: P(j) = x_j exactly for j >= mu, so for any c0 < mu
//     mu = first j >= c0 with x_j == P(j),   P(c0) = y advanced by (lambda - (c - c0) mod lambda) mod lambda.
// a* walks x_j from x_c0 (the shipped checkpoint state, or the problem index when c0 = 0), b*
// walks P(j); three register sets rotate so that no step needs a copy.
// Reference attract.py:287-300 (mu, trajectory_l), attract.py:374-402 (aggregate).
// The lanes are those of word w of found_q.
// `keep` always receives lambda and c0 of the lanes; with KEEP also their canonical keys (from the
// cycle dictionary when it knows every cycle of the word, else by the walk of bsi_key_core -- the
// walked cycles are then entered into the dictionary).
template <bool WRITE_PAD, bool KEEP>
__device__ __forceinline__ void bsi_mu_word(BsiFound *found_q, u64 n_found, u64 w, u32 *aux_s,
                                            BsiCounters *ctr, BsiKeep *keep,
                                            const BsiDict &dict, bool learn, u32 &valid_out) {
    const u32 tid = threadIdx.x;
    {
        BSI_FOR_NODES(BSI_DECL_A)
        BSI_FOR_NODES(BSI_DECL_B)
        BSI_FOR_NODES(BSI_DECL_C)
        u32 valid = 0u, adv_max = 0u, c_max = 0u, from_idx = 0u, lam_max = 0u;
        // one pass over the 32 records of this word: every field that is needed later goes to
        // (thread-interleaved, hence coalesced) local arrays
        u64 cw_[32][BSI_KW], cz_[32][BSI_KW], cx_[32][BSI_KW];
#pragma unroll 1
        for (u32 l = 0; l < 32; ++l) {
            const u64 q = BSI_Q(w, l);
            u64 x_[BSI_KW];
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) { cw_[l][i] = 0ull; cz_[l][i] = 0ull; x_[i] = 0ull; }
            keep->lam[l] = 0u;
            if (q < n_found) {
                const BsiFound r = KEEP ? bsi_found_stream(found_q + q) : found_q[q];
                valid |= 1u << l;
                const u32 adv = r.lam == 1u ? 0u : (r.lam - (r.c - r.c0) % r.lam) % r.lam;
                aux_s[l * BSI_BLOCK + tid] = adv;
                adv_max = max(adv_max, adv);
                c_max = max(c_max, r.c - r.c0);      // the walk is at most c - c0 steps long
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) cw_[l][i] = r.y[i];
                if (r.c0 == 0u) {
                    from_idx |= 1u << l;
                    u64 lp_lane_ = 0;
                    BSI_LANE_SETUP(l, r.idx, x_)
                    (void)lp_lane_;
                } else {
#pragma unroll
                    for (int i = 0; i < BSI_KW; ++i) cz_[l][i] = r.z[i];
                }
                keep->lam[l] = r.lam;
                keep->c0[l] = r.c0;
                lam_max = max(lam_max, r.lam);
                if (KEEP) {
#pragma unroll
                    for (int i = 0; i < BSI_KW; ++i) keep->key[l][i] = r.y[i];
                }
            }
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) cx_[l][i] = x_[i];
        }
        // ---- lanes on cycles (lambda > 1): does the dictionary know every one of them? -----------
        u64 (*const ycw_)[BSI_KW] = cw_;                          // the matched states as codes
        u32 miss = 0u;
        bool use_dict = false;
        u32 pb_[BSI_N];
        // (short cycles are walked: the key walk and the advance of a word take lam_max - 1 steps
        //  each, which is cheaper than 32 lookups and two 32x32 transposes up to about 8 states)
        if (dict.slots && lam_max > BSI_DICT_MIN_LAM) {
            u32 dbase_[32];
            miss = bsi_dict_word(dict, ycw_, keep->lam, aux_s + tid, BSI_BLOCK, KEEP, keep->key, pb_, dbase_);
            use_dict = miss == 0u;
        }
        valid_out = valid;
        if (lam_max > 1u) { if (use_dict) { BSI_COUNT(2) } else { BSI_COUNT(3) } }
        if (use_dict) {
            // keys and P(c0) came straight from the stored orbits
            BSI_LOAD_MEM(b, pb_, 1)
        } else {
            // (throughput paths) the canonical keys, computed here: no separate pass over the records
            if (KEEP) {
                bsi_key_core(keep->key, valid, lam_max);
                if (dict.slots && miss && learn) {
                    // the walked cycles go into the dictionary (once per attractor: the claim decides)
#pragma unroll 1
                    for (u32 l = 0; l < 32; ++l) {
                        u32 base, slot;
                        if (((miss >> l) & 1u) && bsi_dict_claim(dict, keep->key[l], keep->lam[l], base, slot))
                            bsi_dict_learn(dict, keep->key[l], keep->lam[l], base, slot);
                    }
                }
            }
            BSI_PACK(b, cw_)
            // ---- b <- P(c0): y advanced by each lane's own number of steps.  ONE plain walk of adv_max
            //      steps (b -> c -> b, no per-step hold of the lanes that are done); the lanes whose advance
            //      is k are latched into a -- free until x_c0 is built below -- when the walk is k steps
            //      old: an event costs one pass over the nodes, and a word has at most 32 of them ----------
#define BSI_LATCH_A(i) a##i = bsi_lop3<0xCA>(m_, X_(i), a##i);
#define BSI_ADV_EVENT(K)                                                          \
    {                                                                             \
        u32 m_ = 0u, nx_ = 0xFFFFFFFFu;                                           \
        _Pragma("unroll 1")                                                       \
        for (u32 l = 0; l < 32; ++l) {                                            \
            if (!((valid >> l) & 1u)) continue;                                   \
            const u32 adv = aux_s[l * BSI_BLOCK + tid];                           \
            if (adv == (K)) m_ |= 1u << l;                                        \
            else if (adv > (K)) nx_ = min(nx_, adv);                              \
        }                                                                         \
        next_ev = nx_;                                                            \
        BSI_FOR_NODES(BSI_LATCH_A)                                                \
    }
            if (adv_max) {
                u32 next_ev = 0u;
#define X_(i) b##i
                BSI_FOR_NODES(BSI_ZERO_A)
                BSI_ADV_EVENT(0u)                        // (advance 0: P(c0) = y)
#undef X_
                for (u32 k = 1; k <= adv_max; k += 2) {
                    BSI_UPDATE(c, b)
#define X_(i) c##i
                    if (k == next_ev) BSI_ADV_EVENT(k)
#undef X_
                    if (k + 1u > adv_max) break;
                    BSI_UPDATE(b, c)
#define X_(i) b##i
                    if (k + 1u == next_ev) BSI_ADV_EVENT(k + 1u)
#undef X_
                }
#define BSI_COPY_BA(i) b##i = a##i;
                BSI_FOR_NODES(BSI_COPY_BA)
#undef BSI_COPY_BA
            }
#undef BSI_ADV_EVENT
#undef BSI_LATCH_A
        }
        // ---- a <- x_c0: from the problem index (c0 = 0) or the shipped checkpoint state -----------
        {
            // the prefix macro works on (a, b): park P(c0) in c meanwhile
#define BSI_COPY_CB(i) c##i = b##i;
#define BSI_COPY_BC(i) b##i = c##i;
#define BSI_MASK_A(i) a##i &= from_idx;
#define BSI_OR_AB(i) a##i |= b##i;
            BSI_FOR_NODES(BSI_COPY_CB)
            BSI_FOR_NODES(BSI_ZERO_A)
            if (from_idx) {
                BSI_PACK(a, cx_)
                BSI_PREFIX(BSI_LP_ORIGIN)
                BSI_FOR_NODES(BSI_MASK_A)
            }
            if (valid & ~from_idx) {
                BSI_PACK(b, cz_)
                BSI_FOR_NODES(BSI_OR_AB)
            }
            BSI_FOR_NODES(BSI_COPY_BC)
#undef BSI_OR_AB
#undef BSI_MASK_A
#undef BSI_COPY_CB
#undef BSI_COPY_BC
        }
        // ---- lock-step: mu = first j with x_j == P(j) -------------------------------------------
        u32 alive = valid;
#define BSI_LOCK_CMP(A, B, J)                                                     \
    {                                                                             \
        u32 p_ = ~alive, q_ = 0u;                                                 \
        BSI_CMP_FILTER(p_, q_, A, B)                                              \
        p_ |= q_;                                                                 \
        if (p_ != 0xFFFFFFFFu) {                                                  \
            BSI_CMP_REST(p_, A, B)                                                \
            u32 m_ = ~p_;                                                         \
            alive &= ~m_;                                                         \
            while (m_) {                                                          \
                u32 l_ = __ffs(m_) - 1;                                           \
                m_ &= m_ - 1;                                                     \
                aux_s[l_ * BSI_BLOCK + tid] = (J);                                \
            }                                                                     \
        }                                                                         \
    }
#define BSI_LOCK(A, B, Tn, J)                                                     \
    {                                                                             \
        BSI_LOCK_CMP(A, B, J)                                                     \
        if (alive) { BSI_UPDATE(Tn, A) BSI_UPDATE(A, B) }                         \
    }
        if (lam_max <= 1u) {
            // every lane ends in a fixed point: P is constant (= b), only x moves (a <-> c)
            u32 j = 0;
            while (alive && j <= c_max) {
                BSI_LOCK_CMP(a, b, j)
                if (!alive || j + 1 > c_max) break;
                BSI_UPDATE(c, a)
                BSI_LOCK_CMP(c, b, j + 1)
                if (!alive || j + 2 > c_max) break;
                BSI_UPDATE(a, c)
                j += 2;
            }
        } else {
            u32 j = 0;
            while (alive && j <= c_max) {
                BSI_LOCK(a, b, c, j)
                if (!alive || j + 1 > c_max) break;
                BSI_LOCK(c, a, b, j + 1)
                if (!alive || j + 2 > c_max) break;
                BSI_LOCK(b, c, a, j + 2)
                j += 3;
            }
        }
#undef BSI_LOCK
#undef BSI_LOCK_CMP
        if (alive) atomicExch(&ctr->internal_error, 1ull);
        // mu - c0 stays in aux_s; for bsi_resolve_finish (another kernel) mu goes into the record
        if (WRITE_PAD) {
#pragma unroll 1
            for (u32 l = 0; l < 32; ++l)
                if ((valid >> l) & 1u) found_q[BSI_Q(w, l)].pad = found_q[BSI_Q(w, l)].c0 + aux_s[l * BSI_BLOCK + tid];
        }

    }
}

// Acceptance (mu + lambda <= T, lambda <= max_attractor_l; reference attract.py:287-300) and
// aggregation into the attractor table (attract.py:374-402) of the 32 lanes of word w; lanes of
// one thread with equal keys are added once.
template <bool FROM_KEEP, bool WARP_AGG>
__device__ __forceinline__ void bsi_finish_word(const BsiFound *found_q, u64 n_found, u64 w, u64 T,
                                                u64 max_attractor_l, u64 stop_t_cap, BsiSlot *table,
                                                u32 log2_slots, BsiCounters *ctr, u64 &steps,
                                                u64 &n_acc, u64 &n_rej, const u32 *aux_s,
                                                BsiKeep *keep, u32 lanes) {
    {
        u32 todo = 0u;
        if (FROM_KEEP) {
            // same kernel as the mu walk: keys / lambda / c0 were kept, mu - c0 is in aux_s
            todo = lanes;
        } else {
            // another kernel: one pass over the records; mu itself is in the pad field
#pragma unroll 1
            for (u32 l = 0; l < 32; ++l) {
                const u64 q = BSI_Q(w, l);
                if (q < n_found) {
                    const BsiFound r = found_q[q];
                    todo |= 1u << l;
                    keep->lam[l] = r.lam;
                    keep->c0[l] = r.pad;
#pragma unroll
                    for (int i = 0; i < BSI_KW; ++i) keep->key[l][i] = r.key[i];
                }
            }
        }
        // ---- acceptance + aggregation (lanes of one thread with equal keys are added once) ------
        // (WARP_AGG: every thread of the warp is here, idle ones with todo == 0, and the loop is
        // warp-uniform so that threads holding the same key can be summed before the atomics)
        // acceptance of every lane first, each lane on its own (independent iterations: their local-
        // and shared-memory loads overlap, where a loop over "the remaining lanes with this key" makes
        // every load wait for the one before): ok = lanes whose attractor counts, tl_ = their
        // trajectory lengths
        u32 ok = 0u;
        u64 tl_[32];
#pragma unroll 4
        for (u32 l = 0; l < 32; ++l) {
            tl_[l] = 0ull;
            if (!((todo >> l) & 1u)) continue;
            const u64 mu = (u64)keep->c0[l] + (FROM_KEEP ? aux_s[l * BSI_BLOCK + threadIdx.x] : 0u);
            const u64 lam = keep->lam[l];
            if (mu + lam <= T) {                         // the loop stops at t = mu + lambda (model.py:201)
                steps += BSI_LP_ORIGIN + mu + lam;
                if (lam <= max_attractor_l) {            // attract.py:294
                    tl_[l] = BSI_LP_ORIGIN + mu;         // trajectory_l, attract.py:298
                    ok |= 1u << l;
                } else {
                    ++n_rej;
                }
            } else {                                     // closes only after max_t
                steps += stop_t_cap;
                ++n_rej;
            }
        }
        n_acc += __popc(ok);
        todo &= ok;                                      // rejected lanes add nothing to the table
        // then one group per distinct key
        for (;;) {
            if (WARP_AGG ? !__any_sync(0xFFFFFFFFu, todo != 0u) : !todo) break;
            const u32 l0 = todo ? __ffs(todo) - 1 : 0u;
            u64 key0[BSI_KW];
#pragma unroll
            for (int i = 0; i < BSI_KW; ++i) key0[i] = todo ? keep->key[l0][i] : 0ull;
            const u32 lam0 = todo ? keep->lam[l0] : 0u;
            u64 n = 0, s1 = 0, s2 = 0;
            if (todo) {
                u32 grp = 0u;
#pragma unroll 8
                for (u32 l = 0; l < 32; ++l) {
                    bool same = true;
#pragma unroll
                    for (int i = 0; i < BSI_KW; ++i) same = same && (keep->key[l][i] == key0[i]);
                    grp |= (same ? 1u : 0u) << l;
                }
                grp &= todo;
                todo &= ~grp;
                n = __popc(grp);
#pragma unroll 8
                for (u32 l = 0; l < 32; ++l) {
                    const u64 tl = ((grp >> l) & 1u) ? tl_[l] : 0ull;
                    s1 += tl;
                    s2 += tl * tl;
                }
            }
            if (WARP_AGG) {
                __syncwarp();
                bsi_warp_table_add(table, log2_slots, key0, n, s1, s2, lam0, ctr);
            } else if (n) {
                bsi_table_add(table, log2_slots, key0, n, s1, s2, lam0, ctr);
            }
        }
    }
}

// Two ways to run the resolution, chosen ON THE DEVICE from the queue lengths (the host launches
// all the kernels; those that do not apply return at once):
//  * many matched lanes (> few_words words in either queue): throughput matters -- bsi_resolve_mu reads every
//    record once, walks for the keys (when some cycle is longer than 1) and for mu, and finishes
//    its lanes;
//  * few matched lanes: the walks are latency-bound chains that form the tail after every
//    trajectory kernel -- bsi_resolve runs the mu walks and the key walks side by side (first
//    and second half of its task range; tasks are handed out 32 at a time, so a warp is never
//    split between the two), then bsi_resolve_finish aggregates.
#if BSI_UNIT != 3
// ---------------------------------------------------------------------------------------------
// early words: resolution per EVENT and per GROUP, not per lane
// ---------------------------------------------------------------------------------------------
// An entry of the early-word queue is a dense word (all lanes start from the problem-index word, c0 =
// 0), its state y at time j_end -- an on-cycle state of every handed-over lane -- and its match
// events (lanes, lambda).  What the reference computes per trajectory (attract.py:287-300) comes out
// of three bit-sliced passes with no per-lane scalar work:
//   keys    the lexicographic minimum over the cycle, walked from y for max lambda - 1 steps; the same
//           walk latches P(0) = y advanced by (lambda - j_end mod lambda) mod lambda steps, event by
//           event (long cycles: both from the cycle dictionary instead);
//   groups  lanes with equal keys, found by broadcasting one lane's key bits against all lanes
//           (2 ops per node and group; the ordered regime has a handful of groups per word);
//   mu      the lock-step of x_j (from the index word) and P(j); the lanes that agree first at
//           step j add (1, j, j^2) to their groups.
// Only then the table is touched, once per group.
#if BSI_TUNE & 8
#define BSI_EW_DICT_MIN_LAM BSI_DICT_MIN_LAM
#else
#define BSI_EW_DICT_MIN_LAM BSI_EARLY_DICT_MIN_LAM
#endif
struct BsiGroups {
    u32 n_groups;
    u32 mask[32];               // lanes of the group
    u32 lam[32];                // their cycle length
    u64 key[32][BSI_KW];        // their canonical key
    u64 jlim[32];               // a lane of the group that reaches its cycle at relative time j closes by
                                // max_t iff j < jlim (mu + lambda <= T, reference model.py:201)
    u32 n[32], late[32];        // lanes that close in time / only after max_t
    u32 s1[32];                 // sum of j over the former (j < 2^24, at most 32 lanes)
    u64 s2[32];                 // sum of j^2
    unsigned char gid[32];      // lane -> its group
};
__device__ __forceinline__ void bsi_group_open(BsiGroups &gr, u32 g, u32 mask, u32 lam, u64 T) {
    gr.mask[g] = mask;
    gr.lam[g] = lam;
    gr.jlim[g] = (u64)lam <= T ? T - lam + 1ull : 0ull;
    gr.n[g] = 0u; gr.late[g] = 0u; gr.s1[g] = 0u; gr.s2[g] = 0ull;
    while (mask) {
        const u32 l = __ffs(mask) - 1;
        mask &= mask - 1;
        gr.gid[l] = (unsigned char)g;
    }
}

// lanes m reached their cycle at relative time j.  Two regimes: whole words that close together (the
// ordered regime: a few events of many lanes each) are counted per GROUP, one popc each; lanes that
// close one by one (up to 32 events per word, several groups) per LANE through its group number.
__device__ __noinline__ void bsi_early_account(BsiGroups &gr, u32 m, u32 j) {
    if ((u32)__popc(m) < gr.n_groups) {
        while (m) {
            const u32 l = __ffs(m) - 1;
            m &= m - 1;
            const u32 g = gr.gid[l];
            if ((u64)j < gr.jlim[g]) {
                gr.n[g] += 1u;
                gr.s1[g] += j;
                gr.s2[g] += (u64)j * j;
            } else {
                gr.late[g] += 1u;
            }
        }
        return;
    }
#pragma unroll 1
    for (u32 g = 0; g < gr.n_groups; ++g) {
        const u32 cnt = __popc(m & gr.mask[g]);
        if (!cnt) continue;
        if ((u64)j < gr.jlim[g]) {
            const u32 cj = cnt * j;
            gr.n[g] += cnt;
            gr.s1[g] += cj;
            gr.s2[g] += (u64)cj * j;
        } else {
            gr.late[g] += cnt;
        }
    }
}

// Per-thread cache of the attractors its last words fell into: (key, lambda) -> (n, sum, sum of squares)
// of trajectory lengths, added up locally and only written to the table when an entry is evicted
// and when the kernel ends.  In the ordered regime millions of lanes fall into a handful of
// attractors: consecutive words of a thread hit the same few keys, so a group costs a key compare
// and three additions instead of a table probe and three atomics on one of a few hot L2 addresses
// (or the warp-wide aggregation that kept those atomics apart before).
// (16 entries: with 4, the 39 attractors of cambium2 -- five groups per word -- and the 7 of the
//  seed-1 network evicted on most words, and the evictions' atomics queued up on the same few slots:
//  47 % of the kernel's stall samples sat in bsi_table_add.  The victim is the entry with the fewest
//  lanes: an attractor that takes most of the lanes is never written before the end.)
#if BSI_TUNE & 2
#define BSI_TC_WAYS 8
#elif BSI_TUNE & 4
#define BSI_TC_WAYS 32
#else
#define BSI_TC_WAYS 16
#endif
struct BsiTCache {
    u64 key[BSI_TC_WAYS][BSI_KW];
    u64 n[BSI_TC_WAYS], s1[BSI_TC_WAYS], s2[BSI_TC_WAYS];
    u32 lam[BSI_TC_WAYS];
    u32 used;
};
__device__ __forceinline__ void bsi_tcache_init(BsiTCache &tc) { tc.used = 0u; }
__device__ __noinline__ void bsi_tcache_add(BsiTCache &tc, const u64 *key, u64 n, u64 s1, u64 s2, u32 lam,
                                            BsiSlot *table, u32 log2_slots, BsiCounters *ctr) {
#pragma unroll 1
    for (u32 i = 0; i < tc.used; ++i) {
        bool same = true;
#pragma unroll
        for (int k = 0; k < BSI_KW; ++k) same = same && tc.key[i][k] == key[k];
        if (same) {
            const u64 nn = tc.n[i] + n;
            tc.n[i] = nn; tc.s1[i] += s1; tc.s2[i] += s2;
            // (kept in the order of their counts: the attractors that take most of the lanes are found
            //  by the first compare or two; entries swap only when their counts cross)
            if (i && nn > tc.n[i - 1u]) {
#pragma unroll
                for (int k = 0; k < BSI_KW; ++k) { const u64 t = tc.key[i][k]; tc.key[i][k] = tc.key[i - 1u][k]; tc.key[i - 1u][k] = t; }
                u64 t = tc.n[i]; tc.n[i] = tc.n[i - 1u]; tc.n[i - 1u] = t;
                t = tc.s1[i]; tc.s1[i] = tc.s1[i - 1u]; tc.s1[i - 1u] = t;
                t = tc.s2[i]; tc.s2[i] = tc.s2[i - 1u]; tc.s2[i - 1u] = t;
                const u32 tl = tc.lam[i]; tc.lam[i] = tc.lam[i - 1u]; tc.lam[i - 1u] = tl;
            }
            return;
        }
    }
    u32 i = tc.used;
    if (i < BSI_TC_WAYS) {
        tc.used = i + 1u;
    } else {
        i = 0u;
#pragma unroll 1
        for (u32 q = 1; q < BSI_TC_WAYS; ++q)
            if (tc.n[q] < tc.n[i]) i = q;
        if (n <= tc.n[i]) {                              // (the newcomer is the smallest: it goes itself)
            bsi_table_add(table, log2_slots, key, n, s1, s2, lam, ctr);
            return;
        }
        bsi_table_add(table, log2_slots, tc.key[i], tc.n[i], tc.s1[i], tc.s2[i], tc.lam[i], ctr);
    }
#pragma unroll
    for (int k = 0; k < BSI_KW; ++k) tc.key[i][k] = key[k];
    tc.n[i] = n; tc.s1[i] = s1; tc.s2[i] = s2; tc.lam[i] = lam;
}
// the end of the kernel: with WARP_AGG a warp collective (every thread of the warp calls), entry by
// entry, so that the threads of a warp that hold the same key write it once
template <bool WARP_AGG>
__device__ __forceinline__ void bsi_tcache_flush(BsiTCache &tc, BsiSlot *table, u32 log2_slots, BsiCounters *ctr) {
#pragma unroll 1
    for (u32 i = 0; i < BSI_TC_WAYS; ++i) {
        const bool mine = i < tc.used;
        if (WARP_AGG) {
            if (!__any_sync(0xFFFFFFFFu, mine)) break;
            u64 key0[BSI_KW];
#pragma unroll
            for (int k = 0; k < BSI_KW; ++k) key0[k] = mine ? tc.key[i][k] : 0ull;
            bsi_warp_table_add(table, log2_slots, key0, mine ? tc.n[i] : 0ull, mine ? tc.s1[i] : 0ull,
                               mine ? tc.s2[i] : 0ull, mine ? tc.lam[i] : 0u, ctr);
        } else if (mine) {
            bsi_table_add(table, log2_slots, tc.key[i], tc.n[i], tc.s1[i], tc.s2[i], tc.lam[i], ctr);
        }
    }
    tc.used = 0u;
}

// The cycle dictionary's part of an early word (OUT OF LINE: words of short cycles never come here):
// state codes of y by one transpose pass, lookups of the lanes on cycles, then P(0) bit-sliced into
// pb and the groups -- lanes on cycles by where their orbit is stored (one number per attractor),
// fixed points by their state code.  false: some cycle is not in the dictionary (nothing is filled).
__device__ __noinline__ bool bsi_early_dict(const BsiDict &dict, const u32 *y, u32 valid, u32 n_ev, const u32 *evm,
                                            const u32 *evl, const u32 *eva, u32 *pb, BsiGroups &gr, u64 T) {
    u64 cw_[32][BSI_KW];
    u32 lam_[32], adv_[32], dbase_[32];
    {
        BSI_FOR_NODES(BSI_DECL_B)
        BSI_LOAD_STREAM(b, y, 32)
        BSI_UNPACK(cw_, b)
    }
#pragma unroll 1
    for (u32 l = 0; l < 32; ++l) { lam_[l] = 0u; adv_[l] = 0u; }
#pragma unroll 1
    for (u32 k = 0; k < n_ev; ++k) {
        u32 m = evm[k];
        while (m) {
            const u32 l = __ffs(m) - 1;
            m &= m - 1;
            lam_[l] = evl[k];
            adv_[l] = eva[k];
        }
    }
    u64 keys_[32][BSI_KW];
    if (bsi_dict_word(dict, cw_, lam_, adv_, 1u, true, keys_, pb, dbase_)) return false;
    u32 todo = valid, G = 0u;
    while (todo) {
        const u32 l0 = __ffs(todo) - 1;
        const bool cyc = lam_[l0] > 1u;
        u32 g = 0u;
        if (cyc) {
            // (independent loads, all in flight together)
            const u32 b0 = dbase_[l0];
#pragma unroll
            for (u32 l = 0; l < 32; ++l) g |= ((lam_[l] > 1u && dbase_[l] == b0) ? 1u : 0u) << l;
        } else {
#pragma unroll 4
            for (u32 l = 0; l < 32; ++l) {
                bool same = lam_[l] == 1u;
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) same = same && cw_[l][i] == cw_[l0][i];
                g |= (same ? 1u : 0u) << l;
            }
        }
        g &= todo;
        todo &= ~g;
        bsi_group_open(gr, G, g, lam_[l0], T);
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) gr.key[G][i] = cyc ? keys_[l0][i] : cw_[l0][i];
        ++G;
    }
    gr.n_groups = G;
    return true;
}

// One entry of the early-word queue.  The groups' totals go into the thread's attractor cache `tc`
// (the kernel flushes it into the table when it ends).
__device__ __forceinline__ void bsi_early_word(const BsiEarlyQ &eq, u64 e, u64 T, u64 max_attractor_l,
                                               u64 stop_t_cap, BsiSlot *table, u32 log2_slots, BsiCounters *ctr,
                                               const BsiDict &dict, bool learn, BsiTCache &tc, u64 &steps,
                                               u64 &n_acc, u64 &n_rej) {
    BsiGroups gr;
    gr.n_groups = 0u;
    {
        BSI_FOR_NODES(BSI_DECL_A)
        BSI_FOR_NODES(BSI_DECL_B)
        BSI_FOR_NODES(BSI_DECL_C)
        const BsiEarlyEntry ent = bsi_early_entry(eq, e);
        const u32 valid = bsi_ld_stream(ent.mask);
        const u32 info = bsi_ld_stream(ent.info);
        const u32 n_ev = info & 0xFFu, j_end = info >> 8;
        const u64 W_ = bsi_ld_stream(ent.word);
        u32 evm[32], evl[32], eva[32];
        // (all event loads first, independent and in flight together)
#pragma unroll 8
        for (u32 k = 0; k < 32; ++k)
            if (k < n_ev) { evm[k] = bsi_ld_stream(ent.evm + k * 32u); evl[k] = bsi_ld_stream(ent.evl + k * 32u); }
        u32 lam_max = 0u, next_ = 0xFFFFFFFFu;
#pragma unroll 1
        for (u32 k = 0; k < n_ev; ++k) {
            const u32 lam = evl[k];
            // (fixed points, the bulk of the ordered regime, need no division)
            const u32 adv = lam == 1u ? 0u : (lam - j_end % lam) % lam;
            eva[k] = adv;
            lam_max = max(lam_max, lam);
            if (adv) next_ = min(next_, adv);
        }
        u32 pb_[BSI_N];                                  // P(0), bit-sliced (cycles only)
        bool via_dict = false;
        if (lam_max > BSI_EW_DICT_MIN_LAM && dict.slots) {
            via_dict = bsi_early_dict(dict, ent.y, valid, n_ev, evm, evl, eva, pb_, gr, T);
            if (via_dict) { BSI_COUNT(2) } else { BSI_COUNT(3) }
        }
        if (!via_dict) {
            BSI_LOAD_STREAM(b, ent.y, 32)
            BSI_FOR_NODES(BSI_COPY_CB_)
            if (lam_max > 1u) {
                // ---- keys (running minimum in c) and P(0) (latched into pb_) in one walk from y -----
#define BSI_LEX_NODE(i)                                                           \
    lt_ |= bsi_lop3<0x08>(X_(i), c##i, eq_);                                      \
    eq_ = bsi_lop3<0x90>(eq_, X_(i), c##i);
#define BSI_SEL_NODE(i) c##i = bsi_lop3<0xCA>(lt_, X_(i), c##i);
#define BSI_LATCH_NODE(i) pb_[i] = bsi_lop3<0xCA>(m_, X_(i), pb_[i]);
                // state X is y advanced by K steps: minimum, then the events whose P(0) it is
#define BSI_EW_STEP(K)                                                            \
    {                                                                             \
        u32 lt_ = 0u, eq_ = valid;                                                \
        BSI_LEX_CHUNKS                                                            \
        if (lt_) { BSI_FOR_NODES(BSI_SEL_NODE) }                                  \
        if ((K) == next_) {                                                       \
            u32 m_ = 0u, nx_ = 0xFFFFFFFFu;                                       \
            _Pragma("unroll 1")                                                   \
            for (u32 q_ = 0; q_ < n_ev; ++q_) {                                   \
                const u32 adv_ = eva[q_];                                         \
                if (adv_ == (K)) m_ |= evm[q_];                                   \
                else if (adv_ > (K)) nx_ = min(nx_, adv_);                        \
            }                                                                     \
            next_ = nx_;                                                          \
            BSI_FOR_NODES(BSI_LATCH_NODE)                                         \
        }                                                                         \
    }
                BSI_STORE_ALL(b, pb_, 1)                 // (advance 0: P(0) = y)
                for (u32 k = 1; k < lam_max; k += 2) {
                    BSI_UPDATE(a, b)
#define X_(i) a##i
                    BSI_EW_STEP(k)
#undef X_
                    if (k + 1u >= lam_max) break;
                    BSI_UPDATE(b, a)
#define X_(i) b##i
                    BSI_EW_STEP(k + 1u)
#undef X_
                }
#undef BSI_EW_STEP
#undef BSI_LATCH_NODE
#undef BSI_SEL_NODE
#undef BSI_LEX_NODE
            }
            // ---- groups: lanes whose key equals that of the first lane left ---------------------------
            u32 todo = valid, G = 0u;
            while (todo) {
                const u32 l0 = __ffs(todo) - 1;
                const u32 sh_ = 31u - l0;
                u32 diff_ = 0u;
                u32 kc_[2 * BSI_KW];
#pragma unroll
                for (int i = 0; i < 2 * BSI_KW; ++i) kc_[i] = 0u;
                // r_ = the key bit of lane l0 on all lanes (shift it to the sign, then arithmetic shift)
#define BSI_GRP_NODE(i)                                                           \
    {                                                                             \
        const u32 r_ = (u32)((int)(c##i << sh_) >> 31);                           \
        diff_ = bsi_lop3<0xF6>(diff_, c##i, r_);                                  \
        kc_[(i) >> 5] |= r_ & (1u << ((i) & 31));                                 \
    }
                BSI_FOR_NODES(BSI_GRP_NODE)
#undef BSI_GRP_NODE
                const u32 g = ~diff_ & todo;
                todo &= ~g;
                u32 lam0 = 1u;
#pragma unroll 1
                for (u32 k = 0; k < n_ev; ++k)
                    if ((evm[k] >> l0) & 1u) { lam0 = evl[k]; break; }
                bsi_group_open(gr, G, g, lam0, T);
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) gr.key[G][i] = (u64)kc_[2 * i] | ((u64)kc_[2 * i + 1] << 32);
                ++G;
            }
            gr.n_groups = G;
            // the walked cycles go into the dictionary (once per attractor: the claim decides)
            if (learn && dict.slots && lam_max > BSI_EW_DICT_MIN_LAM) {
#pragma unroll 1
                for (u32 g = 0; g < G; ++g) {
                    u32 base, slot;
                    // (short cycles too: a word that holds a long one asks the dictionary for all of them)
                    if (gr.lam[g] > 1u && bsi_dict_claim(dict, gr.key[g], gr.lam[g], base, slot))
                        bsi_dict_learn(dict, gr.key[g], gr.lam[g], base, slot);
                }
            }
        }
        // ---- a <- x_0 from the problem-index word, b <- P(0) -------------------------------------
        BSI_INIT_DENSE(a, W_)
        BSI_PREFIX(BSI_LP_ORIGIN)
        if (lam_max > 1u) { BSI_LOAD_MEM(b, pb_, 1) } else { BSI_LOAD_STREAM(b, ent.y, 32) }
        // ---- lock-step: mu = first j with x_j == P(j) -------------------------------------------
        u32 alive = valid;
#define BSI_LOCK_CMP(A, B, J)                                                     \
    {                                                                             \
        u32 p_ = ~alive, q_ = 0u;                                                 \
        BSI_CMP_FILTER(p_, q_, A, B)                                              \
        p_ |= q_;                                                                 \
        if (p_ != 0xFFFFFFFFu) {                                                  \
            BSI_CMP_REST(p_, A, B)                                                \
            const u32 m_ = ~p_;                                                   \
            if (m_) {                                                             \
                alive &= ~m_;                                                     \
                bsi_early_account(gr, m_, (J));                                   \
            }                                                                     \
        }                                                                         \
    }
#define BSI_LOCK(A, B, Tn, J)                                                     \
    {                                                                             \
        BSI_LOCK_CMP(A, B, J)                                                     \
        if (alive) { BSI_UPDATE(Tn, A) BSI_UPDATE(A, B) }                         \
    }
        if (lam_max <= 1u) {
            // every lane ends in a fixed point: P is constant (= b), only x moves (a <-> c)
            u32 j = 0;
            while (alive && j <= j_end) {
                BSI_LOCK_CMP(a, b, j)
                if (!alive || j + 1 > j_end) break;
                BSI_UPDATE(c, a)
                BSI_LOCK_CMP(c, b, j + 1)
                if (!alive || j + 2 > j_end) break;
                BSI_UPDATE(a, c)
                j += 2;
            }
        } else {
            u32 j = 0;
            while (alive && j <= j_end) {
                BSI_LOCK(a, b, c, j)
                if (!alive || j + 1 > j_end) break;
                BSI_LOCK(c, a, b, j + 1)
                if (!alive || j + 2 > j_end) break;
                BSI_LOCK(b, c, a, j + 2)
                j += 3;
            }
        }
#undef BSI_LOCK
#undef BSI_LOCK_CMP
        if (alive) atomicExch(&ctr->internal_error, 1ull);
    }
    // ---- the groups' totals: acceptance (reference attract.py:287-300) and aggregation ----------
#pragma unroll 1
    for (u32 g = 0; g < gr.n_groups; ++g) {
        const u64 n = gr.n[g], late = gr.late[g], lam = gr.lam[g], s1 = gr.s1[g];
        steps += n * (BSI_LP_ORIGIN + lam) + s1          // the loop stops at t = mu + lambda (model.py:201)
                 + late * stop_t_cap;                    // or runs out at max_t
        n_rej += late;
        if (lam > max_attractor_l) {                     // attract.py:294
            n_rej += n;
        } else if (n) {
            n_acc += n;
            // trajectory_l = lp + j (attract.py:298): sums of (lp + j) and (lp + j)^2 from those of j
            const u64 lp = BSI_LP_ORIGIN;
            bsi_tcache_add(tc, gr.key[g], n, s1 + n * lp, gr.s2[g] + 2ull * lp * s1 + n * lp * lp, (u32)lam,
                           table, log2_slots, ctr);
        }
    }
}
// (the latency-oriented kernel holds three kinds of task: this one stays out of its register budget)
__device__ __noinline__ void bsi_early_word_call(const BsiEarlyQ &eq, u64 e, u64 T, u64 max_attractor_l,
                                                 u64 stop_t_cap, BsiSlot *table, u32 log2_slots, BsiCounters *ctr,
                                                 const BsiDict &dict, BsiTCache &tc, u64 &steps, u64 &n_acc,
                                                 u64 &n_rej) {
    bsi_early_word(eq, e, T, max_attractor_l, stop_t_cap, table, log2_slots, ctr, dict, false, tc, steps, n_acc, n_rej);
}
#endif  // BSI_UNIT != 3

#if BSI_UNIT != 5 && BSI_UNIT != 6
extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_resolve_mu(BsiFound *found_q, const u64 *n_found_p, u64 few_words, u64 T, u64 max_attractor_l,
               u64 stop_t_cap, BsiSlot *table, u32 log2_slots, BsiCounters *ctr, const BsiDict dict) {
    __shared__ u32 aux_s[32 * BSI_BLOCK];
    const u64 n_found = *n_found_p;
    if (bsi_latency_mode(n_found, ctr, few_words)) return;
    const u64 n_words = BSI_QWORDS(n_found);
    u64 steps = 0, n_acc = 0, n_rej = 0;
    BsiKeep keep;
    for (;;) {
        const u64 w = bsi_grab_word(&ctr->next_word3);
        if (__all_sync(0xFFFFFFFFu, w >= n_words)) break;
        u32 lanes = 0u;                  // threads without a word stay for the warp-level aggregation
        if (w < n_words && BSI_Q(w, 0) < n_found)
            bsi_mu_word<false, true>(found_q, n_found, w, aux_s, ctr, &keep, dict, true, lanes);
        __syncwarp();
        bsi_finish_word<true, true>(found_q, n_found, w, T, max_attractor_l, stop_t_cap, table, log2_slots,
                                    ctr, steps, n_acc, n_rej, aux_s, &keep, lanes);
    }
    if (steps) atomicAdd(&ctr->steps, steps);
    if (n_acc) atomicAdd(&ctr->n_accepted, n_acc);
    if (n_rej) atomicAdd(&ctr->n_without, n_rej);
}

#endif  // BSI_UNIT != 5
#if BSI_UNIT == 0 || BSI_UNIT == 6
// The early-word queue of a dense round: keys, mu and aggregation of the lanes that matched an early
// checkpoint, straight from the word's bit-sliced state (see bsi_early_word).
extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_resolve_early(void *early_base, u64 early_cap, const u64 *n_early_p, const u64 *n_found_p,
                  u64 few_words, u64 T, u64 max_attractor_l, u64 stop_t_cap, BsiSlot *table,
                  u32 log2_slots, BsiCounters *ctr, const BsiDict dict) {
    const u64 n_early = *n_early_p;
    if (n_early == 0) return;
    if (bsi_latency_mode(*n_found_p, ctr, few_words)) return;    // bsi_resolve takes the early words too
    const BsiEarlyQ eq = bsi_early_view(early_base, early_cap);
    u64 steps = 0, n_acc = 0, n_rej = 0;
    BsiTCache tc;
    bsi_tcache_init(tc);
    for (;;) {
        const u64 e = bsi_grab_word(&ctr->next_word4);
        if (__all_sync(0xFFFFFFFFu, e >= n_early)) break;
        if (e < n_early)
            bsi_early_word(eq, e, T, max_attractor_l, stop_t_cap, table, log2_slots, ctr, dict, true, tc, steps,
                           n_acc, n_rej);
    }
    __syncwarp();
    bsi_tcache_flush<true>(tc, table, log2_slots, ctr);
    if (steps) atomicAdd(&ctr->steps, steps);
    if (n_acc) atomicAdd(&ctr->n_accepted, n_acc);
    if (n_rej) atomicAdd(&ctr->n_without, n_rej);
}
#endif  // BSI_UNIT 6
#if BSI_UNIT != 3 && BSI_UNIT != 6
// Orbits on request (reference attract.py:22-25: the states of an aggregated attractor, key state
// first): thread i finds keys[i] in the cycle dictionary or stores its orbit there -- the same fast
// walk the resolution kernels use (32 phases of the orbit per step) -- and reports where it starts
// (bases[i]; ~0: not available, e.g. storage exhausted).  The host then copies the orbits out of the
// dictionary's storage, and bsi_orbit_counts can count node activity there without a round trip.
extern "C" __global__ void bsi_dict_learn_keys(const u64 *keys, const u64 *lams, u64 n, u64 *bases,
                                               const BsiDict dict) {
    // one 32-thread block per key: thread 0 claims and walks (the walk is one thread's work: its 32
    // lanes are the parallelism), then the whole warp enters the states -- each entry is a couple of
    // dependent L2 round trips, 32 of them in flight instead of one
    const u64 i = blockIdx.x;
    if (i >= n) return;
    __shared__ u32 sh[3];
    const u32 lam = lams[i] < 2ull || lams[i] > (u64)dict.cap_states ? 0u : (u32)lams[i];
    if (threadIdx.x == 0u) {
        bases[i] = ~0ull;
        sh[0] = 0u;
        if (lam) {
            u64 key[BSI_KW];
#pragma unroll
            for (int k = 0; k < BSI_KW; ++k) key[k] = keys[i * BSI_KW + k];
            u32 idx, base, lamd, slot;
            if (bsi_dict_find(dict, key, idx, base, lamd)) {
                if (idx == base && lamd == lam) bases[i] = base;
            } else if (bsi_dict_claim(dict, key, lam, base, slot)) {
                bsi_dict_store_orbit(dict, key, lam, base);
                bases[i] = base;
                sh[0] = 1u; sh[1] = base; sh[2] = slot;
            }
        }
    }
    __syncwarp();
    if (sh[0]) {
        bsi_dict_put_orbit(dict, lam, sh[1], threadIdx.x & 31u, 32u);
        __syncwarp();
        if (threadIdx.x == 0u) bsi_dict_publish(dict, sh[2]);
    }
}

extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_resolve(BsiFound *found_q, const u64 *n_found_p, u64 few_words, void *early_base, u64 early_cap,
            u64 T, u64 max_attractor_l, u64 stop_t_cap, BsiSlot *table, u32 log2_slots,
            BsiCounters *ctr, const BsiDict dict) {
    __shared__ u32 aux_s[32 * BSI_BLOCK];
    const u64 n_found = *n_found_p;
    if (!bsi_latency_mode(n_found, ctr, few_words)) return;
    const u64 w32 = BSI_QWORDS(n_found);             // a multiple of 32
    // third task range: the early-word queue (its kernel leaves it to this one when the record
    // queue is short, so that all three kinds of walk run side by side)
    const u64 n_early = early_cap ? ctr->n_early : 0ull;
    const u64 e32 = (n_early + 31) & ~31ull;
    const BsiEarlyQ eq = bsi_early_view(early_base, early_cap);
    u64 steps = 0, n_acc = 0, n_rej = 0;
    BsiTCache tc;
    bsi_tcache_init(tc);
    for (;;) {
        const u64 t = bsi_grab_word(&ctr->next_word2);
        if (__all_sync(0xFFFFFFFFu, t >= 2 * w32 + e32)) break;
        if (t < w32) {
            if (BSI_Q(t, 0) < n_found) {
                u32 lanes = 0u;
                BsiKeep keep;
                bsi_mu_word<true, false>(found_q, n_found, t, aux_s, ctr, &keep, dict, false, lanes);
            }
        } else if (t < 2 * w32) {
            if (BSI_Q(t - w32, 0) < n_found) bsi_key_word(found_q, n_found, t - w32, dict);
        } else if (t - 2 * w32 < n_early) {
            // (no learning on the latency-oriented path: see bsi_key_word)
            bsi_early_word_call(eq, t - 2 * w32, T, max_attractor_l, stop_t_cap, table, log2_slots, ctr, dict, tc,
                                steps, n_acc, n_rej);
        }
    }
    bsi_tcache_flush<false>(tc, table, log2_slots, ctr);
    if (steps) atomicAdd(&ctr->steps, steps);
    if (n_acc) atomicAdd(&ctr->n_accepted, n_acc);
    if (n_rej) atomicAdd(&ctr->n_without, n_rej);
}

extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_resolve_finish(const BsiFound *found_q, const u64 *n_found_p, u64 few_words, u64 T,
                   u64 max_attractor_l, u64 stop_t_cap, BsiSlot *table, u32 log2_slots,
                   BsiCounters *ctr) {
    const u64 n_found = *n_found_p;
    if (!bsi_latency_mode(n_found, ctr, few_words)) return;
    const u64 n_words = BSI_QWORDS(n_found);
    u64 steps = 0, n_acc = 0, n_rej = 0;
    BsiKeep keep;
    for (;;) {
        const u64 w = bsi_grab_word(&ctr->next_word3);
        if (__all_sync(0xFFFFFFFFu, w >= n_words)) break;
        if (w >= n_words || BSI_Q(w, 0) >= n_found) continue;
        bsi_finish_word<false, false>(found_q, n_found, w, T, max_attractor_l, stop_t_cap, table, log2_slots,
                               ctr, steps, n_acc, n_rej, nullptr, &keep, 0u);
    }
    if (steps) atomicAdd(&ctr->steps, steps);
    if (n_acc) atomicAdd(&ctr->n_accepted, n_acc);
    if (n_rej) atomicAdd(&ctr->n_without, n_rej);
}

#endif  // BSI_UNIT != 3, 6
#endif  // BSI_UNIT 3, 5, 6

#undef BSI_TRANSPOSE32
#define BSI_TRANSPOSE32(A) bsi_transpose32(A)
#if BSI_UNIT == 0 || BSI_UNIT == 4
// ---- attractor table: export / merge ------------------------------------------------------------
// Compacts the table into out[0 .. cap).  With `header` set, the last block to finish also writes
// one header record there: key = {count, trajectory-steps, accepted, without}, n = 0 (so that a
// merge skips it) -- the per-rank payload of the all_gather needs no host round trip.
extern "C" __global__ void bsi_table_export(const BsiSlot *table, u32 log2_slots, BsiRec *out,
                                            u64 cap, u64 *count, BsiRec *header,
                                            const BsiCounters *ctr, u32 *blocks_done) {
    const u64 n_slots = (u64)1 << log2_slots;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots;
         i += (u64)gridDim.x * blockDim.x) {
        if (table[i].flag == 2u) {
            u64 pos = atomicAdd(count, 1ull);
            if (pos < cap) {
                BsiRec r;
#pragma unroll
                for (int k = 0; k < 4; ++k) r.key[k] = 0ull;
#pragma unroll
                for (int k = 0; k < BSI_KW; ++k) r.key[k] = table[i].key[k];
                r.n = table[i].n; r.s1 = table[i].s1; r.s2 = table[i].s2; r.lam = table[i].lam;
                out[pos] = r;
            }
        }
    }
    if (header) {
        __shared__ u32 last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            last = atomicAdd(blocks_done, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (last && threadIdx.x == 0) {
            BsiRec h;
            h.key[0] = *((volatile u64 *)count);
            h.key[1] = ctr->steps; h.key[2] = ctr->n_accepted; h.key[3] = ctr->n_without;
            h.n = 0ull; h.s1 = ctr->table_full; h.s2 = 0ull; h.lam = 0ull;
            *header = h;
        }
    }
}

extern "C" __global__ void bsi_table_merge(BsiSlot *table, u32 log2_slots, const BsiRec *recs, u64 n,
                                           BsiCounters *ctr) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (u64)gridDim.x * blockDim.x) {
        if (recs[i].n == 0ull) continue;
        u64 key[BSI_KW];
#pragma unroll
        for (int k = 0; k < BSI_KW; ++k) key[k] = recs[i].key[k];
        bsi_table_add(table, log2_slots, key, recs[i].n, recs[i].s1, recs[i].s2, (u32)recs[i].lam, ctr);
    }
}

// Folds the exported tables of other devices into this one (the collection stage of reference
// mpi.py:296-332 as one launch): payload j = header hdrs[j] (record count in key[0]) + records
// recs[j * cap ...], written into this device's memory by the peers' export kernels.  A payload
// that did not fit (count > cap) is skipped as a whole; the host sees the count and asks again.
extern "C" __global__ void bsi_table_merge_payloads(BsiSlot *table, u32 log2_slots, const BsiRec *hdrs,
                                                    const BsiRec *recs, u32 n_payloads, u64 cap,
                                                    BsiCounters *ctr) {
    for (u32 j = 0; j < n_payloads; ++j) {
        const u64 count = hdrs[j].key[0];
        if (count > cap) continue;
        const BsiRec *r = recs + (u64)j * cap;
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
            if (r[i].n == 0ull) continue;
            u64 key[BSI_KW];
#pragma unroll
            for (int k = 0; k < BSI_KW; ++k) key[k] = r[i].key[k];
            bsi_table_add(table, log2_slots, key, r[i].n, r[i].s1, r[i].s2, (u32)r[i].lam, ctr);
        }
    }
}

#endif  // BSI_UNIT 4
#endif  // BSI_MODE == 1

#if BSI_UNIT == 0 || BSI_UNIT == 4
// ---------------------------------------------------------------------------------------------
// measurement: the bare update loop (no compare, no stores) of THIS network -- the practical
// ceiling of the trajectory kernels, to set beside the ideal-chain lop3 peak
// ---------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_update_rate(u32 pairs, u32 seed, u32 *out) {
    BSI_FOR_NODES(BSI_DECL_A)
    BSI_FOR_NODES(BSI_DECL_B)
    BSI_FOR_FIXVAR(BSI_DECL_FIXVAR)
    u32 x_ = seed ^ (blockIdx.x * BSI_BLOCK + threadIdx.x) * 0x9E3779B9u;
#define BSI_SEED_A(i) x_ = x_ * 1664525u + 1013904223u; a##i = x_;
    BSI_FOR_NODES(BSI_SEED_A)
#undef BSI_SEED_A
    for (u32 k = 0; k < pairs; ++k) { BSI_UPDATE(b, a) BSI_UPDATE(a, b) }
    u32 acc = 0u;
#define BSI_FOLD_A(i) acc ^= a##i;
    BSI_FOR_NODES(BSI_FOLD_A)
#undef BSI_FOLD_A
    if (acc == 0x13572468u) out[0] = acc;
}

// ---------------------------------------------------------------------------------------------
// simulate / target / orbits: explicit trajectories
// ---------------------------------------------------------------------------------------------

// Lane l of thread w simulates problem  indices ? indices[32w + l] : idx_begin + 32w + l  for
// n_steps steps (reference simulate.py:97-131 == plain iteration with masks) and stores every
// state as KW code words: out[(q * (n_steps + 1) + t) * KW].
// With codes != 0 the lanes start from explicit state codes instead (orbit regeneration,
// reference attract.py:22-25): no prefix, no perturbations, origin fixed nodes only.
extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_trajectories(u64 idx_begin, const u64 *indices, const u64 *codes, u64 n_lanes, u64 n_steps,
                 u64 *out) {
    const u64 w = (u64)blockIdx.x * BSI_BLOCK + threadIdx.x;
    if ((w << 5) >= n_lanes) return;
    BSI_FOR_NODES(BSI_DECL_A)
    BSI_FOR_NODES(BSI_DECL_B)
    BSI_FOR_FIXVAR(BSI_DECL_FIXVAR)
    BSI_FOR_PVAR(BSI_DECL_PVAR)
    u32 valid = 0u;
    u64 cw_[32][BSI_KW];
#pragma unroll 1
    for (u32 l = 0; l < 32; ++l) {
        u64 q = (w << 5) + l;
        u64 c_[BSI_KW];
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) c_[i] = 0ull;
        if (q < n_lanes) {
            u64 lp_lane_ = 0;
            // (the setup also fills the per-lane fixed-node masks; with explicit codes the index
            //  is irrelevant for them because orbits are only asked for in attract mode)
            BSI_LANE_SETUP(l, (indices ? indices[q] : idx_begin + q), c_)
            (void)lp_lane_;
            if (codes) {
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) c_[i] = codes[q * BSI_KW + i];
            }
            valid |= 1u << l;
        }
#pragma unroll
        for (int i = 0; i < BSI_KW; ++i) cw_[l][i] = c_[i];
    }
    BSI_PACK(a, cw_)
    const bool perturbing = codes == nullptr;
    // every state of all 32 lanes goes out as code words: one transpose per 32 nodes, then
    // KW 64-bit stores per lane
#define BSI_STORE(X, T)                                                           \
    {                                                                             \
        BSI_UNPACK(cw_, X)                                                        \
        _Pragma("unroll 1") for (u32 l = 0; l < 32; ++l) {                        \
            if ((valid >> l) & 1u) {                                              \
                u64 *dst = out + (((w << 5) + l) * (n_steps + 1) + (T)) * BSI_KW; \
                _Pragma("unroll") for (int i = 0; i < BSI_KW; ++i) dst[i] = cw_[l][i]; \
            }                                                                     \
        }                                                                         \
    }
    BSI_STORE(a, 0)
    // two steps per iteration (a -> b -> a): no register copies
    u64 t = 1;
    for (; t + 1 <= n_steps; t += 2) {
        BSI_UPDATE(b, a)
        if (perturbing) { BSI_PERTURB(b, t) }
        BSI_STORE(b, t)
        BSI_UPDATE(a, b)
        if (perturbing) { BSI_PERTURB(a, (t + 1)) }
        BSI_STORE(a, (t + 1))
    }
    if (t <= n_steps) {
        BSI_UPDATE(b, a)
        if (perturbing) { BSI_PERTURB(b, t) }
        BSI_STORE(b, t)
    }
#undef BSI_STORE
}

// Node activity of attractors (reference attractor_analysis.py:49-51: the mean of every node over the
// states of the cycle): block a counts, for every node, the states of orbit a in which it is on.
// states: state codes [..][KW]; orbit a = states[start[a] .. start[a] + lam[a]).
extern "C" __global__ void bsi_orbit_counts(const u64 *states, const u64 *start, const u64 *lam, u64 *counts) {
    const u64 a = blockIdx.x;
    const u64 s0 = start[a], n = lam[a];
    if (s0 == ~0ull) return;                      // (no orbit on the device for this one)
    for (u32 node = threadIdx.x; node < (u32)BSI_N; node += blockDim.x) {
        u64 c = 0;
        for (u64 t = 0; t < n; ++t) c += (states[(s0 + t) * BSI_KW + (node >> 6)] >> (node & 63u)) & 1ull;
        counts[a * BSI_N + node] = c;
    }
}

#if BSI_MODE == 2
// target (reference target.py:109-133 over model.py:152-236): first time t in [lp, T] at which
// the constrained nodes equal the target; lp = the lane's own last perturbation time.
//
// One round with horizon T (absolute time): lanes that hit report (idx, t_hit).  When `check_closure`
// is set (max_t beyond this horizon), lanes without a hit are tested for cycle closure --
// x_T among x_lp .. x_{T-1} -- because after closure no new state, hence no hit, can appear
// (model.py:201: the reference stops there); lanes neither hit nor closed go to the retry queue.
// Lanes that end without a hit add their reference stop time to ctr->steps only in the final
// round (stop time = max_t); closed lanes are listed in closed_q and only COUNTED by the host (their
// exact stop time, mu + lambda, is not worked out in this mode: see bsi_run_stats.steps).
extern "C" __global__ void __launch_bounds__(BSI_BLOCK)
bsi_target_main(u64 idx_begin, const u64 *in_queue, u64 n_lanes, u64 T, int check_closure,
                u64 *hit_idx, u64 *hit_t, u64 hit_cap, u64 *retry_q, u64 *closed_q,
                BsiCounters *ctr) {
    __shared__ u32 th_s[32 * BSI_BLOCK];
    __shared__ u32 lp_s[32 * BSI_BLOCK];
    const u64 w = (u64)blockIdx.x * BSI_BLOCK + threadIdx.x;
    if ((w << 5) >= n_lanes) return;
    BSI_FOR_NODES(BSI_DECL_A)
    BSI_FOR_NODES(BSI_DECL_B)
    BSI_FOR_NODES(BSI_DECL_C)
    BSI_FOR_FIXVAR(BSI_DECL_FIXVAR)
    BSI_FOR_PVAR(BSI_DECL_PVAR)
    u32 valid = 0u, hit = 0u, closed = 0u;
    u64 dev_steps = 0;

    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: hits and c = x_T;  pass 1 (only if needed): closure test against c
        valid = 0u;
        {
            u64 cw_[32][BSI_KW];
#pragma unroll 1
            for (u32 l = 0; l < 32; ++l) {
                u64 q = (w << 5) + l;
                u64 c_[BSI_KW];
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) c_[i] = 0ull;
                if (q < n_lanes) {
                    u64 lp_lane_ = BSI_LP_ORIGIN;
                    // (re-running the setup in pass 1 is harmless: it only ORs the same mask bits)
                    BSI_LANE_SETUP(l, (in_queue ? in_queue[q] : idx_begin + q), c_)
                    lp_s[l * BSI_BLOCK + threadIdx.x] = (u32)lp_lane_;
                    valid |= 1u << l;
                }
#pragma unroll
                for (int i = 0; i < BSI_KW; ++i) cw_[l][i] = c_[i];
            }
            BSI_PACK(a, cw_)
        }
        u32 started = 0u;
        u32 alive = valid & ~hit;
        // one time step of the test on state X at time TT; sets stop_ when the pass is over
#define BSI_TCHECK(X, TT)                                                         \
    {                                                                             \
        const u64 t_ = (TT);                                                      \
        /* lanes whose perturbations are all carried out by time t */               \
        if (BSI_LANES_UNIFORM) {                                                  \
            if (t_ == BSI_LP_ORIGIN) started = valid;                             \
        } else if (t_ <= BSI_LP_MAX) {                                            \
            _Pragma("unroll 1") for (u32 l = 0; l < 32; ++l)                      \
                if (((valid >> l) & 1u) && lp_s[l * BSI_BLOCK + threadIdx.x] == (u32)t_) started |= 1u << l; \
        }                                                                         \
        if (pass == 0) {                                                          \
            u32 acc = ~(alive & started);                                         \
            BSI_CMP_TARGET(acc, X)                                                \
            u32 m_ = ~acc;                                                        \
            if (m_) {                                                             \
                hit |= m_;                                                        \
                alive &= ~m_;                                                     \
                while (m_) {                                                      \
                    u32 l_ = __ffs(m_) - 1;                                       \
                    m_ &= m_ - 1;                                                 \
                    th_s[l_ * BSI_BLOCK + threadIdx.x] = (u32)t_;                 \
                }                                                                 \
            }                                                                     \
            stop_ = t_ >= T || (!alive && t_ >= BSI_LP_MAX);                      \
        } else {                                                                  \
            stop_ = t_ >= T;                                                      \
            if (!stop_) {                                                         \
                u32 p_ = ~(alive & started), q_ = 0u;                             \
                BSI_CMP_FILTER(p_, q_, X, c)                                      \
                p_ |= q_;                                                         \
                if (p_ != 0xFFFFFFFFu) {                                          \
                    BSI_CMP_REST(p_, X, c)                                        \
                    u32 m_ = ~p_;                                                 \
                    closed |= m_;                                                 \
                    alive &= ~m_;                                                 \
                }                                                                 \
                stop_ = !alive;                                                   \
            }                                                                     \
        }                                                                         \
    }
        // two steps per iteration (a -> b -> a), so that no step needs a register copy.
        // General steps (perturbation switch, per-lane start times, 64-bit clock) are only needed up to the
        // last perturbation time of any lane; from there on pass 0 runs a tight loop -- update, target
        // compare, a 32-bit countdown -- in which every lane is started and nothing is perturbed.
        bool stop_ = false, cur_is_b = false;
        const u64 lp_max_ = BSI_LP_MAX;
        const u64 gen_end = (pass == 0 && T > lp_max_ && T - lp_max_ < 0xFFFFFFF0ull) ? lp_max_ : ~0ull;
        u64 t = 0;
        for (; t <= gen_end; t += 2) {
            BSI_TCHECK(a, t)
            if (stop_) { cur_is_b = false; break; }
            BSI_UPDATE(b, a)
            BSI_PERTURB(b, (t + 1))
            BSI_TCHECK(b, t + 1)
            if (stop_) { cur_is_b = true; ++t; break; }
            BSI_UPDATE(a, b)
            BSI_PERTURB(a, (t + 2))
        }
        if (!stop_) {
            // here: pass 0, state a at time t (not yet tested), lp_max < t <= T, every valid lane started
            u32 left = (u32)(T - t);            // updates until the horizon
            const u32 t0 = (u32)t;              // (finite target times are below 2^32: bsi_target_run)
            u32 k = 0;                          // updates done in this loop
#define BSI_TFAST(X)                                                              \
    {                                                                             \
        u32 acc = ~alive;                                                         \
        BSI_CMP_TARGET(acc, X)                                                    \
        u32 m_ = ~acc;                                                            \
        if (m_) {                                                                 \
            alive &= ~m_;                                                         \
            while (m_) {                                                          \
                u32 l_ = __ffs(m_) - 1;                                           \
                m_ &= m_ - 1;                                                     \
                th_s[l_ * BSI_BLOCK + threadIdx.x] = t0 + k;                      \
            }                                                                     \
            if (!alive) break;                                                    \
        }                                                                         \
    }
            for (;;) {
                cur_is_b = false;
                BSI_TFAST(a)
                if (k == left) break;
                BSI_UPDATE_QUIET(b, a)
                ++k;
                cur_is_b = true;
                BSI_TFAST(b)
                if (k == left) break;
                BSI_UPDATE_QUIET(a, b)
                ++k;
            }
#undef BSI_TFAST
            hit = valid & ~alive;               // (pass 0: lanes leave `alive` only by hitting)
            t += k;
        }
        dev_steps += t;
#undef BSI_TCHECK
        if (pass == 0) {
            if (!check_closure || !alive) break;
#define BSI_COPY_CB2(i) c##i = b##i;
            if (cur_is_b) { BSI_FOR_NODES(BSI_COPY_CB2) } else { BSI_FOR_NODES(BSI_COPY_CA) }
#undef BSI_COPY_CB2
        }
    }

    u64 steps = 0;
    // the word's hits take consecutive places of the list, in lane (= index) order: one atomic per word
    u64 pos = (hit & valid) ? atomicAdd(&ctr->n_hits, (u64)__popc(hit & valid)) : 0ull;
#pragma unroll 1
    for (u32 l = 0; l < 32; ++l) {
        if (!((valid >> l) & 1u)) continue;
        u64 q = (w << 5) + l;
        u64 idx = in_queue ? in_queue[q] : idx_begin + q;
        if ((hit >> l) & 1u) {
            if (pos < hit_cap) {
                hit_idx[pos] = idx;
                hit_t[pos] = th_s[l * BSI_BLOCK + threadIdx.x];
            }
            ++pos;
            steps += th_s[l * BSI_BLOCK + threadIdx.x];
        } else if (check_closure && ((closed >> l) & 1u)) {
            closed_q[atomicAdd(&ctr->n_found, 1ull)] = idx;
        } else if (check_closure) {
            retry_q[atomicAdd(&ctr->n_retry, 1ull)] = idx;
        } else {
            atomicAdd(&ctr->n_without, 1ull);
            steps += T > lp_s[l * BSI_BLOCK + threadIdx.x] ? T : lp_s[l * BSI_BLOCK + threadIdx.x];
        }
    }
    if (steps) atomicAdd(&ctr->steps, steps);
    if (dev_steps) atomicAdd(&ctr->device_steps, dev_steps);
}
#endif  // BSI_MODE == 2
#endif  // BSI_UNIT 4
