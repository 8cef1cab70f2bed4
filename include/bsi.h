/*
 * libbsi -- C ABI of the B200-native BoolSi hot path.
 *
 * One shared library, `extern "C"`, plain pointers and sizes, no C++/torch types, no exceptions
 * across the boundary.  It replaces, for the data-parallel path only, what the reference does in
 * Python between the three `*_master` call sites of its CLI (reference boolsi/cli.py:200-204,
 * 256-261, 314-319) and the DB post-conditions its writers read:
 *
 *   rule evaluation + one synchronous step        boolsi/model.py:16-73
 *   perturbation prefix + run-until-stop loop     boolsi/model.py:76-236
 *   per-trajectory attractor extraction           boolsi/attract.py:262-302
 *   aggregation by canonical key                  boolsi/attract.py:374-455  (ZODB store, db.py)
 *   target-substate matching                      boolsi/target.py:109-133
 *   fixed-length simulation                       boolsi/simulate.py:97-131
 *   problem enumeration (index -> problem)        boolsi/batching.py:10-46, 160-282
 *   the task farm / per-rank fan-out              boolsi/mpi.py:77-190, 498-544
 *
 * Semantics follow the reference's DEFAULT detection mode (storing all states).  All results are
 * integers and bit-exact; trajectory-length statistics are returned as the exact integer triple
 * (n, sum, sum of squares) from which the caller derives mean / variation sum.
 *
 * Ownership: the caller owns every buffer it passes (host or device, as documented per function)
 * and the library never retains a caller pointer past the call, except the CUDA stream handle set
 * with bsi_plan_set_stream.  Library state lives behind the opaque bsi_plan (create/destroy).
 * Every function returns BSI_OK (0) or a negative bsi_status; bsi_last_error() gives the message
 * (thread-local).  There is no CPU fallback: without a usable CUDA device calls fail with
 * BSI_E_CUDA.  A plan is bound to one device and is not re-entrant; calls are synchronous unless
 * stated otherwise.
 */
#ifndef BSI_H_
#define BSI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSI_ABI_VERSION 2
#define BSI_KEY_WORDS_MAX 4               /* 64-bit words per state code: n_nodes <= 256 */
#define BSI_INF UINT64_MAX                /* "no cap" for max_t, max_attractor_l, n_to_find */

typedef enum bsi_status {
    BSI_OK = 0,
    BSI_E_ARG = -1,           /* invalid argument / descriptor */
    BSI_E_CUDA = -2,          /* CUDA driver/runtime error, or no device */
    BSI_E_NVRTC = -3,         /* kernel specialisation failed; bsi_last_error() has the log */
    BSI_E_CAPACITY = -4,      /* an output buffer or internal table is too small */
    BSI_E_UNSUPPORTED = -5,   /* e.g. n_nodes > 256, a rule with too many inputs */
    BSI_E_NOMEM = -6
} bsi_status;

typedef enum bsi_mode { BSI_MODE_SIMULATE = 0, BSI_MODE_ATTRACT = 1, BSI_MODE_TARGET = 2 } bsi_mode;

/* Node state ranges of a variation (reference boolsi/constants.py:27-34, batching.py:171-175). */
typedef enum bsi_range {
    BSI_RANGE_TRUE_OR_FALSE = 0,        /* digit 0 -> 0, 1 -> 1                 ("any")  */
    BSI_RANGE_MAYBE_FALSE = 1,          /* digit 0 -> off, 1 -> 0               ("0?")   */
    BSI_RANGE_MAYBE_TRUE = 2,           /* digit 0 -> off, 1 -> 1               ("1?")   */
    BSI_RANGE_MAYBE_TRUE_OR_FALSE = 3   /* digit 0 -> off, 1 -> 0, 2 -> 1       ("any?") */
} bsi_range;

typedef enum bsi_variation_kind { BSI_VAR_INITIAL = 0, BSI_VAR_FIXED = 1, BSI_VAR_PERTURBATION = 2 }
    bsi_variation_kind;

/*
 * The network: what parse_input hands over as 'incoming node lists' and 'truth tables'
 * (reference input.py:124-157).  pred_idx lists each node's predecessors in the reference's
 * order (sorted ascending, input.py:796); tt holds one byte (0/1) per truth-table entry, entry j
 * of node i at tt[tt_off[i] + j], where the FIRST predecessor is the MOST significant bit of j
 * (itertools.product order, input.py:881; lookup model.py:26-27).
 */
typedef struct bsi_network_desc {
    uint32_t n_nodes;
    const uint32_t *pred_off;     /* [n_nodes + 1] */
    const uint32_t *pred_idx;     /* [pred_off[n_nodes]] */
    const uint64_t *tt_off;       /* [n_nodes + 1] */
    const uint8_t *tt;            /* [tt_off[n_nodes]] */
} bsi_network_desc;

/*
 * The problem space: 'origin simulation problem' + 'simulation problem variations'
 * (reference input.py:146-152).  Problem `index` is the mixed-radix number whose digits, least
 * significant first, are the variations in the order given here: initial-state variations, then
 * fixed-node variations, then perturbation variations, each list in the order the reference
 * holds it (batching.py:10-46; do not re-sort).  A perturbation at time t overrides both the
 * update rule and a fixed node (model.py:70-71); fixed nodes act from t = 1 on (model.py:45-47).
 */
typedef struct bsi_space_desc {
    const uint8_t *initial_state;          /* [n_nodes] origin initial state, 0/1 */
    uint32_t n_fixed;
    const uint32_t *fixed_node;
    const uint8_t *fixed_value;
    uint32_t n_perturbations;              /* origin perturbations, any order */
    const uint64_t *pert_t;
    const uint32_t *pert_node;
    const uint8_t *pert_value;
    uint32_t n_variations;
    const uint32_t *var_kind;              /* bsi_variation_kind */
    const uint32_t *var_node;
    const uint64_t *var_t;                 /* perturbation time (kind 2), else 0 */
    const uint32_t *var_range;             /* bsi_range; BSI_VAR_INITIAL must be TRUE_OR_FALSE */
} bsi_space_desc;

/* Target substate for BSI_MODE_TARGET ('target node set' / 'target substate code'). */
typedef struct bsi_target_desc {
    uint64_t mask[BSI_KEY_WORDS_MAX];      /* bit i set <=> node i is constrained */
    uint64_t code[BSI_KEY_WORDS_MAX];      /* required value on the constrained nodes */
} bsi_target_desc;

/* Tuning knobs; zero-initialise for defaults. */
typedef struct bsi_plan_opts {
    uint32_t first_horizon;    /* attract: horizon of the first detection round; 0 = auto (1024) */
    uint32_t filter_nodes;     /* attract: #nodes in the cheap pre-compare; 0 = auto (12) */
    uint32_t table_log2_slots; /* attract: device hash table size; 0 = auto (20) */
    uint32_t chunk_log2;       /* problems per launch wave; 0 = auto (26) */
    uint32_t no_dict;          /* attract: 1 = no cycle dictionary (every matched lane walks its cycle for the
                                  canonical key; for comparison only) */
    uint32_t block_threads;    /* threads per block of the trajectory kernels; 0 = auto (64) */
    uint32_t min_blocks;       /* __launch_bounds__ min blocks per SM (register cap); 0 = auto (1) */
    uint32_t debug_stats;      /* 1: count filter events (stats->reserved[0] = stage-2 filter entries,
                                  [1] = full compares); costs atomics, for tuning only */
    uint32_t resolve_few_words;/* attract: queue length (words of 32 matched lanes) up to which the
                                  latency-oriented resolution kernels are used; 0 = auto (32768) */
    uint32_t no_early_words;   /* attract: 1 = route every matched lane through per-lane records (turns
                                  the bit-sliced early-word queue off; for comparison only) */
    uint32_t smem_slots;       /* attract: how many of the three checkpoint slots live in shared memory
                                  (the others in the thread's L2-resident scratch); 0 = auto (3).
                                  2 with min_blocks 6 buys a third warp per scheduler at N = 64 */
    uint32_t source_only;      /* offline plans only: 1 = generate the specialised source (bsi_plan_source,
                                  bsi_plan_op_counts) and skip NVRTC; the plan has no cubin */
    uint32_t keep_dict;        /* attract: 1 = bsi_attract_table_clear leaves the cycle dictionary as it is (it
                                  holds facts about the network's dynamics, not results) */
    uint32_t tune;             /* experiment switches of the kernel template (BSI_TUNE), 0 = the shipped variant */
    uint32_t target_hit_list;  /* target with a finite max_t: capacity of the hit list that all chunks of a run
                                  append to without a host round trip between them; 0 = auto (2^24 hits), 1 = off
                                  (chunk by chunk, what a run with more hits than the list holds falls back to) */
    uint32_t reserved[1];
} bsi_plan_opts;

/* One aggregated attractor: the replacement for AggregatedAttractor (reference attract.py:18-45). */
typedef struct bsi_attractor_rec {
    uint64_t key[BSI_KEY_WORDS_MAX];  /* min state code over the cycle, bit i = node i */
    uint64_t n;                       /* frequency */
    uint64_t sum_l;                   /* sum of trajectory lengths (absolute entry times) */
    uint64_t sum_l2;                  /* sum of squared trajectory lengths */
    uint64_t lambda;                  /* attractor length */
} bsi_attractor_rec;

typedef struct bsi_run_stats {
    uint64_t n_problems;       /* problems processed by this call */
    uint64_t n_found;          /* attract: problems whose attractor passed both caps; target: hits */
    uint64_t n_without;        /* attract: problems with no attractor within the caps */
    uint64_t steps;            /* sum over problems of the reference's stop time (model.py:235);
                                  target mode: exact for hits; a problem without a hit counts max_t
                                  (an upper bound when its trajectory closes its cycle earlier) when
                                  max_t is finite and NOTHING when max_t = BSI_INF (it stops at its
                                  cycle closure, a time this mode does not work out) */
    uint64_t device_steps;     /* 32-lane word-steps actually executed by the main kernels */
    uint64_t launches;         /* kernels launched by this call */
    uint64_t rounds;           /* attract / target: detection rounds over all chunks */
    double ms_total;           /* device time of the call (CUDA events on the plan's stream) */
    double ms_main;            /* device time of the dominant (trajectory) kernels */
    double ms_resolve;         /* device time of the per-attractor resolution kernels */
    uint64_t reserved[6];
} bsi_run_stats;

/* ---- library (no counterpart in the reference: versioning, diagnostics, device discovery) -- */
int bsi_abi_version(void);
/* Text of the calling thread's last failure; what the Python wrapper raises, as cli.py:146-149
 * logs any exception of a run. */
const char *bsi_last_error(void);
int bsi_device_count(int *count);
/* Plans park freed device buffers for reuse by later plans of the same process (and keep
 * the device's primary context retained); this returns the parked memory to the driver. */
int bsi_release_cached_memory(void);

/* ---- plan: network + problem space + mode, specialised through NVRTC for one device ------- */
/* Compiled kernels are cached per process and, so that later processes skip the few seconds of NVRTC
 * per network, on disk: in $BSI_CACHE_DIR if set ("" or "off" disable the disk level), else in
 * $XDG_CACHE_HOME/boolsi_b200, else in ~/.cache/boolsi_b200. */
/* Replaces the per-run set-up of the reference: configure_solve_simulation_problem
 * (mpi.py:351-376: the mode's per-trajectory function with its caps bound), the fixed-node
 * rewrite of the rules (model.py:31-49, re-done per batch at mpi.py:525-528) and the numeral
 * system of the problem space (batching.py:10-46). */
typedef struct bsi_plan bsi_plan;

int bsi_plan_create(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                    const bsi_target_desc *target /* NULL unless BSI_MODE_TARGET */,
                    const bsi_plan_opts *opts /* NULL for defaults */, int device,
                    bsi_plan **out);
/* Same, but specialise only (NVRTC needs no GPU): the plan exposes source / cubin / op counts and
 * every run function fails on it with BSI_E_CUDA.  Used by the build check and the CPU tests. */
int bsi_plan_create_offline(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                            const bsi_target_desc *target, const bsi_plan_opts *opts,
                            bsi_plan **out);
void bsi_plan_destroy(bsi_plan *plan);
int bsi_plan_n_problems(const bsi_plan *plan, uint64_t *n);   /* 'total combination count', input.py:124-157 */
int bsi_plan_key_words(const bsi_plan *plan, uint32_t *kw);
/* A plan starts on its own non-blocking stream.  cuda_stream: a cudaStream_t / CUstream handle;
 * NULL = the CUDA default (legacy) stream; (void *)-1 = back to the plan's own stream. */
int bsi_plan_set_stream(bsi_plan *plan, void *cuda_stream);
/* Progress of long runs (replaces the reference's TickingProgressBar feed, utils.py:19-54, mpi.py:164):
 * `fn(done, total, user)` is called from the run functions after every finished wave of problems
 * (2^chunk_log2 of them; the host has waited for a wave when the run needs several detection rounds,
 * i.e. always when max_t has no cap) -- from the calling thread for a plan, from the member's host
 * thread for a group.  NULL switches it off.  The callback must not call back into the plan. */
typedef void (*bsi_progress_fn)(uint64_t done, uint64_t total, void *user);
int bsi_plan_set_progress(bsi_plan *plan, bsi_progress_fn fn, void *user);
/* Generated CUDA source / build log (NUL-terminated, owned by the plan; diagnostics, no reference
 * counterpart). */
const char *bsi_plan_source(const bsi_plan *plan);
const char *bsi_plan_build_log(const bsi_plan *plan);
/* The sm_100a cubin NVRTC produced for the plan's dominant kernel unit (owned by the plan):
 * the dense attract trajectory kernel, or the trajectory/target unit in the other modes. */
int bsi_plan_cubin(const bsi_plan *plan, const void **data, uint64_t *bytes);
/* Persistent-grid geometry of the attract trajectory kernel on the plan's device. */
int bsi_plan_launch_config(const bsi_plan *plan, uint32_t *n_sms, uint32_t *block_threads,
                           uint32_t *blocks_per_sm);
/* Static cost model of the specialised step: LOP3 per 32-lane word-step (update, compare). */
int bsi_plan_op_counts(const bsi_plan *plan, uint32_t *update_lop3, uint32_t *compare_lop3,
                       uint32_t *filter_lop3);
/* LOP3 per word-step of the update past the last perturbation time, where nodes fixed for every lane
 * are folded into their readers as constants (model.py:31-49); equals update_lop3 without fixed nodes.
 * Negative = error code. */
int bsi_plan_update_quiet_lop3(const bsi_plan *plan);

/*
 * Specialise a plan's kernels WITHOUT a device (NVRTC only): used by the build check and the CPU
 * test-suite.  Writes the cubin size; the plan cannot run.
 */
int bsi_compile_only(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                     const bsi_target_desc *target, const bsi_plan_opts *opts,
                     uint64_t *cubin_bytes, char *log, uint64_t log_cap);

/* The checkpoint schedule a detection round with this horizon uses (times relative to the last
 * perturbation; see the header comment of bsi_template.cuh): exposed so that the proof obligation
 * -- every trajectory with mu + lambda <= horizon is matched by time `end` -- can be tested. */
int bsi_attract_schedule(uint64_t horizon, uint32_t *ck_time, uint32_t *ck_slot, uint32_t cap,
                         uint32_t *n_ck, uint32_t *end, uint32_t *first_final);

/* ---- attract (replaces attract_master's simulation + collection stages, attract.py:67-138) - */
/* Runs problems [idx_begin, idx_end) and ACCUMULATES into the plan's device attractor table. */
int bsi_attract_run(bsi_plan *plan, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                    uint64_t max_attractor_l, bsi_run_stats *stats);
/* The same, without waiting: everything is enqueued on the plan's stream and the call returns (when
 * the run needs more than one detection round -- max_t beyond the first horizon -- the host must
 * read queue lengths between rounds and the call only returns after the last round is enqueued).
 * Table exports / merges / the next run may be enqueued behind it; bsi_attract_run_finish waits,
 * returns the statistics and reports errors.  One run may be pending per plan. */
int bsi_attract_run_async(bsi_plan *plan, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                          uint64_t max_attractor_l);
int bsi_attract_run_finish(bsi_plan *plan, bsi_run_stats *stats);
int bsi_attract_table_clear(bsi_plan *plan);
int bsi_attract_table_count(bsi_plan *plan, uint64_t *n);
/* Copy the table to HOST memory (unsorted). */
int bsi_attract_table_read(bsi_plan *plan, bsi_attractor_rec *out, uint64_t cap, uint64_t *n);
/* Compact the table into a caller DEVICE buffer of `cap` records (async on the plan's stream;
 * records beyond the count are zeroed, n == 0 marks an empty record): the per-GPU payload of the
 * one NCCL allgather that replaces the MPI collection stage (mpi.py:296-332). */
int bsi_attract_table_export_device(bsi_plan *plan, void *dev_out, uint64_t cap, uint64_t *n);
/* The same without any host round trip (asynchronous on the plan's stream): dev_out[0] becomes a
 * HEADER record -- key = {records in the table, trajectory-steps, accepted, without} of the last
 * bsi_attract_run, n = 0, sum_l = table-overflow flag -- and dev_out[1 .. cap) the records
 * (truncated if the table holds more than cap - 1; the header still carries the true count). */
int bsi_attract_table_export_payload(bsi_plan *plan, void *dev_out, uint64_t cap);
/* Reduce-by-key `n` records in DEVICE memory into the table (integer adds: exact, order-free);
 * records with n == 0 are skipped.  Asynchronous on the plan's stream; a table overflow is
 * reported by the next bsi_attract_table_count / _read / _export_device. */
int bsi_attract_table_merge_device(bsi_plan *plan, const void *dev_records, uint64_t n);
/* Orbit of `key` under the fixed-node-adjusted rules: out[lambda][key_words], key state first
 * (the rotation of attract.py:22-25).  HOST output. */
int bsi_attractor_states(bsi_plan *plan, const uint64_t *key, uint64_t lambda, uint64_t *out);
/* The same for `n` keys in one launch, `n_states` states each (a state sequence longer than the
 * attractor just wraps around): keys[n][key_words] in, out[n][n_states][key_words].  HOST buffers. */
int bsi_attractor_states_batch(bsi_plan *plan, const uint64_t *keys, uint64_t n, uint64_t n_states,
                               uint64_t *out);

/* Orbits of `n` attractors, each exactly its own length: keys[n][key_words], lams[n] in; `states` out,
 * orbit i at state offset lams[0] + .. + lams[i-1], key state first (the rotation of attract.py:22-25).
 * Orbits the run stored in its cycle dictionary are copied from there; the others are entered into it
 * first (one fast walk per attractor).  With `counts` != NULL also their node activity
 * (counts[n][n_nodes], see bsi_attractor_activity), counted on the device where the orbits lie, in the
 * same pass.  HOST buffers. */
int bsi_attractor_orbits(bsi_plan *plan, const uint64_t *keys, const uint64_t *lams, uint64_t n,
                         uint64_t *states, uint64_t *counts);
/* Node activity of attractors (attractor_analysis.py:49-51: each attractor as an observation in [0,1]^N):
 * counts[i][node] = number of states of attractor i's cycle in which `node` is on (its mean activity is
 * counts / lambda).  Counted on the device from the orbits -- the ones the run stored, else regenerated.
 * keys[n][key_words], lams[n] in; counts[n][n_nodes] out.  HOST buffers. */
int bsi_attractor_activity(bsi_plan *plan, const uint64_t *keys, const uint64_t *lams, uint64_t n,
                           uint64_t *counts);

/* ---- simulate (replaces simulate_master's simulation stage, simulate.py:46-78) ------------- */
/* states_out (HOST): [idx_end - idx_begin][max_t + 1][key_words], plain iteration with masks. */
int bsi_simulate_run(bsi_plan *plan, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                     uint64_t *states_out, bsi_run_stats *stats);
/* Same for an explicit list of problem indices (HOST in/out). */
int bsi_simulate_list(bsi_plan *plan, const uint64_t *indices, uint64_t n, uint64_t max_t,
                      uint64_t *states_out, bsi_run_stats *stats);

/* ---- target (replaces target_master's simulation stage, target.py:13-88) ------------------ */
/*
 * Problems [idx_begin, idx_end) whose trajectory meets the target substate at some time in
 * [last perturbation, min(max_t, cycle closure)]; the first `cap` hits in index order go to
 * hit_index / hit_t (HOST); *n_hits is the number written, stats->n_found the total number.
 */
int bsi_target_run(bsi_plan *plan, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                   uint64_t *hit_index, uint64_t *hit_t, uint64_t cap, uint64_t *n_hits,
                   bsi_run_stats *stats);

/* ---- groups: one plan per device in ONE host process (replaces the MPI task farm, mpi.py:193-340) */
/*
 * A group holds one plan per device for the same network / space / mode (kernels are specialised once
 * and loaded on every device).  Its run functions split [idx_begin, idx_end) into contiguous,
 * word-aligned slices, one per device, and drive every device from its own host thread:
 *  - attract: afterwards member 0's table holds the WHOLE job (read it with bsi_group_plan(g, 0) and
 *    bsi_attract_table_read / bsi_attractor_states_batch).  The collection stage (mpi.py:296-332) is
 *    the peers' export kernels writing their compacted tables into device 0's memory over NVLink and
 *    one merge launch there; integer sums, so the result is bit-identical for any number of devices.
 *    Tables start from empty on every call, and grow (job repeated) when the attractors do not fit.
 *  - target: hits of all slices in index order, the first `cap` of them (simulate.py:163-164).
 *  - simulate: every device writes its slice of the host output.
 * devices == NULL: devices 0 .. n_devices-1; n_devices <= 0: every visible device.
 */
typedef struct bsi_group bsi_group;
int bsi_group_create(const bsi_network_desc *network, const bsi_space_desc *space, int mode,
                     const bsi_target_desc *target, const bsi_plan_opts *opts, const int *devices,
                     int n_devices, bsi_group **out);
void bsi_group_destroy(bsi_group *group);
int bsi_group_size(const bsi_group *group, int *n_devices);
/* Member i's plan, owned by the group. */
int bsi_group_plan(bsi_group *group, int i, bsi_plan **plan);
/* total: sums over the members (times: the slowest member); per_device: [n_devices] or NULL. */
int bsi_group_attract_run(bsi_group *group, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                          uint64_t max_attractor_l, bsi_run_stats *total, bsi_run_stats *per_device);
int bsi_group_target_run(bsi_group *group, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                         uint64_t *hit_index, uint64_t *hit_t, uint64_t cap, uint64_t *n_hits,
                         bsi_run_stats *total);
int bsi_group_simulate_run(bsi_group *group, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                           uint64_t *states_out, bsi_run_stats *total);

/* ---- measurement helpers (bench.py's roofline denominators; no reference counterpart) ------ */
/* Integer-logic roofline microbenchmark: independent LOP3 chains on every SM of `device`;
 * returns thread-level lop3.b32 results per second and the SM clock (MHz) seen during the run. */
int bsi_measure_lop3_peak(int device, double *lop3_per_s, double *sm_clock_mhz, int *n_sms);

/* The bare update loop of the plan's own network (no compare, no stores), all SMs busy:
 * 32-lane word-steps per second and the lop3.b32 rate that implies.  The practical ceiling of the
 * trajectory kernels (arbitrary three-register operands), to set beside the ideal-chain peak. */
int bsi_measure_update_rate(bsi_plan *plan, double *word_steps_per_s, double *lop3_per_s);

#ifdef __cplusplus
}
#endif
#endif /* BSI_H_ */
