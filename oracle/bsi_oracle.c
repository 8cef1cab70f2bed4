/*
 * CPU restatement in plain C of the reference's hot path (scalar, one trajectory at a time,
 * truth-table lookups, a hash set of all visited states -- the reference's default
 * "storing all states" mode).
 *
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and only as the checker
 * (or as the timed CPU baseline).  Nothing under boolsi_b200/ links or loads it.
 *
 * Parity status: PINNED -- tests/test_coracle_golden.py checks it against every fixture under
 * tests/golden/ (generated from the unmodified reference by tests/golden/make_golden.py).
 *
 * Deliberately a different algorithm from the product (which is bit-sliced, two-pass and never
 * stores visited states): agreement between the two is evidence, not tautology.
 *
 * Reference lines restated (all under /root/reference/boolsi/):
 *   orc_next_state      model.py:16-28 (apply_update_rules), model.py:31-49 (fixed nodes)
 *   orc_decode_problem  batching.py:212-229, batching.py:160-209, batching.py:10-46
 *   orc_core            model.py:76-128 (perturbation prefix), model.py:152-236 (storing mode)
 *   orc_attract         attract.py:262-302, mpi.py:498-544
 *   orc_target          target.py:109-133
 *   orc_simulate        simulate.py:97-131
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ORC_INF UINT64_MAX

typedef struct {
    uint32_t n_nodes;
    const uint32_t *pred_off;       /* [n_nodes+1] */
    const uint32_t *pred_idx;       /* predecessors, per node sorted ascending (input.py:796) */
    const uint64_t *tt_off;         /* [n_nodes+1] offsets into tt */
    const uint8_t *tt;              /* one byte per table entry; entry j of node i: first
                                       predecessor is the most significant bit of j (input.py:881) */
    const uint8_t *initial_state;   /* [n_nodes] origin initial state */
    uint32_t n_fixed;
    const uint32_t *fixed_node;
    const uint8_t *fixed_value;
    uint32_t n_pert;                /* origin perturbations */
    const uint64_t *pert_t;
    const uint32_t *pert_node;
    const uint8_t *pert_value;
    uint32_t n_var;                 /* variations in enumeration order, least significant first */
    const uint32_t *var_kind;       /* 0 initial state, 1 fixed node, 2 perturbation */
    const uint32_t *var_node;
    const uint64_t *var_t;
    const uint32_t *var_range;      /* 0 TRUE_OR_FALSE 1 MAYBE_FALSE 2 MAYBE_TRUE 3 MAYBE_TRUE_OR_FALSE */
} orc_problem_space;

typedef struct {
    uint64_t t;
    uint32_t node;
    uint8_t value;
} orc_pert;

typedef struct {
    int kw;
    uint64_t *state;      /* [kw] */
    uint64_t *next;       /* [kw] */
    int8_t *fixed;        /* [n_nodes] -1 free, else fixed value */
    orc_pert *perts;      /* this problem's perturbations */
    uint32_t n_perts;
    uint64_t lp;
    /* visited states since lp */
    uint64_t *hist;       /* [hist_cap][kw] */
    uint64_t hist_cap, hist_len;
    uint64_t *slots;      /* hash set: index+1 into hist, 0 empty */
    uint64_t n_slots;
} orc_work;

static int orc_getbit(const uint64_t *s, uint32_t i) { return (int)((s[i >> 6] >> (i & 63)) & 1u); }
static void orc_setbit(uint64_t *s, uint32_t i, int v) {
    if (v) s[i >> 6] |= (uint64_t)1 << (i & 63); else s[i >> 6] &= ~((uint64_t)1 << (i & 63));
}

/* model.py:16-28 with the fixed-node rewrite of model.py:31-49 folded in. */
static void orc_next_state(const orc_problem_space *p, const orc_work *w, const uint64_t *cur,
                           uint64_t *out) {
    memset(out, 0, sizeof(uint64_t) * (size_t)w->kw);
    for (uint32_t i = 0; i < p->n_nodes; ++i) {
        int v;
        if (w->fixed[i] >= 0) {
            v = w->fixed[i];
        } else {
            uint64_t j = 0;
            for (uint32_t e = p->pred_off[i]; e < p->pred_off[i + 1]; ++e)
                j = (j << 1) | (uint64_t)orc_getbit(cur, p->pred_idx[e]);
            v = p->tt[p->tt_off[i] + j];
        }
        if (v) out[i >> 6] |= (uint64_t)1 << (i & 63);
    }
}

/* model.py:52-73: perturbations scheduled for time t override rule and fixed value. */
static void orc_apply_perts(const orc_work *w, uint64_t t, uint64_t *s) {
    for (uint32_t k = 0; k < w->n_perts; ++k)
        if (w->perts[k].t == t) orc_setbit(s, w->perts[k].node, w->perts[k].value);
}

static uint32_t orc_radix(uint32_t range) { return range == 3 ? 3u : 2u; }

/* digit -> value; -1 = variation switched off (batching.py:171-175) */
static int orc_digit_value(uint32_t range, uint32_t digit) {
    switch (range) {
    case 0: return (int)digit;
    case 1: return digit ? 0 : -1;
    case 2: return digit ? 1 : -1;
    default: return digit == 0 ? -1 : (int)digit - 1;
    }
}

static void orc_decode_problem(const orc_problem_space *p, uint64_t index, orc_work *w) {
    memset(w->state, 0, sizeof(uint64_t) * (size_t)w->kw);
    for (uint32_t i = 0; i < p->n_nodes; ++i) {
        if (p->initial_state[i]) orc_setbit(w->state, i, 1);
        w->fixed[i] = -1;
    }
    for (uint32_t k = 0; k < p->n_fixed; ++k) w->fixed[p->fixed_node[k]] = (int8_t)p->fixed_value[k];
    w->n_perts = 0;
    for (uint32_t k = 0; k < p->n_pert; ++k) {
        w->perts[w->n_perts].t = p->pert_t[k];
        w->perts[w->n_perts].node = p->pert_node[k];
        w->perts[w->n_perts].value = p->pert_value[k];
        w->n_perts++;
    }
    for (uint32_t k = 0; k < p->n_var; ++k) {
        uint32_t radix = orc_radix(p->var_range[k]);
        uint32_t digit = (uint32_t)(index % radix);
        index /= radix;
        int v = orc_digit_value(p->var_range[k], digit);
        if (p->var_kind[k] == 0) {
            orc_setbit(w->state, p->var_node[k], (int)digit);
        } else if (v >= 0) {
            if (p->var_kind[k] == 1) {
                w->fixed[p->var_node[k]] = (int8_t)v;
            } else {
                /* perturbed_nodes_by_t[t][node] = v overrides an origin entry at the same place */
                uint32_t q;
                for (q = 0; q < w->n_perts; ++q)
                    if (w->perts[q].t == p->var_t[k] && w->perts[q].node == p->var_node[k]) break;
                w->perts[q].t = p->var_t[k];
                w->perts[q].node = p->var_node[k];
                w->perts[q].value = (uint8_t)v;
                if (q == w->n_perts) w->n_perts++;
            }
        }
    }
    w->lp = 0;                                         /* model.py:125 */
    for (uint32_t q = 0; q < w->n_perts; ++q) if (w->perts[q].t > w->lp) w->lp = w->perts[q].t;
}

static uint64_t orc_hash(const uint64_t *s, int kw) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < kw; ++i) {
        h ^= s[i];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 29;
    }
    return h;
}

static int orc_hist_grow(orc_work *w) {
    uint64_t cap = w->hist_cap ? w->hist_cap * 2 : 1024;
    uint64_t *h = (uint64_t *)realloc(w->hist, sizeof(uint64_t) * cap * (size_t)w->kw);
    if (!h) return -1;
    w->hist = h;
    w->hist_cap = cap;
    free(w->slots);
    w->n_slots = cap * 2;
    w->slots = (uint64_t *)calloc(w->n_slots, sizeof(uint64_t));
    if (!w->slots) return -1;
    for (uint64_t i = 0; i < w->hist_len; ++i) {
        uint64_t pos = orc_hash(w->hist + i * w->kw, w->kw) & (w->n_slots - 1);
        while (w->slots[pos]) pos = (pos + 1) & (w->n_slots - 1);
        w->slots[pos] = i + 1;
    }
    return 0;
}

/* returns index of an equal earlier state, or ORC_INF after inserting the new one */
static uint64_t orc_hist_find_or_add(orc_work *w, const uint64_t *s) {
    if (w->hist_len + 1 > w->hist_cap) { if (orc_hist_grow(w)) return ORC_INF - 1; }
    uint64_t pos = orc_hash(s, w->kw) & (w->n_slots - 1);
    while (w->slots[pos]) {
        uint64_t idx = w->slots[pos] - 1;
        if (!memcmp(w->hist + idx * w->kw, s, sizeof(uint64_t) * (size_t)w->kw)) {
            /* the repeated state is still appended (model.py:213-215) */
            memcpy(w->hist + w->hist_len * w->kw, s, sizeof(uint64_t) * (size_t)w->kw);
            w->hist_len++;
            return idx;
        }
        pos = (pos + 1) & (w->n_slots - 1);
    }
    memcpy(w->hist + w->hist_len * w->kw, s, sizeof(uint64_t) * (size_t)w->kw);
    w->slots[pos] = ++w->hist_len;
    return ORC_INF;
}

static void orc_hist_reset(orc_work *w) {
    if (w->slots) memset(w->slots, 0, sizeof(uint64_t) * w->n_slots);
    w->hist_len = 0;
}

static int orc_masked_equal(const uint64_t *s, const uint64_t *mask, const uint64_t *code, int kw) {
    if (!mask) return 0;                               /* no target: never matches (model.py:200) */
    for (int i = 0; i < kw; ++i) if ((s[i] & mask[i]) != code[i]) return 0;
    return 1;
}

typedef struct {
    uint64_t t;          /* stop time (model.py:235) */
    int found;           /* attractor_is_found */
    int reached;         /* target_substate_is_reached */
    uint64_t first_idx;  /* index since lp of the earlier equal state, when found */
} orc_core_result;

typedef void (*orc_state_sink)(void *ctx, uint64_t t, const uint64_t *state);

/* model.py:152-236, storing mode.  w->state holds the initial state on entry. */
static int orc_core(const orc_problem_space *p, orc_work *w, uint64_t max_t,
                    const uint64_t *target_mask, const uint64_t *target_code,
                    orc_state_sink sink, void *ctx, orc_core_result *r) {
    if (sink) sink(ctx, 0, w->state);
    for (uint64_t t = 0; t < w->lp; ++t) {             /* model.py:93-103 */
        orc_next_state(p, w, w->state, w->next);
        orc_apply_perts(w, t + 1, w->next);
        uint64_t *tmp = w->state; w->state = w->next; w->next = tmp;
        if (sink) sink(ctx, t + 1, w->state);
    }
    orc_hist_reset(w);
    if (orc_hist_find_or_add(w, w->state) == ORC_INF - 1) return -1;
    r->t = w->lp;
    r->found = 0;
    r->first_idx = 0;
    r->reached = orc_masked_equal(w->state, target_mask, target_code, w->kw);
    while (r->t < max_t && !r->found && !r->reached) {  /* model.py:201 */
        r->t++;
        orc_next_state(p, w, w->state, w->next);
        uint64_t *tmp = w->state; w->state = w->next; w->next = tmp;
        r->reached = orc_masked_equal(w->state, target_mask, target_code, w->kw);
        uint64_t idx = orc_hist_find_or_add(w, w->state);
        if (idx == ORC_INF - 1) return -1;
        if (idx != ORC_INF) { r->found = 1; r->first_idx = idx; }
        if (sink) sink(ctx, r->t, w->state);
    }
    return 0;
}

static int orc_work_init(const orc_problem_space *p, orc_work *w) {
    memset(w, 0, sizeof(*w));
    w->kw = (int)((p->n_nodes + 63) / 64);
    if (w->kw == 0) w->kw = 1;
    w->state = (uint64_t *)calloc((size_t)w->kw, sizeof(uint64_t));
    w->next = (uint64_t *)calloc((size_t)w->kw, sizeof(uint64_t));
    w->fixed = (int8_t *)malloc(p->n_nodes ? p->n_nodes : 1);
    w->perts = (orc_pert *)malloc(sizeof(orc_pert) * (size_t)(p->n_pert + p->n_var + 1));
    return (w->state && w->next && w->fixed && w->perts) ? 0 : -1;
}

static void orc_work_free(orc_work *w) {
    free(w->state); free(w->next); free(w->fixed); free(w->perts); free(w->hist); free(w->slots);
}


/* ---- a small pthread work-sharing loop (the image has no libgomp) ------------------------- */
typedef struct orc_job {
    const orc_problem_space *p;
    uint64_t idx_begin, idx_end, chunk;
    _Atomic uint64_t next;
    _Atomic uint64_t steps;
    _Atomic int status;
    void (*body)(struct orc_job *job, orc_work *w, uint64_t index, uint64_t *steps, int *bad);
    /* mode arguments */
    uint64_t max_t, max_attractor_l;
    const uint64_t *target_mask, *target_code;
    uint8_t *flag;
    uint64_t *key, *lambda, *traj_l, *stop_t, *states;
} orc_job;

static void *orc_thread_main(void *arg) {
    orc_job *job = (orc_job *)arg;
    orc_work w;
    int bad = orc_work_init(job->p, &w);
    uint64_t steps = 0;
    for (;;) {
        uint64_t b = atomic_fetch_add(&job->next, job->chunk);
        if (b >= job->idx_end || bad) break;
        uint64_t e = b + job->chunk < job->idx_end ? b + job->chunk : job->idx_end;
        for (uint64_t index = b; index < e && !bad; ++index) job->body(job, &w, index, &steps, &bad);
    }
    atomic_fetch_add(&job->steps, steps);
    if (bad) atomic_store(&job->status, -1);
    orc_work_free(&w);
    return NULL;
}

static int orc_run(orc_job *job, int n_threads, uint64_t *total_steps) {
    if (n_threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        n_threads = n > 0 ? (int)n : 1;
    }
    uint64_t n_items = job->idx_end - job->idx_begin;
    if ((uint64_t)n_threads > n_items / job->chunk + 1) n_threads = (int)(n_items / job->chunk + 1);
    atomic_store(&job->next, job->idx_begin);
    atomic_store(&job->steps, 0);
    atomic_store(&job->status, 0);
    if (n_threads <= 1) {
        orc_thread_main(job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
        int started = 0;
        for (int i = 0; i < n_threads; ++i)
            if (pthread_create(&th[started], NULL, orc_thread_main, job) == 0) started++;
        if (!started) orc_thread_main(job);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    if (total_steps) *total_steps = atomic_load(&job->steps);
    return atomic_load(&job->status);
}

static int orc_code_less(const uint64_t *a, const uint64_t *b, int kw) {
    for (int i = kw - 1; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] < b[i];
    }
    return 0;
}

int orc_key_words(const orc_problem_space *p) {
    int kw = (int)((p->n_nodes + 63) / 64);
    return kw ? kw : 1;
}

/*
 * attract.py:262-302 per problem.  Outputs (any may be NULL), indexed by (index - idx_begin):
 *   found[i]   1 iff an attractor within both caps was found
 *   key[i*kw]  min state code over the cycle           (attract.py:296)
 *   lambda[i]  attractor length                        (attract.py:291)
 *   traj_l[i]  absolute time of the first cycle state  (attract.py:298)
 *   stop_t[i]  t at which the reference's loop stops   (model.py:235) -- the metric's step count
 */
static void orc_attract_body(orc_job *job, orc_work *w, uint64_t index, uint64_t *steps, int *bad) {
    const orc_problem_space *p = job->p;
    int kw = w->kw;
    uint64_t o = index - job->idx_begin;
    orc_core_result r;
    orc_decode_problem(p, index, w);
    if (orc_core(p, w, job->max_t, NULL, NULL, NULL, NULL, &r)) { *bad = 1; return; }
    *steps += r.t;
    if (job->stop_t) job->stop_t[o] = r.t;
    int ok = 0;
    if (r.found) {
        uint64_t mu = r.first_idx;                          /* attract.py:290 */
        uint64_t lam = (w->hist_len - 1) - mu;              /* attract.py:291 */
        if (lam <= job->max_attractor_l) {                  /* attract.py:294 */
            const uint64_t *best = w->hist + mu * kw;
            for (uint64_t q = mu + 1; q < mu + lam; ++q)
                if (orc_code_less(w->hist + q * kw, best, kw)) best = w->hist + q * kw;
            if (job->key) memcpy(job->key + o * kw, best, sizeof(uint64_t) * (size_t)kw);
            if (job->lambda) job->lambda[o] = lam;
            if (job->traj_l) job->traj_l[o] = w->lp + mu;   /* attract.py:298 */
            ok = 1;
        }
    }
    if (job->flag) job->flag[o] = (uint8_t)ok;
    if (!ok) {
        if (job->key) memset(job->key + o * kw, 0, sizeof(uint64_t) * (size_t)kw);
        if (job->lambda) job->lambda[o] = 0;
        if (job->traj_l) job->traj_l[o] = 0;
    }
}

int orc_attract(const orc_problem_space *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                uint64_t max_attractor_l, uint8_t *found, uint64_t *key, uint64_t *lambda,
                uint64_t *traj_l, uint64_t *stop_t, uint64_t *total_steps, int n_threads) {
    orc_job job;
    memset(&job, 0, sizeof(job));
    job.p = p; job.idx_begin = idx_begin; job.idx_end = idx_end; job.chunk = 64;
    job.body = orc_attract_body;
    job.max_t = max_t; job.max_attractor_l = max_attractor_l;
    job.flag = found; job.key = key; job.lambda = lambda; job.traj_l = traj_l; job.stop_t = stop_t;
    return orc_run(&job, n_threads, total_steps);
}

/*
 * target.py:109-133 per problem: reached[i] = 1 iff the target substate is met at some
 * t in [lp, stop]; stop_t[i] = the time of the first hit (or where the loop gave up).
 */
static void orc_target_body(orc_job *job, orc_work *w, uint64_t index, uint64_t *steps, int *bad) {
    orc_core_result r;
    orc_decode_problem(job->p, index, w);
    if (orc_core(job->p, w, job->max_t, job->target_mask, job->target_code, NULL, NULL, &r)) {
        *bad = 1;
        return;
    }
    *steps += r.t;
    if (job->stop_t) job->stop_t[index - job->idx_begin] = r.t;
    if (job->flag) job->flag[index - job->idx_begin] = (uint8_t)r.reached;
}

int orc_target(const orc_problem_space *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
               const uint64_t *target_mask, const uint64_t *target_code, uint8_t *reached,
               uint64_t *stop_t, uint64_t *total_steps, int n_threads) {
    orc_job job;
    memset(&job, 0, sizeof(job));
    job.p = p; job.idx_begin = idx_begin; job.idx_end = idx_end; job.chunk = 64;
    job.body = orc_target_body;
    job.max_t = max_t; job.target_mask = target_mask; job.target_code = target_code;
    job.flag = reached; job.stop_t = stop_t;
    return orc_run(&job, n_threads, total_steps);
}

typedef struct { uint64_t *out; int kw; uint64_t max_t; } orc_sim_ctx;

static void orc_sim_sink(void *c, uint64_t t, const uint64_t *s) {
    orc_sim_ctx *ctx = (orc_sim_ctx *)c;
    if (t <= ctx->max_t) memcpy(ctx->out + t * ctx->kw, s, sizeof(uint64_t) * (size_t)ctx->kw);
}

/*
 * simulate.py:97-131 per problem: states[(i*(max_t+1) + t)*kw] for t = 0..max_t; after the cycle
 * closes the remaining states are the cycle tiled (simulate.py:119-129).
 */
static void orc_simulate_body(orc_job *job, orc_work *w, uint64_t index, uint64_t *steps, int *bad) {
    int kw = w->kw;
    orc_core_result r;
    orc_sim_ctx ctx;
    ctx.out = job->states + (index - job->idx_begin) * (job->max_t + 1) * kw;
    ctx.kw = kw;
    ctx.max_t = job->max_t;
    orc_decode_problem(job->p, index, w);
    if (orc_core(job->p, w, job->max_t, NULL, NULL, orc_sim_sink, &ctx, &r)) { *bad = 1; return; }
    *steps += r.t;
    if (job->stop_t) job->stop_t[index - job->idx_begin] = r.t;
    if (r.t < job->max_t) {
        /* cycle = the last lam states; state(t) for t > r.t repeats with period lam */
        uint64_t lam = (w->hist_len - 1) - r.first_idx;
        for (uint64_t t = r.t + 1; t <= job->max_t; ++t)
            memcpy(ctx.out + t * kw, ctx.out + (t - lam) * kw, sizeof(uint64_t) * (size_t)kw);
    }
}

int orc_simulate(const orc_problem_space *p, uint64_t idx_begin, uint64_t idx_end, uint64_t max_t,
                 uint64_t *states, uint64_t *stop_t, uint64_t *total_steps, int n_threads) {
    orc_job job;
    memset(&job, 0, sizeof(job));
    job.p = p; job.idx_begin = idx_begin; job.idx_end = idx_end; job.chunk = 16;
    job.body = orc_simulate_body;
    job.max_t = max_t; job.states = states; job.stop_t = stop_t;
    return orc_run(&job, n_threads, total_steps);
}
