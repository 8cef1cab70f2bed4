/*
 * Plain C against the libbsi C ABI (include/bsi.h): `attract` for a three-node network over all
 * 2^3 initial states.
 *
 *   gcc -std=c99 -Iinclude examples/attract_example.c -Lboolsi_b200 -lbsi \
 *       -Wl,-rpath,$PWD/boolsi_b200 -o build/attract_example
 *   build/attract_example            # runs on CUDA device 0 (a B200)
 *   build/attract_example --offline  # specialises the kernels only (NVRTC needs no GPU)
 *
 * Network (predecessors sorted ascending, truth-table entry j has the FIRST predecessor as its
 * most significant bit -- reference input.py:796, 881):
 *   node 0:  x0' = x1              node 1:  x1' = x0 AND x2          node 2:  x2' = NOT x2
 *
 * Expected (the oracle's answer): 22 trajectory-steps, two attractors of length 2 -- key 0
 * (states 0, 4) and key 2 (states 2, 5), each with frequency 4 and trajectory lengths summing to 3.
 */
#include <inttypes.h>
#include <stdio.h>
#include <string.h>

#include "bsi.h"

static int fail(const char *what) {
    fprintf(stderr, "%s: %s\n", what, bsi_last_error());
    return 1;
}

int main(int argc, char **argv) {
    const int offline = argc > 1 && strcmp(argv[1], "--offline") == 0;
    if (bsi_abi_version() != BSI_ABI_VERSION) return fail("ABI version mismatch");

    const uint32_t pred_off[4] = {0, 1, 3, 4};
    const uint32_t pred_idx[4] = {1, /* node 1: */ 0, 2, /* node 2: */ 2};
    const uint64_t tt_off[4] = {0, 2, 6, 8};
    const uint8_t tt[8] = {0, 1, /* x0 AND x2, j = 2*x0 + x2: */ 0, 0, 0, 1, /* NOT x2: */ 1, 0};
    bsi_network_desc net = {3, pred_off, pred_idx, tt_off, tt};

    const uint8_t initial_state[3] = {0, 0, 0};
    const uint32_t var_kind[3] = {BSI_VAR_INITIAL, BSI_VAR_INITIAL, BSI_VAR_INITIAL};
    const uint32_t var_node[3] = {0, 1, 2};
    const uint64_t var_t[3] = {0, 0, 0};
    const uint32_t var_range[3] = {BSI_RANGE_TRUE_OR_FALSE, BSI_RANGE_TRUE_OR_FALSE, BSI_RANGE_TRUE_OR_FALSE};
    bsi_space_desc space;
    memset(&space, 0, sizeof space);
    space.initial_state = initial_state;
    space.n_variations = 3;
    space.var_kind = var_kind;
    space.var_node = var_node;
    space.var_t = var_t;
    space.var_range = var_range;

    bsi_plan *plan = NULL;
    int rc = offline ? bsi_plan_create_offline(&net, &space, BSI_MODE_ATTRACT, NULL, NULL, &plan)
                     : bsi_plan_create(&net, &space, BSI_MODE_ATTRACT, NULL, NULL, 0, &plan);
    if (rc != BSI_OK) return fail("bsi_plan_create");
    uint64_t n_problems = 0;
    uint32_t update = 0, compare = 0, filter = 0;
    if (bsi_plan_n_problems(plan, &n_problems) || bsi_plan_op_counts(plan, &update, &compare, &filter))
        return fail("plan queries");
    printf("problems %" PRIu64 ", lop3 per 32-lane step: update %u, compare %u\n", n_problems, update, compare);

    bsi_run_stats stats;
    rc = bsi_attract_run(plan, 0, n_problems, BSI_INF, BSI_INF, &stats);
    if (offline) {
        /* no device behind an offline plan: the run must be refused, loudly */
        printf("offline plan: run refused with status %d (%s)\n", rc, bsi_last_error());
        bsi_plan_destroy(plan);
        return rc == BSI_E_CUDA ? 0 : 1;
    }
    if (rc != BSI_OK) return fail("bsi_attract_run");
    bsi_attractor_rec recs[8];
    uint64_t n = 0;
    if (bsi_attract_table_read(plan, recs, 8, &n) != BSI_OK) return fail("bsi_attract_table_read");
    printf("%" PRIu64 " trajectory-steps, %" PRIu64 " attractors\n", stats.steps, n);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t states[8];
        if (bsi_attractor_states(plan, recs[i].key, recs[i].lambda, states) != BSI_OK) return fail("bsi_attractor_states");
        printf("  key %" PRIu64 " length %" PRIu64 " frequency %" PRIu64 " sum of trajectory lengths %" PRIu64 " states:",
               recs[i].key[0], recs[i].lambda, recs[i].n, recs[i].sum_l);
        for (uint64_t t = 0; t < recs[i].lambda; ++t) printf(" %" PRIu64, states[t]);
        printf("\n");
    }
    bsi_plan_destroy(plan);
    return 0;
}
