/*
 * Test double for libb200stencil.so — TEST INFRASTRUCTURE ONLY (tests/test_zz_cinterface.py).
 *
 * It exports b2_iso_forward / b2_tti_forward with the real signatures but computes NOTHING: each
 * call deep-copies the argument block it received so that the test can check, without a GPU, that the
 * adapter written by `Operator.cinterface()` marshals the reference-style flat argument list into
 * `struct b2_iso_args` / `struct b2_tti_args` correctly.
 */
#include <string.h>
#include "b200stencil.h"

static struct b2_iso_args iso_seen;
static struct b2_tti_args tti_seen;
static struct b2_sparse src_seen, rec_seen;
static float w_seen[6][16];
static int calls = 0;
static int next_rc = 0;

static void keep_sparse(struct b2_sparse **slot, struct b2_sparse *copy) {
    if (*slot) {
        *copy = **slot;
        *slot = copy;
    }
}

static void fill_timers(struct b2_profiler *t) {
    if (!t) return;
    t->section0 += 1.5;
    t->section1 += 0.25;
    t->section2 += 0.125;
}

int b2_iso_forward(const struct b2_iso_args *a) {
    iso_seen = *a;
    for (int d = 0; d < a->ndim; ++d) {
        memcpy(w_seen[d], a->w[d], sizeof(float) * (a->radius + 1));
        iso_seen.w[d] = w_seen[d];
    }
    keep_sparse(&iso_seen.src, &src_seen);
    keep_sparse(&iso_seen.rec, &rec_seen);
    fill_timers(a->timers);
    iso_seen.timers = 0;
    ++calls;
    return next_rc;
}

int b2_tti_forward(const struct b2_tti_args *a) {
    tti_seen = *a;
    for (int d = 0; d < 3; ++d) {
        memcpy(w_seen[d], a->w2[d], sizeof(float) * (a->radius + 1));
        memcpy(w_seen[3 + d], a->w1[d], sizeof(float) * a->radius);
        tti_seen.w2[d] = w_seen[d];
        tti_seen.w1[d] = w_seen[3 + d];
    }
    keep_sparse(&tti_seen.src, &src_seen);
    keep_sparse(&tti_seen.rec, &rec_seen);
    fill_timers(a->timers);
    tti_seen.timers = 0;
    ++calls;
    return next_rc;
}

const struct b2_iso_args *stub_iso_seen(void) { return &iso_seen; }
const struct b2_tti_args *stub_tti_seen(void) { return &tti_seen; }
int stub_calls(void) { return calls; }
void stub_set_rc(int rc) { next_rc = rc; }
