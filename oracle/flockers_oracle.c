/*
 * flockers_oracle.c — CPU restatement of the krABMaga Flockers hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (examples_b200/) never links, imports or calls this file.
 *
 * PARITY UNPINNED: the reference (krABMaga/examples) ships no tests, golden vectors or
 * fixtures for this path, cannot be built here (no Rust toolchain) and the engine crate
 * `krabmaga 0.5.*` (Field2D, toroidal_*, Schedule) is not vendored under /root/reference.
 *   - Bird::step below is a line-by-line restatement of flockers/src/model/bird.rs:27-145,
 *     which IS verifiable (same text in flockers_mpi/src/model/bird.rs:84-183).
 *   - Field2D / toroidal_* / Schedule semantics follow the published upstream algorithm of
 *     krabmaga 0.5 (a port of MASON Continuous2D) as recorded in SURVEY.md Appendix B; every
 *     such assumption is an explicit orc_cfg parameter (E1..E10 of SURVEY.md §8c).
 *
 * Arithmetic: all f32, evaluated in the order written, compile with -ffp-contract=off
 * (Rust never contracts a*b+c into an FMA).
 *
 * Canonical order: the reference's per-cell bag order is the order in which agents stepped in
 * the previous tick, which krABMaga's Schedule (a priority queue with equal priorities) leaves
 * unspecified.  The oracle's schedule steps agents in ascending id (the order `init` enqueues
 * them, flockers/src/model/state.rs:40-51), hence every bag is in ascending-id order.  With
 * cfg.canonical_bags=1 (default) lazy_update / insert_read additionally re-sort bags by id so
 * that threaded and rank-decomposed runs reproduce the single-thread order bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- Bird (flockers/src/model/bird.rs:13-18; MPI wire form flockers_mpi/src/model/bird.rs:31-57) ---- */
typedef struct {
    uint32_t id;
    float x, y;     /* loc: Real2D     */
    float ldx, ldy; /* last_d: Real2D  */
} orc_bird;

/* ---- model constants (flockers/src/main.rs:26-33) + [EXT] switches ---- */
typedef struct {
    float cohesion, avoidance, randomness, consistency, momentum, jump; /* main.rs:26-31 */
    float radius;       /* bird.rs:31 : 10.0 */
    int relax;          /* 1 = get_neighbors_within_relax_distance (shipped), 0 = exact query */
    int window_plus1;   /* E3: 0 = floor(dist/d) (krabmaga), 1 = floor(dist/d)+1 (MASON) */
    int canonical_bags; /* 1 = keep bags sorted by id */
    int wrap_with_slack;/* E4: 0 = window indices wrap modulo ceil(W/d) (recalled), 1 = modulo ceil(W/d)+1 (slack cell on the ring) */
    int slack_to_zero;  /* E5: 0 = a location with cell index ceil(W/d) stays in the slack cell (recalled), 1 = it wraps to cell 0 */
} orc_cfg;

ORC_API void orc_cfg_default(orc_cfg *c) {
    c->cohesion = 0.8f; c->avoidance = 1.0f; c->randomness = 1.1f;
    c->consistency = 0.7f; c->momentum = 1.0f; c->jump = 0.7f;
    c->radius = 10.0f; c->relax = 1; c->window_plus1 = 0; c->canonical_bags = 1;
    c->wrap_with_slack = 0; c->slack_to_zero = 0;
}

/* Rust `f as i32`: saturating, NaN -> 0 */
static int32_t f2i_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

/* ---- toroidal helpers (called at bird.rs:54-55,137-138) [EXT: SURVEY Appendix B, E8] ---- */
ORC_API float orc_toroidal_transform(float v, float dim) {
    if (v >= 0.0f && v < dim) return v;
    v = fmodf(v, dim);
    if (v < 0.0f) v += dim;
    return v; /* may be exactly dim for tiny negative input (E5) */
}

ORC_API float orc_toroidal_distance(float a, float b, float dim) {
    if (fabsf(a - b) <= dim / 2.0f) return a - b;
    float d = orc_toroidal_transform(a, dim) - orc_toroidal_transform(b, dim);
    if (d * 2.0f > dim) return d - dim;
    if (d * 2.0f < -dim) return d + dim;
    return d;
}

/* ---- Philox4x32-10 (Salmon et al., SC'11), own restatement; KATs in tests ---- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

/* rand 0.9 `rng.random::<f32>()` [EXT E10]: 24 high bits * 2^-24, in [0,1) */
static inline float u32_to_unit(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

/* draws for (seed, id, step): counter = (id, step_lo, step_hi, 0), key = (seed_lo, seed_hi) */
ORC_API void orc_draw(uint64_t seed, uint32_t id, uint64_t step, float *r1, float *r2) {
    uint32_t ctr[4] = {id, (uint32_t)step, (uint32_t)(step >> 32), 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    orc_philox4x32_10(ctr, key, o);
    *r1 = u32_to_unit(o[0]);
    *r2 = u32_to_unit(o[1]);
}

/* ---- Field2D<Bird> [EXT: SURVEY Appendix B; call sites state.rs:24,49,55, bird.rs:29-31,142-144] ---- */
typedef struct {
    orc_bird *v;
    uint32_t n, cap;
    volatile int lock;
} orc_bag;

typedef struct {
    float width, height, disc;
    int toroidal;
    int dw, dh;   /* allocated cells: ceil(W/d)+1, ceil(H/d)+1 (E2: one slack row/col) */
    int mx, my;   /* wrap modulus: ceil(W/d), ceil(H/d) */
    orc_bag *bags[2];
    int read, write;
    int canonical;
    int e4_wrap_with_slack, e5_slack_to_zero; /* [EXT] alternatives, see orc_cfg (defaults 0 = recalled upstream behaviour) */
} orc_field;

ORC_API orc_field *orc_field_new(float w, float h, float d, int toroidal, int canonical) {
    orc_field *f = (orc_field *)calloc(1, sizeof(orc_field));
    f->width = w; f->height = h; f->disc = d; f->toroidal = toroidal;
    f->mx = f2i_sat(ceilf(w / d));
    f->my = f2i_sat(ceilf(h / d));
    f->dw = f->mx + 1;
    f->dh = f->my + 1;
    size_t nc = (size_t)f->dw * (size_t)f->dh;
    f->bags[0] = (orc_bag *)calloc(nc, sizeof(orc_bag));
    f->bags[1] = (orc_bag *)calloc(nc, sizeof(orc_bag));
    f->read = 0; f->write = 1;
    f->canonical = canonical;
    return f;
}

ORC_API void orc_field_free(orc_field *f) {
    if (!f) return;
    size_t nc = (size_t)f->dw * (size_t)f->dh;
    for (int b = 0; b < 2; ++b) {
        for (size_t i = 0; i < nc; ++i) free(f->bags[b][i].v);
        free(f->bags[b]);
    }
    free(f);
}

ORC_API void orc_field_dims(const orc_field *f, int *dw, int *dh, int *mx, int *my) {
    *dw = f->dw; *dh = f->dh; *mx = f->mx; *my = f->my;
}

/* E1: cell of a location (E5 alternative: the slack index wraps to 0) */
static inline void discretize(const orc_field *f, float x, float y, int *cx, int *cy) {
    *cx = f2i_sat(floorf(x / f->disc));
    *cy = f2i_sat(floorf(y / f->disc));
    if (f->e5_slack_to_zero) {
        if (*cx == f->mx) *cx = 0;
        if (*cy == f->my) *cy = 0;
    }
}

/* the [EXT] alternatives of E4 / E5 (parity-pinning kit: which setting reproduces a dump of the real engine?) */
ORC_API void orc_field_set_semantics(orc_field *f, int wrap_with_slack, int slack_to_zero) {
    f->e4_wrap_with_slack = wrap_with_slack;
    f->e5_slack_to_zero = slack_to_zero;
}

ORC_API void orc_discretize(const orc_field *f, float x, float y, int *cx, int *cy) { discretize(f, x, y, cx, cy); }

static inline void bag_push(orc_bag *b, const orc_bird *o) {
    while (__sync_lock_test_and_set(&b->lock, 1)) { }
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 4;
        b->v = (orc_bird *)realloc(b->v, (size_t)b->cap * sizeof(orc_bird));
    }
    b->v[b->n++] = *o;
    __sync_lock_release(&b->lock);
}

static int cmp_id(const void *a, const void *b) {
    uint32_t ia = ((const orc_bird *)a)->id, ib = ((const orc_bird *)b)->id;
    return (ia > ib) - (ia < ib);
}

static inline void bag_sort(orc_bag *b) {
    /* insertion sort: bags are tiny and almost sorted */
    for (uint32_t i = 1; i < b->n; ++i) {
        orc_bird t = b->v[i];
        uint32_t j = i;
        while (j > 0 && b->v[j - 1].id > t.id) { b->v[j] = b->v[j - 1]; --j; }
        b->v[j] = t;
    }
    (void)cmp_id;
}

static inline orc_bag *cell_bag(const orc_field *f, int which, int cx, int cy) {
    /* the reference would panic on an out-of-range index; clamp instead (documented) */
    if (cx < 0) cx = 0; if (cx >= f->dw) cx = f->dw - 1;
    if (cy < 0) cy = 0; if (cy >= f->dh) cy = f->dh - 1;
    return &f->bags[which][(size_t)cx * (size_t)f->dh + (size_t)cy];
}

/* a4: set_object_location — appends the whole Bird to the WRITE bag of its cell */
ORC_API void orc_set_object_location(orc_field *f, const orc_bird *o, float x, float y) {
    int cx, cy;
    discretize(f, x, y, &cx, &cy);
    orc_bird b = *o;
    b.x = x; b.y = y;
    bag_push(cell_bag(f, f->write, cx, cy), &b);
}

/* flockers_mpi/src/model/state.rs:128-130: ghost goes straight into the READ buffer */
ORC_API void orc_insert_read(orc_field *f, const orc_bird *o) {
    int cx, cy;
    discretize(f, o->x, o->y, &cx, &cy);
    orc_bag *b = cell_bag(f, f->read, cx, cy);
    bag_push(b, o);
    if (f->canonical) bag_sort(b);
}

/* a11: lazy_update — swap read/write, clear the new write bags */
ORC_API void orc_lazy_update(orc_field *f) {
    int t = f->read; f->read = f->write; f->write = t;
    size_t nc = (size_t)f->dw * (size_t)f->dh;
    orc_bag *w = f->bags[f->write];
    orc_bag *r = f->bags[f->read];
    const int canon = f->canonical;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nc; ++i) {
        w[i].n = 0;
        if (canon && r[i].n > 1) bag_sort(&r[i]);
    }
}

static inline int wrap_mod(int i, int m) { int r = i % m; return r < 0 ? r + m : r; }

/* window half-width (E3) */
static inline int window_k(const orc_field *f, float dist, int plus1) {
    return f2i_sat(floorf(dist / f->disc)) + (plus1 ? 1 : 0);
}

typedef struct { orc_bird *v; size_t n, cap; } orc_vec;
static inline void vec_push(orc_vec *q, const orc_bird *o) {
    if (q->n == q->cap) {
        q->cap = q->cap ? q->cap * 2 : 64;
        q->v = (orc_bird *)realloc(q->v, q->cap * sizeof(orc_bird));
    }
    q->v[q->n++] = *o;
}

/* a5 / a5': the two neighbour queries, READ buffer, x-outer / y-inner, bag order (E4,E6,E7) */
static void query(const orc_field *f, float x, float y, float dist, int relax, int plus1, orc_vec *out) {
    out->n = 0;
    if (dist <= 0.0f) return;
    int k = window_k(f, dist, plus1);
    int cx, cy;
    discretize(f, x, y, &cx, &cy);
    for (int i = cx - k; i <= cx + k; ++i) {
        for (int j = cy - k; j <= cy + k; ++j) {
            int bx = i, by = j;
            if (f->toroidal) {
                bx = wrap_mod(i, f->e4_wrap_with_slack ? f->dw : f->mx);
                by = wrap_mod(j, f->e4_wrap_with_slack ? f->dh : f->my);
            }
            else if (i < 0 || j < 0 || i >= f->dw || j >= f->dh) continue;
            const orc_bag *b = &f->bags[f->read][(size_t)bx * (size_t)f->dh + (size_t)by];
            for (uint32_t e = 0; e < b->n; ++e) {
                if (relax) { vec_push(out, &b->v[e]); continue; }
                float ddx, ddy;
                if (f->toroidal) {
                    ddx = orc_toroidal_distance(x, b->v[e].x, f->width);
                    ddy = orc_toroidal_distance(y, b->v[e].y, f->height);
                } else { ddx = x - b->v[e].x; ddy = y - b->v[e].y; }
                if (sqrtf(ddx * ddx + ddy * ddy) <= dist) vec_push(out, &b->v[e]);
            }
        }
    }
}

ORC_API int64_t orc_get_neighbors(const orc_field *f, float x, float y, float dist, int relax, int plus1,
                                  orc_bird *out, int64_t cap) {
    orc_vec q = {0, 0, 0};
    query(f, x, y, dist, relax, plus1, &q);
    int64_t n = (int64_t)q.n;
    for (int64_t i = 0; i < n && i < cap; ++i) out[i] = q.v[i];
    free(q.v);
    return n;
}

/* read-buffer dump in canonical (cell-major, bag) order; returns count; cell_start has dw*dh+1 entries */
ORC_API int64_t orc_field_dump(const orc_field *f, orc_bird *out, int64_t cap, uint32_t *cell_start) {
    size_t nc = (size_t)f->dw * (size_t)f->dh;
    int64_t n = 0;
    for (size_t c = 0; c < nc; ++c) {
        if (cell_start) cell_start[c] = (uint32_t)n;
        const orc_bag *b = &f->bags[f->read][c];
        for (uint32_t e = 0; e < b->n; ++e) { if (out && n < cap) out[n] = b->v[e]; ++n; }
    }
    if (cell_start) cell_start[nc] = (uint32_t)n;
    return n;
}

/* ---- Bird::step, flockers/src/model/bird.rs:27-145, operation order as written ---- */
static void bird_step(orc_field *f, orc_bird *self, const orc_cfg *c, float r1, float r2, orc_vec *vec, int *drew) {
    query(f, self->x, self->y, c->radius, c->relax, c->window_plus1, vec); /* :29-31 */
    const float width = f->width, height = f->height;                      /* :33-34 */
    float avo_x = 0.0f, avo_y = 0.0f, coh_x = 0.0f, coh_y = 0.0f;          /* :36-40 */
    float rnd_x = 0.0f, rnd_y = 0.0f, con_x = 0.0f, con_y = 0.0f;
    *drew = 0;
    if (vec->n != 0) {                                                     /* :42 */
        float x_avoid = 0.0f, y_avoid = 0.0f, x_cohe = 0.0f, y_cohe = 0.0f, x_cons = 0.0f, y_cons = 0.0f;
        int count = 0;                                                     /* :44-50 */
        for (size_t e = 0; e < vec->n; ++e) {                              /* :52 */
            const orc_bird *el = &vec->v[e];
            if (self->id != el->id) {                                      /* :53 */
                float dx = orc_toroidal_distance(self->x, el->x, width);   /* :54 */
                float dy = orc_toroidal_distance(self->y, el->y, height);  /* :55 */
                count += 1;                                                /* :56 */
                float square = dx * dx + dy * dy;                          /* :59 */
                x_avoid += dx / (square * square + 1.0f);                  /* :60 */
                y_avoid += dy / (square * square + 1.0f);                  /* :61 */
                x_cohe += dx;                                              /* :64 */
                y_cohe += dy;                                              /* :65 */
                x_cons += el->ldx;                                         /* :68 */
                y_cons += el->ldy;                                         /* :69 */
            }
        }
        if (count > 0) {                                                   /* :73 */
            float cf = (float)count;
            x_avoid /= cf; y_avoid /= cf;                                  /* :74-75 */
            x_cohe /= cf; y_cohe /= cf;                                    /* :76-77 */
            x_cons /= cf; y_cons /= cf;                                    /* :78-79 */
            con_x = x_cons / cf;                                           /* :81-84 second division: kept */
            con_y = y_cons / cf;
        } else {
            con_x = x_cons; con_y = y_cons;                                /* :85-90 */
        }
        avo_x = 400.0f * x_avoid; avo_y = 400.0f * y_avoid;                /* :92-95 */
        coh_x = -x_cohe / 10.0f; coh_y = -y_cohe / 10.0f;                  /* :97-100 */
        float x_rand = r1 * 2.0f - 1.0f;                                   /* :104-105 */
        float y_rand = r2 * 2.0f - 1.0f;                                   /* :106-107 */
        float square = sqrtf(x_rand * x_rand + y_rand * y_rand);           /* :109 */
        rnd_x = 0.05f * x_rand / square;                                   /* :110-113 */
        rnd_y = 0.05f * y_rand / square;
        *drew = 1;
    }
    float mom_x = self->ldx, mom_y = self->ldy;                            /* :116 */
    float dx = c->cohesion * coh_x + c->avoidance * avo_x + c->consistency * con_x
             + c->randomness * rnd_x + c->momentum * mom_x;                /* :118-122 */
    float dy = c->cohesion * coh_y + c->avoidance * avo_y + c->consistency * con_y
             + c->randomness * rnd_y + c->momentum * mom_y;                /* :123-127 */
    float dis = sqrtf(dx * dx + dy * dy);                                  /* :129 */
    if (dis > 0.0f) {                                                      /* :130 */
        dx = dx / dis * c->jump;                                           /* :131 */
        dy = dy / dis * c->jump;                                           /* :132 */
    }
    self->ldx = dx; self->ldy = dy;                                        /* :135 */
    float loc_x = orc_toroidal_transform(self->x + dx, width);             /* :137 */
    float loc_y = orc_toroidal_transform(self->y + dy, height);            /* :138 */
    self->x = loc_x; self->y = loc_y;                                      /* :140 */
    orc_set_object_location(f, self, loc_x, loc_y);                        /* :142-144 */
}

/* single Bird::step with explicit draws; returns 1 if the two draws were consumed */
ORC_API int orc_bird_step(orc_field *f, orc_bird *self, const orc_cfg *c, float r1, float r2) {
    orc_vec v = {0, 0, 0};
    int drew = 0;
    bird_step(f, self, c, r1, r2, &v, &drew);
    free(v.v);
    return drew;
}

/* ---- Schedule + Flocker state (flockers/src/model/state.rs:12-56; Schedule::step [EXT Appendix B]) ---- */
typedef struct {
    orc_field *field;
    orc_cfg cfg;
    orc_bird *agents; /* the Schedule's own copies, ascending id order */
    int64_t n;
    uint64_t step;
} orc_sim;

ORC_API orc_sim *orc_sim_new(float w, float h, float d, int toroidal, const orc_cfg *cfg) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    s->cfg = *cfg;
    s->field = orc_field_new(w, h, d, toroidal, cfg->canonical_bags);
    orc_field_set_semantics(s->field, cfg->wrap_with_slack, cfg->slack_to_zero);
    return s;
}

ORC_API void orc_sim_free(orc_sim *s) {
    if (!s) return;
    orc_field_free(s->field);
    free(s->agents);
    free(s);
}

ORC_API orc_field *orc_sim_field(orc_sim *s) { return s->field; }
ORC_API int64_t orc_sim_num_agents(const orc_sim *s) { return s->n; }
ORC_API uint64_t orc_sim_steps_done(const orc_sim *s) { return s->step; }

/* Flocker::init (state.rs:37-52) with the placement supplied by the caller: insert into the
 * WRITE buffer and enqueue a copy in the schedule.  Birds must be given in ascending id. */
ORC_API void orc_sim_init(orc_sim *s, const orc_bird *birds, int64_t n) {
    s->agents = (orc_bird *)realloc(s->agents, (size_t)(n > 0 ? n : 1) * sizeof(orc_bird));
    memcpy(s->agents, birds, (size_t)n * sizeof(orc_bird));
    s->n = n;
    s->step = 0;
    for (int64_t i = 0; i < n; ++i) orc_set_object_location(s->field, &birds[i], birds[i].x, birds[i].y);
}

/* Flocker::init placement itself (state.rs:41-48): loc = (W*r1, H*r2), last_d = 0, draws from
 * Philox(seed, id, ctr=0) instead of the unseeded thread RNG */
ORC_API void orc_make_uniform(uint64_t seed, float w, float h, int64_t n, orc_bird *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float r1, r2;
        orc_draw(seed, (uint32_t)i, 0, &r1, &r2);
        out[i].id = (uint32_t)i;
        out[i].x = w * r1; out[i].y = h * r2;
        out[i].ldx = 0.0f; out[i].ldy = 0.0f;
    }
}

/* one Schedule::step: (step==0: update) ; every agent's step ; update.
 * injected (nullable): float pairs indexed by id; else Philox(seed, id, step). */
ORC_API void orc_sim_step(orc_sim *s, uint64_t seed, const float *injected, int threads) {
#ifdef _OPENMP
    /* every parallel region of the tick (the agents' loop AND lazy_update's bag sort) uses the requested thread count,
     * whatever OMP_NUM_THREADS said when libgomp was loaded (torchrun exports OMP_NUM_THREADS=1 to its ranks) */
    if (threads > 1) omp_set_num_threads(threads);
#endif
    if (s->step == 0) orc_lazy_update(s->field); /* state.update(0) */
    const uint64_t step = s->step;
    orc_field *f = s->field;
    const orc_cfg *c = &s->cfg;
    if (threads <= 1) {
        orc_vec v = {0, 0, 0};
        for (int64_t i = 0; i < s->n; ++i) {
            orc_bird *a = &s->agents[i];
            float r1, r2; int drew;
            if (injected) { r1 = injected[2 * (size_t)a->id]; r2 = injected[2 * (size_t)a->id + 1]; }
            else orc_draw(seed, a->id, step, &r1, &r2);
            bird_step(f, a, c, r1, r2, &v, &drew);
        }
        free(v.v);
    } else {
#pragma omp parallel num_threads(threads)
        {
            orc_vec v = {0, 0, 0};
#pragma omp for schedule(static)
            for (int64_t i = 0; i < s->n; ++i) {
                orc_bird *a = &s->agents[i];
                float r1, r2; int drew;
                if (injected) { r1 = injected[2 * (size_t)a->id]; r2 = injected[2 * (size_t)a->id + 1]; }
                else orc_draw(seed, a->id, step, &r1, &r2);
                bird_step(f, a, c, r1, r2, &v, &drew);
            }
            free(v.v);
        }
    }
    s->step += 1;
    orc_lazy_update(f); /* state.update(step+1) */
}

ORC_API void orc_sim_agents(const orc_sim *s, orc_bird *out) { memcpy(out, s->agents, (size_t)s->n * sizeof(orc_bird)); }

/* ---- rank-decomposed variant (flockers_mpi): the caller (a test, over gloo) moves the birds ---- */

/* remove every agent for which keep[i]==0, then append `arrivals`; keeps ascending-id order.
 * Mirrors after_step (flockers_mpi/src/model/state.rs:139-175): leavers are dequeued, arrivals
 * scheduled and inserted (into the WRITE buffer, visible after the next update). */
ORC_API void orc_sim_replace_agents(orc_sim *s, const orc_bird *agents, int64_t n) {
    s->agents = (orc_bird *)realloc(s->agents, (size_t)(n > 0 ? n : 1) * sizeof(orc_bird));
    memcpy(s->agents, agents, (size_t)n * sizeof(orc_bird));
    qsort(s->agents, (size_t)n, sizeof(orc_bird), cmp_id);
    s->n = n;
}

/* the agents-only part of a tick for the rank-decomposed protocol: no implicit update calls.
 * Each stepped bird is inserted in the WRITE buffer only if owner(new loc)==this strip; else it is
 * reported in out_leavers (flockers_mpi/src/model/bird.rs:190-195).  Strip = global cell columns
 * [col0,col1) plus, for the strip holding column 0, the slack column mx. */
ORC_API int64_t orc_sim_step_strip(orc_sim *s, uint64_t step, uint64_t seed, const float *injected,
                                   int col0, int col1, orc_bird *out_leavers, int64_t cap) {
    orc_field *f = s->field;
    const orc_cfg *c = &s->cfg;
    int64_t nl = 0, kept = 0;
    orc_vec v = {0, 0, 0};
    for (int64_t i = 0; i < s->n; ++i) {
        orc_bird *a = &s->agents[i];
        float r1, r2; int drew;
        if (injected) { r1 = injected[2 * (size_t)a->id]; r2 = injected[2 * (size_t)a->id + 1]; }
        else orc_draw(seed, a->id, step, &r1, &r2);
        /* bird_step inserts into the write bag; pop it again if the bird left the strip */
        bird_step(f, a, c, r1, r2, &v, &drew);
        int cx, cy;
        discretize(f, a->x, a->y, &cx, &cy);
        int mine = (cx >= col0 && cx < col1) || (cx >= f->mx && col0 == 0);
        if (!mine) {
            orc_bag *b = cell_bag(f, f->write, cx, cy);
            b->n -= 1; /* the push just made is the last element */
            if (nl < cap) out_leavers[nl] = *a;
            ++nl;
        } else {
            s->agents[kept++] = *a;
        }
    }
    free(v.v);
    s->n = kept;
    return nl;
}

/* the same for a 2-D block (tile): global cell columns [col0,col1) x rows [row0,row1); the slack column / row belong
 * to the blocks holding column 0 / row 0 (flockers_mpi partitions the field in 2-D blocks, state.rs:33,75) */
ORC_API int64_t orc_sim_step_tile(orc_sim *s, uint64_t step, uint64_t seed, const float *injected,
                                  int col0, int col1, int row0, int row1, orc_bird *out_leavers, int64_t cap) {
    orc_field *f = s->field;
    const orc_cfg *c = &s->cfg;
    int64_t nl = 0, kept = 0;
    orc_vec v = {0, 0, 0};
    for (int64_t i = 0; i < s->n; ++i) {
        orc_bird *a = &s->agents[i];
        float r1, r2; int drew;
        if (injected) { r1 = injected[2 * (size_t)a->id]; r2 = injected[2 * (size_t)a->id + 1]; }
        else orc_draw(seed, a->id, step, &r1, &r2);
        bird_step(f, a, c, r1, r2, &v, &drew);
        int cx, cy;
        discretize(f, a->x, a->y, &cx, &cy);
        int mine = ((cx >= col0 && cx < col1) || (cx >= f->mx && col0 == 0)) &&
                   ((cy >= row0 && cy < row1) || (cy >= f->my && row0 == 0));
        if (!mine) {
            orc_bag *b = cell_bag(f, f->write, cx, cy);
            b->n -= 1; /* the push just made is the last element */
            if (nl < cap) out_leavers[nl] = *a;
            ++nl;
        } else {
            s->agents[kept++] = *a;
        }
    }
    free(v.v);
    s->n = kept;
    return nl;
}

/* birds of the READ buffer in the cell rectangle [c0,c1) x [r0,r1) (no wrap: the caller passes wrapped pieces) */
ORC_API int64_t orc_field_cells(const orc_field *f, int c0, int c1, int r0, int r1, orc_bird *out, int64_t cap) {
    int64_t n = 0;
    for (int cx = c0; cx < c1; ++cx)
        for (int cy = r0; cy < r1; ++cy) {
            const orc_bag *b = &f->bags[f->read][(size_t)cx * (size_t)f->dh + (size_t)cy];
            for (uint32_t e = 0; e < b->n; ++e) { if (n < cap) out[n] = b->v[e]; ++n; }
        }
    return n;
}

/* arrivals after migration: schedule + insert into WRITE (state.rs:167-174) */
ORC_API void orc_sim_add_agents(orc_sim *s, const orc_bird *arr, int64_t n) {
    s->agents = (orc_bird *)realloc(s->agents, (size_t)(s->n + n > 0 ? s->n + n : 1) * sizeof(orc_bird));
    for (int64_t i = 0; i < n; ++i) {
        s->agents[s->n + i] = arr[i];
        orc_set_object_location(s->field, &arr[i], arr[i].x, arr[i].y);
    }
    s->n += n;
    qsort(s->agents, (size_t)s->n, sizeof(orc_bird), cmp_id);
}

/* birds of the READ buffer whose cell column is in [c0,c1) (halo source, state.rs:122-124) */
ORC_API int64_t orc_field_columns(const orc_field *f, int c0, int c1, orc_bird *out, int64_t cap) {
    int64_t n = 0;
    for (int cx = c0; cx < c1; ++cx)
        for (int cy = 0; cy < f->dh; ++cy) {
            const orc_bag *b = &f->bags[f->read][(size_t)cx * (size_t)f->dh + (size_t)cy];
            for (uint32_t e = 0; e < b->n; ++e) { if (n < cap) out[n] = b->v[e]; ++n; }
        }
    return n;
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
