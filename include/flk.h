/*
 * flk.h — C ABI of the B200-native Flockers hot path (libflk.so).
 *
 * This is the drop-in boundary: the entry points are what a krABMaga `Field2D<Bird>` /
 * `Flocker` binding needs (Rust `extern "C"`, see INTEGRATION.md and rust/src/ffi.rs).
 * Plain C types only: opaque handle, pointers, sizes.  No torch types, no callbacks.
 *
 * Citations are relative to the reference checkout (krABMaga/examples):
 *   flockers/src/model/bird.rs      Bird::step                      (a5..a10 in SURVEY.md §8)
 *   flockers/src/model/state.rs     Flocker::{new,init,update}       (a2,a3,a11)
 *   flockers_mpi/src/model/*.rs     halo / migration protocol        (a13)
 * Engine methods (Field2D::*) live in the external crate krabmaga 0.5.*; their call sites are cited.
 *
 * Conventions
 *   - every call returns FLK_OK (0) or a negative FLK_E_* code; nothing throws or aborts across
 *     the boundary; flk_last_error() returns a thread-local description of the last failure.
 *   - a handle is driven by ONE host thread (like krABMaga's Schedule); distinct handles are independent.
 *   - the library owns all device memory; host arrays passed in stay owned by the caller.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with FLK_E_CUDA.
 *   - `float2`-shaped arguments are plain `const float*` with x,y interleaved (Real2D{x,y}).
 *   - host buffer lifetime: calls that take HOST arrays (flk_set_objects*, flk_step's injected_r, flk_query_batch)
 *     enqueue their copies on the handle's stream.  From pageable memory the copy is staged before the call returns;
 *     from PINNED memory (flk_host_alloc, cudaHostAlloc, torch pin_memory) it runs asynchronously, so such a buffer
 *     must stay untouched until the next flk_synchronize on that handle returns.
 *   - device-side failures (a rank or a message over its capacity) are reported by the next flk_synchronize,
 *     flk_dist_num_owned or snapshot call as FLK_E_CAPACITY; the field contents are undefined afterwards.
 */
#ifndef FLK_H
#define FLK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define FLK_OK 0
#define FLK_E_INVALID (-1)     /* bad argument / null handle */
#define FLK_E_CUDA (-2)        /* CUDA runtime error (message in flk_last_error) */
#define FLK_E_CAPACITY (-3)    /* more objects than the capacity given at create / output truncated */
#define FLK_E_UNSUPPORTED (-4) /* configuration outside what the kernels implement */
#define FLK_E_NCCL (-5)        /* NCCL error / library not found */
#define FLK_E_STATE (-6)       /* call not valid in the current state */

typedef struct flk_field flk_field;

/* ---- lifecycle ----------------------------------------------------------------------------- */

/* Field2D::new(width, height, discretization, toroidal)           flockers/src/model/state.rs:24
 * capacity = max number of objects the field will ever hold; device = CUDA ordinal. */
int flk_field_create(float width, float height, float discretization, int toroidal,
                     uint64_t capacity, int device, flk_field **out);
/* Field2D<O> for an object type that carries more than Bird's {id, loc, last_d} (template/src/model/crab.rs:11-19,
 * virusnetwork/src/model/node.rs): payload_bytes (multiple of 4, <= 256) opaque bytes per object travel with it through
 * set_object_location, Bird::step (copied unchanged to the successor) and lazy_update; on strips and tiles
 * (flk_dist_create_ex / flk_mgpu_create_ex) they also travel in the halo and migration messages. */
int flk_field_create_ex(float width, float height, float discretization, int toroidal, uint64_t capacity, int device,
                        uint32_t payload_bytes, flk_field **out);
int flk_field_destroy(flk_field *f);

/* State::reset re-creates the field (flockers/src/model/state.rs:32-35): drop every object, keep memory. */
int flk_field_clear(flk_field *f);

/* model constants COHESION..JUMP (flockers/src/main.rs:26-31), query radius (bird.rs:31), and which
 * query Bird::step uses: relax=1 get_neighbors_within_relax_distance (shipped, bird.rs:29-31),
 * relax=0 get_neighbors_within_distance.  Defaults are the reference values. */
int flk_set_params(flk_field *f, float cohesion, float avoidance, float randomness, float consistency,
                   float momentum, float jump, float radius, int relax);

/* Engine semantics that could not be read off the reference (the krabmaga crate is not vendored; SURVEY.md §8c lists them
 * as E1..E11).  window_plus1 (E3): 0 (default) = the query window spans floor(dist / discretization) cells on either side,
 * 1 = one more (MASON's Continuous2D).  The parity-pinning kit finds out which one the real engine uses;
 * the CPU restatement under oracle/ carries the same switch plus the ones for E4 / E5. */
int flk_set_engine_semantics(flk_field *f, int window_plus1);

/* 0 = exact (default): every f32 operation of Bird::step is one IEEE-754 round-to-nearest operation, no FMA
 *     contraction, evaluated in the reference's order with the neighbours of a cell visited in ascending id (the
 *     canonical bag order).  Results are bit-identical to the repository's CPU oracle (oracle/), which
 *     restates bird.rs line by line; they equal a krABMaga run bit for bit only when its bags are in that order too
 *     (the engine's bag order is the schedule's execution order, which it leaves unspecified) — otherwise the f32
 *     sums differ in their last bits (quantified by tests/test_oracle.py::test_bag_order_sensitivity_is_within_tolerance).
 * 1 = fast: reciprocal-based division in the neighbour loop (<= 2 ulp per term, same order). */
int flk_set_math_mode(flk_field *f, int mode);

/* Storage layout of the double buffer.  All forms produce the same results bit for bit.
 *   0  sorted runs, classic WRITE buffer: READ is one array sorted by (cell, id); Bird::step appends successors in arrival
 *      order and lazy_update is a counting sort (scan, scatter, rank).  Any density, any handle.
 *   2  sorted runs, slot-staged WRITE buffer: Bird::step stores each successor straight into one of a fixed number of
 *      slots of its cell and lazy_update is scan + ONE pass that ranks and compacts every cell.  Single-GPU handles
 *      without payload; arrivals beyond a cell's slots go to a spill list (nothing is dropped).
 *   1  slots as the READ buffer too (fixed-capacity cells read by the step kernel directly): kept as a measured
 *      alternative, slower on B200, toroidal fields with the shipped query only.
 *  -1  (default) automatic = form 0: on B200 form 2 measured within 2 % of it and form 1 slower; forms 1
 *      and 2 are entered on request only, when the next upload is published, and left again (for form 0) when the spill
 *      count of a tick exceeds a handful of birds (forms requested through the environment are kept regardless).
 * flk_get_layout reports the form in use, the slots per cell and {objects, largest column (form 1), largest cell, objects
 * beyond their cell's slots} of the last lazy_update. */
int flk_set_layout(flk_field *f, int layout);
int flk_get_layout(flk_field *f, int32_t *layout, int32_t *slots_per_cell, uint32_t totals[4]);

/* grid geometry: allocated cells dw x dh (= ceil(W/d)+1, one slack row/col) and wrap modulus mx x my */
int flk_field_dims(const flk_field *f, int32_t *dw, int32_t *dh, int32_t *mx, int32_t *my);

/* ---- Field2D write side ------------------------------------------------------------------------- */

/* Field2D::set_object_location(bird, loc)   flockers/src/model/state.rs:49, bird.rs:142-144
 * Appends Bird{id, loc, last_d} to the WRITE buffer; visible to queries after flk_lazy_update. */
int flk_set_object_location(flk_field *f, uint32_t id, float x, float y, float last_dx, float last_dy);

/* bulk form of the above (n birds; HOST arrays; loc/last_d interleaved x,y) */
int flk_set_objects(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *last_d);

/* flk_set_objects with n * payload_bytes bytes of payload (NULL = zero-filled) */
int flk_set_objects_ex(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *last_d,
                       const void *payload);

/* Field::lazy_update()   flockers/src/model/state.rs:54-56
 * WRITE becomes READ (re-binned by cell, ascending id inside a cell), the new WRITE buffer is empty.
 * Asynchronous on the handle's stream. */
int flk_lazy_update(flk_field *f);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* One tick of `for every scheduled Bird: Bird::step(state)` (flockers/src/model/bird.rs:27-145):
 * reads the READ buffer, appends every bird's successor to the WRITE buffer.
 * Randomness (bird.rs:103-107): if injected_r != NULL it is a HOST array of n_injected (r1,r2) pairs,
 * the pair of bird `id` at [2*id],[2*id+1] (ids >= n_injected fall back to Philox);
 * else Philox4x32-10 with key=seed, counter=(id,step,0), r = (out >> 8) * 2^-24.
 * Asynchronous on the handle's stream (injected_r: see "host buffer lifetime" above). */
int flk_step(flk_field *f, uint64_t step, uint64_t seed, const float *injected_r, uint64_t n_injected);

/* `simulate!(state, n_steps, 1)` inner loop (flockers/src/main.rs:47): n_steps x { step ; lazy_update },
 * tick numbers first_step, first_step+1, ...; Philox randomness.  Replays a captured CUDA graph. */
int flk_run(flk_field *f, uint64_t first_step, uint64_t n_steps, uint64_t seed);

/* ---- Field2D read side ------------------------------------------------------------------------- */

/* Field2D::get_neighbors_within_relax_distance(loc, dist)   flockers/src/model/bird.rs:29-31
 * Field2D::get_neighbors_within_distance(loc, dist)         (named by the engine API; not called by bird.rs)
 * Synchronous.  Writes up to `cap` ids in query order (x-outer, y-inner, ascending id per cell);
 * *n = full result size; returns FLK_E_CAPACITY if *n > cap (first cap ids are valid). */
int flk_get_neighbors_within_relax_distance(flk_field *f, float x, float y, float dist,
                                            uint32_t *ids, uint32_t cap, uint32_t *n);
int flk_get_neighbors_within_distance(flk_field *f, float x, float y, float dist,
                                      uint32_t *ids, uint32_t cap, uint32_t *n);

/* The same queries returning what the reference returns: the objects themselves (bird.rs:29-31 iterates `Vec<Bird>` and
 * reads elem.loc and elem.last_d, bird.rs:52-71).  flk_bird is Bird {id, loc, last_d} (bird.rs:13-18) in its 20-byte wire
 * form (flockers_mpi/src/model/bird.rs:31-57), as stored in the READ buffer; same order and error behaviour as above. */
typedef struct flk_bird {
    uint32_t id;
    float x, y;             /* loc */
    float last_dx, last_dy; /* last_d */
} flk_bird;
int flk_get_neighbors_within_relax_distance_objects(flk_field *f, float x, float y, float dist,
                                                    flk_bird *objects, uint32_t cap, uint32_t *n);
int flk_get_neighbors_within_distance_objects(flk_field *f, float x, float y, float dist,
                                              flk_bird *objects, uint32_t cap, uint32_t *n);

/* batched query for nq HOST locations (interleaved x,y): offsets[nq+1], ids[cap] (CSR); _objects: flk_bird objects[cap] */
int flk_query_batch(flk_field *f, uint64_t nq, const float *loc, float dist, int relax,
                    uint64_t *offsets, uint32_t *ids, uint64_t cap);
int flk_query_batch_objects(flk_field *f, uint64_t nq, const float *loc, float dist, int relax,
                            uint64_t *offsets, flk_bird *objects, uint64_t cap);

/* number of objects in the READ buffer */
int flk_num_objects(const flk_field *f, uint64_t *n);

/* Field2D::get / get_location for every object (flockers/src/visualization/vis_state.rs:52, bird_vis.rs:23):
 * copies the READ buffer, in (cell, id) order, to HOST arrays (any of them may be NULL). Synchronous. */
int flk_snapshot(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n);

/* Field2D::get(&object) / Field2D::get_location(object)   flockers/src/visualization/vis_state.rs:52, bird_vis.rs:23
 * Looks one object up by id in the READ buffer (the reference hashes Bird by id, bird.rs:148-163).  Synchronous.
 * *found = 0 when no object has that id (the reference returns None); out = {x, y, last_d.x, last_d.y}. */
int flk_get_object(flk_field *f, uint32_t id, float out[4], int32_t *found);

/* flk_snapshot without the final synchronisation: returns once the copies are enqueued on the handle's stream.
 * The HOST arrays must be pinned (flk_host_alloc) and must not be read before flk_synchronize returns. */
int flk_snapshot_async(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n);

/* Dense host interface.  The reference's host copy of the agents is a list indexed by id: Flocker::init numbers the
 * birds 0..N in list order (flockers/src/model/state.rs:40-51) and Bird is hashed by id alone (bird.rs:148-163).
 * With the id implied by the position only Bird's 16 bytes {loc.x, loc.y, last_d.x, last_d.y} move per object.
 *   flk_set_objects_dense   set_object_location for ids id0 .. id0+n-1, rec4[4*k..4*k+3] = object id0+k
 *   flk_snapshot_dense      rec4[4*k..] = READ copy of object id0+k (all-ones NaN when no such object);
 *                           FLK_E_INVALID if the field holds an id outside [id0, id0+n)
 *   ..._async               as flk_snapshot_async; the range check is reported by the next flk_synchronize
 * Single-GPU handles only (a strip owns an arbitrary subset of the ids). */
int flk_set_objects_dense(flk_field *f, uint64_t n, uint32_t id0, const float *rec4);
int flk_snapshot_dense(flk_field *f, uint32_t id0, uint64_t n, float *rec4);
int flk_snapshot_dense_async(flk_field *f, uint32_t id0, uint64_t n, float *rec4);

/* Frames for a front-end (Field2D::get / get_location for every object, once per rendered frame:
 * flockers/src/visualization/vis_state.rs:52, bird_vis.rs:23) without stopping the simulation: flk_frame_begin copies the
 * READ buffer aside on the device and lets a side stream move that copy to the HOST arrays (pinned; id[n], rec4[4n] =
 * {loc.x, loc.y, last_d.x, last_d.y}, (cell, id) order; either may be NULL) while flk_run / flk_step calls made afterwards
 * proceed; the arrays are valid after flk_frame_wait.  One frame in flight per handle (a second begin queues behind it). */
int flk_frame_begin(flk_field *f, uint32_t *id, float *rec4, uint64_t *n);
int flk_frame_wait(flk_field *f);

/* flk_snapshot plus the payload bytes, same order */
int flk_snapshot_ex(flk_field *f, uint32_t *id, float *loc, float *last_d, void *payload, uint64_t *n);

/* the binning itself: cell_start[dw*dh+1] (flat cell = cx*dh+cy) and ids in sorted order */
int flk_get_binning(flk_field *f, uint32_t *cell_start, uint32_t *sorted_ids);

/* ---- timing / sync ----------------------------------------------------------------------------- */
int flk_synchronize(flk_field *f);
/* device time (CUDA events on the handle's stream) of the last flk_step / flk_lazy_update / flk_run call */
int flk_elapsed_ms(flk_field *f, float *ms);
/* the CUDA stream (cudaStream_t) all work of this handle is enqueued on */
int flk_stream(flk_field *f, void **stream);
/* kernels launched by this handle since creation */
int flk_launch_count(const flk_field *f, uint64_t *n);
/* on = 1: CUDA events around every step-kernel launch (flk_profile_step_kernel); on = 2: also around the three
 * lazy_update phases (flk_profile_rebin); 0 = off.  While on, flk_run launches directly instead of replaying its graph. */
int flk_profile_enable(flk_field *f, int on);
/* flk_run replays a captured CUDA graph per tick (default on); off = plain launches */
int flk_use_graph(flk_field *f, int on);
int flk_profile_step_kernel(flk_field *f, double *total_ms, uint64_t *launches);
/* total device time (ms) of the three lazy_update phases {scan, scatter, rank} and the number of lazy_updates timed */
int flk_profile_rebin(flk_field *f, double ms[3], uint64_t *updates);

/* pinned host memory helpers (fast flk_set_objects / flk_snapshot) */
int flk_host_alloc(void **p, uint64_t bytes);
int flk_host_free(void *p);

/* diagnostics: compares the kernels' shared-reciprocal IEEE division against __fdiv_rn on n operand pairs */
int flk_selftest_div2(int device, uint64_t n, uint64_t seed, uint64_t *mismatches);

const char *flk_last_error(void);
const char *flk_version(void);

/* ---- multi-GPU: one rank per GPU, x-strip decomposition (flockers_mpi) --------------------------- */

/* NCCL unique id (128 bytes) created on rank 0 and distributed by the caller (e.g. torch.distributed) */
int flk_dist_unique_id(uint8_t id128[128]);

/* Kdtree::create_tree(...) (flockers_mpi/src/model/state.rs:33) replaced by a 1-D periodic strip
 * decomposition over cell columns: rank r owns global columns [r*mx/world, (r+1)*mx/world).
 * capacity is per rank.  The global field is width x height as in flk_field_create. */
int flk_dist_create(float width, float height, float discretization, uint64_t capacity, int device,
                    int rank, int world, const uint8_t id128[128], flk_field **out);

/* The same with explicit strip boundaries: col_bounds[world+1] global cell columns, col_bounds[0] = 0,
 * col_bounds[world] = ceil(width/discretization), every strip at least 2 columns wide; identical on every rank.
 * NULL = equal widths.  This is the Kdtree partition of flockers_mpi (state.rs:33) restricted to one axis. */
int flk_dist_create_cols(float width, float height, float discretization, uint64_t capacity, int device,
                         int rank, int world, const uint8_t id128[128], const int32_t *col_bounds, flk_field **out);

/* Tiles: the ranks form a grid of tile_rows rows x (world / tile_rows) columns, rank = row * columns + column, and the
 * cell rows are split like the cell columns (row_bounds[tile_rows+1], NULL = equal heights).  Every tile exchanges with
 * its four edge neighbours only; birds of the diagonal neighbours travel in two hops (left/right first, then the
 * arrivals that sit in a border row are forwarded down/up).  Needs at least 2 x 2 tiles; tile_rows = 1 is the strip
 * decomposition.  This is the 2-D block partition of flockers_mpi's Kdtree (state.rs:33) on a regular grid. */
int flk_dist_create_tiles(float width, float height, float discretization, uint64_t capacity, int device,
                          int rank, int world, const uint8_t id128[128], int tile_rows, const int32_t *col_bounds,
                          const int32_t *row_bounds, flk_field **out);
/* flk_dist_create_tiles with payload_bytes opaque bytes per object (see flk_field_create_ex); tile_rows = 1: strips */
int flk_dist_create_ex(float width, float height, float discretization, uint64_t capacity, int device,
                       int rank, int world, const uint8_t id128[128], int tile_rows, const int32_t *col_bounds,
                       const int32_t *row_bounds, uint32_t payload_bytes, flk_field **out);
/* global cell range [col0,col1) x [row0,row1) owned by this rank (strips: all rows) */
int flk_dist_tile(const flk_field *f, int32_t *col0, int32_t *col1, int32_t *row0, int32_t *row1);

/* Boundaries that give every strip about the same number of birds: x[i*stride] (HOST) is the x coordinate of sampled
 * bird i (stride in floats: 1 for a plain array, 2 for interleaved loc, 4 for rec4).  Host-side, no device needed. */
int flk_balance_columns(const float *x, uint64_t n, uint64_t stride, float width, float discretization, int world,
                        int32_t *col_bounds);

/* get_block_by_location (flockers_mpi/src/model/state.rs:75, bird.rs:190): owner rank of a location */
int flk_dist_owner(const flk_field *f, float x, float y, int32_t *rank);
/* global cell-column range [col0,col1) owned by this rank */
int flk_dist_strip(const flk_field *f, int32_t *col0, int32_t *col1);

/* Publish the uploaded birds and build the first halo (before_step of tick 0, state.rs:113-133).
 * flk_set_objects on a dist handle keeps only the birds this rank owns (the init scatter, state.rs:51-103,
 * may simply hand every rank the full list).  Collective: every rank calls it. */
int flk_dist_commit(flk_field *f);

/* Distributed tick: every owned Bird::step + ONE message per neighbour carrying the migrants (after_step,
 * state.rs:139-175) and next tick's ghosts (before_step, state.rs:113-133) + lazy_update (state.rs:105-107).
 * Collective.  flk_run on a dist handle loops this with Philox randomness. */
int flk_dist_step(flk_field *f, uint64_t step, uint64_t seed, const float *injected_r, uint64_t n_injected);

/* number of birds this rank owns (ghost copies excluded) — synchronising; reports capacity overflows */
int flk_dist_num_owned(flk_field *f, uint64_t *n);
/* flk_snapshot restricted to the birds this rank owns.  Strips: (cell, id) order.  Tiles: the owned birds are compacted
 * out of the READ buffer in no particular order (sort by id if an order is needed). */
int flk_dist_snapshot_owned(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n);

/* flk_dist_snapshot_owned plus the payload bytes of the owned objects, same order */
int flk_dist_snapshot_owned_ex(flk_field *f, uint32_t *id, float *loc, float *last_d, void *payload, uint64_t *n);

/* as above, but the copies of the records are only enqueued: *n is exact on return (the call waits for the tick that
 * produced the READ buffer), the HOST arrays (pinned) are valid after flk_synchronize */
int flk_dist_snapshot_owned_async(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n);

/* ---- multi-GPU, one host thread driving n strips (single process; peer copies instead of NCCL) -------- */
/* out[n] receives one handle per strip; devices[r] is the CUDA ordinal of strip r (ordinals may repeat). */
int flk_mgpu_create(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                    const int *devices, flk_field **out);
int flk_mgpu_create_cols(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                         const int *devices, const int32_t *col_bounds, flk_field **out);
int flk_mgpu_create_tiles(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                          const int *devices, int tile_rows, const int32_t *col_bounds, const int32_t *row_bounds,
                          flk_field **out);
int flk_mgpu_create_ex(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                       const int *devices, int tile_rows, const int32_t *col_bounds, const int32_t *row_bounds,
                       uint32_t payload_bytes, flk_field **out);
int flk_mgpu_commit(flk_field **handles, int n);
int flk_mgpu_step(flk_field **handles, int n, uint64_t step, uint64_t seed, const float *injected_r,
                  uint64_t n_injected);
int flk_mgpu_run(flk_field **handles, int n, uint64_t first_step, uint64_t n_steps, uint64_t seed);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FLK_H */
