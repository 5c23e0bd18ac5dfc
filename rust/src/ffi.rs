//! Raw bindings of include/flk.h (one `extern "C"` item per declaration; keep in the header's order).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)]
pub struct flk_field {
    _private: [u8; 0],
}

/// Bird {id, loc, last_d} (flockers/src/model/bird.rs:13-18) in its 20-byte wire form (flockers_mpi/src/model/bird.rs:31-57)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct flk_bird {
    pub id: u32,
    pub x: c_float,
    pub y: c_float,
    pub last_dx: c_float,
    pub last_dy: c_float,
}

pub const FLK_OK: c_int = 0;
pub const FLK_E_CAPACITY: c_int = -3;

extern "C" {
    pub fn flk_field_create(width: c_float, height: c_float, discretization: c_float, toroidal: c_int, capacity: u64,
                            device: c_int, out: *mut *mut flk_field) -> c_int;
    pub fn flk_field_create_ex(width: c_float, height: c_float, discretization: c_float, toroidal: c_int, capacity: u64,
                               device: c_int, payload_bytes: u32, out: *mut *mut flk_field) -> c_int;
    pub fn flk_field_destroy(f: *mut flk_field) -> c_int;
    pub fn flk_field_clear(f: *mut flk_field) -> c_int;
    pub fn flk_set_params(f: *mut flk_field, cohesion: c_float, avoidance: c_float, randomness: c_float,
                          consistency: c_float, momentum: c_float, jump: c_float, radius: c_float, relax: c_int) -> c_int;
    pub fn flk_set_engine_semantics(f: *mut flk_field, window_plus1: c_int) -> c_int;
    pub fn flk_set_math_mode(f: *mut flk_field, mode: c_int) -> c_int;
    pub fn flk_set_layout(f: *mut flk_field, layout: c_int) -> c_int;
    pub fn flk_get_layout(f: *mut flk_field, layout: *mut i32, slots_per_cell: *mut i32, totals: *mut u32) -> c_int;
    pub fn flk_field_dims(f: *const flk_field, dw: *mut i32, dh: *mut i32, mx: *mut i32, my: *mut i32) -> c_int;
    pub fn flk_set_object_location(f: *mut flk_field, id: u32, x: c_float, y: c_float, ldx: c_float, ldy: c_float) -> c_int;
    pub fn flk_set_objects(f: *mut flk_field, n: u64, id: *const u32, loc: *const c_float, last_d: *const c_float) -> c_int;
    pub fn flk_set_objects_ex(f: *mut flk_field, n: u64, id: *const u32, loc: *const c_float, last_d: *const c_float,
                              payload: *const c_void) -> c_int;
    pub fn flk_snapshot_ex(f: *mut flk_field, id: *mut u32, loc: *mut c_float, last_d: *mut c_float, payload: *mut c_void,
                           n: *mut u64) -> c_int;
    pub fn flk_lazy_update(f: *mut flk_field) -> c_int;
    pub fn flk_step(f: *mut flk_field, step: u64, seed: u64, injected_r: *const c_float, n_injected: u64) -> c_int;
    pub fn flk_run(f: *mut flk_field, first_step: u64, n_steps: u64, seed: u64) -> c_int;
    pub fn flk_get_neighbors_within_relax_distance(f: *mut flk_field, x: c_float, y: c_float, dist: c_float,
                                                   ids: *mut u32, cap: u32, n: *mut u32) -> c_int;
    pub fn flk_get_neighbors_within_distance(f: *mut flk_field, x: c_float, y: c_float, dist: c_float,
                                             ids: *mut u32, cap: u32, n: *mut u32) -> c_int;
    pub fn flk_get_neighbors_within_relax_distance_objects(f: *mut flk_field, x: c_float, y: c_float, dist: c_float,
                                                           objects: *mut flk_bird, cap: u32, n: *mut u32) -> c_int;
    pub fn flk_get_neighbors_within_distance_objects(f: *mut flk_field, x: c_float, y: c_float, dist: c_float,
                                                     objects: *mut flk_bird, cap: u32, n: *mut u32) -> c_int;
    pub fn flk_query_batch(f: *mut flk_field, nq: u64, loc: *const c_float, dist: c_float, relax: c_int,
                           offsets: *mut u64, ids: *mut u32, cap: u64) -> c_int;
    pub fn flk_query_batch_objects(f: *mut flk_field, nq: u64, loc: *const c_float, dist: c_float, relax: c_int,
                                   offsets: *mut u64, objects: *mut flk_bird, cap: u64) -> c_int;
    pub fn flk_num_objects(f: *const flk_field, n: *mut u64) -> c_int;
    pub fn flk_snapshot(f: *mut flk_field, id: *mut u32, loc: *mut c_float, last_d: *mut c_float, n: *mut u64) -> c_int;
    pub fn flk_snapshot_async(f: *mut flk_field, id: *mut u32, loc: *mut c_float, last_d: *mut c_float, n: *mut u64) -> c_int;
    pub fn flk_set_objects_dense(f: *mut flk_field, n: u64, id0: u32, rec4: *const c_float) -> c_int;
    pub fn flk_snapshot_dense(f: *mut flk_field, id0: u32, n: u64, rec4: *mut c_float) -> c_int;
    pub fn flk_snapshot_dense_async(f: *mut flk_field, id0: u32, n: u64, rec4: *mut c_float) -> c_int;
    pub fn flk_frame_begin(f: *mut flk_field, id: *mut u32, rec4: *mut c_float, n: *mut u64) -> c_int;
    pub fn flk_frame_wait(f: *mut flk_field) -> c_int;
    pub fn flk_get_object(f: *mut flk_field, id: u32, out: *mut c_float, found: *mut i32) -> c_int;
    pub fn flk_get_binning(f: *mut flk_field, cell_start: *mut u32, sorted_ids: *mut u32) -> c_int;
    pub fn flk_synchronize(f: *mut flk_field) -> c_int;
    pub fn flk_elapsed_ms(f: *mut flk_field, ms: *mut c_float) -> c_int;
    pub fn flk_stream(f: *mut flk_field, stream: *mut *mut c_void) -> c_int;
    pub fn flk_launch_count(f: *const flk_field, n: *mut u64) -> c_int;
    pub fn flk_profile_enable(f: *mut flk_field, on: c_int) -> c_int;
    pub fn flk_use_graph(f: *mut flk_field, on: c_int) -> c_int;
    pub fn flk_profile_step_kernel(f: *mut flk_field, total_ms: *mut c_double, launches: *mut u64) -> c_int;
    pub fn flk_profile_rebin(f: *mut flk_field, ms: *mut c_double, updates: *mut u64) -> c_int;
    pub fn flk_host_alloc(p: *mut *mut c_void, bytes: u64) -> c_int;
    pub fn flk_host_free(p: *mut c_void) -> c_int;
    pub fn flk_selftest_div2(device: c_int, n: u64, seed: u64, mismatches: *mut u64) -> c_int;
    pub fn flk_last_error() -> *const c_char;
    pub fn flk_version() -> *const c_char;
    pub fn flk_dist_unique_id(id128: *mut u8) -> c_int;
    pub fn flk_dist_create(width: c_float, height: c_float, discretization: c_float, capacity: u64, device: c_int,
                           rank: c_int, world: c_int, id128: *const u8, out: *mut *mut flk_field) -> c_int;
    pub fn flk_dist_create_cols(width: c_float, height: c_float, discretization: c_float, capacity: u64, device: c_int,
                                rank: c_int, world: c_int, id128: *const u8, col_bounds: *const i32, out: *mut *mut flk_field) -> c_int;
    pub fn flk_dist_create_tiles(width: c_float, height: c_float, discretization: c_float, capacity: u64, device: c_int,
                                 rank: c_int, world: c_int, id128: *const u8, tile_rows: c_int, col_bounds: *const i32,
                                 row_bounds: *const i32, out: *mut *mut flk_field) -> c_int;
    pub fn flk_dist_create_ex(width: c_float, height: c_float, discretization: c_float, capacity: u64, device: c_int,
                              rank: c_int, world: c_int, id128: *const u8, tile_rows: c_int, col_bounds: *const i32,
                              row_bounds: *const i32, payload_bytes: u32, out: *mut *mut flk_field) -> c_int;
    pub fn flk_dist_tile(f: *const flk_field, col0: *mut i32, col1: *mut i32, row0: *mut i32, row1: *mut i32) -> c_int;
    pub fn flk_mgpu_create_tiles(width: c_float, height: c_float, discretization: c_float, capacity_per_rank: u64, n: c_int,
                                 devices: *const c_int, tile_rows: c_int, col_bounds: *const i32, row_bounds: *const i32,
                                 out: *mut *mut flk_field) -> c_int;
    pub fn flk_balance_columns(x: *const c_float, n: u64, stride: u64, width: c_float, discretization: c_float, world: c_int,
                               col_bounds: *mut i32) -> c_int;
    pub fn flk_mgpu_create_cols(width: c_float, height: c_float, discretization: c_float, capacity_per_rank: u64, n: c_int,
                                devices: *const c_int, col_bounds: *const i32, out: *mut *mut flk_field) -> c_int;
    pub fn flk_dist_owner(f: *const flk_field, x: c_float, y: c_float, rank: *mut i32) -> c_int;
    pub fn flk_dist_strip(f: *const flk_field, col0: *mut i32, col1: *mut i32) -> c_int;
    pub fn flk_dist_commit(f: *mut flk_field) -> c_int;
    pub fn flk_dist_step(f: *mut flk_field, step: u64, seed: u64, injected_r: *const c_float, n_injected: u64) -> c_int;
    pub fn flk_dist_num_owned(f: *mut flk_field, n: *mut u64) -> c_int;
    pub fn flk_dist_snapshot_owned(f: *mut flk_field, id: *mut u32, loc: *mut c_float, last_d: *mut c_float, n: *mut u64) -> c_int;
    pub fn flk_dist_snapshot_owned_async(f: *mut flk_field, id: *mut u32, loc: *mut c_float, last_d: *mut c_float, n: *mut u64) -> c_int;
    pub fn flk_dist_snapshot_owned_ex(f: *mut flk_field, id: *mut u32, loc: *mut c_float, last_d: *mut c_float,
                                      payload: *mut c_void, n: *mut u64) -> c_int;
    pub fn flk_mgpu_create(width: c_float, height: c_float, discretization: c_float, capacity_per_rank: u64, n: c_int,
                           devices: *const c_int, out: *mut *mut flk_field) -> c_int;
    pub fn flk_mgpu_create_ex(width: c_float, height: c_float, discretization: c_float, capacity_per_rank: u64, n: c_int,
                              devices: *const c_int, tile_rows: c_int, col_bounds: *const i32, row_bounds: *const i32,
                              payload_bytes: u32, out: *mut *mut flk_field) -> c_int;
    pub fn flk_mgpu_commit(handles: *mut *mut flk_field, n: c_int) -> c_int;
    pub fn flk_mgpu_step(handles: *mut *mut flk_field, n: c_int, step: u64, seed: u64, injected_r: *const c_float,
                         n_injected: u64) -> c_int;
    pub fn flk_mgpu_run(handles: *mut *mut flk_field, n: c_int, first_step: u64, n_steps: u64, seed: u64) -> c_int;
}
