//! `GpuDistField2D` — the flockers_mpi analogue (feature `distributed`): one process per GPU, each owning a block of the
//! torus; halo and migration travel inside `flk_dist_step` (NCCL over NVLink), so the `State` hooks of the MPI example
//! shrink to one call each.
//!
//! Reference call sites (krABMaga/examples):
//!   Kdtree::create_tree(0, 0., 0., W, H, DISCRETIZATION, 25.)   flockers_mpi/src/model/state.rs:33   -> `new` / `new_tiles`
//!   field1.get_block_by_location(x, y)                          flockers_mpi/src/model/bird.rs:190   -> `get_block_by_location`
//!   before_step: message_exchange(prec_neighbors) + insert_read flockers_mpi/src/model/state.rs:113-133 }
//!   Bird::step for every local bird                             flockers_mpi/src/model/bird.rs:66-195 }  -> `step_all`
//!   after_step: message_exchange(agents_to_send) + insert       flockers_mpi/src/model/state.rs:139-175 }
//!   update: lazy_update                                         flockers_mpi/src/model/state.rs:105-107 }
//! The NCCL unique id is created on rank 0 (`unique_id`) and broadcast by the caller with the MPI world krABMaga
//! already owns (`UNIVERSE.world()`, flockers_mpi/src/model/bird.rs:70).
use crate::{check, ffi, FlockObject};
use krabmaga::engine::location::Real2D;
use std::ptr;

pub struct GpuDistField2D<O: FlockObject> {
    h: *mut ffi::flk_field,
    pub width: f32,
    pub height: f32,
    pub rank: i32,
    pub world: i32,
    _o: std::marker::PhantomData<O>,
}

/// 128 opaque bytes; rank 0 creates them, every rank passes the same bytes to `new`.
pub fn unique_id() -> [u8; 128] {
    let mut id = [0u8; 128];
    check(unsafe { ffi::flk_dist_unique_id(id.as_mut_ptr()) });
    id
}

impl<O: FlockObject> GpuDistField2D<O> {
    /// x-strips of equal width (`col_bounds = None`) or with the given `world + 1` cell-column boundaries
    /// (`balance_columns` computes population-balanced ones, which is what the kd-tree of the reference is for).
    pub fn new(width: f32, height: f32, discretization: f32, capacity_per_rank: u64, device: i32, rank: i32, world: i32,
               nccl_id: &[u8; 128], col_bounds: Option<&[i32]>) -> Self {
        Self::new_tiles(width, height, discretization, capacity_per_rank, device, rank, world, nccl_id, 1, col_bounds, None)
    }

    /// 2-D blocks: `tile_rows` rows of `world / tile_rows` tiles, rank = row * columns + column.
    #[allow(clippy::too_many_arguments)]
    pub fn new_tiles(width: f32, height: f32, discretization: f32, capacity_per_rank: u64, device: i32, rank: i32,
                     world: i32, nccl_id: &[u8; 128], tile_rows: i32, col_bounds: Option<&[i32]>,
                     row_bounds: Option<&[i32]>) -> Self {
        let mut h = ptr::null_mut();
        let cb = col_bounds.map_or(ptr::null(), |b| b.as_ptr());
        let rb = row_bounds.map_or(ptr::null(), |b| b.as_ptr());
        check(unsafe {
            ffi::flk_dist_create_tiles(width, height, discretization, capacity_per_rank, device, rank, world,
                                       nccl_id.as_ptr(), tile_rows, cb, rb, &mut h)
        });
        GpuDistField2D { h, width, height, rank, world, _o: std::marker::PhantomData }
    }

    /// `field1.insert(bird, loc)` of the init scatter (state.rs:82,98): every rank may offer the whole list, a handle keeps
    /// the birds whose location it owns.
    pub fn insert(&self, object: O, loc: Real2D) {
        let d = object.last_d();
        check(unsafe { ffi::flk_set_object_location(self.h, object.id(), loc.x, loc.y, d.x, d.y) });
    }

    /// publish the inserted birds and build the first halo (before_step of tick 0, state.rs:113-133).  Collective.
    pub fn commit(&self) { check(unsafe { ffi::flk_dist_commit(self.h) }); }

    /// `get_block_by_location` (bird.rs:190): rank owning a location
    pub fn get_block_by_location(&self, x: f32, y: f32) -> u32 {
        let mut r = 0i32;
        check(unsafe { ffi::flk_dist_owner(self.h, x, y, &mut r) });
        r as u32
    }

    /// one distributed tick: before_step + every local `Bird::step` + after_step + update.  Collective.
    pub fn step_all(&self, step: u64, seed: u64) {
        check(unsafe { ffi::flk_dist_step(self.h, step, seed, ptr::null(), 0) });
    }

    /// birds this rank owns (ghost copies excluded)
    pub fn num_owned(&self) -> u64 {
        let mut n = 0u64;
        check(unsafe { ffi::flk_dist_num_owned(self.h, &mut n) });
        n
    }

    /// the birds this rank owns, for output or visualisation (`Field2D::get` per object in the reference)
    pub fn get_owned(&self) -> Vec<O> {
        let cap = self.num_owned() as usize;
        let (mut id, mut loc, mut ld) = (vec![0u32; cap], vec![0f32; 2 * cap], vec![0f32; 2 * cap]);
        let mut n = 0u64;
        check(unsafe { ffi::flk_dist_snapshot_owned(self.h, id.as_mut_ptr(), loc.as_mut_ptr(), ld.as_mut_ptr(), &mut n) });
        (0..n as usize)
            .map(|i| O::from_parts(id[i], Real2D { x: loc[2 * i], y: loc[2 * i + 1] }, Real2D { x: ld[2 * i], y: ld[2 * i + 1] }))
            .collect()
    }
}

impl<O: FlockObject> Drop for GpuDistField2D<O> {
    fn drop(&mut self) { unsafe { ffi::flk_field_destroy(self.h) }; }
}

/// `world + 1` cell-column boundaries that give every strip about the same number of the sampled x coordinates
pub fn balance_columns(xs: &[f32], width: f32, discretization: f32, world: i32) -> Vec<i32> {
    let mut b = vec![0i32; world as usize + 1];
    check(unsafe { ffi::flk_balance_columns(xs.as_ptr(), xs.len() as u64, 1, width, discretization, world, b.as_mut_ptr()) });
    b
}
