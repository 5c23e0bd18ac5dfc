//! `GpuField2D` — a drop-in for `krabmaga::engine::fields::field_2d::Field2D<Bird>` on the Flockers hot path,
//! plus `FlockStepper`, the single scheduled agent that replaces the per-bird `Agent::step` dispatch.
//!
//! Method names and argument meaning follow the reference call sites (krABMaga/examples):
//!   new / set_object_location / lazy_update      flockers/src/model/state.rs:24,49,55
//!   get_neighbors_within_relax_distance           flockers/src/model/bird.rs:29-31
//!   Bird::step for every bird                     flockers/src/model/bird.rs:27-145  -> `step_all`
//! The reference methods return no `Result`; a non-zero libflk code panics with `flk_last_error()`.
pub mod ffi;
#[cfg(feature = "distributed")]
pub mod dist;

use krabmaga::engine::agent::Agent;
use krabmaga::engine::location::{Location2D, Real2D};
use krabmaga::engine::state::State;
use std::cell::Cell;
use std::ffi::CStr;
use std::fmt::Display;
use std::hash::Hash;
use std::ptr;

/// What the field stores per object.  The bounds are the reference's own for `Field2D<O>`
/// (`O: Location2D<Real2D> + Clone + Copy + Hash + Eq + Display`, as `impl`ed for Bird at bird.rs:148-176); the three
/// methods are what the device needs beyond them: the u32 the object is hashed and compared by (bird.rs:148-163), its
/// heading, and a way to rebuild the object from the record the device keeps (Bird is exactly {id, loc, last_d}).
pub trait FlockObject: Location2D<Real2D> + Copy + Hash + Eq + Display {
    fn id(&self) -> u32;
    fn last_d(&self) -> Real2D;
    fn from_parts(id: u32, loc: Real2D, last_d: Real2D) -> Self;
}

pub(crate) fn check(rc: i32) {
    if rc != ffi::FLK_OK {
        let msg = unsafe { CStr::from_ptr(ffi::flk_last_error()) }.to_string_lossy().into_owned();
        panic!("libflk error {rc}: {msg}");
    }
}

pub struct GpuField2D<O: FlockObject> {
    h: *mut ffi::flk_field,
    pub width: f32,
    pub height: f32,
    _o: std::marker::PhantomData<O>,
}

fn object_of<O: FlockObject>(b: &ffi::flk_bird) -> O {
    O::from_parts(b.id, Real2D { x: b.x, y: b.y }, Real2D { x: b.last_dx, y: b.last_dy })
}

impl<O: FlockObject> GpuField2D<O> {
    /// `Field2D::new(w, h, discretization, toroidal)` + the capacity the device store is sized for.
    pub fn new(width: f32, height: f32, discretization: f32, toroidal: bool, capacity: u64) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::flk_field_create(width, height, discretization, toroidal as i32, capacity, 0, &mut h) });
        GpuField2D { h, width, height, _o: std::marker::PhantomData }
    }
    /// `&self`, like the reference (called through `downcast_ref`, bird.rs:28,142): interior mutability lives in the library.
    pub fn set_object_location(&self, object: O, loc: Real2D) {
        let d = object.last_d();
        check(unsafe { ffi::flk_set_object_location(self.h, object.id(), loc.x, loc.y, d.x, d.y) });
    }
    /// `Field2D::get_neighbors_within_relax_distance(loc, dist) -> Vec<O>` (bird.rs:29-31): the objects themselves, in
    /// query order, as the READ buffer holds them (Bird::step reads `elem.loc` and `elem.last_d`, bird.rs:52-71).
    pub fn get_neighbors_within_relax_distance(&self, loc: Real2D, dist: f32) -> Vec<O> { self.query(loc, dist, true) }
    /// `Field2D::get_neighbors_within_distance(loc, dist) -> Vec<O>`
    pub fn get_neighbors_within_distance(&self, loc: Real2D, dist: f32) -> Vec<O> { self.query(loc, dist, false) }
    fn query(&self, loc: Real2D, dist: f32, relax: bool) -> Vec<O> {
        let mut buf = vec![ffi::flk_bird::default(); 256];
        loop {
            let mut n = 0u32;
            let rc = unsafe {
                if relax { ffi::flk_get_neighbors_within_relax_distance_objects(self.h, loc.x, loc.y, dist, buf.as_mut_ptr(), buf.len() as u32, &mut n) }
                else { ffi::flk_get_neighbors_within_distance_objects(self.h, loc.x, loc.y, dist, buf.as_mut_ptr(), buf.len() as u32, &mut n) }
            };
            if rc == ffi::FLK_E_CAPACITY && n as usize > buf.len() { buf.resize(n as usize, ffi::flk_bird::default()); continue; }
            check(rc);
            return buf[..n as usize].iter().map(object_of::<O>).collect();
        }
    }
    /// `Field2D::get(&object) -> Option<&O>` (visualization/vis_state.rs:52); by value here: the stored copy lives on the device
    pub fn get(&self, object: &O) -> Option<O> {
        let (mut out, mut found) = ([0f32; 4], 0i32);
        check(unsafe { ffi::flk_get_object(self.h, object.id(), out.as_mut_ptr(), &mut found) });
        if found == 0 { return None; }
        Some(O::from_parts(object.id(), Real2D { x: out[0], y: out[1] }, Real2D { x: out[2], y: out[3] }))
    }
    /// `Field2D::get_location(object) -> Option<Real2D>` (visualization/bird_vis.rs:23)
    pub fn get_location(&self, object: O) -> Option<Real2D> {
        self.get(&object).map(|o| o.get_location())
    }
    /// `Field2D::num_objects()`
    pub fn num_objects(&self) -> usize {
        let mut n = 0u64;
        check(unsafe { ffi::flk_num_objects(self.h, &mut n) });
        n as usize
    }
    /// `Field::lazy_update` (state.rs:54-56)
    pub fn lazy_update(&mut self) { check(unsafe { ffi::flk_lazy_update(self.h) }); }
    /// every bird's `Bird::step` for tick `step` (bird.rs:27-145)
    pub fn step_all(&self, step: u64, seed: u64) { check(unsafe { ffi::flk_step(self.h, step, seed, ptr::null(), 0) }); }
    /// `Field2D::get` for every object (visualization/vis_state.rs:52)
    pub fn get_all(&self) -> Vec<O> {
        let mut n = 0u64;
        check(unsafe { ffi::flk_num_objects(self.h, &mut n) });
        let (mut id, mut loc, mut ld) = (vec![0u32; n as usize], vec![0f32; 2 * n as usize], vec![0f32; 2 * n as usize]);
        check(unsafe { ffi::flk_snapshot(self.h, id.as_mut_ptr(), loc.as_mut_ptr(), ld.as_mut_ptr(), &mut n) });
        (0..n as usize).map(|i| O::from_parts(id[i], Real2D { x: loc[2 * i], y: loc[2 * i + 1] }, Real2D { x: ld[2 * i], y: ld[2 * i + 1] })).collect()
    }
    /// The agents as the Schedule keeps them: a list indexed by id (state.rs:40-51 numbers birds 0..N in list order).
    /// `rec4[4*id..]` = `{loc.x, loc.y, last_d.x, last_d.y}`; 16 bytes per bird cross the bus instead of 20.
    pub fn set_all_dense(&self, rec4: &[f32]) {
        check(unsafe { ffi::flk_set_objects_dense(self.h, (rec4.len() / 4) as u64, 0, rec4.as_ptr()) });
    }
    pub fn get_all_dense(&self, n: usize) -> Vec<f32> {
        let mut rec4 = vec![0f32; 4 * n];
        check(unsafe { ffi::flk_snapshot_dense(self.h, 0, n as u64, rec4.as_mut_ptr()) });
        rec4
    }
    pub fn clear(&mut self) { check(unsafe { ffi::flk_field_clear(self.h) }); }
}

impl<O: FlockObject> Drop for GpuField2D<O> {
    fn drop(&mut self) { unsafe { ffi::flk_field_destroy(self.h) }; }
}

/// States that own a `GpuField2D` expose it (and the current tick) to the stepper.
pub trait HasGpuField<O: FlockObject> {
    fn gpu_field(&self) -> &GpuField2D<O>;
    fn tick(&self) -> u64;
}

/// The one agent scheduled instead of N birds (idiom of schelling/src/model/updater.rs, forestfire/src/model/spread.rs).
pub struct FlockStepper<S, O> {
    pub seed: u64,
    calls: Cell<u64>,
    _s: std::marker::PhantomData<fn() -> (S, O)>,
}

impl<S, O> FlockStepper<S, O> {
    pub fn new(seed: u64) -> Self { FlockStepper { seed, calls: Cell::new(0), _s: std::marker::PhantomData } }
}

// krabmaga boxes agents as `Box<dyn Agent>` and clones them (Agent: DynClone); written by hand so that the state and
// object types need not be Clone themselves (a derive would ask for `S: Clone, O: Clone`)
impl<S, O> Clone for FlockStepper<S, O> {
    fn clone(&self) -> Self { FlockStepper { seed: self.seed, calls: self.calls.clone(), _s: std::marker::PhantomData } }
}

impl<S: State + HasGpuField<O> + 'static, O: FlockObject + 'static> Agent for FlockStepper<S, O> {
    fn step(&mut self, state: &mut dyn State) {
        let s = state.as_any().downcast_ref::<S>().unwrap();
        s.gpu_field().step_all(s.tick(), self.seed);
        self.calls.set(self.calls.get() + 1);
    }
}
