//! flockers/src/{main.rs, model/state.rs, model/bird.rs} with `field1` on the GPU.  Diff against the reference:
//!   - `field1: Field2D<Bird>`            ->  `field1: GpuField2D<Bird>`
//!   - N x `schedule_repeating(bird)`     ->  1 x `schedule_repeating(FlockStepper)`
//! `main`, `Flocker::new`, `init`, `update`, `simulate_old!` are unchanged.
use krabmaga::engine::{location::{Location2D, Real2D}, schedule::Schedule, state::State};
use krabmaga::rand::{self, Rng};
use krabmaga::*;
use krabmaga_flk::{FlockObject, FlockStepper, GpuField2D, HasGpuField};
use std::any::Any;
use std::fmt;
use std::hash::{Hash, Hasher};

pub static DISCRETIZATION: f32 = 10.0 / 1.5;
pub static TOROIDAL: bool = true;

#[derive(Clone, Copy)]
pub struct Bird { pub id: u32, pub loc: Real2D, pub last_d: Real2D }

// bird.rs:148-176, unchanged: identity, hash and equality by id; Location2D
impl Hash for Bird { fn hash<H: Hasher>(&self, state: &mut H) { self.id.hash(state); } }
impl PartialEq for Bird { fn eq(&self, other: &Bird) -> bool { self.id == other.id } }
impl Eq for Bird {}
impl Location2D<Real2D> for Bird {
    fn get_location(self) -> Real2D { self.loc }
    fn set_location(&mut self, loc: Real2D) { self.loc = loc; }
}
impl fmt::Display for Bird { fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result { write!(f, "{}", self.id) } }

impl FlockObject for Bird {
    fn id(&self) -> u32 { self.id }
    fn last_d(&self) -> Real2D { self.last_d }
    fn from_parts(id: u32, loc: Real2D, last_d: Real2D) -> Self { Bird { id, loc, last_d } }
}

pub struct Flocker { pub step: u64, pub field1: GpuField2D<Bird>, pub initial_flockers: u32, pub dim: (f32, f32) }

impl Flocker {
    pub fn new(dim: (f32, f32), initial_flockers: u32) -> Self {
        Flocker { step: 0, field1: GpuField2D::new(dim.0, dim.1, DISCRETIZATION, TOROIDAL, initial_flockers as u64), initial_flockers, dim }
    }
}

impl HasGpuField<Bird> for Flocker {
    fn gpu_field(&self) -> &GpuField2D<Bird> { &self.field1 }
    fn tick(&self) -> u64 { self.step }
}

impl State for Flocker {
    fn reset(&mut self) { self.step = 0; self.field1.clear(); }
    fn init(&mut self, schedule: &mut Schedule) {
        let mut rng = rand::rng();
        for bird_id in 0..self.initial_flockers {
            let (r1, r2): (f32, f32) = (rng.random(), rng.random());
            let loc = Real2D { x: self.dim.0 * r1, y: self.dim.1 * r2 };
            let bird = Bird { id: bird_id, loc, last_d: Real2D { x: 0., y: 0. } };
            self.field1.set_object_location(bird, loc);
        }
        schedule.schedule_repeating(Box::new(FlockStepper::<Flocker, Bird>::new(0xB12D)), 0., 0);
    }
    fn update(&mut self, step: u64) { self.step = step; self.field1.lazy_update(); }
    fn as_any(&self) -> &dyn Any { self }
    fn as_any_mut(&mut self) -> &mut dyn Any { self }
    fn as_state_mut(&mut self) -> &mut dyn State { self }
    fn as_state(&self) -> &dyn State { self }
}

fn main() {
    let state = Flocker::new((100., 100.), 1000);
    let _ = simulate_old!(state, 200, 1, krabmaga::Info::Normal);
}
