//! Parity-pinning kit, reference side (INTEGRATION.md "Pinning parity to the real engine").
//!
//! Runs the UNMODIFIED reference model — `flockers/src/model/{bird,state}.rs` of krABMaga/examples, included by path, on
//! the real `krabmaga` engine — over the committed input `tests/golden/pin_input.txt` and writes what the engine answers:
//!   * both neighbour queries on the published initial state (ids in the order returned): E1-E7 of SURVEY.md §8c
//!   * `toroidal_transform` / `toroidal_distance` on probe values: E8
//!   * every agent's state after each of T ticks of `Schedule::step`: E9, bag order, and Bird::step itself
//! `tests/test_reference_golden.py` compares that file with the CPU restatement (oracle/) and with the CUDA path.
//!
//! The reference draws its randomness from the unseeded thread RNG (bird.rs:103-107), so a run cannot be replayed; this
//! crate root therefore defines the model constants the way `flockers/src/main.rs:26-33` does, except RANDOMNESS = 0.0:
//! the random term is then +-0 whatever was drawn, and everything else of Bird::step is exercised unchanged.
//!
//! Not compiled in the repository's image (no cargo there).  One command on a machine with cargo and CUDA-less is fine:
//!   ln -s $KRABMAGA_EXAMPLES/flockers/src/model rust/examples/ref_model
//!   (cd rust && cargo run --release --example dump_golden -- ../tests/golden/pin_input.txt ../tests/golden/reference_pin.txt)
//!   python -m pytest tests/test_reference_golden.py -q
#[path = "ref_model/mod.rs"]
mod model;

use krabmaga::engine::fields::field::Field;
use krabmaga::engine::fields::field_2d::{toroidal_distance, toroidal_transform, Field2D};
use krabmaga::engine::location::Real2D;
use krabmaga::engine::schedule::Schedule;
use krabmaga::engine::state::State;
use model::bird::Bird;
use model::state::Flocker;
use std::fmt::Write as _;
use std::{env, fs};

// flockers/src/main.rs:26-33 (bird.rs and state.rs import these from the crate root)
pub static COHESION: f32 = 0.8;
pub static AVOIDANCE: f32 = 1.0;
pub static RANDOMNESS: f32 = 0.0; // 1.1 in the reference: see the header
pub static CONSISTENCY: f32 = 0.7;
pub static MOMENTUM: f32 = 1.0;
pub static JUMP: f32 = 0.7;
pub static DISCRETIZATION: f32 = 10.0 / 1.5;
pub static TOROIDAL: bool = true;

fn f(tok: &str) -> f32 { f32::from_bits(u32::from_str_radix(tok, 16).expect("hex float")) }

fn main() {
    let args: Vec<String> = env::args().collect();
    let (src, dst) = (&args[1], &args[2]);
    let text = fs::read_to_string(src).expect("pin input");
    let mut lines = text.lines().filter(|l| !l.trim().is_empty() && !l.starts_with('#'));
    let (mut w, mut h, mut d) = (0f32, 0f32, 0f32);
    let (mut birds, mut probes, mut tprobes, mut ticks) = (Vec::new(), Vec::new(), Vec::new(), 0usize);
    while let Some(l) = lines.next() {
        let t: Vec<&str> = l.split_whitespace().collect();
        match t[0] {
            "field" => { w = f(t[1]); h = f(t[2]); d = f(t[3]); }
            "birds" => for _ in 0..t[1].parse::<usize>().unwrap() {
                let b: Vec<&str> = lines.next().unwrap().split_whitespace().collect();
                birds.push(Bird::new(b[0].parse().unwrap(), Real2D { x: f(b[1]), y: f(b[2]) }, Real2D { x: f(b[3]), y: f(b[4]) }));
            },
            "ticks" => ticks = t[1].parse().unwrap(),
            "probes" => for _ in 0..t[1].parse::<usize>().unwrap() {
                let p: Vec<&str> = lines.next().unwrap().split_whitespace().collect();
                probes.push((f(p[0]), f(p[1]), f(p[2]), p[3] == "1"));
            },
            "tprobes" => for _ in 0..t[1].parse::<usize>().unwrap() {
                let p: Vec<&str> = lines.next().unwrap().split_whitespace().collect();
                tprobes.push((p[0] == "1", f(p[1]), f(p[2]), f(p[3])));
            },
            _ => {}
        }
    }
    assert!((d - DISCRETIZATION).abs() == 0.0, "the input was made for the reference discretization");
    let mut out = String::new();
    writeln!(out, "# flk pin output v1 krabmaga (real engine), reference model unmodified, RANDOMNESS = 0").unwrap();

    // ---- queries on the published initial state: a field of its own, like Flocker::new builds it (state.rs:24)
    let mut field: Field2D<Bird> = Field2D::new(w, h, DISCRETIZATION, TOROIDAL);
    for b in &birds { field.set_object_location(*b, b.loc); }
    field.lazy_update();
    for (i, (x, y, dist, relax)) in probes.iter().enumerate() {
        let loc = Real2D { x: *x, y: *y };
        let v = if *relax { field.get_neighbors_within_relax_distance(loc, *dist) } else { field.get_neighbors_within_distance(loc, *dist) };
        write!(out, "query {} {}", i, v.len()).unwrap();
        for e in &v { write!(out, " {}", e.id).unwrap(); }
        writeln!(out).unwrap();
    }
    for (i, (is_dist, a, b, dim)) in tprobes.iter().enumerate() {
        let r = if *is_dist { toroidal_distance(*a, *b, *dim) } else { toroidal_transform(*a, *dim) };
        writeln!(out, "tresult {} {:08x}", i, r.to_bits()).unwrap();
    }

    // ---- the model itself: what Flocker::init does (state.rs:40-51) with the given birds instead of random ones
    let mut state = Flocker::new((w, h), birds.len() as u32);
    let mut schedule = Schedule::new();
    for b in &birds {
        state.field1.set_object_location(*b, b.loc);
        schedule.schedule_repeating(Box::new(*b), 0., 0);
    }
    for t in 1..=ticks {
        schedule.step(&mut state);
        writeln!(out, "state {} {}", t, birds.len()).unwrap();
        for b in &birds {
            // Field2D::get(&object) (visualization/vis_state.rs:52): the stored copy, found by id (bird.rs:148-163)
            let cur = state.field1.get(b).expect("every bird stays in the field");
            writeln!(out, "{} {:08x} {:08x} {:08x} {:08x}", cur.id, cur.loc.x.to_bits(), cur.loc.y.to_bits(),
                     cur.last_d.x.to_bits(), cur.last_d.y.to_bits()).unwrap();
        }
    }
    fs::write(dst, out).expect("write pin output");
    eprintln!("wrote {dst}");
}
