// krabmaga_gpu.hpp — C++ host mirror of the krABMaga surface the Flockers model is written against, backed by
// libflk.so (include/flk.h).  The reference is Rust (compiled code) and no Rust toolchain exists in this image,
// so the host side above the C ABI is C++; rust/ holds the equivalent Rust shim as source.
//
// Mirrors (names, argument meaning, call order):
//   Real2D, Bird                       flockers/src/model/bird.rs:13-24
//   Field2D<Bird>                      call sites flockers/src/model/state.rs:24,49,55; bird.rs:29-31,142-144
//   State / Agent / Schedule::step     krabmaga::engine::{state,agent,schedule};
//                                      hand-written outer loop: forestfire_bayesian/src/main.rs:63-72
//   simulate_old!                      flockers/src/main.rs:47
// Error behaviour: the reference's Field2D methods return no Result, so a non-zero FLK code throws
// std::runtime_error (Rust shim: panic!).
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "flk.h"

namespace krabmaga_gpu {

struct Real2D {
    float x, y;
};

struct Bird {
    uint32_t id;
    Real2D loc;
    Real2D last_d;
};

inline void check(int rc) {
    if (rc != FLK_OK) throw std::runtime_error(std::string("libflk: ") + flk_last_error());
}

// flockers/src/main.rs:26-33
constexpr float COHESION = 0.8f, AVOIDANCE = 1.0f, RANDOMNESS = 1.1f, CONSISTENCY = 0.7f, MOMENTUM = 1.0f, JUMP = 0.7f;
constexpr float DISCRETIZATION = 10.0f / 1.5f;
constexpr bool TOROIDAL = true;

class Field2D {
public:
    Field2D(float width, float height, float discretization, bool toroidal, uint64_t capacity, int device = 0)
        : width(width), height(height) {
        check(flk_field_create(width, height, discretization, toroidal ? 1 : 0, capacity, device, &h_));
    }
    ~Field2D() { flk_field_destroy(h_); }
    Field2D(const Field2D &) = delete;
    Field2D &operator=(const Field2D &) = delete;

    void set_object_location(const Bird &b, Real2D loc) {
        check(flk_set_object_location(h_, b.id, loc.x, loc.y, b.last_d.x, b.last_d.y));
    }
    // like the reference (bird.rs:29-31): the objects themselves, in query order — Bird::step reads elem.loc / elem.last_d
    std::vector<Bird> get_neighbors_within_relax_distance(Real2D loc, float dist) { return query(loc, dist, true); }
    std::vector<Bird> get_neighbors_within_distance(Real2D loc, float dist) { return query(loc, dist, false); }
    // Field2D::get(&object) / get_location(object) (visualization/vis_state.rs:52, bird_vis.rs:23); false = not in the field
    bool get(const Bird &b, Bird *out) {
        float r[4];
        int32_t found = 0;
        check(flk_get_object(h_, b.id, r, &found));
        if (found && out) *out = Bird{b.id, {r[0], r[1]}, {r[2], r[3]}};
        return found != 0;
    }
    bool get_location(const Bird &b, Real2D *loc) {
        Bird cur;
        if (!get(b, &cur)) return false;
        if (loc) *loc = cur.loc;
        return true;
    }
    void lazy_update() { check(flk_lazy_update(h_)); }
    void clear() { check(flk_field_clear(h_)); }
    void step_all_birds(uint64_t step, uint64_t seed) { check(flk_step(h_, step, seed, nullptr, 0)); }
    uint64_t num_objects() const {
        uint64_t n = 0;
        check(flk_num_objects(h_, &n));
        return n;
    }
    // the agents as the Schedule keeps them, a list indexed by id (state.rs:40-51): 16 bytes per bird each way
    void set_all_dense(const std::vector<float> &rec4) {
        check(flk_set_objects_dense(h_, rec4.size() / 4, 0, rec4.data()));
    }
    std::vector<float> get_all_dense(size_t n) {
        std::vector<float> rec4(4 * n);
        check(flk_snapshot_dense(h_, 0, n, rec4.data()));
        return rec4;
    }
    std::vector<Bird> get_all() {   // Field2D::get / get_location for every object (vis_state.rs:52)
        const uint64_t n = num_objects();
        std::vector<uint32_t> id(n);
        std::vector<float> loc(2 * n), ld(2 * n);
        uint64_t got = 0;
        check(flk_snapshot(h_, id.data(), loc.data(), ld.data(), &got));
        std::vector<Bird> out(got);
        for (uint64_t i = 0; i < got; ++i) out[i] = Bird{id[i], {loc[2 * i], loc[2 * i + 1]}, {ld[2 * i], ld[2 * i + 1]}};
        return out;
    }
    void synchronize() { check(flk_synchronize(h_)); }
    flk_field *raw() { return h_; }
    const float width, height;

private:
    std::vector<Bird> query(Real2D loc, float dist, bool relax) {
        static_assert(sizeof(flk_bird) == sizeof(Bird), "Bird is flk_bird: {id, loc, last_d}");
        std::vector<Bird> out(256);
        uint32_t n = 0;
        for (;;) {
            flk_bird *p = reinterpret_cast<flk_bird *>(out.data());
            int rc = relax ? flk_get_neighbors_within_relax_distance_objects(h_, loc.x, loc.y, dist, p, (uint32_t)out.size(), &n)
                           : flk_get_neighbors_within_distance_objects(h_, loc.x, loc.y, dist, p, (uint32_t)out.size(), &n);
            if (rc == FLK_E_CAPACITY && n > out.size()) { out.resize(n); continue; }
            check(rc);
            out.resize(n);
            return out;
        }
    }
    flk_field *h_ = nullptr;
};

class Schedule;

struct State {   // krabmaga::engine::state::State
    virtual ~State() = default;
    virtual void init(Schedule &schedule) = 0;
    virtual void update(uint64_t step) = 0;
    virtual void reset() = 0;
    virtual void before_step(Schedule &) {}
    virtual void after_step(Schedule &) {}
    virtual bool end_condition(Schedule &) { return false; }
};

struct Agent {   // krabmaga::engine::agent::Agent
    virtual ~Agent() = default;
    virtual void step(State &state) = 0;
    virtual bool is_stopped(State &) { return false; }
};

class Schedule {   // krabmaga::engine::schedule::Schedule, repeating agents only
public:
    uint64_t step = 0;
    float time = 0.0f;
    void schedule_repeating(std::shared_ptr<Agent> a, float /*time*/, int /*ordering*/) { events_.push_back(std::move(a)); }
    void run_step(State &state) {   // Schedule::step (Appendix B)
        if (step == 0) state.update(0);
        state.before_step(*this);
        for (auto &a : events_) a->step(state);
        state.after_step(*this);
        step += 1;
        time += 1.0f;
        state.update(step);
    }

private:
    std::vector<std::shared_ptr<Agent>> events_;
};

// simulate_old!(state, n_step, reps, Info::Normal): returns seconds per repetition and prints the table line
template <class S>
std::vector<double> simulate(S &state, uint64_t n_step, int reps, bool verbose = true) {
    std::vector<double> out;
    for (int r = 0; r < reps; ++r) {
        Schedule schedule;
        state.reset();
        state.init(schedule);
        auto t0 = std::chrono::steady_clock::now();
        for (uint64_t i = 0; i < n_step; ++i) {
            schedule.run_step(state);
            if (state.end_condition(schedule)) break;
        }
        state.sync();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out.push_back(dt);
        if (verbose) std::printf("Time elapsed: %.6f s   steps/s: %.2f\n", dt, n_step / dt);
    }
    return out;
}

}  // namespace krabmaga_gpu
