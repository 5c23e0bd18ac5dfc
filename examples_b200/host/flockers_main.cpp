// flockers_main.cpp — the Flockers model of flockers/src/{main.rs,model/state.rs,model/bird.rs} written against
// the C++ host mirror.  Per-bird Agent::step dispatch is replaced by ONE scheduled FlockStepper agent (the repo's
// own whole-field-agent idiom: schelling/src/model/updater.rs, forestfire/src/model/spread.rs); everything else
// keeps the reference's shape:  main -> Flocker::new -> simulate(state, step, 1).
//
//   ./flockers_demo [num_agents=1000] [dim=100] [steps=200]
#include <cstdlib>
#include <random>

#include "krabmaga_gpu.hpp"

using namespace krabmaga_gpu;

struct Flocker;

struct FlockStepper : Agent {   // stands in for the 1000 scheduled Birds (bird.rs:26-146)
    uint64_t seed;
    explicit FlockStepper(uint64_t s) : seed(s) {}
    void step(State &state) override;
};

struct Flocker : State {   // flockers/src/model/state.rs:12-73
    uint64_t step = 0;
    Field2D field1;
    uint32_t initial_flockers;
    float dim_w, dim_h;
    Flocker(float w, float h, uint32_t n) : field1(w, h, DISCRETIZATION, TOROIDAL, n), initial_flockers(n), dim_w(w), dim_h(h) {}
    void reset() override {   // state.rs:32-35
        step = 0;
        field1.clear();
    }
    void init(Schedule &schedule) override {   // state.rs:37-52
        std::mt19937 rng(0xF10C);
        std::uniform_real_distribution<float> u(0.0f, 1.0f);
        for (uint32_t id = 0; id < initial_flockers; ++id) {
            float r1 = u(rng), r2 = u(rng);
            if (r1 >= 1.0f) r1 = 0.0f;
            if (r2 >= 1.0f) r2 = 0.0f;
            Bird bird{id, {dim_w * r1, dim_h * r2}, {0.0f, 0.0f}};
            field1.set_object_location(bird, bird.loc);
        }
        schedule.schedule_repeating(std::make_shared<FlockStepper>(0xB12D), 0.0f, 0);
    }
    void update(uint64_t s) override {   // state.rs:54-56
        step = s;
        field1.lazy_update();
    }
    void sync() { field1.synchronize(); }
};

void FlockStepper::step(State &state) {
    auto &s = static_cast<Flocker &>(state);
    s.field1.step_all_birds(s.step, seed);
}

int main(int argc, char **argv) {   // flockers/src/main.rs:37-48
    const uint32_t num_agents = argc > 1 ? (uint32_t)std::atoll(argv[1]) : 1000;
    const float dim = argc > 2 ? (float)std::atof(argv[2]) : 100.0f;
    const uint64_t step = argc > 3 ? (uint64_t)std::atoll(argv[3]) : 200;
    try {
        Flocker state(dim, dim, num_agents);
        simulate(state, step, 1);
        auto birds = state.field1.get_all();
        double sx = 0, sy = 0;
        for (auto &b : birds) { sx += b.loc.x; sy += b.loc.y; }
        std::printf("birds: %zu  mean loc: (%.4f, %.4f)  neighbours of (50,50): %zu\n", birds.size(), sx / birds.size(),
                    sy / birds.size(), state.field1.get_neighbors_within_relax_distance({50.0f, 50.0f}, 10.0f).size());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
