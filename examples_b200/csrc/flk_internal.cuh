// flk_internal.cuh — the handle behind `flk_field*` and helpers shared by flk_api.cu / flk_dist.cu.
#pragma once
#include <cstdarg>
#include <string>
#include <vector>

#include "flk_kernels.cuh"
#include "flk_slots.cuh"

struct flk_dist_state;  // flk_dist.cu

struct flk_field {
    int device = 0;
    cudaStream_t stream = nullptr;
    flk::Geom g{};
    flk::Params p{};
    int math_mode = 0;
    uint64_t cap = 0;
    uint32_t ncell = 0;      // local cells = ncols*dh
    flk::Store s{};
    uint8_t *scratch = nullptr;
    uint64_t n_read = 0;     // objects in the READ buffer (host view; an upper bound on dist handles)
    uint64_t w_count = 0;    // WRITE slots in use
    std::vector<uint32_t> pend_id;   // buffered single set_object_location calls
    std::vector<float> pend_loc, pend_ld;
    uint32_t *d_paystage = nullptr;   // upload staging of payload words
    uint64_t paystage_cap = 0;
    float2 *d_injected = nullptr;
    // flk_frame_*: a device-side copy of the READ buffer that a side stream downloads while the simulation goes on
    flk::Rec *frame_rec = nullptr;
    uint32_t *frame_ids = nullptr;
    cudaStream_t frame_stream = nullptr;
    cudaEvent_t ev_frame_ready = nullptr, ev_frame_done = nullptr;
    bool frame_pending = false;
    uint32_t *h_err = nullptr;        // pinned word: device error bits (ERR_*) as of the last flk_synchronize
    uint32_t *h_dense = nullptr;      // pinned word: ids outside the range of the last dense snapshot
    bool dense_pending = false;
    uint64_t injected_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // flk_run graph cache
    bool use_graph = true, graph_valid = false, capturing = false;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_n = 0, graph_seed = 0, graph_launches = 0;
    uint64_t launches = 0;
    bool profile = false;
    int profile_level = 0;
    std::vector<cudaEvent_t> prof_events;      // pairs around every k_step launch (set)
    std::vector<cudaEvent_t> prof_rebin;       // quadruples per lazy_update: start, after scan, after scatter, after rank
    flk_dist_state *dist = nullptr;
    // fixed-capacity-cell layout (flk_slots.cuh).  layout: 0 = sorted runs (rec / ids / cell_start), 1 = slots.
    flk::SlotStore S{};
    int layout = 0;
    bool slot_capable = false;     // geometry / payload / capacity allow the slotted layout (arrays allocated)
    int layout_pref = -1;          // -1 automatic, 0 classic, 1 slots, 2 slot-staged WRITE buffer (flk_set_layout / FLK_LAYOUT)
    bool layout_forced = false;    // FLK_LAYOUT: keep the requested form whatever the spill count says (tests)
    bool view_valid = false;       // layout 1: rec / ids / cell_start hold the READ buffer too (read-side calls)
    bool uploaded = false;         // objects were uploaded since the last lazy_update (layout decision is due)
    uint32_t *h_totals = nullptr;  // pinned: {birds, largest column, largest cell, spilled} of a recent READ buffer
    uint32_t *d_stats = nullptr;   // 2 words: cell statistics of the sorted-runs READ buffer
    uint32_t *h_sw = nullptr;      // pinned: {largest cell, spilled arrivals} of a recent slot-staged lazy_update
    cudaGraphExec_t sgraph[2] = {nullptr, nullptr};   // layout 1: two ticks, first one reading buffer 0 / 1
    bool sgraph_valid[2] = {false, false};
    uint64_t sgraph_n[2] = {0, 0}, sgraph_seed[2] = {0, 0}, sgraph_launches[2] = {0, 0};
};

namespace flk {

int fail(int code, const char *fmt, ...);

#define CUDA_TRY(x)                                                                                     \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess) return flk::fail(FLK_E_CUDA, "%s: %s", #x, cudaGetErrorString(e_));      \
    } while (0)

#define CHECK_HANDLE(f)                                                                                 \
    do {                                                                                                \
        if (!(f)) return flk::fail(FLK_E_INVALID, "null handle");                                       \
        cudaError_t e_ = cudaSetDevice((f)->device);                                                    \
        if (e_ != cudaSuccess) return flk::fail(FLK_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); \
    } while (0)

int setup_geometry(flk_field *f, float w, float h, float d, int toroidal);
int update_window(flk_field *f);
int field_alloc(flk_field *f);
void field_free(flk_field *f);
int flush_pending(flk_field *f);
int write_append(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *ld,
                 const uint32_t *payload = nullptr);
int upload_injected(flk_field *f, const float *r, uint64_t n);
int enqueue_step(flk_field *f, uint64_t step, int use_step_dev, uint64_t seed, const float2 *d_inj, uint32_t n_inj);
int enqueue_lazy_update(flk_field *f, int bump_step);

// flk_dist.cu
void dist_free(flk_field *f);
int dist_set_objects(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *ld,
                     const uint32_t *payload = nullptr);
int dist_run(flk_field *f, uint64_t first_step, uint64_t n_steps, uint64_t seed);
int dist_refresh_counts(flk_field *f);
void dist_reset_counts(flk_field *f);
int report_device_errors(flk_field *f, uint32_t flag);
// layout management (flk_api.cu)
int slots_alloc(flk_field *f);
bool slots_params_ok(const flk_field *f);
int slots_publish(flk_field *f, int bump_step);
uint32_t slot_col_bound(const flk_field *f);
int ensure_view(flk_field *f);
int demote_to_runs(flk_field *f);
int maybe_promote_to_slots(flk_field *f);
int decide_write_mode(flk_field *f);
int set_slotw(flk_field *f, int on);
int run_slots(flk_field *f, uint64_t first_step, uint64_t n_steps, uint64_t seed);

}  // namespace flk
