// flk_api.cu — C-ABI layer of libflk.so (include/flk.h): handle, device memory, stream, CUDA graph, timing.
// There is no CPU fallback anywhere in this file: a missing device or a failed launch is an error code.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flk.h"
#include "flk_internal.cuh"

namespace flk {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

static float f32_ceil_div(float a, float b) { return std::ceil(a / b); }  // f32 divide, like the reference

int field_alloc(flk_field *f) {
    const uint64_t cap = f->cap;
    const uint64_t m = (uint64_t)f->ncell + 1;
    Store &s = f->s;
    s.cap = (uint32_t)cap;
    if (f->g.mode == 0) s.capx = 0;
    s.nrecv = f->g.mode == 2 ? 4u : 2u;
    const uint64_t capw = cap + (uint64_t)s.nrecv * s.capx;   // WRITE slots: stepped/uploaded + received from the neighbours
    CUDA_TRY(cudaMalloc(&s.rec, cap * sizeof(Rec)));
    CUDA_TRY(cudaMalloc(&s.ids, cap * 4));
    const uint64_t mp = ((m + 2047) / 2048) * 2048;   // padded to the scan tile (k_scan_* use whole-tile vector accesses)
    CUDA_TRY(cudaMalloc(&s.cell_start, mp * 4));
    CUDA_TRY(cudaMalloc(&s.trec, capw * sizeof(Rec)));
    CUDA_TRY(cudaMalloc(&s.tid, capw * 4));
    CUDA_TRY(cudaMalloc(&s.tkey, capw * 4));
    CUDA_TRY(cudaMalloc(&s.tslot, capw * 4));
    CUDA_TRY(cudaMalloc(&s.cnt, mp * 4));
    // one scratch block: the ranking stage (16 B/bird) during lazy_update; upload / snapshot staging otherwise
    CUDA_TRY(cudaMalloc(&f->scratch, cap * 20 + 128));
    s.stage = reinterpret_cast<uint4 *>(f->scratch);
    const uint64_t nblk = (m + 2047) / 2048 + 1;
    CUDA_TRY(cudaMalloc(&s.blocksum, nblk * 4));
    CUDA_TRY(cudaMalloc(&s.step_dev, 8));
    CUDA_TRY(cudaMalloc(&s.n_read_dev, 4));
    CUDA_TRY(cudaMalloc(&s.err_flag, 4));
    CUDA_TRY(cudaMemsetAsync(s.err_flag, 0, 4, f->stream));
    CUDA_TRY(cudaHostAlloc(&f->h_err, 4, cudaHostAllocDefault));
    *f->h_err = 0;
    if (s.pw) {
        CUDA_TRY(cudaMalloc(&s.pay, cap * s.pw * 4));
        CUDA_TRY(cudaMalloc(&s.tpay, capw * s.pw * 4));
    }
    CUDA_TRY(cudaMemsetAsync(s.cell_start, 0, mp * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(s.cnt, 0, mp * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(s.step_dev, 0, 8, f->stream));
    CUDA_TRY(cudaMemsetAsync(s.n_read_dev, 0, 4, f->stream));
    return FLK_OK;
}

void field_free(flk_field *f) {
    Store &s = f->s;
    cudaFree(s.rec); cudaFree(s.ids); cudaFree(s.cell_start); cudaFree(s.trec); cudaFree(s.tid);
    cudaFree(s.tkey); cudaFree(s.tslot); cudaFree(s.cnt); cudaFree(f->scratch); cudaFree(s.blocksum);
    cudaFree(s.step_dev); cudaFree(s.n_read_dev); cudaFree(s.err_flag);
    if (s.pay) cudaFree(s.pay);
    if (s.tpay) cudaFree(s.tpay);
    if (f->d_paystage) cudaFree(f->d_paystage);
    if (f->d_injected) cudaFree(f->d_injected);
    if (f->h_dense) cudaFreeHost(f->h_dense);
    if (f->h_err) cudaFreeHost(f->h_err);
    if (f->frame_rec) cudaFree(f->frame_rec);
    if (f->frame_ids) cudaFree(f->frame_ids);
    if (f->ev_frame_ready) cudaEventDestroy(f->ev_frame_ready);
    if (f->ev_frame_done) cudaEventDestroy(f->ev_frame_done);
    if (f->frame_stream) cudaStreamDestroy(f->frame_stream);
    if (f->graph_exec) cudaGraphExecDestroy(f->graph_exec);
    for (auto &e : f->prof_events) cudaEventDestroy(e);
    for (auto &e : f->prof_rebin) cudaEventDestroy(e);
    if (f->ev0) cudaEventDestroy(f->ev0);
    if (f->ev1) cudaEventDestroy(f->ev1);
    if (f->stream) cudaStreamDestroy(f->stream);
}

int setup_geometry(flk_field *f, float w, float h, float d, int toroidal) {
    if (!(w > 0.0f) || !(h > 0.0f) || !(d > 0.0f)) return fail(FLK_E_INVALID, "width, height, discretization must be > 0");
    Geom &g = f->g;
    g.W = w; g.H = h; g.d = d; g.toroidal = toroidal ? 1 : 0;
    const float cw = f32_ceil_div(w, d), ch = f32_ceil_div(h, d);
    if (cw > 46000.0f || ch > 46000.0f) return fail(FLK_E_UNSUPPORTED, "grid too large (%g x %g cells)", cw, ch);
    g.mx = (int)cw; g.my = (int)ch;
    g.dh = g.my + 1;
    g.dhg = g.dh; g.row0 = 0; g.rown = g.my; g.has_slack_row = 1;
    g.ncols = g.mx + 1;
    g.mode = 0; g.col0 = 0; g.nown = g.mx; g.has_slack = 1;
    return FLK_OK;
}

int update_window(flk_field *f) {
    const float q = std::floor(f->p.radius / f->g.d);
    int k = (q != q) ? 0 : (int)q;
    if (k < 0) k = 0;
    if (f->p.radius > 0.0f && (2 * k + 1 > f->g.mx || 2 * k + 1 > f->g.my) && f->g.toroidal)
        return fail(FLK_E_UNSUPPORTED, "query window (%d cells) wider than the grid (%d x %d): duplicate visits are not implemented",
                    2 * k + 1, f->g.mx, f->g.my);
    if (f->g.mode != 0 && k > 1) return fail(FLK_E_UNSUPPORTED, "strip decomposition implements a one-column halo (k<=1), got k=%d", k);
    if (f->g.mode != 0 && !(std::fabs(f->p.jump) < f->g.d))
        return fail(FLK_E_UNSUPPORTED, "strip decomposition needs |jump| < discretisation (a bird may cross one column per tick)");
    f->p.k = k;
    return FLK_OK;
}

// flush host-side single set_object_location calls into the WRITE buffer
int flush_pending(flk_field *f) {
    if (f->pend_id.empty()) return FLK_OK;
    std::vector<uint32_t> id;
    std::vector<float> loc, ld;
    id.swap(f->pend_id); loc.swap(f->pend_loc); ld.swap(f->pend_ld);
    if (f->g.mode != 0) return dist_set_objects(f, id.size(), id.data(), loc.data(), ld.data());
    return write_append(f, id.size(), id.data(), loc.data(), ld.data());
}

int write_append(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *ld,
                 const uint32_t *payload) {
    if (n == 0) return FLK_OK;
    const uint32_t *d_pay = nullptr;
    if (payload && f->s.pw) {
        const uint64_t words = n * f->s.pw;
        if (words > f->paystage_cap) {
            if (f->d_paystage) { CUDA_TRY(cudaStreamSynchronize(f->stream)); cudaFree(f->d_paystage); f->d_paystage = nullptr; }
            CUDA_TRY(cudaMalloc(&f->d_paystage, words * 4));
            f->paystage_cap = words;
        }
        CUDA_TRY(cudaMemcpyAsync(f->d_paystage, payload, words * 4, cudaMemcpyHostToDevice, f->stream));
        d_pay = f->d_paystage;
    }
    if (f->w_count + n > f->cap)
        return fail(FLK_E_CAPACITY, "WRITE buffer would hold %llu objects, capacity is %llu",
                    (unsigned long long)(f->w_count + n), (unsigned long long)f->cap);
    // staging inside the scratch block (free outside lazy_update; stream order protects it)
    uint32_t *d_id = reinterpret_cast<uint32_t *>(f->scratch);
    float2 *d_loc = reinterpret_cast<float2 *>(f->scratch + n * 4 + ((n * 4) % 8));
    float2 *d_ld = d_loc + n;
    CUDA_TRY(cudaMemcpyAsync(d_id, id, n * 4, cudaMemcpyHostToDevice, f->stream));
    CUDA_TRY(cudaMemcpyAsync(d_loc, loc, n * 8, cudaMemcpyHostToDevice, f->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ld, ld, n * 8, cudaMemcpyHostToDevice, f->stream));
    launch_ingest(f->g, f->s, (uint32_t)n, d_id, d_loc, d_ld, d_pay, (uint32_t)f->w_count, 1, f->stream);
    f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    f->w_count += n;
    return FLK_OK;
}

int enqueue_lazy_update(flk_field *f, int bump_step) {
    const uint32_t m = f->ncell + 1;
    const bool prof = f->profile && f->profile_level >= 2 && !f->capturing;
    auto mark = [&]() -> int {
        if (!prof) return FLK_OK;
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        f->prof_rebin.push_back(e);
        CUDA_TRY(cudaEventRecord(e, f->stream));
        return FLK_OK;
    };
    if (int rc = mark()) return rc;
    launch_scan(f->s, m, f->stream);
    if (int rc = mark()) return rc;
    launch_scatter(f->s, (uint32_t)f->w_count, f->g.mode != 0, f->stream);
    if (int rc = mark()) return rc;
    launch_rank(f->s, (uint32_t)f->w_count, f->ncell, bump_step, f->stream);
    if (int rc = mark()) return rc;
    f->launches += 3 + (f->g.mode != 0 ? 2 : (f->w_count ? 1 : 0)) + 1;
    CUDA_TRY(cudaGetLastError());
    f->n_read = f->w_count;  // single GPU: every WRITE entry is kept
    f->w_count = 0;
    return FLK_OK;
}

int enqueue_step(flk_field *f, uint64_t step, int use_step_dev, uint64_t seed, const float2 *d_inj, uint32_t n_inj) {
    if (f->w_count + f->n_read > f->cap)
        return fail(FLK_E_CAPACITY, "WRITE buffer would hold %llu objects, capacity is %llu",
                    (unsigned long long)(f->w_count + f->n_read), (unsigned long long)f->cap);
    StepArgs a;
    a.cell_lo = 0; a.cell_hi = f->ncell;
    a.step = step; a.use_step_dev = use_step_dev; a.seed = seed;
    a.injected = d_inj; a.n_injected = n_inj;
    a.out_base = (uint32_t)f->w_count;
    a.math_mode = f->math_mode;
    a.identity = 0;
    a.skip_bands = 0;
    const bool prof = f->profile && !f->capturing;
    if (prof) {
        cudaEvent_t e0, e1;
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        f->prof_events.push_back(e0); f->prof_events.push_back(e1);
        CUDA_TRY(cudaEventRecord(e0, f->stream));
    }
    launch_step(f->g, f->p, f->s, a, (uint32_t)f->n_read, f->stream);
    if (prof) CUDA_TRY(cudaEventRecord(f->prof_events.back(), f->stream));
    if (f->n_read) f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    f->w_count += f->n_read;
    return FLK_OK;
}

}  // namespace flk

using namespace flk;

extern "C" {

const char *flk_last_error(void) { return g_err.c_str(); }
const char *flk_version(void) { return "flk 0.1 (sm_100a)"; }

int flk_field_create(float width, float height, float discretization, int toroidal, uint64_t capacity, int device,
                     flk_field **out) {
    return flk_field_create_ex(width, height, discretization, toroidal, capacity, device, 0, out);
}

int flk_field_create_ex(float width, float height, float discretization, int toroidal, uint64_t capacity, int device,
                        uint32_t payload_bytes, flk_field **out) {
    if (!out) return fail(FLK_E_INVALID, "out is null");
    if (payload_bytes % 4 != 0 || payload_bytes > 256) return fail(FLK_E_INVALID, "payload_bytes must be a multiple of 4, <= 256");
    *out = nullptr;
    if (capacity == 0 || capacity > 0x7fffffffull) return fail(FLK_E_INVALID, "capacity must be in [1, 2^31)");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(FLK_E_CUDA, "no CUDA device (%s): libflk has no CPU fallback", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(FLK_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    flk_field *f = new flk_field();
    f->device = device;
    int rc = setup_geometry(f, width, height, discretization, toroidal);
    if (rc) { delete f; return rc; }
    f->p = Params{0.8f, 1.0f, 1.1f, 0.7f, 1.0f, 0.7f, 10.0f, 1, 1};  // flockers/src/main.rs:26-31, bird.rs:31
    rc = update_window(f);
    if (rc) { delete f; return rc; }
    f->cap = capacity;
    f->ncell = (uint32_t)f->g.ncols * (uint32_t)f->g.dh;
    f->s.pw = payload_bytes / 4;
    rc = [&]() -> int {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&f->ev0));
        CUDA_TRY(cudaEventCreate(&f->ev1));
        return field_alloc(f);
    }();
    if (rc) { field_free(f); delete f; return rc; }   // whatever was created so far (cudaFree(nullptr) is a no-op)
    *out = f;
    return FLK_OK;
}

int flk_field_destroy(flk_field *f) {
    if (!f) return FLK_OK;
    cudaSetDevice(f->device);
    cudaStreamSynchronize(f->stream);
    dist_free(f);
    field_free(f);
    delete f;
    return FLK_OK;
}

int flk_field_clear(flk_field *f) {
    CHECK_HANDLE(f);
    const uint64_t m = (uint64_t)f->ncell + 1;
    CUDA_TRY(cudaMemsetAsync(f->s.cell_start, 0, m * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(f->s.cnt, 0, m * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(f->s.n_read_dev, 0, 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(f->s.err_flag, 0, 4, f->stream));
    f->n_read = 0; f->w_count = 0;
    f->pend_id.clear(); f->pend_loc.clear(); f->pend_ld.clear();
    f->graph_valid = false;
    if (f->dist) dist_reset_counts(f);
    return FLK_OK;
}

int flk_set_params(flk_field *f, float cohesion, float avoidance, float randomness, float consistency, float momentum,
                   float jump, float radius, int relax) {
    CHECK_HANDLE(f);
    Params old = f->p;
    f->p.cohesion = cohesion; f->p.avoidance = avoidance; f->p.randomness = randomness;
    f->p.consistency = consistency; f->p.momentum = momentum; f->p.jump = jump;
    f->p.radius = radius; f->p.relax = relax ? 1 : 0;
    int rc = update_window(f);
    if (rc) { f->p = old; return rc; }
    f->graph_valid = false;
    return FLK_OK;
}

int flk_set_math_mode(flk_field *f, int mode) {
    CHECK_HANDLE(f);
    if (mode != 0 && mode != 1) return fail(FLK_E_INVALID, "math mode must be 0 (exact) or 1 (fast)");
    f->math_mode = mode;
    f->graph_valid = false;
    return FLK_OK;
}

int flk_field_dims(const flk_field *f, int32_t *dw, int32_t *dh, int32_t *mx, int32_t *my) {
    if (!f) return fail(FLK_E_INVALID, "null handle");
    if (dw) *dw = f->g.mx + 1;
    if (dh) *dh = f->g.dh;
    if (mx) *mx = f->g.mx;
    if (my) *my = f->g.my;
    return FLK_OK;
}

int flk_set_object_location(flk_field *f, uint32_t id, float x, float y, float last_dx, float last_dy) {
    CHECK_HANDLE(f);
    if (f->w_count + f->pend_id.size() + 1 > f->cap) return fail(FLK_E_CAPACITY, "capacity %llu exceeded", (unsigned long long)f->cap);
    f->pend_id.push_back(id);
    f->pend_loc.push_back(x); f->pend_loc.push_back(y);
    f->pend_ld.push_back(last_dx); f->pend_ld.push_back(last_dy);
    return FLK_OK;
}

int flk_set_objects(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *last_d) {
    CHECK_HANDLE(f);
    if (n && (!id || !loc || !last_d)) return fail(FLK_E_INVALID, "null array");
    int rc = flush_pending(f);
    if (rc) return rc;
    if (f->g.mode != 0) return dist_set_objects(f, n, id, loc, last_d);
    return write_append(f, n, id, loc, last_d);
}

int flk_set_objects_ex(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *last_d,
                       const void *payload) {
    CHECK_HANDLE(f);
    if (n && (!id || !loc || !last_d)) return fail(FLK_E_INVALID, "null array");
    if (f->g.mode != 0) return fail(FLK_E_UNSUPPORTED, "payloads are not carried by strip handles");
    int rc = flush_pending(f);
    if (rc) return rc;
    return write_append(f, n, id, loc, last_d, static_cast<const uint32_t *>(payload));
}

int flk_snapshot_ex(flk_field *f, uint32_t *id, float *loc, float *last_d, void *payload, uint64_t *n) {
    CHECK_HANDLE(f);
    if (payload && f->s.pw && f->n_read)
        CUDA_TRY(cudaMemcpyAsync(payload, f->s.pay, f->n_read * f->s.pw * 4, cudaMemcpyDeviceToHost, f->stream));
    return flk_snapshot(f, id, loc, last_d, n);
}

int flk_lazy_update(flk_field *f) {
    CHECK_HANDLE(f);
    int rc = flush_pending(f);
    if (rc) return rc;
    if (f->g.mode != 0) return fail(FLK_E_STATE, "dist handles fold lazy_update into flk_dist_step / flk_dist_commit");
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    rc = enqueue_lazy_update(f, 0);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    return FLK_OK;
}

int flk_step(flk_field *f, uint64_t step, uint64_t seed, const float *injected_r, uint64_t n_injected) {
    CHECK_HANDLE(f);
    int rc = flush_pending(f);
    if (rc) return rc;
    if (f->g.mode != 0) return fail(FLK_E_STATE, "use flk_dist_step on a dist handle");
    const float2 *d_inj = nullptr;
    if (injected_r) {
        rc = upload_injected(f, injected_r, n_injected);
        if (rc) return rc;
        d_inj = f->d_injected;
    }
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    rc = enqueue_step(f, step, 0, seed, d_inj, (uint32_t)n_injected);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    return FLK_OK;
}

int flk_run(flk_field *f, uint64_t first_step, uint64_t n_steps, uint64_t seed) {
    CHECK_HANDLE(f);
    int rc = flush_pending(f);
    if (rc) return rc;
    if (f->g.mode != 0) return dist_run(f, first_step, n_steps, seed);
    if (f->w_count != 0) return fail(FLK_E_STATE, "flk_run needs an empty WRITE buffer (call flk_lazy_update first)");
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    // a captured tick pays off from a handful of replays on; short runs without a cached graph launch directly
    const bool have_graph = f->graph_valid && f->graph_n == f->n_read && f->graph_seed == seed;
    if (f->profile || !f->use_graph || (!have_graph && n_steps < 8)) {
        for (uint64_t t = 0; t < n_steps; ++t) {
            rc = enqueue_step(f, first_step + t, 0, seed, nullptr, 0);
            if (rc) return rc;
            rc = enqueue_lazy_update(f, 0);
            if (rc) return rc;
        }
    } else {
        if (!f->graph_valid || f->graph_n != f->n_read || f->graph_seed != seed) {
            if (f->graph_exec) { cudaGraphExecDestroy(f->graph_exec); f->graph_exec = nullptr; }
            cudaGraph_t graph;
            const uint64_t launches_before = f->launches;
            CUDA_TRY(cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal));
            f->capturing = true;
            rc = enqueue_step(f, 0, 1, seed, nullptr, 0);
            if (!rc) rc = enqueue_lazy_update(f, 1);
            f->capturing = false;
            cudaError_t ce = cudaStreamEndCapture(f->stream, &graph);
            f->graph_launches = f->launches - launches_before;
            f->launches = launches_before;
            if (rc) { if (ce == cudaSuccess) cudaGraphDestroy(graph); return rc; }
            CUDA_TRY(ce);
            ce = cudaGraphInstantiate(&f->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            CUDA_TRY(ce);
            f->graph_valid = true; f->graph_n = f->n_read; f->graph_seed = seed;
        }
        CUDA_TRY(cudaMemcpyAsync(f->s.step_dev, &first_step, 8, cudaMemcpyHostToDevice, f->stream));
        for (uint64_t t = 0; t < n_steps; ++t) CUDA_TRY(cudaGraphLaunch(f->graph_exec, f->stream));
        f->launches += f->graph_launches * n_steps;
    }
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    return FLK_OK;
}

int flk_num_objects(const flk_field *f, uint64_t *n) {
    if (!f || !n) return fail(FLK_E_INVALID, "null argument");
    *n = f->n_read;
    return FLK_OK;
}

static int snapshot_impl(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n, bool sync) {
    CHECK_HANDLE(f);
    if (f->g.mode != 0) { int rc = dist_refresh_counts(f); if (rc) return rc; }
    const uint64_t nr = f->n_read;
    if (n) *n = nr;
    if (nr) {
        float2 *d_loc = reinterpret_cast<float2 *>(f->scratch);
        float2 *d_ld = d_loc + nr;
        if (loc || last_d) {
            launch_split(f->s, (uint32_t)nr, d_loc, d_ld, f->stream);
            f->launches += 1;
            CUDA_TRY(cudaGetLastError());
        }
        if (id) CUDA_TRY(cudaMemcpyAsync(id, f->s.ids, nr * 4, cudaMemcpyDeviceToHost, f->stream));
        if (loc) CUDA_TRY(cudaMemcpyAsync(loc, d_loc, nr * 8, cudaMemcpyDeviceToHost, f->stream));
        if (last_d) CUDA_TRY(cudaMemcpyAsync(last_d, d_ld, nr * 8, cudaMemcpyDeviceToHost, f->stream));
    }
    if (sync) CUDA_TRY(cudaStreamSynchronize(f->stream));
    return FLK_OK;
}

int flk_snapshot(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n) {
    return snapshot_impl(f, id, loc, last_d, n, true);
}

int flk_snapshot_async(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n) {
    return snapshot_impl(f, id, loc, last_d, n, false);
}

// ---- dense (id-indexed) host interface: 16 bytes per object each way ---------------------------------------------
static uint32_t *dense_counter(flk_field *f) {
    return reinterpret_cast<uint32_t *>(f->scratch + ((f->cap * 20 + 15) & ~(uint64_t)15) + 64);
}

int flk_set_objects_dense(flk_field *f, uint64_t n, uint32_t id0, const float *rec4) {
    CHECK_HANDLE(f);
    if (n && !rec4) return fail(FLK_E_INVALID, "null array");
    if (f->g.mode != 0) return fail(FLK_E_UNSUPPORTED, "dense upload is for single-GPU handles (a strip owns arbitrary ids)");
    if ((uint64_t)id0 + n > 0x100000000ull) return fail(FLK_E_INVALID, "id range exceeds u32");
    int rc = flush_pending(f);
    if (rc) return rc;
    if (n == 0) return FLK_OK;
    if (f->w_count + n > f->cap)
        return fail(FLK_E_CAPACITY, "WRITE buffer would hold %llu objects, capacity is %llu",
                    (unsigned long long)(f->w_count + n), (unsigned long long)f->cap);
    float4 *d_rec = reinterpret_cast<float4 *>(f->scratch);   // staging (free outside lazy_update; stream order protects it)
    CUDA_TRY(cudaMemcpyAsync(d_rec, rec4, n * 16, cudaMemcpyHostToDevice, f->stream));
    launch_ingest_dense(f->g, f->s, (uint32_t)n, id0, d_rec, (uint32_t)f->w_count, f->stream);
    f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    f->w_count += n;
    return FLK_OK;
}

static int snapshot_dense_impl(flk_field *f, uint32_t id0, uint64_t n, float *rec4, bool sync) {
    CHECK_HANDLE(f);
    if (n && !rec4) return fail(FLK_E_INVALID, "null array");
    if (f->g.mode != 0) return fail(FLK_E_UNSUPPORTED, "dense snapshot is for single-GPU handles (a strip owns arbitrary ids)");
    if (n > f->cap) return fail(FLK_E_CAPACITY, "id range %llu larger than the capacity %llu", (unsigned long long)n, (unsigned long long)f->cap);
    if ((uint64_t)id0 + n > 0x100000000ull) return fail(FLK_E_INVALID, "id range exceeds u32");
    if (n == 0) return FLK_OK;
    if (!f->h_dense) CUDA_TRY(cudaHostAlloc(&f->h_dense, 4, cudaHostAllocDefault));
    float4 *d_out = reinterpret_cast<float4 *>(f->scratch);
    uint32_t *d_cnt = dense_counter(f);
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 4, f->stream));
    if (f->n_read != n) CUDA_TRY(cudaMemsetAsync(d_out, 0xFF, n * 16, f->stream));   // ids not in the field read back as NaN
    launch_gather_dense(f->s, (uint32_t)f->n_read, id0, (uint32_t)n, d_out, d_cnt, f->stream);
    f->launches += f->n_read ? 1 : 0;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(rec4, d_out, n * 16, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaMemcpyAsync(f->h_dense, d_cnt, 4, cudaMemcpyDeviceToHost, f->stream));
    f->dense_pending = true;
    if (sync) return flk_synchronize(f);
    return FLK_OK;
}

int flk_snapshot_dense(flk_field *f, uint32_t id0, uint64_t n, float *rec4) {
    return snapshot_dense_impl(f, id0, n, rec4, true);
}

int flk_snapshot_dense_async(flk_field *f, uint32_t id0, uint64_t n, float *rec4) {
    return snapshot_dense_impl(f, id0, n, rec4, false);
}

// ---- frames for a front-end: the READ buffer is copied aside on the handle's stream (device to device), then a side
// stream moves that copy to the host while the handle's stream is free to run the next ticks
int flk_frame_begin(flk_field *f, uint32_t *id, float *rec4, uint64_t *n) {
    CHECK_HANDLE(f);
    if (f->g.mode != 0) return fail(FLK_E_UNSUPPORTED, "frames are taken from single-GPU handles (use flk_dist_snapshot_owned_async on a strip)");
    if (!f->frame_stream) {
        CUDA_TRY(cudaMalloc(&f->frame_rec, f->cap * sizeof(Rec)));
        CUDA_TRY(cudaMalloc(&f->frame_ids, f->cap * 4));
        CUDA_TRY(cudaStreamCreateWithFlags(&f->frame_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&f->ev_frame_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&f->ev_frame_done, cudaEventDisableTiming));
    }
    const uint64_t nr = f->n_read;
    if (n) *n = nr;
    // the previous frame may still be on its way to the host: do not overwrite the copy it is reading
    if (f->frame_pending) CUDA_TRY(cudaStreamWaitEvent(f->stream, f->ev_frame_done, 0));
    if (nr) {
        if (rec4) CUDA_TRY(cudaMemcpyAsync(f->frame_rec, f->s.rec, nr * sizeof(Rec), cudaMemcpyDeviceToDevice, f->stream));
        if (id) CUDA_TRY(cudaMemcpyAsync(f->frame_ids, f->s.ids, nr * 4, cudaMemcpyDeviceToDevice, f->stream));
    }
    CUDA_TRY(cudaEventRecord(f->ev_frame_ready, f->stream));
    CUDA_TRY(cudaStreamWaitEvent(f->frame_stream, f->ev_frame_ready, 0));
    if (nr) {
        if (rec4) CUDA_TRY(cudaMemcpyAsync(rec4, f->frame_rec, nr * sizeof(Rec), cudaMemcpyDeviceToHost, f->frame_stream));
        if (id) CUDA_TRY(cudaMemcpyAsync(id, f->frame_ids, nr * 4, cudaMemcpyDeviceToHost, f->frame_stream));
    }
    CUDA_TRY(cudaEventRecord(f->ev_frame_done, f->frame_stream));
    f->frame_pending = true;
    return FLK_OK;
}

int flk_frame_wait(flk_field *f) {
    CHECK_HANDLE(f);
    if (!f->frame_pending) return FLK_OK;
    CUDA_TRY(cudaEventSynchronize(f->ev_frame_done));
    f->frame_pending = false;
    return FLK_OK;
}

int flk_get_object(flk_field *f, uint32_t id, float out[4], int32_t *found) {
    CHECK_HANDLE(f);
    if (!out || !found) return fail(FLK_E_INVALID, "null argument");
    *found = 0;
    // result slot at the tail of the scratch block (free outside lazy_update; stream order protects it)
    float4 *d_out = reinterpret_cast<float4 *>(f->scratch + ((f->cap * 20 + 15) & ~(uint64_t)15));
    uint32_t *d_found = reinterpret_cast<uint32_t *>(d_out + 1);
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 32, f->stream));
    launch_find(f->s, (uint32_t)f->n_read, id, d_out, d_found, f->stream);
    f->launches += f->n_read ? 1 : 0;
    CUDA_TRY(cudaGetLastError());
    float h[8];
    CUDA_TRY(cudaMemcpyAsync(h, d_out, 32, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    uint32_t cnt;
    memcpy(&cnt, &h[4], 4);
    if (cnt) { memcpy(out, h, 16); *found = 1; }
    return FLK_OK;
}

int flk_get_binning(flk_field *f, uint32_t *cell_start, uint32_t *sorted_ids) {
    CHECK_HANDLE(f);
    if (f->g.mode != 0) { int rc = dist_refresh_counts(f); if (rc) return rc; }
    if (cell_start)
        CUDA_TRY(cudaMemcpyAsync(cell_start, f->s.cell_start, ((uint64_t)f->ncell + 1) * 4, cudaMemcpyDeviceToHost, f->stream));
    if (sorted_ids && f->n_read)
        CUDA_TRY(cudaMemcpyAsync(sorted_ids, f->s.ids, f->n_read * 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    return FLK_OK;
}

int flk_query_batch(flk_field *f, uint64_t nq, const float *loc, float dist, int relax, uint64_t *offsets,
                    uint32_t *ids, uint64_t cap) {
    CHECK_HANDLE(f);
    if (!offsets || (nq && !loc)) return fail(FLK_E_INVALID, "null array");
    if (nq > 0x7fffffffull) return fail(FLK_E_INVALID, "too many queries");
    offsets[0] = 0;
    if (nq == 0) return FLK_OK;
    if (dist > 0.0f) {
        const int k = (int)std::floor(dist / f->g.d);
        if (f->g.toroidal && (2 * k + 1 > f->g.mx || 2 * k + 1 > f->g.my))
            return fail(FLK_E_UNSUPPORTED, "query window wider than the grid");
        if (f->g.mode != 0 && k > 1) return fail(FLK_E_UNSUPPORTED, "dist handles answer queries with k<=1 only");
    }
    float2 *d_loc = nullptr;
    uint32_t *d_counts = nullptr, *d_out = nullptr;
    uint64_t *d_off = nullptr;
    int rc = FLK_OK;
    std::vector<uint32_t> counts(nq);
#define Q_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(FLK_E_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); goto done; } } while (0)
    Q_TRY(cudaMalloc(&d_loc, nq * 8));
    Q_TRY(cudaMalloc(&d_counts, nq * 4));
    Q_TRY(cudaMalloc(&d_off, (nq + 1) * 8));
    Q_TRY(cudaMemcpyAsync(d_loc, loc, nq * 8, cudaMemcpyHostToDevice, f->stream));
    launch_query_count(f->g, f->s, (uint32_t)nq, d_loc, dist, relax, d_counts, f->stream);
    f->launches += 1;
    Q_TRY(cudaGetLastError());
    Q_TRY(cudaMemcpyAsync(counts.data(), d_counts, nq * 4, cudaMemcpyDeviceToHost, f->stream));
    Q_TRY(cudaStreamSynchronize(f->stream));
    for (uint64_t q = 0; q < nq; ++q) offsets[q + 1] = offsets[q] + counts[q];
    if (offsets[nq] > cap || (offsets[nq] && !ids)) {
        rc = fail(FLK_E_CAPACITY, "query result has %llu ids, buffer holds %llu", (unsigned long long)offsets[nq], (unsigned long long)cap);
        goto done;
    }
    if (offsets[nq]) {
        Q_TRY(cudaMalloc(&d_out, offsets[nq] * 4));
        Q_TRY(cudaMemcpyAsync(d_off, offsets, (nq + 1) * 8, cudaMemcpyHostToDevice, f->stream));
        launch_query_fill(f->g, f->s, (uint32_t)nq, d_loc, dist, relax, d_off, d_out, f->stream);
        f->launches += 1;
        Q_TRY(cudaGetLastError());
        Q_TRY(cudaMemcpyAsync(ids, d_out, offsets[nq] * 4, cudaMemcpyDeviceToHost, f->stream));
        Q_TRY(cudaStreamSynchronize(f->stream));
    }
done:
#undef Q_TRY
    cudaFree(d_loc); cudaFree(d_counts); cudaFree(d_off); cudaFree(d_out);
    return rc;
}

static int query_one(flk_field *f, float x, float y, float dist, int relax, uint32_t *ids, uint32_t cap, uint32_t *n) {
    CHECK_HANDLE(f);
    if (!n) return fail(FLK_E_INVALID, "n is null");
    const float loc[2] = {x, y};
    uint64_t off[2] = {0, 0};
    // first try with the caller's buffer; on overflow report the full size with the first `cap` ids
    int rc = flk_query_batch(f, 1, loc, dist, relax, off, ids, cap);
    *n = (uint32_t)off[1];
    if (rc == FLK_E_CAPACITY && off[1] > cap) {
        std::vector<uint32_t> tmp(off[1]);
        int rc2 = flk_query_batch(f, 1, loc, dist, relax, off, tmp.data(), tmp.size());
        if (rc2) return rc2;
        if (ids && cap) memcpy(ids, tmp.data(), (size_t)cap * 4);
        return fail(FLK_E_CAPACITY, "query result has %u ids, buffer holds %u", *n, cap);
    }
    return rc;
}

int flk_get_neighbors_within_relax_distance(flk_field *f, float x, float y, float dist, uint32_t *ids, uint32_t cap,
                                            uint32_t *n) {
    return query_one(f, x, y, dist, 1, ids, cap, n);
}

int flk_get_neighbors_within_distance(flk_field *f, float x, float y, float dist, uint32_t *ids, uint32_t cap,
                                      uint32_t *n) {
    return query_one(f, x, y, dist, 0, ids, cap, n);
}

int flk_synchronize(flk_field *f) {
    CHECK_HANDLE(f);
    // device-side error bits (a message or a rank over its capacity, a launch that did not cover its range) are
    // reported here, whatever call enqueued the work that raised them
    CUDA_TRY(cudaMemcpyAsync(f->h_err, f->s.err_flag, 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    if (*f->h_err) return report_device_errors(f, *f->h_err);
    if (f->dense_pending) {   // outcome of the last (possibly asynchronous) dense snapshot
        f->dense_pending = false;
        if (*f->h_dense)
            return fail(FLK_E_INVALID, "dense snapshot: %u objects have ids outside the requested range", *f->h_dense);
    }
    return FLK_OK;
}

int flk_elapsed_ms(flk_field *f, float *ms) {
    CHECK_HANDLE(f);
    if (!ms) return fail(FLK_E_INVALID, "ms is null");
    CUDA_TRY(cudaEventSynchronize(f->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, f->ev0, f->ev1));
    return FLK_OK;
}

int flk_stream(flk_field *f, void **stream) {
    if (!f || !stream) return fail(FLK_E_INVALID, "null argument");
    *stream = (void *)f->stream;
    return FLK_OK;
}

int flk_launch_count(const flk_field *f, uint64_t *n) {
    if (!f || !n) return fail(FLK_E_INVALID, "null argument");
    *n = f->launches;
    return FLK_OK;
}

int flk_profile_enable(flk_field *f, int on) {
    CHECK_HANDLE(f);
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    for (auto &e : f->prof_events) cudaEventDestroy(e);
    f->prof_events.clear();
    for (auto &e : f->prof_rebin) cudaEventDestroy(e);
    f->prof_rebin.clear();
    f->profile = on != 0;
    f->profile_level = on;   // 1 = events around k_step only; 2 = also around the three lazy_update phases
    return FLK_OK;
}

int flk_profile_step_kernel(flk_field *f, double *total_ms, uint64_t *launches) {
    CHECK_HANDLE(f);
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < f->prof_events.size(); i += 2) {
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, f->prof_events[i], f->prof_events[i + 1]));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = f->prof_events.size() / 2;
    return FLK_OK;
}

int flk_profile_rebin(flk_field *f, double ms[3], uint64_t *updates) {
    CHECK_HANDLE(f);
    if (!ms) return fail(FLK_E_INVALID, "ms is null");
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    ms[0] = ms[1] = ms[2] = 0.0;
    uint64_t n = 0;
    for (size_t i = 0; i + 3 < f->prof_rebin.size(); i += 4, ++n)
        for (int k = 0; k < 3; ++k) {
            float t = 0.0f;
            CUDA_TRY(cudaEventElapsedTime(&t, f->prof_rebin[i + k], f->prof_rebin[i + k + 1]));
            ms[k] += t;
        }
    if (updates) *updates = n;
    return FLK_OK;
}

int flk_use_graph(flk_field *f, int on) {
    CHECK_HANDLE(f);
    f->use_graph = on != 0;
    return FLK_OK;
}

int flk_selftest_div2(int device, uint64_t n, uint64_t seed, uint64_t *mismatches) {
    if (!mismatches || n == 0 || n > (1ull << 32)) return fail(FLK_E_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(device));
    unsigned long long *d = nullptr, h = 0;
    CUDA_TRY(cudaMalloc(&d, 8));
    CUDA_TRY(cudaMemset(d, 0, 8));
    launch_selftest_div2(n, seed, d, 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(FLK_E_CUDA, "selftest: %s", cudaGetErrorString(e));
    *mismatches = h;
    return FLK_OK;
}

int flk_host_alloc(void **p, uint64_t bytes) {
    if (!p) return fail(FLK_E_INVALID, "p is null");
    CUDA_TRY(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return FLK_OK;
}

int flk_host_free(void *p) {
    CUDA_TRY(cudaFreeHost(p));
    return FLK_OK;
}

}  // extern "C"

namespace flk {

int report_device_errors(flk_field *f, uint32_t flag) {
    if (flag & ERR_CAPACITY)
        return fail(FLK_E_CAPACITY, "the field holds more objects than its capacity %llu (objects were dropped; the state is invalid)",
                    (unsigned long long)f->cap);
    if (flag & ERR_EXCHANGE)
        return fail(FLK_E_CAPACITY, "a halo/migration message exceeded %u birds or a launch did not cover its range "
                                    "(birds were dropped or not stepped; the state is invalid)", f->s.capx);
    return fail(FLK_E_STATE, "device error bits 0x%x", flag);
}

int upload_injected(flk_field *f, const float *r, uint64_t n) {
    if (n == 0) return FLK_OK;
    if (n > 0xffffffffull) return fail(FLK_E_INVALID, "n_injected exceeds the u32 id space");
    if (n > f->injected_cap) {
        if (f->d_injected) { CUDA_TRY(cudaStreamSynchronize(f->stream)); cudaFree(f->d_injected); f->d_injected = nullptr; }
        CUDA_TRY(cudaMalloc(&f->d_injected, n * 8));
        f->injected_cap = n;
    }
    CUDA_TRY(cudaMemcpyAsync(f->d_injected, r, n * 8, cudaMemcpyHostToDevice, f->stream));
    return FLK_OK;
}

}  // namespace flk
