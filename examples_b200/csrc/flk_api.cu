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
    const uint64_t mp = ((m + 4095) / 4096) * 4096;   // padded to the widest scan tile (k_scan_onepass uses whole-tile vector accesses)
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
    CUDA_TRY(cudaMalloc(&s.scan_state, (nblk + 2) * 8));
    CUDA_TRY(cudaMemsetAsync(s.scan_state, 0, (nblk + 2) * 8, f->stream));
    if (f->g.mode == 0) {   // column-aligned k_step launches (single-GPU handles)
        CUDA_TRY(cudaMalloc(&s.colchunk, 2 * ((uint64_t)f->g.ncols + 1) * 4));   // chunk table + seam-bird table
        CUDA_TRY(cudaMemsetAsync(s.colchunk, 0, 2 * ((uint64_t)f->g.ncols + 1) * 4, f->stream));
    }
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
    // the slot forms of the double buffer are allocated when they are asked for (flk_set_layout / FLK_LAYOUT)
    if (const char *e = getenv("FLK_LAYOUT")) {
        if (!strcmp(e, "runs")) f->layout_pref = 0;
        else if (!strcmp(e, "slots")) f->layout_pref = 1;
        else if (!strcmp(e, "staged")) f->layout_pref = 2;
        f->layout_forced = f->layout_pref >= 1;
    }
    if (f->layout_pref >= 1) return slots_alloc(f);
    return FLK_OK;
}

// ---- the fixed-capacity-cell layout (flk_slots.cuh): allocation and the switches between the two layouts -----------
static int slot_env_C() {
    const char *e = getenv("FLK_SLOT_C");
    const int c = e ? atoi(e) : 16;
    return (c == 4 || c == 8 || c == 12 || c == 16) ? c : 16;
}

int slots_alloc(flk_field *f) {
    Geom &g = f->g;
    if (f->slot_capable) return FLK_OK;
    f->layout = 0;
    // single-GPU handles with Bird-sized objects (strips and tiles keep the classic form)
    if (g.mode != 0 || f->s.pw != 0) return FLK_OK;
    const int C = slot_env_C();
    SlotStore &S = f->S;
    const uint64_t ncell = f->ncell, ncols = (uint64_t)g.ncols;
    const uint64_t per_col = (f->cap + (uint64_t)(g.mode == 0 ? g.mx : g.nown) - 1) / (uint64_t)(g.mode == 0 ? g.mx : g.nown);
    uint64_t stride = 8 * ((per_col + SLOT_CHUNK - 1) / SLOT_CHUNK) + 8;
    if (stride < 16) stride = 16;
    if (stride > 65536) stride = 65536;
    S.C = C; S.rd = 0; S.stride = (uint32_t)stride; S.spill_cap = (uint32_t)f->cap;
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaMalloc(&S.rec[b], ncell * C * sizeof(float4)));
        CUDA_TRY(cudaMalloc(&S.id[b], ncell * C * 4));
        CUDA_TRY(cudaMalloc(&S.cnt[b], (ncell + 1) * 4));
        CUDA_TRY(cudaMalloc(&S.spill[b], (uint64_t)S.spill_cap * sizeof(SpillEntry)));
        CUDA_TRY(cudaMemsetAsync(S.cnt[b], 0, (ncell + 1) * 4, f->stream));
        // slots are written 16 and 4 bytes at a time, sectors partially: start from defined memory
        CUDA_TRY(cudaMemsetAsync(S.rec[b], 0, ncell * C * sizeof(float4), f->stream));
        CUDA_TRY(cudaMemsetAsync(S.id[b], 0, ncell * C * 4, f->stream));
    }
    CUDA_TRY(cudaMalloc(&S.spill_cnt, 8));
    CUDA_TRY(cudaMalloc(&S.perm, ncell * 8));
    CUDA_TRY(cudaMalloc(&S.colcount, ncols * 4));
    CUDA_TRY(cudaMalloc(&S.colmax, ncols * 4));
    CUDA_TRY(cudaMalloc(&S.colchunk, (ncols + 1) * 4));
    CUDA_TRY(cudaMalloc(&S.coltab, ncols * stride * 4));
    CUDA_TRY(cudaMalloc(&S.totals, 16));
    CUDA_TRY(cudaMalloc(&S.ticket, 4));
    CUDA_TRY(cudaMalloc(&f->d_stats, 8));
    CUDA_TRY(cudaHostAlloc(&f->h_totals, 16, cudaHostAllocDefault));
    memset(f->h_totals, 0, 16);
    CUDA_TRY(cudaMemsetAsync(S.spill_cnt, 0, 8, f->stream));
    CUDA_TRY(cudaMemsetAsync(S.colcount, 0, ncols * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(S.colmax, 0, ncols * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(S.colchunk, 0, (ncols + 1) * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(S.totals, 0, 16, f->stream));
    CUDA_TRY(cudaMemsetAsync(S.ticket, 0, 4, f->stream));
    // the slot-staged WRITE buffer of the sorted-runs layout shares buffer 0 (only one of the two uses is active)
    Store &s = f->s;
    s.slotw = 0; s.sw_C = C; s.sw_rec = S.rec[0]; s.sw_id = S.id[0]; s.sw_spill = S.spill[0]; s.sw_spill_cnt = S.spill_cnt;
    s.sw_spill_cap = S.spill_cap;
    CUDA_TRY(cudaMalloc(&s.sw_totals, 8));
    CUDA_TRY(cudaMemsetAsync(s.sw_totals, 0, 8, f->stream));
    CUDA_TRY(cudaHostAlloc(&f->h_sw, 8, cudaHostAllocDefault));
    memset(f->h_sw, 0, 8);
    f->slot_capable = true;
    return FLK_OK;
}

static void slots_free(flk_field *f) {
    SlotStore &S = f->S;
    for (int b = 0; b < 2; ++b) { cudaFree(S.rec[b]); cudaFree(S.id[b]); cudaFree(S.cnt[b]); cudaFree(S.spill[b]); }
    cudaFree(S.spill_cnt); cudaFree(S.perm); cudaFree(S.colcount); cudaFree(S.colmax); cudaFree(S.colchunk);
    cudaFree(S.coltab); cudaFree(S.totals); cudaFree(S.ticket);
    if (f->d_stats) cudaFree(f->d_stats);
    if (f->h_totals) cudaFreeHost(f->h_totals);
    if (f->h_sw) cudaFreeHost(f->h_sw);
    if (f->s.sw_totals) cudaFree(f->s.sw_totals);
    for (int b = 0; b < 2; ++b) if (f->sgraph[b]) cudaGraphExecDestroy(f->sgraph[b]);
}

// the parameters the slotted step kernel's tile path is written for (anything else runs on the sorted-runs layout)
bool slots_params_ok(const flk_field *f) {
    const Geom &g = f->g;
    const Params &p = f->p;
    return p.relax && p.k == 1 && p.radius > 0.0f && g.toroidal && g.mx >= 6 && g.my >= 6 && g.d >= 0x1p-20f && g.d <= 4096.0f;
}

// loose bound of the largest column's population: the largest column of a recent READ buffer (copied back after every
// lazy_update, possibly a few ticks old) or the mean, plus 3 %.  It only sizes the step kernel's grid.
uint32_t slot_col_bound(const flk_field *f) {
    const uint64_t cols = (uint64_t)(f->g.mode == 0 ? f->g.mx : f->g.nown);
    uint64_t est = (f->n_read + cols - 1) / (cols ? cols : 1);
    if (f->h_totals && f->h_totals[1] > est) est = f->h_totals[1];
    return (uint32_t)(est + est / 32 + 128);
}

// Field2D::lazy_update on the slotted store: the WRITE buffer becomes READ
int slots_publish(flk_field *f, int bump_step) {
    launch_slot_colscan(f->g, f->S, f->s, (uint32_t)f->g.ncols, bump_step, f->stream);
    f->S.rd ^= 1;
    CUDA_TRY(cudaMemcpyAsync(f->h_totals, f->S.totals, 16, cudaMemcpyDeviceToHost, f->stream));
    f->launches += 1;
    f->view_valid = false;
    CUDA_TRY(cudaGetLastError());
    return FLK_OK;
}

// read-side calls work on the sorted runs: build them from the slotted READ buffer when they are stale
int ensure_view(flk_field *f) {
    if (f->layout != 1 || f->view_valid) return FLK_OK;
    CUDA_TRY(cudaMemcpyAsync(f->s.cnt, f->S.cnt[f->S.rd], (uint64_t)f->ncell * 4, cudaMemcpyDeviceToDevice, f->stream));
    launch_scan(f->s, f->ncell + 1, f->stream);            // -> cell_start; leaves s.cnt zeroed again
    launch_slot_materialize(f->g, slot_args(f->S), f->s, f->ncell, f->stream);
    f->launches += 2;
    CUDA_TRY(cudaGetLastError());
    f->view_valid = true;
    return FLK_OK;
}

// leave the slotted layout (too many spilled birds, or parameters its step kernel is not written for).  Only between a
// lazy_update and the next write: the slotted WRITE buffer must be empty.
int demote_to_runs(flk_field *f) {
    if (f->layout != 1) return FLK_OK;
    if (f->w_count != 0) return fail(FLK_E_STATE, "layout change with a non-empty WRITE buffer (call flk_lazy_update first)");
    int rc = ensure_view(f);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(f->S.cnt[f->S.rd], 0, ((uint64_t)f->ncell + 1) * 4, f->stream));
    CUDA_TRY(cudaMemsetAsync(f->S.spill_cnt, 0, 8, f->stream));
    f->layout = 0;
    f->view_valid = false;
    f->graph_valid = false;
    return FLK_OK;
}

// switch the WRITE buffer of the sorted-runs layout between its two forms (only while it is empty)
int set_slotw(flk_field *f, int on) {
    if (on && !f->slot_capable) on = 0;
    if (f->s.slotw == on) return FLK_OK;
    if (f->w_count != 0) return fail(FLK_E_STATE, "WRITE buffer form change with a non-empty WRITE buffer");
    f->s.slotw = on;
    f->graph_valid = false;
    return FLK_OK;
}

// An upload was just published through the classic WRITE buffer: choose how the following ticks write.
//   slot-staged WRITE buffer (default where it fits): arrivals go straight to their cell's slots, lazy_update = scan + k_compact
//   classic WRITE buffer: any density (clustered flocks overflow the slots of their cells)
//   slotted READ layout: only on request (flk_set_layout(1))
int decide_write_mode(flk_field *f) {
    // automatic = classic: on B200 the two slot forms measured within 2 % of it (form 2) or slower (form 1), DESIGN.md §3
    if (!f->slot_capable || f->g.mode != 0 || f->w_count != 0 || f->layout != 0 || f->layout_pref <= 0) return set_slotw(f, 0);
    if (f->n_read == 0) return FLK_OK;
    uint32_t st[2] = {0, 0};
    CUDA_TRY(cudaMemsetAsync(f->d_stats, 0, 8, f->stream));
    launch_cell_stats(f->s, f->ncell, f->S.C, f->d_stats, f->stream);
    CUDA_TRY(cudaMemcpyAsync(st, f->d_stats, 8, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->launches += 1;
    const uint64_t limit = f->n_read / 4096 > 64 ? f->n_read / 4096 : 64;
    const bool fits = st[1] <= limit || f->layout_forced;
    if (!fits) return set_slotw(f, 0);
    if (f->layout_pref == 1) {
        if (f->g.toroidal && slots_params_ok(f)) return maybe_promote_to_slots(f);
        return set_slotw(f, 0);
    }
    return set_slotw(f, 1);
}

// enter the slotted layout when the freshly published sorted-runs READ buffer fits it (single GPU: decided with one small
// synchronising read right after an upload was published)
int maybe_promote_to_slots(flk_field *f) {
    if (!f->slot_capable || f->layout != 0 || f->w_count != 0 || f->layout_pref != 1 || !f->g.toroidal || !slots_params_ok(f))
        return FLK_OK;
    if (f->n_read == 0) return FLK_OK;
    if (int rc = set_slotw(f, 0)) return rc;                   // buffer 0 is about to hold the slotted READ / WRITE buffers
    launch_slot_ingest_dense(f->g, slot_args(f->S), f->s, (uint32_t)f->n_read, 0, reinterpret_cast<const float4 *>(f->s.rec),
                             f->s.ids, f->stream);
    f->launches += 2;
    int rc = slots_publish(f, 0);
    if (rc) return rc;
    f->layout = 1;
    f->view_valid = true;      // rec / ids / cell_start still describe this very READ buffer
    f->graph_valid = false;
    return FLK_OK;
}

void field_free(flk_field *f) {
    Store &s = f->s;
    cudaFree(s.rec); cudaFree(s.ids); cudaFree(s.cell_start); cudaFree(s.trec); cudaFree(s.tid);
    cudaFree(s.tkey); cudaFree(s.tslot); cudaFree(s.cnt); cudaFree(f->scratch); cudaFree(s.scan_state);
    cudaFree(s.step_dev); cudaFree(s.n_read_dev); cudaFree(s.err_flag);
    if (s.colchunk) cudaFree(s.colchunk);
    if (s.pay) cudaFree(s.pay);
    if (s.tpay) cudaFree(s.tpay);
    if (f->d_paystage) cudaFree(f->d_paystage);
    if (f->d_injected) cudaFree(f->d_injected);
    if (f->h_dense) cudaFreeHost(f->h_dense);
    if (f->h_err) cudaFreeHost(f->h_err);
    slots_free(f);
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
    k += f->g.wplus1;                                   // [EXT] E3 alternative (flk_set_engine_semantics)
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
    // a bulk upload into an empty WRITE buffer goes through the classic form (any density); the form of the following
    // ticks is chosen when it is published (decide_write_mode)
    if (f->layout == 0 && f->s.slotw && f->w_count == 0 && f->g.mode == 0) { int rc = set_slotw(f, 0); if (rc) return rc; }
    if (f->layout == 1) launch_slot_ingest(f->g, slot_args(f->S), f->s, (uint32_t)n, d_id, d_loc, d_ld, 1, f->stream);
    else launch_ingest(f->g, f->s, (uint32_t)n, d_id, d_loc, d_ld, d_pay, (uint32_t)f->w_count, 1, f->stream);
    f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    f->w_count += n;
    f->uploaded = true;
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
    if (f->layout == 1) {   // slotted: one pass over the cell counters (profile: all of it counted as "scan")
        if (int rc = mark()) return rc;
        if (int rc = slots_publish(f, bump_step)) return rc;
        for (int k = 0; k < 3; ++k) if (int rc = mark()) return rc;
        f->n_read = f->w_count;
        f->w_count = 0;
        if (f->uploaded && !f->capturing && f->g.mode == 0) {
            // an upload was published: is the slotted layout still the right one? (a clustered population spills)
            f->uploaded = false;
            CUDA_TRY(cudaStreamSynchronize(f->stream));
            const uint64_t limit = f->n_read / 4096 > 64 ? f->n_read / 4096 : 64;
            if (!f->layout_forced && f->h_totals[3] > limit) return demote_to_runs(f);
        }
        return FLK_OK;
    }
    if (f->s.slotw) {   // slot-staged WRITE buffer: scan + one pass (profile: the pass is counted as "scatter")
        if (int rc = mark()) return rc;
        launch_scan(f->s, m, f->stream);
        if (int rc = mark()) return rc;
        CUDA_TRY(cudaMemsetAsync(f->s.sw_totals, 0, 8, f->stream));
        launch_compact(f->s, f->ncell, bump_step, f->stream, (uint32_t)f->g.ncols, (uint32_t)f->g.dh);
        CUDA_TRY(cudaMemsetAsync(f->s.sw_spill_cnt, 0, 4, f->stream));
        if (int rc = mark()) return rc;
        if (int rc = mark()) return rc;
        CUDA_TRY(cudaMemcpyAsync(f->h_sw, f->s.sw_totals, 8, cudaMemcpyDeviceToHost, f->stream));
        f->launches += 2;
        CUDA_TRY(cudaGetLastError());
        f->n_read = f->w_count;
        f->w_count = 0;
        const uint64_t limit = f->n_read / 4096 > 64 ? f->n_read / 4096 : 64;
        if (f->uploaded && !f->capturing) {       // births were published: look at the spill count now
            f->uploaded = false;
            CUDA_TRY(cudaStreamSynchronize(f->stream));
        }
        // the spill count of a recent tick (copied back asynchronously): a field that grew dense leaves the slots
        if (!f->capturing && !f->layout_forced && f->h_sw[1] > limit) return set_slotw(f, 0);
        return FLK_OK;
    }
    static const uint64_t small_max = [] {
        const char *e = getenv("FLK_REBIN_SMALL");
        return e ? strtoull(e, nullptr, 10) : (uint64_t)SMALL_REBIN_MAX;
    }();
    if (f->g.mode == 0 && f->w_count <= small_max && m <= 16 * small_max) {
        // a small field: one launch, one CTA (profile: all of it counted as "scan")
        if (int rc = mark()) return rc;
        launch_rebin_small(f->s, f->ncell, (uint32_t)f->w_count, bump_step, f->stream, (uint32_t)f->g.ncols, (uint32_t)f->g.dh);
        for (int k = 0; k < 3; ++k) if (int rc = mark()) return rc;
        f->launches += 1;
    } else {
        if (int rc = mark()) return rc;
        launch_scan(f->s, m, f->stream);
        if (int rc = mark()) return rc;
        launch_scatter(f->s, (uint32_t)f->w_count, f->g.mode != 0, f->stream);
        if (int rc = mark()) return rc;
        launch_rank(f->s, (uint32_t)f->w_count, f->ncell, bump_step, f->stream, (uint32_t)f->g.ncols, (uint32_t)f->g.dh);
        if (int rc = mark()) return rc;
        f->launches += 1 + ((f->g.mode != 0 || f->w_count) ? 1 : 0) + 1;
    }
    CUDA_TRY(cudaGetLastError());
    f->n_read = f->w_count;  // single GPU: every WRITE entry is kept
    f->w_count = 0;
    if (f->uploaded && !f->capturing && f->g.mode == 0) {
        f->uploaded = false;
        return decide_write_mode(f);
    }
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
    a.nseam = 0; a.gen_threads = 0;
    const bool prof = f->profile && !f->capturing;
    if (prof) {
        cudaEvent_t e0, e1;
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        f->prof_events.push_back(e0); f->prof_events.push_back(e1);
        CUDA_TRY(cudaEventRecord(e0, f->stream));
    }
    if (f->layout == 1) launch_slot_step(f->g, f->p, slot_args(f->S), f->s, a, 0, f->g.ncols, slot_col_bound(f), f->stream);
    else launch_step(f->g, f->p, f->s, a, (uint32_t)f->n_read, f->stream);
    if (prof) CUDA_TRY(cudaEventRecord(f->prof_events.back(), f->stream));
    if (f->n_read) f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    f->w_count += f->n_read;
    return FLK_OK;
}

// flk_run on the slotted layout.  The two buffers swap roles every tick, so a captured tick only replays from the same
// parity: the graph holds TWO ticks (one per parity of the READ buffer, `step_dev` bumped by each), an odd tick left over
// is launched directly.
int run_slots(flk_field *f, uint64_t first_step, uint64_t n_steps, uint64_t seed) {
    uint64_t t = 0;
    if (!f->profile && f->use_graph && n_steps >= 8) {
        const int par = f->S.rd;
        if (!f->sgraph_valid[par] || f->sgraph_n[par] != f->n_read || f->sgraph_seed[par] != seed) {
            if (f->sgraph[par]) { cudaGraphExecDestroy(f->sgraph[par]); f->sgraph[par] = nullptr; }
            f->sgraph_valid[par] = false;
            cudaGraph_t graph;
            const uint64_t launches_before = f->launches;
            CUDA_TRY(cudaStreamBeginCapture(f->stream, cudaStreamCaptureModeThreadLocal));
            f->capturing = true;
            int rc = FLK_OK;
            for (int k = 0; k < 2 && !rc; ++k) {
                rc = enqueue_step(f, 0, 1, seed, nullptr, 0);
                if (!rc) rc = enqueue_lazy_update(f, 1);
            }
            f->capturing = false;
            cudaError_t ce = cudaStreamEndCapture(f->stream, &graph);
            f->sgraph_launches[par] = f->launches - launches_before;
            f->launches = launches_before;
            if (rc) { if (ce == cudaSuccess) cudaGraphDestroy(graph); return rc; }
            CUDA_TRY(ce);
            ce = cudaGraphInstantiate(&f->sgraph[par], graph, 0);
            cudaGraphDestroy(graph);
            CUDA_TRY(ce);
            f->sgraph_valid[par] = true; f->sgraph_n[par] = f->n_read; f->sgraph_seed[par] = seed;
        }
        CUDA_TRY(cudaMemcpyAsync(f->s.step_dev, &first_step, 8, cudaMemcpyHostToDevice, f->stream));
        for (; t + 2 <= n_steps; t += 2) CUDA_TRY(cudaGraphLaunch(f->sgraph[par], f->stream));
        f->launches += f->sgraph_launches[par] * (t / 2);
        f->view_valid = false;
    }
    for (; t < n_steps; ++t) {
        int rc = enqueue_step(f, first_step + t, 0, seed, nullptr, 0);
        if (!rc) rc = enqueue_lazy_update(f, 0);
        if (rc) return rc;
    }
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
    if (f->slot_capable) {
        for (int b = 0; b < 2; ++b) CUDA_TRY(cudaMemsetAsync(f->S.cnt[b], 0, ((uint64_t)f->ncell + 1) * 4, f->stream));
        CUDA_TRY(cudaMemsetAsync(f->S.spill_cnt, 0, 8, f->stream));
        CUDA_TRY(cudaMemsetAsync(f->S.colcount, 0, (uint64_t)f->g.ncols * 4, f->stream));
        CUDA_TRY(cudaMemsetAsync(f->S.colchunk, 0, ((uint64_t)f->g.ncols + 1) * 4, f->stream));
    }
    f->layout = 0; f->view_valid = false; f->uploaded = false;
    f->s.slotw = 0;
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
    f->sgraph_valid[0] = f->sgraph_valid[1] = false;
    // the slotted step kernel is written for the shipped query (relax, 3x3 window); anything else steps on the sorted runs
    if (f->layout == 1 && !slots_params_ok(f)) {
        rc = demote_to_runs(f);
        if (rc) { f->p = old; return rc; }
    }
    return FLK_OK;
}

int flk_set_engine_semantics(flk_field *f, int window_plus1) {
    CHECK_HANDLE(f);
    if (window_plus1 != 0 && window_plus1 != 1) return fail(FLK_E_INVALID, "window_plus1 must be 0 or 1");
    const int old = f->g.wplus1;
    f->g.wplus1 = window_plus1;
    int rc = update_window(f);
    if (rc) { f->g.wplus1 = old; update_window(f); return rc; }
    f->graph_valid = false;
    f->sgraph_valid[0] = f->sgraph_valid[1] = false;
    if (f->layout == 1 && !slots_params_ok(f)) return demote_to_runs(f);
    return FLK_OK;
}

int flk_set_math_mode(flk_field *f, int mode) {
    CHECK_HANDLE(f);
    if (mode != 0 && mode != 1) return fail(FLK_E_INVALID, "math mode must be 0 (exact) or 1 (fast)");
    f->math_mode = mode;
    f->graph_valid = false;
    f->sgraph_valid[0] = f->sgraph_valid[1] = false;
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
    int rc = flush_pending(f);
    if (rc) return rc;
    if (f->g.mode != 0) return dist_set_objects(f, n, id, loc, last_d, static_cast<const uint32_t *>(payload));
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
    if (f->layout == 1) {
        rc = run_slots(f, first_step, n_steps, seed);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
        return FLK_OK;
    }
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
    if (int rc = ensure_view(f)) return rc;
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
    if (f->layout == 0 && f->s.slotw && f->w_count == 0) { int rc2 = set_slotw(f, 0); if (rc2) return rc2; }
    if (f->layout == 1) launch_slot_ingest_dense(f->g, slot_args(f->S), f->s, (uint32_t)n, id0, d_rec, nullptr, f->stream);
    else launch_ingest_dense(f->g, f->s, (uint32_t)n, id0, d_rec, (uint32_t)f->w_count, f->stream);
    f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    f->w_count += n;
    f->uploaded = true;
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
    if (f->layout == 1) launch_slot_gather_dense(slot_args(f->S), f->ncell, id0, (uint32_t)n, d_out, d_cnt, f->stream);
    else launch_gather_dense(f->s, (uint32_t)f->n_read, id0, (uint32_t)n, d_out, d_cnt, f->stream);
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
    if (int rc = ensure_view(f)) return rc;
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
    if (int rc = ensure_view(f)) return rc;
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
    if (int rc = ensure_view(f)) return rc;
    if (cell_start)
        CUDA_TRY(cudaMemcpyAsync(cell_start, f->s.cell_start, ((uint64_t)f->ncell + 1) * 4, cudaMemcpyDeviceToHost, f->stream));
    if (sorted_ids && f->n_read)
        CUDA_TRY(cudaMemcpyAsync(sorted_ids, f->s.ids, f->n_read * 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    return FLK_OK;
}

static int query_batch_impl(flk_field *f, uint64_t nq, const float *loc, float dist, int relax, uint64_t *offsets,
                            uint32_t *ids, uint64_t cap, int objects);

int flk_query_batch(flk_field *f, uint64_t nq, const float *loc, float dist, int relax, uint64_t *offsets,
                    uint32_t *ids, uint64_t cap) {
    return query_batch_impl(f, nq, loc, dist, relax, offsets, ids, cap, 0);
}

int flk_query_batch_objects(flk_field *f, uint64_t nq, const float *loc, float dist, int relax, uint64_t *offsets,
                            flk_bird *objects, uint64_t cap) {
    static_assert(sizeof(flk_bird) == 20, "flk_bird is Bird's 20-byte wire form");
    return query_batch_impl(f, nq, loc, dist, relax, offsets, reinterpret_cast<uint32_t *>(objects), cap, 1);
}

// `ids` receives ids (objects = 0) or 5-word objects (objects = 1); cap counts entries, not words
static int query_batch_impl(flk_field *f, uint64_t nq, const float *loc, float dist, int relax, uint64_t *offsets,
                            uint32_t *ids, uint64_t cap, int objects) {
    CHECK_HANDLE(f);
    const uint64_t words = objects ? 5 : 1;
    if (!offsets || (nq && !loc)) return fail(FLK_E_INVALID, "null array");
    if (nq > 0x7fffffffull) return fail(FLK_E_INVALID, "too many queries");
    offsets[0] = 0;
    if (nq == 0) return FLK_OK;
    if (int rc0 = ensure_view(f)) return rc0;
    if (dist > 0.0f) {
        const int k = (int)std::floor(dist / f->g.d) + f->g.wplus1;
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
        Q_TRY(cudaMalloc(&d_out, offsets[nq] * 4 * words));
        Q_TRY(cudaMemcpyAsync(d_off, offsets, (nq + 1) * 8, cudaMemcpyHostToDevice, f->stream));
        launch_query_fill(f->g, f->s, (uint32_t)nq, d_loc, dist, relax, d_off, d_out, f->stream, objects);
        f->launches += 1;
        Q_TRY(cudaGetLastError());
        Q_TRY(cudaMemcpyAsync(ids, d_out, offsets[nq] * 4 * words, cudaMemcpyDeviceToHost, f->stream));
        Q_TRY(cudaStreamSynchronize(f->stream));
    }
done:
#undef Q_TRY
    cudaFree(d_loc); cudaFree(d_counts); cudaFree(d_off); cudaFree(d_out);
    return rc;
}

static int query_one(flk_field *f, float x, float y, float dist, int relax, uint32_t *ids, uint32_t cap, uint32_t *n,
                     int objects = 0) {
    CHECK_HANDLE(f);
    if (!n) return fail(FLK_E_INVALID, "n is null");
    const float loc[2] = {x, y};
    uint64_t off[2] = {0, 0};
    const size_t words = objects ? 5 : 1;
    // first try with the caller's buffer; on overflow report the full size with the first `cap` entries
    int rc = query_batch_impl(f, 1, loc, dist, relax, off, ids, cap, objects);
    *n = (uint32_t)off[1];
    if (rc == FLK_E_CAPACITY && off[1] > cap) {
        std::vector<uint32_t> tmp(off[1] * words);
        int rc2 = query_batch_impl(f, 1, loc, dist, relax, off, tmp.data(), off[1], objects);
        if (rc2) return rc2;
        if (ids && cap) memcpy(ids, tmp.data(), (size_t)cap * 4 * words);
        return fail(FLK_E_CAPACITY, "query result has %u entries, buffer holds %u", *n, cap);
    }
    return rc;
}

int flk_get_neighbors_within_relax_distance_objects(flk_field *f, float x, float y, float dist, flk_bird *objects,
                                                    uint32_t cap, uint32_t *n) {
    return query_one(f, x, y, dist, 1, reinterpret_cast<uint32_t *>(objects), cap, n, 1);
}

int flk_get_neighbors_within_distance_objects(flk_field *f, float x, float y, float dist, flk_bird *objects, uint32_t cap,
                                              uint32_t *n) {
    return query_one(f, x, y, dist, 0, reinterpret_cast<uint32_t *>(objects), cap, n, 1);
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

int flk_set_layout(flk_field *f, int layout) {
    CHECK_HANDLE(f);
    if (layout < -1 || layout > 2)
        return fail(FLK_E_INVALID, "layout must be -1 (automatic), 0 (sorted runs, classic WRITE buffer), 1 (slots) or 2 (sorted runs, slot-staged WRITE buffer)");
    if (layout >= 1 && !f->slot_capable) { int rc = slots_alloc(f); if (rc) return rc; }
    if (layout >= 1 && !f->slot_capable)
        return fail(FLK_E_UNSUPPORTED, "this handle cannot use the slot forms (strips, tiles or a payload)");
    f->layout_pref = layout;
    if (layout != 1 && f->layout == 1) { int rc = demote_to_runs(f); if (rc) return rc; }
    if (f->w_count != 0 || f->g.mode != 0) return FLK_OK;      // takes effect at the next lazy_update after an upload
    if (layout <= 0) return set_slotw(f, 0);
    return f->n_read ? decide_write_mode(f) : FLK_OK;          // an empty field: decided when the first upload is published
}

int flk_get_layout(flk_field *f, int32_t *layout, int32_t *slots_per_cell, uint32_t totals[4]) {
    CHECK_HANDLE(f);
    if (layout) *layout = f->layout == 1 ? 1 : (f->s.slotw ? 2 : 0);
    if (slots_per_cell) *slots_per_cell = f->slot_capable ? f->S.C : 0;
    if (totals) {
        totals[0] = totals[1] = totals[2] = totals[3] = 0;
        if (f->layout == 1) {
            CUDA_TRY(cudaMemcpyAsync(f->h_totals, f->S.totals, 16, cudaMemcpyDeviceToHost, f->stream));
            CUDA_TRY(cudaStreamSynchronize(f->stream));
            memcpy(totals, f->h_totals, 16);
        } else if (f->s.slotw) {
            uint32_t st[2] = {0, 0};
            CUDA_TRY(cudaMemsetAsync(f->d_stats, 0, 8, f->stream));
            launch_cell_stats(f->s, f->ncell, f->S.C, f->d_stats, f->stream);
            CUDA_TRY(cudaMemcpyAsync(st, f->d_stats, 8, cudaMemcpyDeviceToHost, f->stream));
            CUDA_TRY(cudaMemcpyAsync(f->h_sw, f->s.sw_totals, 8, cudaMemcpyDeviceToHost, f->stream));
            CUDA_TRY(cudaStreamSynchronize(f->stream));
            totals[0] = (uint32_t)f->n_read; totals[2] = st[0]; totals[3] = f->h_sw[1];
        }
    }
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
