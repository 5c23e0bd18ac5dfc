// flk_slots.cuh — the fixed-capacity-cell ("slotted") form of the Field2D double buffer.
//
// Every cell owns C record slots, so a successor's place is `cell * C + arrival slot`, known inside the step kernel:
// Bird::step writes the next READ buffer directly and Field2D::lazy_update (flockers/src/model/state.rs:54-56) shrinks
// to one pass over the per-cell counters (k_slot_colscan) instead of scan + scatter + rank over the records.
//
//   READ  buffer   rec[cell*C + slot], id[cell*C + slot], cnt[cell]      slots filled in ARRIVAL order
//                  perm[cell]   the slots of the cell in ascending-id order, 4 bits each (the canonical bag order the
//                               neighbour sums are evaluated in) — written by k_slot_colscan
//                  colcount[c], colchunk[c], coltab[c*stride + m]        which birds a CTA of the step kernel owns
//   WRITE buffer   the other (rec, id, cnt) triple; cnt doubles as the arrival counter (atomicAdd)
//   spill          arrivals beyond C slots of a cell: {rec, id, cell} appended to a list as large as the capacity, so no
//                  object is ever dropped; readers merge them (slow path).  A field that spills more than a handful of
//                  birds is moved to the sorted-runs layout by the host (flk_api.cu: demote_to_runs).
#pragma once
#include "flk_kernels.cuh"

namespace flk {

constexpr int SLOT_C_MAX = 16;         // 4-bit slot numbers in perm
constexpr int SLOT_CHUNK = 128;        // birds per CTA of the slotted step kernel (= FLK_STEP_BS)

// the slotted store of one handle (both buffers)
struct SlotStore {
    float4 *rec[2] = {nullptr, nullptr};
    uint32_t *id[2] = {nullptr, nullptr};
    uint32_t *cnt[2] = {nullptr, nullptr};
    SpillEntry *spill[2] = {nullptr, nullptr};
    uint32_t *spill_cnt = nullptr;     // [2]
    uint64_t *perm = nullptr;          // [ncell]  of the READ buffer
    uint32_t *colcount = nullptr;      // [ncols]  birds per column of the READ buffer (spilled ones included)
    uint32_t *colmax = nullptr;        // [ncols]  largest cell count of the column
    uint32_t *colchunk = nullptr;      // [ncols+1] exclusive prefix of ceil(colcount / SLOT_CHUNK)
    uint32_t *coltab = nullptr;        // [ncols*stride] chunk m of column c starts at row (lo 16 bits), `off` birds into it
    uint32_t *totals = nullptr;        // [4] READ buffer: birds, largest column, largest cell, spilled birds
    uint32_t *ticket = nullptr;        // last-block-done counter of k_slot_colscan
    uint32_t stride = 0;
    uint32_t spill_cap = 0;
    int C = 0;
    int rd = 0;                        // which buffer is READ (host side)
};

// what one launch sees: READ side const, WRITE side mutable
struct SlotArgs {
    const float4 *rrec;
    const uint32_t *rid;
    const uint32_t *rcnt;
    const uint64_t *rperm;
    const uint32_t *colcount;
    const uint32_t *colchunk;
    const uint32_t *coltab;
    const SpillEntry *rspill;
    const uint32_t *rspill_cnt;
    float4 *wrec;
    uint32_t *wid;
    uint32_t *wcnt;
    SpillEntry *wspill;
    uint32_t *wspill_cnt;
    uint32_t stride;
    uint32_t spill_cap;
    int C;
};

inline SlotArgs slot_args(const SlotStore &S) {
    const int r = S.rd, w = 1 - S.rd;
    SlotArgs a;
    a.rrec = S.rec[r]; a.rid = S.id[r]; a.rcnt = S.cnt[r]; a.rperm = S.perm;
    a.colcount = S.colcount; a.colchunk = S.colchunk; a.coltab = S.coltab;
    a.rspill = S.spill[r]; a.rspill_cnt = S.spill_cnt + r;
    a.wrec = S.rec[w]; a.wid = S.id[w]; a.wcnt = S.cnt[w];
    a.wspill = S.spill[w]; a.wspill_cnt = S.spill_cnt + w;
    a.stride = S.stride; a.spill_cap = S.spill_cap; a.C = S.C;
    return a;
}

// Field2D::lazy_update on the slotted store: the WRITE buffer (index 1 - S.rd) becomes READ.  One CTA per column:
// column-local prefix sums of the counters -> coltab / colcount, ascending-id slot order -> perm, the old READ counters
// are zeroed (the new WRITE buffer starts empty); the last CTA to finish builds colchunk and totals.  The caller flips
// S.rd afterwards.
void launch_slot_colscan(const Geom &g, const SlotStore &S, const Store &s, uint32_t ncols, int bump_step, cudaStream_t st);

// Bird::step for the birds of local columns [col_lo, col_hi) of the READ buffer; max_col_birds = a loose bound of the
// largest column's population (sizes the grid; never affects correctness)
void launch_slot_step(const Geom &g, const Params &p, const SlotArgs &S, const Store &s, const StepArgs &a, int col_lo,
                      int col_hi, uint32_t max_col_birds, cudaStream_t st);

// bulk set_object_location into the WRITE buffer (state.rs:49)
void launch_slot_ingest(const Geom &g, const SlotArgs &S, const Store &s, uint32_t n, const uint32_t *id, const float2 *loc,
                        const float2 *ld, int owned_only, cudaStream_t st);
void launch_slot_ingest_dense(const Geom &g, const SlotArgs &S, const Store &s, uint32_t n, uint32_t id0, const float4 *rec4,
                              const uint32_t *ids /* null: id0 + j */, cudaStream_t st);
// strips: arrivals from the neighbours' messages join the WRITE buffer
void launch_slot_unpack(const Geom &g, const SlotArgs &S, const Store &s, cudaStream_t st);

// READ buffer -> the sorted-runs arrays (s.rec, s.ids) given s.cell_start = exclusive scan of the READ counters
void launch_slot_materialize(const Geom &g, const SlotArgs &S, const Store &s, uint32_t ncell, cudaStream_t st);
// READ record of object `id` -> out[id - id0]
void launch_slot_gather_dense(const SlotArgs &S, uint32_t ncell, uint32_t id0, uint32_t n, float4 *out, uint32_t *outside,
                              cudaStream_t st);
// sorted-runs READ buffer -> statistics for the layout decision: out[0] = largest cell, out[1] = birds beyond C per cell
void launch_cell_stats(const Store &s, uint32_t ncell, int C, uint32_t *out, cudaStream_t st);

}  // namespace flk
