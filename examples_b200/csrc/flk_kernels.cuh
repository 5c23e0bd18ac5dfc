// flk_kernels.cuh — launch interface between the C-ABI layer (flk_api.cu) and the kernels (flk_kernels.cu).
#pragma once
#include "flk_device.cuh"

namespace flk {

constexpr uint32_t KEY_INVALID = 0xFFFFFFFFu;

// an arrival beyond the slots of its cell (slot-staged WRITE buffer / slotted layout): appended to a list
struct __align__(16) SpillEntry {
    float4 rec;
    uint32_t id, cell, pad0, pad1;
};

// All device arrays of one handle (DESIGN.md §data layout).  READ buffer = (rec, ids, cell_start), sorted by
// (cell, id).  WRITE buffer = (trec, tid, tkey, tslot) + per-cell counters cnt, in arrival order.
struct Store {
    Rec *rec;              // READ records
    uint32_t *ids;         // READ ids
    uint32_t *cell_start;  // READ: ncell+1 offsets
    Rec *trec;             // WRITE records
    uint32_t *tid;         // WRITE ids
    uint32_t *tkey;        // WRITE local cell key (KEY_INVALID = not kept on this GPU)
    uint32_t *tslot;       // WRITE arrival slot inside its cell
    uint32_t *cnt;         // WRITE per-cell counters (ncell+1)
    uint4 *stage;          // staging for the (cell,id) ranking: {id, WRITE index, cell key, -} in cell order
    unsigned long long *scan_state;   // one-pass scan: ticket, done counter, one status word per tile (kept zeroed between launches)
    uint64_t *step_dev;    // tick counter used by graph replays
    uint32_t *n_read_dev;  // number of READ objects (device copy)
    uint32_t *err_flag;    // device-side error bits (ERR_*)
    // generic Field2D<O> payload: pw 32-bit words per object carried along with its record (0 = none)
    uint32_t *pay;         // READ payload [cap*pw]
    uint32_t *tpay;        // WRITE payload
    uint32_t pw;
    uint32_t cap;          // READ / stepped-WRITE capacity
    // strip decomposition only: fixed-capacity exchange buffers, XHdr | Rec[capx] | id[capx]
    uint32_t capx;
    uint8_t *sendL, *sendR, *recvL, *recvR;
    // tiles (mode 2) add the row direction; received birds occupy WRITE slots [cap, cap + nrecv*capx)
    uint8_t *sendD, *sendU, *recvD, *recvU;
    uint32_t nrecv;        // 2 strips, 4 tiles
    // slot-staged WRITE buffer (slotw != 0): an arrival goes to sw_rec / sw_id [key * sw_C + arrival slot] instead of
    // trec / tid / tkey / tslot [w], so that lazy_update is scan + ONE pass (k_compact) instead of scan + scatter + rank;
    // arrivals beyond sw_C slots of a cell are appended to sw_spill (as large as the capacity: nothing is ever dropped)
    int slotw;
    int sw_C;
    float4 *sw_rec;
    uint32_t *sw_id;
    SpillEntry *sw_spill;
    uint32_t *sw_spill_cnt;
    uint32_t sw_spill_cap;
    uint32_t *sw_totals;   // [2] of the last k_compact: largest cell, spilled arrivals
    // single-GPU handles: colchunk[c] = number of k_step CTAs (chunks of FLK_STEP_BS birds) that the tile columns before c
    // need, colchunk[ncols] = all of them; colchunk[ncols + 1 + c] = seam-row birds of the tile columns before c (written with
    // cell_start by lazy_update: k_rank / k_compact / k_rebin_small); null = not maintained
    uint32_t *colchunk;
};

constexpr uint32_t ERR_CAPACITY = 1u;   // more objects than `cap` after an exchange
constexpr uint32_t ERR_EXCHANGE = 2u;   // more border/migrating birds than `capx`

struct XHdr {
    uint32_t count, pad0, pad1, pad2;
};
// message layout: XHdr | Rec[capx] | id[capx] | payload[capx * pw] (Field2D<O>: whatever else the object carries)
__host__ __device__ inline size_t xbuf_bytes(uint32_t capx, uint32_t pw = 0) { return sizeof(XHdr) + (size_t)capx * (20 + 4 * (size_t)pw); }
__host__ __device__ inline uint32_t *xbuf_pay(uint8_t *b, uint32_t capx) {
    return reinterpret_cast<uint32_t *>(b + sizeof(XHdr) + (size_t)capx * 20);
}
__host__ __device__ inline Rec *xbuf_rec(uint8_t *b) { return reinterpret_cast<Rec *>(b + sizeof(XHdr)); }
__host__ __device__ inline uint32_t *xbuf_id(uint8_t *b, uint32_t capx) {
    return reinterpret_cast<uint32_t *>(b + sizeof(XHdr) + (size_t)capx * 16);
}

struct StepArgs {
    uint32_t cell_lo, cell_hi;  // step birds of cells [cell_lo, cell_hi)
    uint64_t step;              // tick number (ignored when use_step_dev)
    int use_step_dev;
    uint64_t seed;
    const float2 *injected;     // device, indexed by id, or null
    uint32_t n_injected;        // pairs in `injected`
    uint32_t out_base;          // WRITE index of READ index 0
    int math_mode;              // 0 exact, 1 fast
    int identity;               // 1 = re-stage every owned bird unchanged (dist commit: builds the first halo)
    int skip_bands;             // tiles: the two owned rows next to each row boundary are stepped by k_step_bands, skip them
    uint32_t nseam;             // column-aligned launches (COLS): CTAs at the head of the grid that step the general-path birds
    int gen_threads;            // COLS: those CTAs use a thread per bird (large fields) instead of a warp per bird
};

void launch_ingest(const Geom &g, const Store &s, uint32_t n, const uint32_t *id, const float2 *loc,
                   const float2 *ld, const uint32_t *payload, uint32_t base, int owned_only, cudaStream_t st);
void launch_ingest_dense(const Geom &g, const Store &s, uint32_t n, uint32_t id0, const float4 *rec4, uint32_t base,
                         cudaStream_t st);
void launch_gather_dense(const Store &s, uint32_t n_upper, uint32_t id0, uint32_t n, float4 *out, uint32_t *outside,
                         cudaStream_t st);
// phase 0: messages from the left/right neighbours (tiles: arrivals in a border row are forwarded to the lower/upper
// neighbour's message); phase 1 (tiles): messages from the lower/upper neighbours
void launch_unpack(const Geom &g, const Store &s, cudaStream_t st, int phase = 0);
// tiles: the birds this rank owns are not an index range of the READ buffer: compact them (any of the outputs may be null)
void launch_compact_owned(const Geom &g, const Store &s, uint32_t n_upper, uint32_t *out_id, float2 *out_loc, float2 *out_ld,
                          uint32_t *counter, cudaStream_t st, uint32_t *out_pay = nullptr);
void launch_step(const Geom &g, const Params &p, const Store &s, const StepArgs &a, uint32_t n_upper,
                 cudaStream_t st);
// tiles: Bird::step for the two owned rows next to each row boundary of local columns [col_lo, col_hi) — the only
// interior-column birds whose successors can enter the lower/upper messages
void launch_step_bands(const Geom &g, const Params &p, const Store &s, const StepArgs &a, int col_lo, int col_hi,
                       cudaStream_t st);
// exclusive scan of cnt -> cell_start, zeroing cnt; 3 launches
void launch_scan(const Store &s, uint32_t ncell_plus1, cudaStream_t st);
void launch_scatter(const Store &s, uint32_t n_write_extent, int dist, cudaStream_t st);
void launch_rank(const Store &s, uint32_t n_upper, uint32_t ncell, int bump_step, cudaStream_t st, uint32_t ncols = 0, uint32_t dh = 0);
// slot-staged WRITE buffer -> READ buffer in (cell, ascending id) order, given cell_start (one pass, 8 lanes per cell)
// the whole rebin of a small single-GPU field (classic WRITE buffer) in one launch
constexpr uint32_t SMALL_REBIN_MAX = 4096;   // birds: up to here one CTA beats the five-launch rebin
uint32_t step_warp_max();   // populations up to this many birds are stepped by k_step_warp
void launch_rebin_small(const Store &s, uint32_t ncell, uint32_t n_write, int bump_step, cudaStream_t st, uint32_t ncols,
                        uint32_t dh);
void launch_compact(const Store &s, uint32_t ncell, int bump_step, cudaStream_t st, uint32_t ncols = 0, uint32_t dh = 0);
void launch_selftest_div2(uint64_t n, uint64_t seed, unsigned long long *mismatches, cudaStream_t st);
void launch_fill_u32(uint32_t *p, uint32_t v, uint64_t n, cudaStream_t st);
void launch_find(const Store &s, uint32_t n_upper, uint32_t id, float4 *out, uint32_t *found, cudaStream_t st);
void launch_split(const Store &s, uint32_t n, float2 *loc, float2 *ld, cudaStream_t st);
void launch_query_count(const Geom &g, const Store &s, uint32_t nq, const float2 *loc, float dist, int relax,
                        uint32_t *counts, cudaStream_t st);
// objects = 0: out[offsets[q] + k] = id; 1: out[5 * (offsets[q] + k) ..] = {id, loc.x, loc.y, last_d.x, last_d.y}
void launch_query_fill(const Geom &g, const Store &s, uint32_t nq, const float2 *loc, float dist, int relax,
                       const uint64_t *offsets, uint32_t *out, cudaStream_t st, int objects = 0);

}  // namespace flk
