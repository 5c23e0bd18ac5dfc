// flk_kernels.cu — sm_100a kernels of the Flockers hot path.
//
//   k_step      Bird::step for every bird of the READ buffer (flockers/src/model/bird.rs:27-145), fused with
//               Field2D::set_object_location (cell key + per-cell arrival slot of the successor, bird.rs:142-144)
//   k_scan_onepass  exclusive scan of the per-cell counters -> cell_start          } Field2D::lazy_update
//   k_scatter   counting-sort scatter of (id, source index) into cell order       } (flockers/src/model/state.rs:54-56)
//   k_rank      ascending-id rank inside each cell + gather of the records        }
//   k_rebin_small   the three of them in one CTA (small fields)
//   k_step_warp a warp per bird (small fields)
//   k_ingest    bulk Field2D::set_object_location from host arrays (state.rs:49)
//   k_query_*   Field2D::get_neighbors_within_{relax_,}distance (bird.rs:29-31)
//
// No tensor cores: the path has no dense contraction.
#include <cstdlib>

#include "flk_kernels.cuh"
#include "flk_step.cuh"

namespace flk {

// ------------------------------------------------------------------------------------------------
// neighbour window enumeration: calls f(s, e) for each run [s,e) of READ indices, in the reference's
// query order (x-outer, y-inner, bag order; SURVEY Appendix B "relax query")
// ------------------------------------------------------------------------------------------------
template <int ROWS = -1, class F>   // ROWS: 0 = rows not split (compiled out), 1 = split, -1 = decided at run time
__device__ __forceinline__ void for_each_window_run(const Geom &g, const uint32_t *__restrict__ cell_start,
                                                    int lc, int cy, int k, F &&f) {
    const bool rows_split = ROWS < 0 ? g.mode == 2 : ROWS != 0;
    for (int ox = -k; ox <= k; ++ox) {
        const int col = window_col(g, lc, ox);
        if (col < 0) continue;
        int r0 = cy - k, r1 = cy + k;
        bool contiguous;
        if (rows_split) {
            // cy is a LOCAL row; the ghost rows materialise the wrap.  Slack row (global row my): window {my-1, 0, 1}
            if (cy == g.rown + 2) { r0 = 0; r1 = 2 * k; }
            r0 = max(r0, 0);
            r1 = min(r1, g.rown + 1);
            contiguous = true;
        } else if (g.toroidal) {
            contiguous = (r0 >= 0) && (r1 < g.my);
        } else {
            r0 = max(r0, 0);
            r1 = min(r1, g.dh - 1);
            contiguous = true;
        }
        const int ngroups = contiguous ? 1 : (2 * k + 1);
        const uint32_t *cs = cell_start + (size_t)col * g.dh;
        for (int gi = 0; gi < ngroups; ++gi) {
            int ra, rb;
            if (contiguous) {
                ra = r0; rb = r1;
            } else {
                ra = rb = wrap_row(g, cy - k + gi);
                if (ra < 0) continue;
            }
            f(__ldg(cs + ra), __ldg(cs + rb + 1));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// neighbour accumulation (bird.rs:44-71), candidates visited in query order
// ------------------------------------------------------------------------------------------------

// Generic form: any window width, wrap, seam, either query.  Used by border birds and by the exact-distance query.
// MATH: 0 = exact (division guarded per pair, div2_rn); 1 = fast (approximate reciprocal)
template <bool RELAX, int MATH, int ROWS = -1>
__device__ __forceinline__ void accumulate_window(const Geom &g, const Params &p, const Store &s, const Rec &me,
                                                  uint32_t i, int lc, int cy, Acc &acc) {
    const float W = g.W, H = g.H;
    const float halfW = __fdiv_rn(W, 2.0f), halfH = __fdiv_rn(H, 2.0f);
    float xa = 0.0f, ya = 0.0f, xc = 0.0f, yc = 0.0f, xs = 0.0f, ys = 0.0f;
    uint32_t nvec = 0;
    int count = 0;
    const float4 *__restrict__ rec = reinterpret_cast<const float4 *>(s.rec);
    for_each_window_run<ROWS>(g, s.cell_start, lc, cy, p.k, [&](uint32_t rs, uint32_t re) {
        uint32_t j = rs;
        while (true) {
            // the caller itself is part of the query result and skipped by id (bird.rs:53): split the run at i
            const uint32_t stop = (RELAX && i >= j && i < re) ? i : re;
            for (; j < stop; ++j) {
                const float4 o = __ldg(rec + j);
                const float dx = toroidal_distance(me.x, o.x, W, halfW);   // bird.rs:54
                const float dy = toroidal_distance(me.y, o.y, H, halfH);   // bird.rs:55
                const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // :59
                if (!RELAX) {
                    // get_neighbors_within_distance keeps elem iff sqrt(dx^2+dy^2) <= dist
                    float qs = sq;
                    if (!g.toroidal) {
                        const float px = __fadd_rn(me.x, -o.x), py = __fadd_rn(me.y, -o.y);
                        qs = __fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py));
                    }
                    if (!(__fsqrt_rn(qs) <= p.radius)) continue;
                    ++nvec;
                    if (j == i) continue;
                    ++count;
                }
                const float den = __fadd_rn(__fmul_rn(sq, sq), 1.0f);
                float qa, qb;
                if (MATH == 0) {
                    div2_rn(dx, dy, den, qa, qb);             // :60-61, two IEEE divisions
                } else {
                    float inv;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(den));
                    qa = __fmul_rn(dx, inv);
                    qb = __fmul_rn(dy, inv);
                }
                xa = __fadd_rn(xa, qa);      // :60
                ya = __fadd_rn(ya, qb);      // :61
                xc = __fadd_rn(xc, dx);      // :64
                yc = __fadd_rn(yc, dy);      // :65
                xs = __fadd_rn(xs, o.z);     // :68
                ys = __fadd_rn(ys, o.w);     // :69
            }
            if (stop == re) break;
            j = stop + 1;
        }
        if (RELAX) {
            nvec += re - rs;
            count += (int)(re - rs) - ((i >= rs && i < re) ? 1 : 0);
        }
    });
    acc.xa = xa; acc.ya = ya; acc.xc = xc; acc.yc = yc; acc.xs = xs; acc.ys = ys;
    acc.nvec = nvec; acc.count = count;
}

// Interior form (k == 1, no wrap, no seam): the three column runs are known without any modulo or branching, all six
// cell_start loads are issued up front, and each run is a tight 2x-unrolled loop of 25 instructions per candidate.
// The bird itself sits in the middle run and is neutralised there (bird.rs:53, see accumulate_run_smem).  Exact
// arithmetic, see div2_rn_interior.
template <int MATH, bool SELF>
__device__ __forceinline__ void accumulate_run(const float4 *__restrict__ rec, uint32_t j, uint32_t e, uint32_t self,
                                               float mx_, float my_, float &xa, float &ya, float &xc, float &yc,
                                               float &xs, float &ys) {
#pragma unroll 2
    for (; j < e; ++j) {
        float4 o = __ldg(rec + j);
        if (SELF && j == self) { o.z = 0.0f; o.w = 0.0f; }   // see accumulate_run_smem
        const float dx = __fadd_rn(mx_, -o.x);                              // bird.rs:54 (no seam in reach)
        const float dy = __fadd_rn(my_, -o.y);                              // :55
        const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));   // :59
        const float den = __fadd_rn(__fmul_rn(sq, sq), 1.0f);
        float qa, qb;
        if (MATH == 1) {
            float inv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(den));
            qa = __fmul_rn(dx, inv);
            qb = __fmul_rn(dy, inv);
        } else {
            div2_rn_interior(dx, dy, den, qa, qb);                          // :60-61
        }
        xa = __fadd_rn(xa, qa);
        ya = __fadd_rn(ya, qb);
        xc = __fadd_rn(xc, dx);      // :64
        yc = __fadd_rn(yc, dy);      // :65
        xs = __fadd_rn(xs, o.z);     // :68
        ys = __fadd_rn(ys, o.w);     // :69
    }
}

template <int MATH>
__device__ __forceinline__ void accumulate_interior(const Geom &g, const Store &s, const Rec &me, uint32_t i, int lc,
                                                    int cy, Acc &acc) {
    const float4 *__restrict__ rec = reinterpret_cast<const float4 *>(s.rec);
    const uint32_t *__restrict__ cs = s.cell_start + (uint32_t)(lc - 1) * (uint32_t)g.dh + (uint32_t)(cy - 1);
    const uint32_t a0 = __ldg(cs), b0 = __ldg(cs + 3);
    const uint32_t a1 = __ldg(cs + g.dh), b1 = __ldg(cs + g.dh + 3);
    const uint32_t a2 = __ldg(cs + 2 * g.dh), b2 = __ldg(cs + 2 * g.dh + 3);
    float xa = 0.0f, ya = 0.0f, xc = 0.0f, yc = 0.0f, xs = 0.0f, ys = 0.0f;
    accumulate_run<MATH, false>(rec, a0, b0, 0, me.x, me.y, xa, ya, xc, yc, xs, ys);
    accumulate_run<MATH, true>(rec, a1, b1, i, me.x, me.y, xa, ya, xc, yc, xs, ys);   // own cell is in the middle run
    accumulate_run<MATH, false>(rec, a2, b2, 0, me.x, me.y, xa, ya, xc, yc, xs, ys);
    acc.xa = xa; acc.ya = ya; acc.xc = xc; acc.yc = yc; acc.xs = xs; acc.ys = ys;
    acc.nvec = (b0 - a0) + (b1 - a1) + (b2 - a2);
    acc.count = (int)acc.nvec - 1;
}

// ------------------------------------------------------------------------------------------------
// Field2D::set_object_location(*self, loc) (bird.rs:142-144): WRITE entry w + per-cell arrival slot.
// Strip mode adds the flockers_mpi routing (bird.rs:190-195, state.rs:113-175): a successor landing in one of
// the two columns next to a strip boundary is also appended to the message for that neighbour (it is either a
// migrant or next tick's ghost there); one landing outside the local columns is not kept here.
// ------------------------------------------------------------------------------------------------
// one WRITE entry: arrival `slot` of cell `key` (KEY_INVALID: not kept on this GPU), WRITE index w
__device__ __forceinline__ void write_entry(const Store &s, uint32_t w, uint32_t key, uint32_t slot, const float4 &r,
                                            uint32_t id) {
    if (s.slotw) {   // slot-staged: the record goes straight to its cell
        if (key == KEY_INVALID) return;
        if (slot < (uint32_t)s.sw_C) {
            const size_t q = (size_t)key * (uint32_t)s.sw_C + slot;
            s.sw_rec[q] = r;
            s.sw_id[q] = id;
        } else {
            const uint32_t k = atomicAdd(s.sw_spill_cnt, 1u);
            if (k < s.sw_spill_cap) {
                SpillEntry e;
                e.rec = r; e.id = id; e.cell = key; e.pad0 = 0; e.pad1 = 0;
                s.sw_spill[k] = e;
            } else {
                atomicOr(s.err_flag, ERR_CAPACITY);
            }
        }
        return;
    }
    reinterpret_cast<float4 *>(s.trec)[w] = r;
    s.tid[w] = id;
    s.tkey[w] = key;
    s.tslot[w] = slot;
}

template <bool NOPAY = false, int ROWS = -1>   // ROWS: 0 = rows not split (compiled out), 1 = split, -1 = decided at run time
__device__ __forceinline__ void emit_successor(const Geom &g, const Store &s, const Rec &out, uint32_t id, uint32_t w,
                                               bool d_ok, const Recip &Rd, uint32_t src) {
    // Field2D<O>: whatever else the object carries moves with it (set_object_location stores the whole object)
    if (!NOPAY)
        for (uint32_t k = 0; k < s.pw; ++k) s.tpay[(size_t)w * s.pw + k] = s.pay[(size_t)src * s.pw + k];
    const int ncx = d_ok ? cell_coord_fast(out.x, Rd, g.mx + 1) : cell_coord(out.x, g.d, g.mx + 1);
    const bool rows = ROWS < 0 ? g.mode == 2 : ROWS != 0;
    const int nrows = rows ? g.dhg : g.dh;   // the same number unless the rows are split
    const int ncy = d_ok ? cell_coord_fast(out.y, Rd, nrows) : cell_coord(out.y, g.d, nrows);
    const int nlc = local_col(g, ncx);
    const int nlr = rows ? local_row(g, ncy) : ncy;
    uint32_t key = KEY_INVALID, slot = 0;
    if (nlc != FLK_DROP && (!rows || nlr != FLK_DROP)) {
        key = (uint32_t)nlc * (uint32_t)g.dh + (uint32_t)nlr;
        slot = atomicAdd(s.cnt + key, 1u);
    }
    write_entry(s, w, key, slot, make_float4(out.x, out.y, out.ldx, out.ldy), id);
    const uint32_t *pay = (!NOPAY && s.pw) ? s.pay + (size_t)src * s.pw : nullptr;   // the object's payload travels with it
    if (g.mode != 0) {
        if (nlc == 0 || nlc == 1) xbuf_append(s, s.sendL, out, id, pay);
        if (nlc == g.nown || nlc == g.nown + 1 || (nlc == FLK_DROP && ncx >= g.mx)) xbuf_append(s, s.sendR, out, id, pay);
    }
    if (rows) {   // same rule along y; a corner successor goes into both (and is forwarded diagonally, k_unpack)
        if (nlr == 0 || nlr == 1) xbuf_append(s, s.sendD, out, id, pay);
        if (nlr == g.rown || nlr == g.rown + 1 || (nlr == FLK_DROP && ncy >= g.my)) xbuf_append(s, s.sendU, out, id, pay);
    }
}

// ------------------------------------------------------------------------------------------------
// k_step
// ------------------------------------------------------------------------------------------------
// everything of Bird::step after the neighbour loop (bird.rs:73-145) + set_object_location
template <int MATH, bool NOPAY = false, int ROWS = -1>
__device__ __forceinline__ void finish_bird(const Geom &g, const Params &p, const Store &s, const StepArgs &a, const Rec &me,
                                            uint32_t my_id, uint32_t w, const Acc &acc, bool d_ok, const Recip &Rd,
                                            uint32_t src) {
    const Rec out = bird_finish<MATH>(g, p, a, s.step_dev, me, my_id, acc);
    emit_successor<NOPAY, ROWS>(g, s, out, my_id, w, d_ok, Rd, src);   // :142-144
}

// Everything that is not the common case — a bird next to the seam, a CTA that straddles two cell columns, a window
// wider than 3x3, the exact-distance query, a non-toroidal field — finishes its bird here.  Kept out of line so
// that the register allocation of the tile path in k_step is not shaped by these loops (same arithmetic, same order).
template <bool RELAX, int MATH, bool TILES>
__device__ __noinline__ void step_bird_general(const Geom &g, const Params &p, const Store &s, const StepArgs &a,
                                               uint32_t i, uint32_t w) {
    const Rec me = s.rec[i];
    const uint32_t my_id = s.ids[i];
    const bool d_ok = g.d >= 0x1p-20f && g.d <= 4096.0f;
    const Recip Rd = make_recip(g.d);
    if (a.identity) {   // dist commit: re-stage the bird unchanged (builds the first halo)
        emit_successor<false, TILES ? 1 : 0>(g, s, me, my_id, w, d_ok, Rd, i);
        return;
    }
    const int gcx = d_ok ? cell_coord_fast(me.x, Rd, g.mx + 1) : cell_coord(me.x, g.d, g.mx + 1);
    const int nrows = TILES ? g.dhg : g.dh;
    const int cy = d_ok ? cell_coord_fast(me.y, Rd, nrows) : cell_coord(me.y, g.d, nrows);
    const int lc = local_col(g, gcx);
    const int lr = TILES ? local_row(g, cy) : cy;   // row of the local cell arrays (== cy unless the rows are split)
    Acc acc;
    // Birds whose 3x3 window cannot touch the seam never take the |a-b| > dim/2 branch of toroidal_distance:
    // cells are < 2d apart and dim >= 6d, so the plain difference is the toroidal one (exactly).
    const bool interior = g.toroidal && p.k == 1 && g.mx >= 6 && g.my >= 6 && gcx >= 1 && gcx <= g.mx - 2 &&
                          cy >= 1 && cy <= g.my - 2 && d_ok;
    // both queries return nothing for dist <= 0 (`if dist <= 0 -> []`; k_query does the same):
    // the bird then sees no neighbour, draws nothing and keeps only its momentum (bird.rs:42)
    if (p.radius <= 0.0f) { /* acc stays empty */ }
    else if (interior && RELAX) accumulate_interior<MATH>(g, s, me, i, lc, lr, acc);
    else accumulate_window<RELAX, MATH, TILES ? 1 : 0>(g, p, s, me, i, lc, lr, acc);
    finish_bird<MATH, false, TILES ? 1 : 0>(g, p, s, a, me, my_id, w, acc, d_ok, Rd, i);
}

// ------------------------------------------------------------------------------------------------
// k_step_warp: Bird::step with ONE WARP PER BIRD, for fields too small to fill the GPU with a thread per bird (the
// shipped run: 1000 birds = 8 CTAs of k_step, one warp per scheduler, ~8 000 dependent instructions each: 32 us).
// The lanes share the bird's candidates: every window run is taken 32 candidates at a time, each lane computes the pair
// terms of one candidate (bird.rs:54-61: toroidal distances, square, the two exact quotients) and parks them in shared
// memory; then six lanes — one per accumulator of bird.rs:44-49 — add the parked terms in candidate order, so every f32
// sum is evaluated in exactly the order of the thread-per-bird kernels.  Lane 0 finishes the bird.  Any window width,
// either query, seam or not, strips too (rows not split).
// ------------------------------------------------------------------------------------------------
constexpr int WARP_RUNS = 32;   // window runs a warp keeps in shared memory (3 or 9 for the shipped 3x3 window; 49 for 7x7 -> general kernel)

// Bird::step of READ bird i by the 32 lanes of the calling warp (all of them); term / run_start / run_pre: the warp's own
// shared-memory scratch (6 x 32 floats, 2 x (WARP_RUNS + 1) words)
template <bool RELAX, int MATH>
__device__ __forceinline__ void step_bird_warp(const Geom &g, const Params &p, const Store &s, const StepArgs &a, uint32_t i,
                                               float (*term)[32], uint32_t *run_start, uint32_t *run_pre) {
    const int lane = threadIdx.x & 31;
    const uint32_t w = a.out_base + i;
    const Rec me = s.rec[i];
    const uint32_t my_id = s.ids[i];
    const bool d_ok = g.d >= 0x1p-20f && g.d <= 4096.0f;
    const Recip Rd = make_recip(g.d);
    const int gcx = cell_coord(me.x, g.d, g.mx + 1);
    const int cy = cell_coord(me.y, g.d, g.dh);
    const int lc = local_col(g, gcx);
    if (g.mode != 0 && (lc == 0 || lc == g.nown + 1)) {    // ghost copy: stepped by its owner, dropped here
        if (lane == 0 && !s.slotw) s.tkey[w] = KEY_INVALID;
        return;
    }
    if (a.identity) {                                      // dist commit: re-stage the bird unchanged
        if (lane == 0) emit_successor<false, 0>(g, s, me, my_id, w, d_ok, Rd, i);
        return;
    }
    const float W = g.W, H = g.H;
    const float halfW = __fdiv_rn(W, 2.0f), halfH = __fdiv_rn(H, 2.0f);
    float acc_l = 0.0f;                                    // lanes 0..5: x_avoid, y_avoid, x_cohe, y_cohe, x_cons, y_cons
    uint32_t nvec = 0;
    int count = 0;
    const float4 *__restrict__ rec = reinterpret_cast<const float4 *>(s.rec);
    if (p.radius > 0.0f) {                                 // both queries return nothing for dist <= 0
        // ---- the window's runs in query order (all their bounds are requested before any is used)
        int nruns = 0;
        uint32_t total = 0;
        for_each_window_run<0>(g, s.cell_start, lc, cy, p.k, [&](uint32_t rs, uint32_t re) {
            if (lane == 0 && nruns < WARP_RUNS) { run_start[nruns] = rs; run_pre[nruns] = total; }
            total += re - rs;
            ++nruns;
        });
        if (lane == 0 && nruns <= WARP_RUNS) run_pre[nruns] = total;
        __syncwarp();
        // ---- candidates 32 at a time, by their position in the query result
        for (uint32_t c0 = 0; c0 < total; c0 += 32) {
            const uint32_t c = c0 + (uint32_t)lane;
            bool keep = c < total;
            bool self = false;
            if (keep) {
                int r = 0;
                for (int q = 1; q < nruns; ++q) r += run_pre[q] <= c ? 1 : 0;     // the run that holds candidate c
                const uint32_t j = run_start[r] + (c - run_pre[r]);
                self = j == i;                                                     // bird.rs:53
                const float4 o = __ldg(rec + j);
                const float dx = toroidal_distance(me.x, o.x, W, halfW);           // bird.rs:54
                const float dy = toroidal_distance(me.y, o.y, H, halfH);           // bird.rs:55
                const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // :59
                if (!RELAX) {   // get_neighbors_within_distance keeps elem iff sqrt(dx^2+dy^2) <= dist
                    float qs = sq;
                    if (!g.toroidal) {
                        const float px = __fadd_rn(me.x, -o.x), py = __fadd_rn(me.y, -o.y);
                        qs = __fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py));
                    }
                    keep = __fsqrt_rn(qs) <= p.radius;
                }
                const float den = __fadd_rn(__fmul_rn(sq, sq), 1.0f);
                float qa, qb;
                if (MATH == 0) {
                    div2_rn(dx, dy, den, qa, qb);          // :60-61, two IEEE divisions
                } else {
                    float inv;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(den));
                    qa = __fmul_rn(dx, inv);
                    qb = __fmul_rn(dy, inv);
                }
                term[0][lane] = qa; term[1][lane] = qb;       // :60-61
                term[2][lane] = dx; term[3][lane] = dy;       // :64-65
                term[4][lane] = o.z; term[5][lane] = o.w;     // :68-69
            }
            const uint32_t kept = __ballot_sync(0xffffffffu, keep);
            const uint32_t add = kept & ~__ballot_sync(0xffffffffu, self);
            nvec += (uint32_t)__popc(kept);
            count += __popc(add);                          // :56
            __syncwarp();
            if (lane < 6) {
                uint32_t m = add;
                while (m) {                                // candidates in order, one rounded addition each
                    const int e = __ffs(m) - 1;
                    m &= m - 1;
                    acc_l = __fadd_rn(acc_l, term[lane][e]);
                }
            }
            __syncwarp();
        }
    }
    Acc acc;
    acc.xa = __shfl_sync(0xffffffffu, acc_l, 0); acc.ya = __shfl_sync(0xffffffffu, acc_l, 1);
    acc.xc = __shfl_sync(0xffffffffu, acc_l, 2); acc.yc = __shfl_sync(0xffffffffu, acc_l, 3);
    acc.xs = __shfl_sync(0xffffffffu, acc_l, 4); acc.ys = __shfl_sync(0xffffffffu, acc_l, 5);
    acc.nvec = nvec; acc.count = count;
    if (lane == 0) finish_bird<MATH, false, 0>(g, p, s, a, me, my_id, w, acc, d_ok, Rd, i);
}

template <bool RELAX, int MATH>
__global__ void __launch_bounds__(128) k_step_warp(const __grid_constant__ Geom g, const __grid_constant__ Params p,
                                                   const __grid_constant__ Store s, const __grid_constant__ StepArgs a) {
    __shared__ float term_s[4][6][32];
    __shared__ uint32_t run_s[4][2][WARP_RUNS + 1];        // [0]: first READ index of each run, [1]: candidates before it
    const int wid = threadIdx.x >> 5;
    const uint32_t lo = a.cell_lo ? __ldg(s.cell_start + a.cell_lo) : 0u;
    const uint32_t hi = __ldg(s.cell_start + a.cell_hi);
    if (blockIdx.x == 0 && threadIdx.x == 0 && hi - lo > gridDim.x * 4u) atomicOr(s.err_flag, ERR_EXCHANGE);
    const uint32_t i = lo + blockIdx.x * 4u + (uint32_t)wid;
    if (i >= hi) return;                                   // warp-uniform
    step_bird_warp<RELAX, MATH>(g, p, s, a, i, term_s[wid], run_s[wid][0], run_s[wid][1]);
}

// out of line for k_step's column-aligned launches (own register allocation, like step_bird_general)
template <bool RELAX, int MATH>
__device__ __noinline__ void step_bird_warp_ool(const Geom &g, const Params &p, const Store &s, const StepArgs &a, uint32_t i,
                                                float (*term)[32], uint32_t *run_start, uint32_t *run_pre) {
    step_bird_warp<RELAX, MATH>(g, p, s, a, i, term, run_start, run_pre);
}

// k_step: one thread per READ bird.  TILE = stage the CTA's three window-column runs in shared memory with TMA bulk
// copies when the CTA's birds sit in one interior column (the common case); other CTAs take step_bird_general.
// FASTG = the host checked what does not depend on the bird (launch_step): toroidal field, 3x3 window, relax query,
// grid >= 6x6 cells, discretisation in the shared-reciprocal range, no payload, not the identity pass.
#ifndef FLK_STEP_MINBLOCKS
#define FLK_STEP_MINBLOCKS 9   // 56 registers; with the prefetching loop: 0.851 ms (9), 0.856 (10), 0.884 (12)
#endif
// TILES = the rows are split too (mode 2): the row of the local cell arrays differs from the global row
// COLS (single-GPU handles, with TILE and FASTG): the grid is aligned to the cell columns.  After the a.nseam seam CTAs, CTA
// b steps one chunk of FLK_STEP_BS birds of ONE column (s.colchunk maps chunks to columns), so no CTA straddles two
// columns, and a CTA that starts or ends a column stages the rows it has (clamped at the seam) instead of leaving the tile
// path.  The birds that need the full toroidal distance — the seam rows 0, my-1 and my (slack) of those columns and all of
// columns 0, mx-1 and mx (slack) — are stepped by the a.nseam CTAs at the head of the grid: with ONE WARP PER BIRD
// (step_bird_warp) while they fit about one wave — a thread per bird on the general path takes ~25 us, which in a launch
// of a few waves is the launch — and with a thread per bird, all lanes busy, in larger fields (a.gen_threads).
template <bool RELAX, int MATH, bool TILE, bool FASTG, bool TILES = false, bool COLS = false>
__global__ void __launch_bounds__(FLK_STEP_BS, FLK_STEP_MINBLOCKS)
k_step(const __grid_constant__ Geom g, const __grid_constant__ Params p, const __grid_constant__ Store s,
       const __grid_constant__ StepArgs a) {
    __shared__ __align__(128) float4 tile_s[TILE ? 3 * TILE_CAP + 1 : 1];
    float4 *const tile[3] = {tile_s, tile_s + (TILE ? TILE_CAP : 0), tile_s + (TILE ? 2 * TILE_CAP : 0)};
    __shared__ __align__(8) uint64_t bar;
    __shared__ int shc[TILES ? 7 : 5];
    __shared__ float term_w[COLS ? FLK_STEP_BS / 32 : 1][6][32];                  // COLS: scratch of step_bird_warp
    __shared__ uint32_t run_w[COLS ? FLK_STEP_BS / 32 : 1][2][WARP_RUNS + 1];

    // cell_start[0] is 0 by construction: a whole-field launch knows its first bird without a load, so the bird's own
    // record (read speculatively: any index below cap is allocated) is in flight together with the range end
    uint32_t lo, hi, first;
    if (COLS) {
        const uint32_t dhc = (uint32_t)g.dh, ncols = (uint32_t)g.ncols;
        if (blockIdx.x < a.nseam) {
            // ---- the birds that need the general path, ONE WARP EACH (they are few, and a thread per bird takes ~25 us
            // for one of them: in a launch of a few waves that would be the launch): column 0, then columns mx-1 and
            // mx (slack), then the seam rows 0, my-1, my of the tile columns
            const int wid = threadIdx.x >> 5;
            const uint32_t *cs = s.cell_start;
            const uint32_t nA = __ldg(cs + dhc), bB = __ldg(cs + (ncols - 2) * dhc), nB = __ldg(cs + ncols * dhc) - bB;
            const uint32_t *colseam = s.colchunk + ncols + 1;
            const uint32_t nS = __ldg(colseam + ncols), ngen = nA + nB + nS;
            auto general_bird = [&](uint32_t wq) -> uint32_t {   // READ index of general bird number wq
                if (wq < nA) return wq;
                if (wq < nA + nB) return bB + (wq - nA);
                const uint32_t q = wq - nA - nB;
                uint32_t c = min(ncols - 1, (uint32_t)(((uint64_t)q * ncols) / nS));
                while (__ldg(colseam + c) > q) --c;
                while (__ldg(colseam + c + 1) <= q) ++c;
                const uint32_t r = q - __ldg(colseam + c);
                const uint32_t c0 = __ldg(cs + c * dhc), n0 = __ldg(cs + c * dhc + 1) - c0;
                return r < n0 ? c0 + r : __ldg(cs + (c + 1) * dhc - 2) + (r - n0);
            };
            if (a.gen_threads) {
                // a field of many waves: a THREAD per general bird (the warps would need several waves of their own;
                // here the slow threads start first and finish inside the launch)
                for (uint32_t tq = blockIdx.x * FLK_STEP_BS + threadIdx.x; tq < ngen; tq += a.nseam * FLK_STEP_BS) {
                    const uint32_t i = general_bird(tq);
                    step_bird_general<RELAX, MATH, TILES>(g, p, s, a, i, a.out_base + i);
                }
                return;
            }
            for (uint32_t wq = blockIdx.x * (FLK_STEP_BS / 32) + (uint32_t)wid; wq < ngen; wq += a.nseam * (FLK_STEP_BS / 32)) {
                const uint32_t i = general_bird(wq);
                step_bird_warp_ool<RELAX, MATH>(g, p, s, a, i, term_w[wid], run_w[wid][0], run_w[wid][1]);
            }
            return;
        }
        const uint32_t nch = __ldg(s.colchunk + ncols);
        const uint32_t B = blockIdx.x - a.nseam;
        if (B >= nch) return;
        uint32_t c = min(ncols - 1, (uint32_t)(((uint64_t)B * ncols) / nch));
        while (__ldg(s.colchunk + c) > B) --c;
        while (__ldg(s.colchunk + c + 1) <= B) ++c;
        lo = __ldg(s.cell_start + c * dhc);
        hi = __ldg(s.cell_start + (c + 1) * dhc);
        first = lo + (B - __ldg(s.colchunk + c)) * blockDim.x;
    } else {
        lo = a.cell_lo ? __ldg(s.cell_start + a.cell_lo) : 0u;
        first = lo + blockIdx.x * blockDim.x;
    }
    const uint32_t i = first + threadIdx.x;
    float2 xy = make_float2(0.0f, 0.0f);
    if (TILE && i < s.cap) xy = *reinterpret_cast<const float2 *>(s.rec + i);
    if (!COLS) {
        hi = __ldg(s.cell_start + a.cell_hi);
        if (blockIdx.x == 0 && threadIdx.x == 0 && hi - lo > gridDim.x * blockDim.x)
            atomicOr(s.err_flag, ERR_EXCHANGE);    // the launch does not cover the range (border launches use capx)
    }
    if (first >= hi) return;                       // whole CTA out of range
    const bool valid = i < hi;
    const uint32_t w = a.out_base + i;

    if (TILE) {
        // cell of this bird (same arithmetic as the binning that placed it)
        const bool d_ok = FASTG || (g.d >= 0x1p-20f && g.d <= 4096.0f);
        const Recip Rd = make_recip(g.d);
        const uint32_t dh = (uint32_t)g.dh;
        float mex = 0.0f, mey = 0.0f;
        int gcx = 0, cy = 0, lc = 0, lr = 0;   // cy: global row (seam test); lr: row of the local cell arrays
        if (valid) {
            mex = xy.x; mey = xy.y;
            gcx = d_ok ? cell_coord_fast(mex, Rd, g.mx + 1) : cell_coord(mex, g.d, g.mx + 1);
            cy = d_ok ? cell_coord_fast(mey, Rd, TILES ? g.dhg : g.dh) : cell_coord(mey, g.d, TILES ? g.dhg : g.dh);
            lc = local_col(g, gcx);
            lr = TILES ? local_row(g, cy) : cy;
        }
        const uint32_t last = min(first + blockDim.x, hi) - 1;
        if (threadIdx.x == 0) {
            shc[0] = lc; shc[1] = cy; shc[4] = gcx;
            if (TILES) shc[5] = lr;
            mbar_init(&bar, 1);
        }
        if (i == last) { shc[2] = lc; shc[3] = cy; if (TILES) shc[6] = lr; }
        __syncthreads();
        const int lcA = shc[0], cyA = shc[1], lcB = shc[2], cyB = shc[3], gcxA = shc[4];
        const int lrA = TILES ? shc[5] : cyA, lrB = TILES ? shc[6] : cyB;
        // COLS: the seam rows are somebody else's, so any row range of a tile column will do (staged rows clamped to 0 .. my-1)
        bool eligible = (FASTG || (RELAX && !a.identity && g.toroidal && p.k == 1 && g.mx >= 6 && g.my >= 6 && d_ok &&
                                   s.pw == 0)) &&
                        lcA == lcB && gcxA >= 1 && gcxA <= g.mx - 2 && (COLS || (cyA >= 1 && cyB <= g.my - 2)) && cyA <= cyB &&
                        (g.mode == 0 || (lcA >= 1 && lcA <= g.nown)) &&
                        (!TILES || (lrA >= 1 && lrB <= g.rown && lrA <= lrB));
        uint32_t sA0 = 0, sA1 = 0, sA2 = 0, len0 = 0, len1 = 0, len2 = 0;
        if (eligible) {
            const int rlo = COLS ? max(lrA - 1, 0) : lrA - 1, rhi = COLS ? min(lrB + 1, g.my - 1) : lrB + 1;
            const uint32_t *cs = s.cell_start + ((uint32_t)(lcA - 1) * dh + (uint32_t)rlo);
            const uint32_t span = (uint32_t)(rhi - rlo) + 1u;
            sA0 = __ldg(cs);          len0 = __ldg(cs + span) - sA0;
            sA1 = __ldg(cs + dh);     len1 = __ldg(cs + dh + span) - sA1;
            sA2 = __ldg(cs + 2 * dh); len2 = __ldg(cs + 2 * dh + span) - sA2;
            eligible = len0 <= TILE_CAP && len1 <= TILE_CAP && len2 <= TILE_CAP;
        }
        if (eligible) {   // CTA-uniform
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, (len0 + len1 + len2) * 16u);
                const float4 *rec = reinterpret_cast<const float4 *>(s.rec);
                if (len0) tma_load_1d(&tile[0][0], rec + sA0, len0 * 16u, &bar);
                if (len1) tma_load_1d(&tile[1][0], rec + sA1, len1 * 16u, &bar);
                if (len2) tma_load_1d(&tile[2][0], rec + sA2, len2 * 16u, &bar);
            }
            // this bird's three runs, as offsets into the staged column runs (loads in flight during the copy)
            uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0, a2 = 0, b2 = 0;
            if (valid && !(COLS && (cy == 0 || cy >= g.my - 1))) {
                const uint32_t *cs = s.cell_start + ((uint32_t)(lc - 1) * dh + (uint32_t)(lr - 1));
                a0 = __ldg(cs);          b0 = __ldg(cs + 3);
                a1 = __ldg(cs + dh);     b1 = __ldg(cs + dh + 3);
                a2 = __ldg(cs + 2 * dh); b2 = __ldg(cs + 2 * dh + 3);
            }
            mbar_wait(&bar, 0);
            if (!valid) return;
            if (COLS && (cy == 0 || cy >= g.my - 1)) return;                       // a seam row: stepped by the seam CTAs
            if (TILES && a.skip_bands && (lr <= 2 || lr >= g.rown - 1)) return;   // stepped by k_step_bands
            float xa = 0.0f, ya = 0.0f, xc = 0.0f, yc = 0.0f, xs = 0.0f, ys = 0.0f;
            if (MATH == 0 && FLK_PACKED) {
                const f32x2 ME = pk(mex, mey);
                f32x2 A = 0ull, Cc = 0ull, Ss = 0ull;
                accumulate_run_smem_packed<false>(tile[0], a0 - sA0, b0 - sA0, 0, ME, A, Cc, Ss);
                accumulate_run_smem_packed<true>(tile[1], a1 - sA1, b1 - sA1, i - sA1, ME, A, Cc, Ss);
                accumulate_run_smem_packed<false>(tile[2], a2 - sA2, b2 - sA2, 0, ME, A, Cc, Ss);
                unpk(A, xa, ya); unpk(Cc, xc, yc); unpk(Ss, xs, ys);
            } else {
                accumulate_run_smem<MATH, false>(tile[0], a0 - sA0, b0 - sA0, 0, mex, mey, xa, ya, xc, yc, xs, ys);
                accumulate_run_smem<MATH, true>(tile[1], a1 - sA1, b1 - sA1, i - sA1, mex, mey, xa, ya, xc, yc, xs, ys);
                accumulate_run_smem<MATH, false>(tile[2], a2 - sA2, b2 - sA2, 0, mex, mey, xa, ya, xc, yc, xs, ys);
            }
            Acc acc;
            acc.xa = xa; acc.ya = ya; acc.xc = xc; acc.yc = yc; acc.xs = xs; acc.ys = ys;
            acc.nvec = (b0 - a0) + (b1 - a1) + (b2 - a2);
            acc.count = (int)acc.nvec - 1;
            // the heading and the id are only needed from here on: fetched now (L1/L2 hits) instead of being carried
            // through the loops
            const float2 ld = *(reinterpret_cast<const float2 *>(s.rec + i) + 1);
            const Rec me = Rec{mex, mey, ld.x, ld.y};
            finish_bird<MATH, FASTG, TILES ? 1 : 0>(g, p, s, a, me, s.ids[i], w, acc, d_ok, Rd, i);
            return;
        }
    }
    if (!valid) return;
    if (COLS) {   // a tile column's seam-row birds belong to the seam CTAs whatever path this CTA takes
        const Rec me = s.rec[i];
        const int cx = cell_coord(me.x, g.d, g.mx + 1), cy = cell_coord(me.y, g.d, g.dh);
        if (cx >= 1 && cx <= g.mx - 2 && (cy == 0 || cy >= g.my - 1)) return;
    }
    if (g.mode != 0) {
        const Rec me = s.rec[i];
        const int lc = local_col(g, cell_coord(me.x, g.d, g.mx + 1));
        const int lr = TILES ? local_row(g, cell_coord(me.y, g.d, g.dhg)) : 1;
        if (lc == 0 || lc == g.nown + 1 || (TILES && (lr == 0 || lr == g.rown + 1))) {
            // ghost copy (flockers_mpi: received_neighbors): stepped by its owner, dropped here
            if (!s.slotw) s.tkey[w] = KEY_INVALID;
            return;
        }
        if (TILES && a.skip_bands && (lr <= 2 || lr >= g.rown - 1)) return;   // stepped by k_step_bands (slack row too)
    }
    step_bird_general<RELAX, MATH, TILES>(g, p, s, a, i, w);   // also the identity pass (re-stage unchanged)
}

// k_step_bands (tiles): one warp per (column, band); band 0 = local rows {1,2}, band 1 = rows {rown-1, rown}, band 2 = the slack row.  The birds
// of a band are a contiguous index range of the READ buffer; each is finished by step_bird_general (a few thousand
// birds per tick: the point is that they are done EARLY, so that the lower/upper exchange can run beside the
// interior step)
template <bool RELAX, int MATH>
__global__ void __launch_bounds__(32) k_step_bands(const __grid_constant__ Geom g, const __grid_constant__ Params p,
                                                   const __grid_constant__ Store s, const __grid_constant__ StepArgs a,
                                                   int col_lo) {
    const uint32_t col = (uint32_t)col_lo + blockIdx.y, dh = (uint32_t)g.dh;
    // band 2 = the slack row (a bird at y == H steps into row my-1 or row 0, i.e. into a border row of some tile)
    const uint32_t r_lo = blockIdx.x == 0 ? 1u : (blockIdx.x == 1 ? (uint32_t)g.rown - 1u : (uint32_t)g.rown + 2u);
    const uint32_t r_hi = blockIdx.x == 2 ? r_lo + 1u : r_lo + 2u;
    const uint32_t lo = s.cell_start[col * dh + r_lo], hi = s.cell_start[col * dh + r_hi];
    for (uint32_t i = lo + threadIdx.x; i < hi; i += 32) step_bird_general<RELAX, MATH, true>(g, p, s, a, i, a.out_base + i);
}

void launch_step_bands(const Geom &g, const Params &p, const Store &s, const StepArgs &a, int col_lo, int col_hi,
                       cudaStream_t st) {
    if (col_hi <= col_lo) return;
    const dim3 grid(3, (unsigned)(col_hi - col_lo));
    if (p.relax) {
        if (a.math_mode == 1) k_step_bands<true, 1><<<grid, 32, 0, st>>>(g, p, s, a, col_lo);
        else k_step_bands<true, 0><<<grid, 32, 0, st>>>(g, p, s, a, col_lo);
    } else {
        if (a.math_mode == 1) k_step_bands<false, 1><<<grid, 32, 0, st>>>(g, p, s, a, col_lo);
        else k_step_bands<false, 0><<<grid, 32, 0, st>>>(g, p, s, a, col_lo);
    }
}

constexpr uint32_t COLS_STEP_MAX = 12000000;   // birds: up to here the column-aligned grid pays (8 M birds on a square field: 587 vs 607 us per tick;
                                               // 16 M birds on 672 x 5376 cells: 1.131 ms either way, long columns have few straddling CTAs)
constexpr uint32_t GENWARP_MAX = 8000;     // general birds (seam rows + seam columns) of a column-aligned launch: up to here a warp each, a thread
                                           // each beyond (B200: 250 000 birds 41.0 vs 51.2 us per tick, 500 000 birds 64.4 vs 60.8)
constexpr uint32_t WARP_STEP_MAX = 5000;    // birds: up to here a warp per bird beats the column-aligned grid (B200: 4000 birds 22.1 vs 23.5 us per tick,
                                            // 6000 birds 24.8 vs 22.7, 12 000 birds 34.8 vs 24.6)

uint32_t step_warp_max() {
    static const uint32_t v = [] { const char *e = getenv("FLK_STEP_WARP"); return e ? (uint32_t)atoi(e) : WARP_STEP_MAX; }();   // A/B runs
    return v;
}

void launch_step(const Geom &g, const Params &p, const Store &s, const StepArgs &a, uint32_t n_upper,
                 cudaStream_t st) {
    if (n_upper == 0) return;
    const uint32_t warp_max = step_warp_max();
    if (n_upper <= warp_max && g.mode != 2 && (2 * p.k + 1) * (2 * p.k + 1) <= WARP_RUNS) {
        const dim3 blk(128), grd((n_upper + 3) / 4);
        if (p.relax) {
            if (a.math_mode == 1) k_step_warp<true, 1><<<grd, blk, 0, st>>>(g, p, s, a);
            else k_step_warp<true, 0><<<grd, blk, 0, st>>>(g, p, s, a);
        } else {
            if (a.math_mode == 1) k_step_warp<false, 1><<<grd, blk, 0, st>>>(g, p, s, a);
            else k_step_warp<false, 0><<<grd, blk, 0, st>>>(g, p, s, a);
        }
        return;
    }
    static const int tile = [] { const char *e = getenv("FLK_STEP_TILE"); return e ? atoi(e) : 1; }();
    static const uint32_t colgrid_max = [] { const char *e = getenv("FLK_STEP_COLS"); return e ? (uint32_t)atoi(e) : COLS_STEP_MAX; }();   // A/B runs
    const int bs = FLK_STEP_BS;
    const dim3 block(bs), grid((n_upper + bs - 1) / bs);
    const bool fastg = p.relax && g.toroidal && p.k == 1 && g.mx >= 6 && g.my >= 6 && g.d >= 0x1p-20f && g.d <= 4096.0f &&
                       s.pw == 0 && !a.identity;
    // single-GPU handles of moderate size: column-aligned grid + seam CTAs (whole-field launches only)
    if (g.mode == 0 && tile && fastg && n_upper <= colgrid_max && s.colchunk && a.cell_lo == 0 && a.math_mode == 0) {
        StepArgs ac = a;
        // head of the grid: a warp per general bird (three columns + three rows of the field), sized for a uniform field
        // with a margin; the warps stride over whatever the population really is
        static const double genwarp_max = [] { const char *e = getenv("FLK_STEP_GENWARP"); return e ? atof(e) : (double)GENWARP_MAX; }();   // A/B runs
        const double gen = (double)n_upper * (3.0 / (double)g.ncols + 3.0 / (double)g.dh);
        ac.gen_threads = gen > genwarp_max ? 1 : 0;
        const double per_cta = ac.gen_threads ? (double)bs : (double)(bs / 32);
        ac.nseam = (uint32_t)fmin(fmax(gen * 1.25 / per_cta + 8.0, 16.0), 32768.0);
        const dim3 gridc(ac.nseam + (n_upper + bs - 1) / bs + (uint32_t)g.ncols);   // every column ends with one partial chunk at most
        k_step<true, 0, true, true, false, true><<<gridc, block, 0, st>>>(g, p, s, ac);
        return;
    }
#define FLK_LAUNCH(R, M, T, F) k_step<R, M, T, F><<<grid, block, 0, st>>>(g, p, s, a)
#define FLK_LAUNCH_T(R, M, T) k_step<R, M, T, false, true><<<grid, block, 0, st>>>(g, p, s, a)
    if (g.mode == 2) {   // rows split: every variant is the TILES instantiation (ghost rows, local row arithmetic)
        if (p.relax) {
            if (a.math_mode == 1) { if (tile) FLK_LAUNCH_T(true, 1, true); else FLK_LAUNCH_T(true, 1, false); }
            else if (tile && fastg) k_step<true, 0, true, true, true><<<grid, block, 0, st>>>(g, p, s, a);
            else { if (tile) FLK_LAUNCH_T(true, 0, true); else FLK_LAUNCH_T(true, 0, false); }
        } else {
            if (a.math_mode == 1) FLK_LAUNCH_T(false, 1, false);
            else FLK_LAUNCH_T(false, 0, false);
        }
    } else if (p.relax) {
        if (a.math_mode == 1) {
            if (tile && fastg) FLK_LAUNCH(true, 1, true, true);
            else if (tile) FLK_LAUNCH(true, 1, true, false);
            else FLK_LAUNCH(true, 1, false, false);
        } else {
            if (tile && fastg) FLK_LAUNCH(true, 0, true, true);
            else if (tile) FLK_LAUNCH(true, 0, true, false);
            else FLK_LAUNCH(true, 0, false, false);
        }
    } else {
        if (a.math_mode == 1) FLK_LAUNCH(false, 1, false, false);
        else FLK_LAUNCH(false, 0, false, false);
    }
#undef FLK_LAUNCH
#undef FLK_LAUNCH_T
}

// ------------------------------------------------------------------------------------------------
// k_ingest: bulk set_object_location from (already uploaded) arrays; appends at WRITE index base+j
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ingest(const Geom g, const Store s, uint32_t n, const uint32_t *__restrict__ id,
                                                const float2 *__restrict__ loc, const float2 *__restrict__ ld,
                                                const uint32_t *__restrict__ payload, uint32_t base, int owned_only) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float2 l = loc[j], d = ld[j];
    const int cx = cell_coord(l.x, g.d, g.mx + 1);
    const int cy = local_row(g, cell_coord(l.y, g.d, g.dhg));
    int lc = local_col(g, cx);
    // strip / tile mode: only birds this rank owns are accepted (flockers_mpi/src/model/state.rs:75-85); ghosts arrive
    // through the halo exchange
    if (owned_only && g.mode != 0 && (lc == 0 || lc == g.nown + 1)) lc = FLK_DROP;
    if (owned_only && g.mode == 2 && (cy == 0 || cy == g.rown + 1)) lc = FLK_DROP;
    uint32_t key = KEY_INVALID, slot = 0;
    if (lc != FLK_DROP && cy != FLK_DROP) {
        key = (uint32_t)lc * (uint32_t)g.dh + (uint32_t)cy;
        slot = atomicAdd(s.cnt + key, 1u);
    }
    write_entry(s, base + j, key, slot, make_float4(l.x, l.y, d.x, d.y), id[j]);
    for (uint32_t k = 0; k < s.pw; ++k)
        s.tpay[(size_t)(base + j) * s.pw + k] = payload ? payload[(size_t)j * s.pw + k] : 0u;
}

void launch_ingest(const Geom &g, const Store &s, uint32_t n, const uint32_t *id, const float2 *loc,
                   const float2 *ld, const uint32_t *payload, uint32_t base, int owned_only, cudaStream_t st) {
    if (n == 0) return;
    k_ingest<<<(n + 255) / 256, 256, 0, st>>>(g, s, n, id, loc, ld, payload, base, owned_only);
}

// dense form: the host keeps its agents as a list indexed by id (Flocker::init numbers them 0..N in list order,
// flockers/src/model/state.rs:40-51), so the id is the position and one 16-byte {loc, last_d} record is all that moves
__global__ void __launch_bounds__(256) k_ingest_dense(const Geom g, const Store s, uint32_t n, uint32_t id0,
                                                      const float4 *__restrict__ rec4, uint32_t base) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 r = rec4[j];
    const int cx = cell_coord(r.x, g.d, g.mx + 1);
    const int cy = cell_coord(r.y, g.d, g.dhg);
    const uint32_t key = (uint32_t)cx * (uint32_t)g.dh + (uint32_t)cy;
    const uint32_t slot = atomicAdd(s.cnt + key, 1u);
    write_entry(s, base + j, key, slot, r, id0 + j);
    for (uint32_t k = 0; k < s.pw; ++k) s.tpay[(size_t)(base + j) * s.pw + k] = 0u;
}

void launch_ingest_dense(const Geom &g, const Store &s, uint32_t n, uint32_t id0, const float4 *rec4, uint32_t base,
                         cudaStream_t st) {
    if (n == 0) return;
    k_ingest_dense<<<(n + 255) / 256, 256, 0, st>>>(g, s, n, id0, rec4, base);
}

// ------------------------------------------------------------------------------------------------
// k_unpack: birds received from the two neighbouring strips (migrants + next tick's ghosts) join the WRITE
// buffer at [cap, cap+capx) (from the left) and [cap+capx, cap+2capx) (from the right)
// (flockers_mpi/src/model/state.rs:128-130,167-174)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack(const Geom g, const Store s, int phase) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * s.capx) return;
    const int side = t >= s.capx;
    const uint32_t j = side ? t - s.capx : t;
    uint8_t *buf = phase == 0 ? (side ? s.recvR : s.recvL) : (side ? s.recvU : s.recvD);
    const uint32_t n = min(reinterpret_cast<const XHdr *>(buf)->count, s.capx);
    if (j >= n) return;
    const float4 r = reinterpret_cast<const float4 *>(xbuf_rec(buf))[j];
    const uint32_t id = xbuf_id(buf, s.capx)[j];
    const int cx = cell_coord(r.x, g.d, g.mx + 1);
    const int cy = cell_coord(r.y, g.d, g.dhg);
    const int lc = local_col(g, cx);
    const int lr = local_row(g, cy);
    uint32_t key = KEY_INVALID, slot = 0;
    if (lc != FLK_DROP && lr != FLK_DROP) {
        key = (uint32_t)lc * (uint32_t)g.dh + (uint32_t)lr;
        slot = atomicAdd(s.cnt + key, 1u);
    }
    const uint32_t w = s.cap + (uint32_t)phase * 2u * s.capx + t;
    write_entry(s, w, key, slot, r, id);
    const uint32_t *pay = s.pw ? xbuf_pay(buf, s.capx) + (size_t)j * s.pw : nullptr;
    for (uint32_t k = 0; k < s.pw; ++k) s.tpay[(size_t)w * s.pw + k] = pay[k];
    if (g.mode == 2 && phase == 0) {
        // tiles: a bird that came in from the left/right neighbour and sits in one of my border rows is a ghost (or a
        // migrant) of my lower/upper neighbour too, whose columns are mine: forward it (the diagonal hop)
        const Rec rr = Rec{r.x, r.y, r.z, r.w};
        if (lr == 0 || lr == 1) xbuf_append(s, s.sendD, rr, id, pay);
        if (lr == g.rown || lr == g.rown + 1 || (lr == FLK_DROP && cy >= g.my)) xbuf_append(s, s.sendU, rr, id, pay);
    }
}

void launch_unpack(const Geom &g, const Store &s, cudaStream_t st, int phase) {
    k_unpack<<<(2 * s.capx + 255) / 256, 256, 0, st>>>(g, s, phase);
}

// ------------------------------------------------------------------------------------------------
// k_compact_owned (tiles): the birds a tile owns — columns 1..nown (+ slack column) x rows 1..rown (+ slack row) — are
// scattered over the READ buffer; append them to the output arrays (order: arrival; callers sort by id)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compact_owned(const Geom g, const Store s, uint32_t n_upper, uint32_t *out_id,
                                                       float2 *out_loc, float2 *out_ld, uint32_t *counter, uint32_t *out_pay) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_upper || i >= *s.n_read_dev) return;
    const float4 r = reinterpret_cast<const float4 *>(s.rec)[i];
    const int lc = local_col(g, cell_coord(r.x, g.d, g.mx + 1));
    const int lr = local_row(g, cell_coord(r.y, g.d, g.dhg));
    const bool col_owned = (lc >= 1 && lc <= g.nown) || lc == g.nown + 2;
    const bool row_owned = g.mode != 2 || (lr >= 1 && lr <= g.rown) || lr == g.rown + 2;
    if (!col_owned || !row_owned) return;
    const uint32_t k = atomicAdd(counter, 1u);
    if (out_id) out_id[k] = s.ids[i];
    if (out_loc) out_loc[k] = make_float2(r.x, r.y);
    if (out_ld) out_ld[k] = make_float2(r.z, r.w);
    if (out_pay) for (uint32_t q = 0; q < s.pw; ++q) out_pay[(size_t)k * s.pw + q] = s.pay[(size_t)i * s.pw + q];
}

void launch_compact_owned(const Geom &g, const Store &s, uint32_t n_upper, uint32_t *out_id, float2 *out_loc, float2 *out_ld,
                          uint32_t *counter, cudaStream_t st, uint32_t *out_pay) {
    if (n_upper == 0) return;
    k_compact_owned<<<(n_upper + 255) / 256, 256, 0, st>>>(g, s, n_upper, out_id, out_loc, out_ld, counter, out_pay);
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of cnt[0..m) -> cell_start[0..m), zeroing cnt
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr uint32_t SCAN_WIDE_MIN = 1u << 21;   // cells: above, 16 counters per thread (cnt / cell_start are padded to 2 * SCAN_TILE)

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[THREADS / 32];
    __shared__ uint32_t wtot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < THREADS / 32 ? wsum[lane] : 0;
        uint32_t xi = warp_inclusive_scan(x);
        if (lane < THREADS / 32) wsum[lane] = xi - x;
        if (lane == 31) wtot = xi;
    }
    __syncthreads();
    const uint32_t r = inc - v + wsum[w];
    *total = wtot;
    __syncthreads();
    return r;
}

// One launch (decoupled look-back; cnt / cell_start are allocated padded to a multiple of the widest tile, 2 * SCAN_TILE,
// zero-filled, so whole-tile uint4 accesses are safe): a CTA takes the next tile by ticket, publishes its tile's total,
// sums the totals / prefixes its predecessors have published (they took their tickets earlier, so they are running) and
// publishes its own inclusive prefix.  state[0] = ticket, state[1] = CTAs done, state[2 + t] = {status : value} of tile t
// (status 0 = nothing yet, 1 = tile total, 2 = inclusive prefix); the last CTA to finish clears the state for the next
// launch.  Two launches fewer per tick than reduce / spine / down-sweep (round 1): 2 us of a 40-48 us tick at 60-250 k
// birds, neutral at 16 M.
template <int ITEMS>   // counters per thread: 8 (tiles of 2048 cells) or 16 (4096: half as many look-backs for large fields)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(uint32_t *__restrict__ cnt, uint32_t *__restrict__ cell_start,
                                                               unsigned long long *state, uint32_t nblk) {
    constexpr int Q = ITEMS / 4;
    __shared__ uint32_t sh_tile, sh_excl, sh_last;
    if (threadIdx.x == 0) sh_tile = (uint32_t)atomicAdd(state, 1ull);
    __syncthreads();
    const uint32_t tile = sh_tile;
    volatile unsigned long long *stt = state + 2;
    uint4 *pc = reinterpret_cast<uint4 *>(cnt + (size_t)tile * (SCAN_THREADS * ITEMS)) + threadIdx.x * Q;
    uint4 *ps = reinterpret_cast<uint4 *>(cell_start + (size_t)tile * (SCAN_THREADS * ITEMS)) + threadIdx.x * Q;
    uint4 c[Q];
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        c[q] = pc[q];
        v += c[q].x + c[q].y + c[q].z + c[q].w;
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan<SCAN_THREADS>(v, &total);
    if (threadIdx.x == 0) {
        stt[tile] = ((tile == 0 ? 2ull : 1ull) << 32) | total;
        if (tile == 0) sh_excl = 0;
    }
    if (tile > 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t excl = 0;
        for (int idx = (int)tile - 1;; idx -= 32) {
            const int j = idx - lane;
            unsigned long long w = 2ull << 32;          // before tile 0: an inclusive prefix of zero
            if (j >= 0) do { w = stt[j]; } while ((w >> 32) == 0);
            const unsigned pm = __ballot_sync(0xffffffffu, (w >> 32) == 2);
            const int first = pm ? __ffs(pm) - 1 : 31;  // the nearest predecessor that already knows its prefix
            uint32_t val = lane <= first ? (uint32_t)w : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            excl += val;
            if (pm) break;
        }
        if (lane == 0) {
            stt[tile] = (2ull << 32) | (excl + total);
            sh_excl = excl;
        }
    }
    __syncthreads();
    run += sh_excl;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        uint4 o;
        o.x = run; o.y = o.x + c[q].x; o.z = o.y + c[q].y; o.w = o.z + c[q].z;
        run = o.w + c[q].w;
        ps[q] = o;
        pc[q] = make_uint4(0, 0, 0, 0);   // the new WRITE buffer starts empty
    }
    if (threadIdx.x == 0) {
        // release: this thread's status stores must have landed before the CTA counts as done — the last CTA clears the
        // state words, and a status store arriving after that would survive into the next launch (seen as a stale
        // inclusive prefix there: with four strips sharing one GPU this happened once in a few hundred ticks)
        __threadfence();
        sh_last = atomicAdd(state + 1, 1ull) == (unsigned long long)nblk - 1;
    }
    __syncthreads();
    if (sh_last)   // every CTA is past its look-back
        for (uint32_t i = threadIdx.x; i < nblk + 2; i += SCAN_THREADS) state[i] = 0;
}

void launch_scan(const Store &s, uint32_t m, cudaStream_t st) {
    if (m > SCAN_WIDE_MIN) {
        const uint32_t nblk = (m + 2 * SCAN_TILE - 1) / (2 * SCAN_TILE);
        k_scan_onepass<2 * SCAN_ITEMS><<<nblk, SCAN_THREADS, 0, st>>>(s.cnt, s.cell_start, s.scan_state, nblk);
    } else {
        const uint32_t nblk = (m + SCAN_TILE - 1) / SCAN_TILE;
        k_scan_onepass<SCAN_ITEMS><<<nblk, SCAN_THREADS, 0, st>>>(s.cnt, s.cell_start, s.scan_state, nblk);
    }
}

// ------------------------------------------------------------------------------------------------
// k_scatter / k_rank: WRITE buffer -> READ buffer in (cell, ascending id) order
// ------------------------------------------------------------------------------------------------
// strips / tiles: the arrivals unpacked into the receive regions of the WRITE buffer, [cap + b*capx, cap + b*capx +
// count_b) for message b; one slot per thread, t = index into the receive regions
__device__ __forceinline__ void scatter_received(const Store &s, uint32_t t) {
    if (t >= s.nrecv * s.capx) return;
    const uint32_t b = t / s.capx;
    const uint8_t *buf = b == 0 ? s.recvL : (b == 1 ? s.recvR : (b == 2 ? s.recvD : s.recvU));
    if (t - b * s.capx >= min(reinterpret_cast<const XHdr *>(buf)->count, s.capx)) return;
    const uint32_t i = s.cap + t;
    const uint32_t key = s.tkey[i];
    if (key == KEY_INVALID) return;
    const uint32_t p = s.cell_start[key] + s.tslot[i];
    if (p >= s.cap) {
        atomicOr(s.err_flag, ERR_CAPACITY);
        return;
    }
    s.stage[p] = make_uint4(s.tid[i], i, key, 0u);   // one 16-byte store instead of three scattered 4-byte ones
}

// four consecutive stepped/uploaded slots per thread: three 16-byte loads, then four independent gather -> store
// chains in flight (the one-slot-per-thread form reached 42 % of DRAM bandwidth: too few bytes in flight per thread)
#ifndef FLK_SCATTER_E
#define FLK_SCATTER_E 4   // WRITE slots per thread (multiple of 4)
#endif
__global__ void __launch_bounds__(256) k_scatter4(const Store s, uint32_t n_limit, int use_dev_count, uint32_t nb_stepped) {
    constexpr int E = FLK_SCATTER_E;
    if (blockIdx.x >= nb_stepped) {   // strips / tiles: the CTAs behind the stepped slots take the receive regions
        scatter_received(s, (blockIdx.x - nb_stepped) * blockDim.x + threadIdx.x);
        return;
    }
    const uint32_t n = use_dev_count ? min(*s.n_read_dev, n_limit) : n_limit;
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * (uint32_t)E;
    if (i0 >= n) return;
    uint32_t key[E], slot[E], id[E];
    if (i0 + E <= n) {
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
            const uint4 k4 = *reinterpret_cast<const uint4 *>(s.tkey + i0 + 4 * q);
            const uint4 s4 = *reinterpret_cast<const uint4 *>(s.tslot + i0 + 4 * q);
            const uint4 d4 = *reinterpret_cast<const uint4 *>(s.tid + i0 + 4 * q);
            key[4 * q] = k4.x; key[4 * q + 1] = k4.y; key[4 * q + 2] = k4.z; key[4 * q + 3] = k4.w;
            slot[4 * q] = s4.x; slot[4 * q + 1] = s4.y; slot[4 * q + 2] = s4.z; slot[4 * q + 3] = s4.w;
            id[4 * q] = d4.x; id[4 * q + 1] = d4.y; id[4 * q + 2] = d4.z; id[4 * q + 3] = d4.w;
        }
    } else {
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const bool ok = i0 + t < n;
            key[t] = ok ? s.tkey[i0 + t] : KEY_INVALID;
            slot[t] = ok ? s.tslot[i0 + t] : 0u;
            id[t] = ok ? s.tid[i0 + t] : 0u;
        }
    }
    uint32_t base[E];
#pragma unroll
    for (int t = 0; t < E; ++t) base[t] = key[t] != KEY_INVALID ? __ldg(s.cell_start + key[t]) : 0u;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        if (key[t] == KEY_INVALID) continue;
        const uint32_t p = base[t] + slot[t];
        if (p >= s.cap) {
            atomicOr(s.err_flag, ERR_CAPACITY);
            continue;
        }
        s.stage[p] = make_uint4(id[t], i0 + t, key[t], 0u);
    }
}

void launch_scatter(const Store &s, uint32_t n, int dist, cudaStream_t st) {
    if (dist) {
        // stepped slots [0, n) (n = host upper bound of the READ population), then the two receive regions
        const uint32_t nn = n < s.cap ? n : s.cap;
        const uint32_t nb = ((nn + FLK_SCATTER_E - 1) / FLK_SCATTER_E + 255) / 256;
        k_scatter4<<<nb + (s.nrecv * s.capx + 255) / 256, 256, 0, st>>>(s, nn, 1, nb);
        return;
    }
    if (n == 0) return;
    const uint32_t nb = ((n + FLK_SCATTER_E - 1) / FLK_SCATTER_E + 255) / 256;
    k_scatter4<<<nb, 256, 0, st>>>(s, n, 0, nb);
}

// Tables of the column-aligned k_step launches (single-GPU handles; one CTA, after cell_start is final).  A TILE column is
// one whose 3x3 windows never touch the x seam: 1 .. mx-2 (ncols = mx + 1 with the slack column).
//   colchunk[c]              k_step CTAs (chunks of FLK_STEP_BS birds) needed by the tile columns before c; [ncols] = all
//   colchunk[ncols + 1 + c]  birds in the seam rows (0, my-1 and the slack row my) of the tile columns before c; [.. + ncols] = all
template <int THREADS = 256>
__device__ __forceinline__ void build_colchunk(const Store &s, uint32_t ncols, uint32_t dh) {
    __shared__ uint32_t sh_carry[2];
    if (threadIdx.x == 0) { sh_carry[0] = 0; sh_carry[1] = 0; }
    __syncthreads();
    uint32_t *colseam = s.colchunk + ncols + 1;
    for (uint32_t c0 = 0; c0 < ncols; c0 += blockDim.x) {
        const uint32_t c = c0 + threadIdx.x;
        uint32_t ch = 0, sm = 0;
        if (c >= 1 && c + 2 < ncols) {   // tile column
            const uint32_t *cs = s.cell_start + (size_t)c * dh;
            ch = (cs[dh] - cs[0] + FLK_STEP_BS - 1) / FLK_STEP_BS;
            sm = (cs[1] - cs[0]) + (cs[dh] - cs[dh - 2]);
        }
        uint32_t total, total2;
        const uint32_t ex = block_exclusive_scan<THREADS>(ch, &total);
        const uint32_t ex2 = block_exclusive_scan<THREADS>(sm, &total2);
        if (c < ncols) { s.colchunk[c] = sh_carry[0] + ex; colseam[c] = sh_carry[1] + ex2; }
        __syncthreads();
        if (threadIdx.x == 0) { sh_carry[0] += total; sh_carry[1] += total2; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { s.colchunk[ncols] = sh_carry[0]; colseam[ncols] = sh_carry[1]; }
}

// two staged entries per thread: both record gathers and both cell-range loads are in flight before either member
// loop starts (the one-entry form ran at 61 % of DRAM bandwidth)
#ifndef FLK_RANK_MINBLOCKS
#define FLK_RANK_MINBLOCKS 8   // 32 registers, full occupancy: 0.138 ms against 0.148 (6 CTAs/SM) and 0.158 (4-5)
#endif
__global__ void __launch_bounds__(256, FLK_RANK_MINBLOCKS) k_rank(const Store s, uint32_t ncell, int bump_step, uint32_t ncols,
                                                                  uint32_t dh) {
    if (blockIdx.x == gridDim.x - 1 && ncols) {   // an extra CTA at the end of the grid
        build_colchunk(s, ncols, dh);
        return;
    }
    const uint32_t n = min(s.cell_start[ncell], s.cap);
    const uint32_t p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2u;
    if (p0 == 0) {
        *s.n_read_dev = n;
        if (bump_step) *s.step_dev += 1;
    }
    if (p0 >= n) return;
    const bool two = p0 + 1 < n;
    const uint4 m0 = s.stage[p0];
    const uint4 m1 = two ? s.stage[p0 + 1] : m0;
    const float4 r0 = reinterpret_cast<const float4 *>(s.trec)[m0.y];
    const float4 r1 = reinterpret_cast<const float4 *>(s.trec)[m1.y];
    const uint32_t lo0 = __ldg(s.cell_start + m0.z), hi0 = min(__ldg(s.cell_start + m0.z + 1), s.cap);
    const uint32_t lo1 = __ldg(s.cell_start + m1.z), hi1 = min(__ldg(s.cell_start + m1.z + 1), s.cap);
    uint32_t k0 = 0, k1 = 0;
    for (uint32_t q = lo0; q < hi0; ++q) {
        const uint32_t other = s.stage[q].x;
        k0 += (other < m0.x) || (other == m0.x && q < p0);
    }
    if (two) {
        for (uint32_t q = lo1; q < hi1; ++q) {
            const uint32_t other = s.stage[q].x;
            k1 += (other < m1.x) || (other == m1.x && q < p0 + 1);
        }
    }
    const uint32_t d0 = lo0 + k0;
    reinterpret_cast<float4 *>(s.rec)[d0] = r0;
    s.ids[d0] = m0.x;
    for (uint32_t k = 0; k < s.pw; ++k) s.pay[(size_t)d0 * s.pw + k] = s.tpay[(size_t)m0.y * s.pw + k];
    if (two) {
        const uint32_t d1 = lo1 + k1;
        reinterpret_cast<float4 *>(s.rec)[d1] = r1;
        s.ids[d1] = m1.x;
        for (uint32_t k = 0; k < s.pw; ++k) s.pay[(size_t)d1 * s.pw + k] = s.tpay[(size_t)m1.y * s.pw + k];
    }
}

void launch_rank(const Store &s, uint32_t n_upper, uint32_t ncell, int bump_step, cudaStream_t st, uint32_t ncols, uint32_t dh) {
    const uint32_t nb = n_upper == 0 ? 1 : ((n_upper + 1) / 2 + 255) / 256;
    if (!s.colchunk) ncols = 0;
    k_rank<<<nb + (ncols ? 1 : 0), 256, 0, st>>>(s, ncell, bump_step, ncols, dh);
}

// ------------------------------------------------------------------------------------------------
// k_rebin_small: the whole rebin (scan, scatter, rank, colchunk) of a small single-GPU field in ONE CTA — the shipped
// 1000-bird run (flockers/src/main.rs:40-44) spends its tick in launch latencies, not in work: one launch instead of five.
// Same arithmetic and the same (cell, ascending id) result as k_scan_onepass + k_scatter4 + k_rank.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_rebin_small(const Store s, uint32_t ncell, uint32_t n_write, int bump_step,
                                                      uint32_t ncols, uint32_t dh) {
    const uint32_t m = ncell + 1;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < m; base += 1024) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < m ? s.cnt[idx] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan<1024>(v, &total);
        if (idx < m) {
            s.cell_start[idx] = ex + carry;
            s.cnt[idx] = 0;                    // the new WRITE buffer starts empty
        }
        carry += total;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_write; i += 1024) {
        const uint32_t key = s.tkey[i];
        if (key == KEY_INVALID) continue;
        const uint32_t p = s.cell_start[key] + s.tslot[i];
        if (p >= s.cap) {
            atomicOr(s.err_flag, ERR_CAPACITY);
            continue;
        }
        s.stage[p] = make_uint4(s.tid[i], i, key, 0u);
    }
    __syncthreads();
    const uint32_t n = min(carry, s.cap);
    if (threadIdx.x == 0) {
        *s.n_read_dev = n;
        if (bump_step) *s.step_dev += 1;
    }
    for (uint32_t p0 = threadIdx.x; p0 < n; p0 += 1024) {
        const uint4 m0 = s.stage[p0];
        const float4 r0 = reinterpret_cast<const float4 *>(s.trec)[m0.y];
        const uint32_t lo = s.cell_start[m0.z], hi = min(s.cell_start[m0.z + 1], s.cap);
        uint32_t k0 = 0;
        for (uint32_t q = lo; q < hi; ++q) {
            const uint32_t other = s.stage[q].x;
            k0 += (other < m0.x) || (other == m0.x && q < p0);
        }
        const uint32_t d0 = lo + k0;
        reinterpret_cast<float4 *>(s.rec)[d0] = r0;
        s.ids[d0] = m0.x;
        for (uint32_t k = 0; k < s.pw; ++k) s.pay[(size_t)d0 * s.pw + k] = s.tpay[(size_t)m0.y * s.pw + k];
    }
    if (ncols) build_colchunk<1024>(s, ncols, dh);
}

void launch_rebin_small(const Store &s, uint32_t ncell, uint32_t n_write, int bump_step, cudaStream_t st, uint32_t ncols,
                        uint32_t dh) {
    // the column tables are only read by k_step's column-aligned launches, which a population this small never gets
    // (k_step_warp takes it)
    if (!s.colchunk || n_write <= step_warp_max()) ncols = 0;
    k_rebin_small<<<1, 1024, 0, st>>>(s, ncell, n_write, bump_step, ncols, dh);
}

// ------------------------------------------------------------------------------------------------
// k_compact: slot-staged WRITE buffer -> READ buffer.  Eight lanes per cell, one staged slot (two when a cell of the warp
// holds more than eight arrivals) per lane: the lane ranks its id among the cell's ids with shuffles and stores its
// record at cell_start[cell] + rank.  One pass over the records instead of k_scatter + k_rank: the record is read where
// Bird::step left it and written once, in place.
// ------------------------------------------------------------------------------------------------
// a cell with spilled arrivals (rare): one thread merges the slots and the cell's spill entries by (id, source)
__device__ __noinline__ void compact_spilled_cell(const Store &s, uint32_t cell, uint32_t cs, uint32_t n) {
    const uint32_t C = (uint32_t)s.sw_C;
    const size_t base = (size_t)cell * C;
    const uint32_t nsp = min(*s.sw_spill_cnt, s.sw_spill_cap);
    long long prev_id = -1;
    uint32_t prev_src = 0;
    for (uint32_t k = 0; k < n; ++k) {
        long long best_id = 0x100000000ll;
        uint32_t best_src = 0xffffffffu;
        for (uint32_t q = 0; q < C; ++q) {
            const long long id = s.sw_id[base + q];
            const bool after = id > prev_id || (id == prev_id && q > prev_src);
            if (after && (id < best_id || (id == best_id && q < best_src))) { best_id = id; best_src = q; }
        }
        for (uint32_t j = 0; j < nsp; ++j) {
            if (s.sw_spill[j].cell != cell) continue;
            const long long id = s.sw_spill[j].id;
            const uint32_t src = C + j;
            const bool after = id > prev_id || (id == prev_id && src > prev_src);
            if (after && (id < best_id || (id == best_id && src < best_src))) { best_id = id; best_src = src; }
        }
        if (best_src == 0xffffffffu) return;
        const uint32_t d = cs + k;
        if (d < s.cap) {
            reinterpret_cast<float4 *>(s.rec)[d] = best_src < C ? s.sw_rec[base + best_src] : s.sw_spill[best_src - C].rec;
            s.ids[d] = (uint32_t)best_id;
        } else {
            atomicOr(s.err_flag, ERR_CAPACITY);
        }
        prev_id = best_id; prev_src = best_src;
    }
}

// (__grid_constant__: the out-of-line spill merge takes the store by reference; without it every thread would first copy
// the whole parameter struct to its stack)
__global__ void __launch_bounds__(256) k_compact(const __grid_constant__ Store s, uint32_t ncell, int bump_step, uint32_t ncols,
                                                 uint32_t dh) {
    if (blockIdx.x == gridDim.x - 1 && ncols) {   // an extra CTA at the end of the grid
        build_colchunk(s, ncols, dh);
        return;
    }
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cell = gid >> 3, l8 = gid & 7u;
    const int lane = threadIdx.x & 31, grp = lane & 24;
    if (gid == 0) {
        *s.n_read_dev = min(s.cell_start[ncell], s.cap);
        if (bump_step) *s.step_dev += 1;
        s.sw_totals[1] = min(*s.sw_spill_cnt, s.sw_spill_cap);
    }
    uint32_t cs = 0, n = 0;
    if (cell < ncell) {
        cs = __ldg(s.cell_start + cell);
        n = __ldg(s.cell_start + cell + 1) - cs;
    }
    const uint32_t C = (uint32_t)s.sw_C;
    const bool spilled = n > C;
    const uint32_t ns = spilled ? 0u : n;                  // slots this group ranks (a spilled cell is merged by one thread)
    uint32_t nmax = ns;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    if (nmax) {   // warp-uniform
        const size_t base = (size_t)cell * C;
        const bool two = nmax > 8;
        // a lane without a slot holds the largest id: it never counts as "smaller" for anybody
        uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu;
        float4 r0 = make_float4(0, 0, 0, 0), r1 = r0;
        const bool h0 = l8 < ns, h1 = two && l8 + 8 < ns;
        if (h0) { id0 = s.sw_id[base + l8]; r0 = s.sw_rec[base + l8]; }
        if (h1) { id1 = s.sw_id[base + l8 + 8]; r1 = s.sw_rec[base + l8 + 8]; }
        // rank among the group's first eight slots (equal ids: lower slot first, so the ranks stay a permutation)
        uint32_t k0 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t o = __shfl_sync(0xffffffffu, id0, grp | j);
            k0 += (o < id0 || (o == id0 && (uint32_t)j < l8)) ? 1u : 0u;
        }
        if (!h0) k0 = 0;
        uint32_t k1 = 0;
        if (two) {   // warp-uniform, rare: a cell of the warp holds more than eight arrivals
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t o0 = __shfl_sync(0xffffffffu, id0, grp | j);   // first set seen by my second slot
                const uint32_t o1 = __shfl_sync(0xffffffffu, id1, grp | j);   // second set
                k1 += (o0 < id1 || o0 == id1) ? 1u : 0u;                      // slot j < 8 <= my second slot
                k0 += (o1 < id0) ? 1u : 0u;                                   // slot j + 8 > my first slot
                k1 += (o1 < id1 || (o1 == id1 && (uint32_t)j < l8)) ? 1u : 0u;
            }
            // (a lane without a slot holds 0xffffffff and sits at a higher slot number than any real one of its group,
            //  so it is never counted, not even against a real id 0xffffffff)
        }
        if (h0) {
            const uint32_t d = cs + k0;
            if (d < s.cap) { reinterpret_cast<float4 *>(s.rec)[d] = r0; s.ids[d] = id0; }
            else atomicOr(s.err_flag, ERR_CAPACITY);
        }
        if (h1) {
            const uint32_t d = cs + k1;
            if (d < s.cap) { reinterpret_cast<float4 *>(s.rec)[d] = r1; s.ids[d] = id1; }
            else atomicOr(s.err_flag, ERR_CAPACITY);
        }
    }
    if (spilled && l8 == 0) compact_spilled_cell(s, cell, cs, n);
}

void launch_compact(const Store &s, uint32_t ncell, int bump_step, cudaStream_t st, uint32_t ncols, uint32_t dh) {
    const uint64_t threads = (uint64_t)ncell * 8;
    if (!s.colchunk) ncols = 0;
    k_compact<<<(unsigned)((threads + 255) / 256) + (ncols ? 1 : 0), 256, 0, st>>>(s, ncell, bump_step, ncols, dh);
}

// ------------------------------------------------------------------------------------------------
// diagnostics: div2_rn against __fdiv_rn on pseudo-random operands shaped like the hot loop's
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_selftest_div2(uint64_t n, uint64_t seed, unsigned long long *mismatches) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t o0, o1;
    philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 0x5e1f, 0, (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1);
    uint32_t o2, o3;
    philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 0x5e20, 0, (uint32_t)seed, (uint32_t)(seed >> 32), o2, o3);
    // numerators: random sign/mantissa, exponent in [2^-70, 2^16] (covers the guard edge), some zeros/denormals
    auto mk = [](uint32_t u, int emin, int span) {
        const uint32_t e = (uint32_t)(127 + emin) + (u >> 23) % (uint32_t)span;
        return __uint_as_float((u & 0x807fffffu) | (e << 23));
    };
    float a = mk(o0, -70, 87), b = mk(o1, -70, 87);
    if ((o2 & 0xff) == 0) a = 0.0f;
    if ((o2 & 0xff) == 1) b = __uint_as_float(o3 & 0x807fffffu);   // denormal
    float d;
    if (o2 & 0x100) {                       // d = sq*sq+1 from the numerators, as in bird.rs:59-61
        const float sq = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
        d = __fadd_rn(__fmul_rn(sq, sq), 1.0f);
    } else {
        d = mk(o3, 0, 64);                  // [1, 2^64): crosses the 2^60 guard
    }
    float qa, qb;
    div2_rn(a, b, d, qa, qb);
    const float ra = __fdiv_rn(a, d), rb = __fdiv_rn(b, d);
    const bool ok = (__float_as_uint(qa) == __float_as_uint(ra) || (qa != qa && ra != ra)) &&
                    (__float_as_uint(qb) == __float_as_uint(rb) || (qb != qb && rb != rb));
    if (!ok) atomicAdd(mismatches, 1ull);
}

void launch_selftest_div2(uint64_t n, uint64_t seed, unsigned long long *mismatches, cudaStream_t st) {
    k_selftest_div2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seed, mismatches);
}

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t *p, uint32_t v, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

void launch_fill_u32(uint32_t *p, uint32_t v, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    k_fill_u32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
}

// ------------------------------------------------------------------------------------------------
// Field2D::get / get_location: look one id up in the READ buffer
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_find(const Store s, uint32_t id, float4 *out, uint32_t *found) {
    const uint32_t n = *s.n_read_dev;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || s.ids[i] != id) return;
    if (atomicAdd(found, 1u) == 0) *out = reinterpret_cast<const float4 *>(s.rec)[i];
}

void launch_find(const Store &s, uint32_t n_upper, uint32_t id, float4 *out, uint32_t *found, cudaStream_t st) {
    if (n_upper == 0) return;
    k_find<<<(n_upper + 255) / 256, 256, 0, st>>>(s, id, out, found);
}

// ------------------------------------------------------------------------------------------------
// snapshot helper: split READ records into loc / last_d arrays
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split(const Store s, uint32_t n, float2 *loc, float2 *ld) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r = reinterpret_cast<const float4 *>(s.rec)[i];
    if (loc) loc[i] = make_float2(r.x, r.y);
    if (ld) ld[i] = make_float2(r.z, r.w);
}

void launch_split(const Store &s, uint32_t n, float2 *loc, float2 *ld, cudaStream_t st) {
    if (n == 0) return;
    k_split<<<(n + 255) / 256, 256, 0, st>>>(s, n, loc, ld);
}

// dense snapshot: READ record of object `id` -> out[id - id0] (ids outside [id0, id0+n) are counted, not written)
__global__ void __launch_bounds__(256) k_gather_dense(const Store s, uint32_t n_upper, uint32_t id0, uint32_t n,
                                                      float4 *__restrict__ out, uint32_t *__restrict__ outside) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_upper || i >= *s.n_read_dev) return;
    const uint32_t k = s.ids[i] - id0;
    if (k < n) out[k] = reinterpret_cast<const float4 *>(s.rec)[i];
    else atomicAdd(outside, 1u);
}

void launch_gather_dense(const Store &s, uint32_t n_upper, uint32_t id0, uint32_t n, float4 *out, uint32_t *outside,
                         cudaStream_t st) {
    if (n_upper == 0) return;
    k_gather_dense<<<(n_upper + 255) / 256, 256, 0, st>>>(s, n_upper, id0, n, out, outside);
}

// ------------------------------------------------------------------------------------------------
// queries (Field2D::get_neighbors_within_{relax_,}distance): one thread per query location
// ------------------------------------------------------------------------------------------------
// FILL: 0 = count, 1 = ids, 2 = whole objects {id, loc.x, loc.y, last_d.x, last_d.y} (20 bytes, Bird's wire form)
template <int FILL>
__global__ void __launch_bounds__(128) k_query(const Geom g, const Store s, uint32_t nq, const float2 *__restrict__ loc,
                                               float dist, int relax, uint32_t *__restrict__ counts,
                                               const uint64_t *__restrict__ offsets, uint32_t *__restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float2 l = loc[q];
    uint32_t n = 0;
    if (dist > 0.0f) {
        const int k = __float2int_rz(floorf(__fdiv_rn(dist, g.d))) + g.wplus1;
        const int gcx = cell_coord(l.x, g.d, g.mx + 1);
        const int cy = local_row(g, cell_coord(l.y, g.d, g.dhg));   // row of the local cell arrays
        const int lc = local_col(g, gcx);
        const float halfW = __fdiv_rn(g.W, 2.0f), halfH = __fdiv_rn(g.H, 2.0f);
        uint32_t *dst = FILL ? out + offsets[q] * (FILL == 2 ? 5u : 1u) : nullptr;
        if (lc != FLK_DROP && cy != FLK_DROP) {
            for_each_window_run(g, s.cell_start, lc, cy, k, [&](uint32_t rs, uint32_t re) {
                for (uint32_t j = rs; j < re; ++j) {
                    if (!relax) {
                        const float4 o = reinterpret_cast<const float4 *>(s.rec)[j];
                        float dx, dy;
                        if (g.toroidal) {
                            dx = toroidal_distance(l.x, o.x, g.W, halfW);
                            dy = toroidal_distance(l.y, o.y, g.H, halfH);
                        } else {
                            dx = __fadd_rn(l.x, -o.x);
                            dy = __fadd_rn(l.y, -o.y);
                        }
                        const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                        if (!(__fsqrt_rn(sq) <= dist)) continue;
                    }
                    if (FILL == 1) dst[n] = s.ids[j];
                    if (FILL == 2) {
                        const float4 o = reinterpret_cast<const float4 *>(s.rec)[j];
                        uint32_t *d5 = dst + 5u * n;
                        d5[0] = s.ids[j];
                        d5[1] = __float_as_uint(o.x); d5[2] = __float_as_uint(o.y);
                        d5[3] = __float_as_uint(o.z); d5[4] = __float_as_uint(o.w);
                    }
                    ++n;
                }
            });
        }
    }
    if (FILL == 0) counts[q] = n;
}

void launch_query_count(const Geom &g, const Store &s, uint32_t nq, const float2 *loc, float dist, int relax,
                        uint32_t *counts, cudaStream_t st) {
    if (nq == 0) return;
    k_query<0><<<(nq + 127) / 128, 128, 0, st>>>(g, s, nq, loc, dist, relax, counts, nullptr, nullptr);
}

void launch_query_fill(const Geom &g, const Store &s, uint32_t nq, const float2 *loc, float dist, int relax,
                       const uint64_t *offsets, uint32_t *out, cudaStream_t st, int objects) {
    if (nq == 0) return;
    if (objects) k_query<2><<<(nq + 127) / 128, 128, 0, st>>>(g, s, nq, loc, dist, relax, nullptr, offsets, out);
    else k_query<1><<<(nq + 127) / 128, 128, 0, st>>>(g, s, nq, loc, dist, relax, nullptr, offsets, out);
}

}  // namespace flk
