// flk_slots.cu — sm_100a kernels of the fixed-capacity-cell ("slotted") double buffer (flk_slots.cuh).
//
//   k_slot_step     Bird::step (flockers/src/model/bird.rs:27-145) for the birds of a range of cell columns; the successor
//                   is stored straight into its cell of the next READ buffer (Field2D::set_object_location, bird.rs:142-144)
//   k_slot_colscan  Field2D::lazy_update (flockers/src/model/state.rs:54-56): one pass over the per-cell counters
//   k_slot_ingest*  bulk set_object_location (state.rs:49);  k_slot_unpack: arrivals of the strip exchange
//   k_slot_materialize / k_slot_gather_dense: the read side (snapshots, queries go through the sorted-runs view)
//
// No tensor cores: the path has no dense contraction.
#include <cstdlib>

#include "flk_slots.cuh"
#include "flk_step.cuh"

namespace flk {


// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan of one value per thread (returns the exclusive prefix, the total in *total)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[THREADS / 32];
    __shared__ uint32_t wtot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t x = lane < THREADS / 32 ? wsum[lane] : 0;
        const uint32_t xi = warp_incl_scan(x);
        if (lane < THREADS / 32) wsum[lane] = xi - x;
        if (lane == 31) wtot = xi;
    }
    __syncthreads();
    const uint32_t r = inc - v + wsum[w];
    *total = wtot;
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------------
// WRITE side: Field2D::set_object_location — claim the next slot of the cell, or spill
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void slot_insert(const SlotArgs &S, uint32_t *err_flag, uint32_t cell, const float4 &r,
                                            uint32_t id) {
    const uint32_t slot = atomicAdd(S.wcnt + cell, 1u);
    if (slot < (uint32_t)S.C) {
        const size_t w = (size_t)cell * (uint32_t)S.C + slot;
        S.wrec[w] = r;
        S.wid[w] = id;
    } else {
        const uint32_t k = atomicAdd(S.wspill_cnt, 1u);
        if (k < S.spill_cap) {
            SpillEntry e;
            e.rec = r; e.id = id; e.cell = cell; e.pad0 = 0; e.pad1 = 0;
            S.wspill[k] = e;
        } else {
            atomicOr(err_flag, ERR_CAPACITY);
        }
    }
}

// successor of a stepped bird: its cell in the local grid (+ the flockers_mpi routing of strip mode, see emit_successor
// in flk_kernels.cu: bird.rs:190-195, state.rs:113-175)
__device__ __forceinline__ void slot_emit(const Geom &g, const SlotArgs &S, const Store &s, const Rec &out, uint32_t id,
                                          bool d_ok, const Recip &Rd) {
    const int ncx = d_ok ? cell_coord_fast(out.x, Rd, g.mx + 1) : cell_coord(out.x, g.d, g.mx + 1);
    const int ncy = d_ok ? cell_coord_fast(out.y, Rd, g.dh) : cell_coord(out.y, g.d, g.dh);
    const int nlc = local_col(g, ncx);
    if (nlc != FLK_DROP)
        slot_insert(S, s.err_flag, (uint32_t)nlc * (uint32_t)g.dh + (uint32_t)ncy, make_float4(out.x, out.y, out.ldx, out.ldy), id);
    if (g.mode != 0) {
        if (nlc == 0 || nlc == 1) xbuf_append(s, s.sendL, out, id);
        if (nlc == g.nown || nlc == g.nown + 1 || (nlc == FLK_DROP && ncx >= g.mx)) xbuf_append(s, s.sendR, out, id);
    }
}

// ------------------------------------------------------------------------------------------------
// READ side: the entries of one cell in ascending-id order (the canonical bag order)
// ------------------------------------------------------------------------------------------------
// slow path: the cell holds more than C entries, the excess sits in the spill list.  Selection by (id, source index):
// O(n * (C + spilled)) per visit, reached only by the handful of cells that overflow before the host changes the layout.
template <class F>
__device__ __noinline__ void for_each_entry_spilled(const SlotArgs &S, uint32_t cell, uint32_t n, F &f) {
    const uint32_t C = (uint32_t)S.C;
    const size_t base = (size_t)cell * C;
    const uint32_t nsp = min(*S.rspill_cnt, S.spill_cap);
    long long prev_id = -1;
    uint32_t prev_src = 0;
    for (uint32_t k = 0; k < n; ++k) {
        long long best_id = 0x100000000ll;
        uint32_t best_src = 0xffffffffu;
        for (uint32_t q = 0; q < C; ++q) {
            const long long id = S.rid[base + q];
            const bool after = id > prev_id || (id == prev_id && q > prev_src);
            if (after && (id < best_id || (id == best_id && q < best_src))) { best_id = id; best_src = q; }
        }
        for (uint32_t j = 0; j < nsp; ++j) {
            if (S.rspill[j].cell != cell) continue;
            const long long id = S.rspill[j].id;
            const uint32_t src = C + j;
            const bool after = id > prev_id || (id == prev_id && src > prev_src);
            if (after && (id < best_id || (id == best_id && src < best_src))) { best_id = id; best_src = src; }
        }
        if (best_src == 0xffffffffu) return;
        if (best_src < C) f(S.rrec[base + best_src], (uint32_t)best_id);
        else f(S.rspill[best_src - C].rec, (uint32_t)best_id);
        prev_id = best_id; prev_src = best_src;
    }
}

template <class F>
__device__ __forceinline__ void for_each_entry(const SlotArgs &S, uint32_t cell, F &&f) {
    const uint32_t n = __ldg(S.rcnt + cell);
    if (n == 0) return;
    const uint32_t C = (uint32_t)S.C;
    if (n <= C) {
        const size_t base = (size_t)cell * C;
        const unsigned long long perm = __ldg(reinterpret_cast<const unsigned long long *>(S.rperm) + cell);
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t slot = (uint32_t)(perm >> (4 * k)) & 15u;
            f(__ldg(S.rrec + base + slot), __ldg(S.rid + base + slot));
        }
        return;
    }
    for_each_entry_spilled(S, cell, n, f);
}

// the kk-th bird of a cell for the thread <-> bird map.  A cell may be split between two CTAs, one on the tile path and
// one here, so both must number its birds the same way: ascending id (perm) while the cell fits its slots; a spilled cell
// (never on the tile path) uses raw slot order, then its spilled entries in list order.
__device__ __forceinline__ bool slot_fetch(const SlotArgs &S, uint32_t cell, uint32_t n, uint32_t kk, float4 &r, uint32_t &id) {
    const uint32_t C = (uint32_t)S.C;
    if (n <= C) {
        const unsigned long long perm = __ldg(reinterpret_cast<const unsigned long long *>(S.rperm) + cell);
        const uint32_t slot = (uint32_t)(perm >> (4 * kk)) & 15u;
        r = S.rrec[(size_t)cell * C + slot];
        id = S.rid[(size_t)cell * C + slot];
        return true;
    }
    if (kk < C) {
        r = S.rrec[(size_t)cell * C + kk];
        id = S.rid[(size_t)cell * C + kk];
        return true;
    }
    const uint32_t nsp = min(*S.rspill_cnt, S.spill_cap);
    uint32_t seen = C;
    for (uint32_t j = 0; j < nsp; ++j) {
        if (S.rspill[j].cell != cell) continue;
        if (seen == kk) { r = S.rspill[j].rec; id = S.rspill[j].id; return true; }
        ++seen;
    }
    return false;
}

// cells of the (2k+1)^2 window of (lc, cy) in the reference's query order (x-outer, y-inner)
template <class F>
__device__ __forceinline__ void for_each_window_cell(const Geom &g, int lc, int cy, int k, F &&f) {
    for (int ox = -k; ox <= k; ++ox) {
        const int col = window_col(g, lc, ox);
        if (col < 0) continue;
        for (int oy = -k; oy <= k; ++oy) {
            int r = cy + oy;
            if (g.toroidal) {
                r = wrap_row(g, r);
            } else if (r < 0 || r > g.dh - 1) {
                continue;
            }
            if (r < 0) continue;
            f((uint32_t)col * (uint32_t)g.dh + (uint32_t)r);
        }
    }
}

// generic neighbour accumulation (bird.rs:44-71): any window width, seam, either query; the bird itself is left out by
// id (bird.rs:53)
template <bool RELAX, int MATH>
__device__ __forceinline__ void slot_accumulate_window(const Geom &g, const Params &p, const SlotArgs &S, const Rec &me,
                                                       uint32_t my_id, int lc, int cy, Acc &acc) {
    const float W = g.W, H = g.H;
    const float halfW = __fdiv_rn(W, 2.0f), halfH = __fdiv_rn(H, 2.0f);
    float xa = 0.0f, ya = 0.0f, xc = 0.0f, yc = 0.0f, xs = 0.0f, ys = 0.0f;
    uint32_t nvec = 0;
    int count = 0;
    for_each_window_cell(g, lc, cy, p.k, [&](uint32_t cell) {
        for_each_entry(S, cell, [&](const float4 &o, uint32_t oid) {
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
                if (!(__fsqrt_rn(qs) <= p.radius)) return;
            }
            ++nvec;
            if (oid == my_id) return;                                  // :53
            ++count;                                                   // :56
            const float den = __fadd_rn(__fmul_rn(sq, sq), 1.0f);
            float qa, qb;
            if (MATH == 0) {
                div2_rn(dx, dy, den, qa, qb);                          // :60-61, two IEEE divisions
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
        });
    });
    acc.xa = xa; acc.ya = ya; acc.xc = xc; acc.yc = yc; acc.xs = xs; acc.ys = ys;
    acc.nvec = nvec; acc.count = count;
}

// Everything that is not the common case (seam, slack cells, spilled cells, other windows / queries, non-toroidal fields,
// the identity pass of a strip commit): one bird, out of line
template <bool RELAX, int MATH>
__device__ __noinline__ void slot_step_bird_general(const Geom &g, const Params &p, const SlotArgs &S, const Store &s,
                                                    const StepArgs &a, int lc, int row, uint32_t n, uint32_t kk) {
    const uint32_t cell = (uint32_t)lc * (uint32_t)g.dh + (uint32_t)row;
    float4 r;
    uint32_t my_id;
    if (!slot_fetch(S, cell, n, kk, r, my_id)) return;
    const Rec me = Rec{r.x, r.y, r.z, r.w};
    const bool d_ok = g.d >= 0x1p-20f && g.d <= 4096.0f;
    const Recip Rd = make_recip(g.d);
    if (a.identity) {   // dist commit: re-stage the bird unchanged (builds the first halo)
        slot_emit(g, S, s, me, my_id, d_ok, Rd);
        return;
    }
    Acc acc;
    if (p.radius > 0.0f) slot_accumulate_window<RELAX, MATH>(g, p, S, me, my_id, lc, row, acc);   // dist <= 0: empty result
    const Rec out = bird_finish<MATH>(g, p, a, s.step_dev, me, my_id, acc);
    slot_emit(g, S, s, out, my_id, d_ok, Rd);
}

// global cell column of a local one (strips: ghosts materialise the wrap; the slack column is global column mx)
__device__ __forceinline__ int global_col(const Geom &g, int lc) {
    if (g.mode == 0) return lc;
    if (lc == g.nown + 2) return g.mx;
    int c = g.col0 + lc - 1;
    if (c < 0) c += g.mx;
    else if (c >= g.mx) c -= g.mx;
    return c;
}

// ------------------------------------------------------------------------------------------------
// k_slot_step: grid (1 + chunk bound, columns).  blockIdx.y picks the cell column; blockIdx.x = 0 steps the column's
// seam-row cells (FAST launches), blockIdx.x = m + 1 steps chunk m: SLOT_CHUNK consecutive birds of the column
// (column-local order: row, id).  The last CTA of a column also takes the chunks beyond the grid's bound, if any.
// FAST = the host checked what does not depend on the bird (toroidal, 3x3 window, relax query, grid >= 6x6, discretisation
// in the shared-reciprocal range, not the identity pass): such chunks stage the three window columns of their rows in
// shared memory — every staged cell copied in ascending-id order (perm) with asynchronous copies — and run the packed
// two-lane candidate loop of flk_step.cuh on it.  The birds of the seam rows (0, my-1 and the slack row my) need the
// full toroidal distance: they are stepped one warp per cell by the blockIdx.x = 0 CTAs, so a chunk that starts or ends
// a column still runs on the tile.  Seam COLUMNS, chunks with a spilled cell in reach and chunks whose rows do not fit
// the tile take slot_step_bird_general bird by bird.
// ------------------------------------------------------------------------------------------------
#ifndef FLK_SLOT_MINBLOCKS
#define FLK_SLOT_MINBLOCKS 9
#endif
constexpr int SLOT_NR = 42;   // rows staged per window column: 3 x 42 cells = 126 of the CTA's 128 threads

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// columns whose 3x3 windows never touch the x seam and that this rank steps
__device__ __forceinline__ bool tile_column(const Geom &g, int c) {
    const int gcx = global_col(g, c);
    return gcx >= 1 && gcx <= g.mx - 2 && (g.mode == 0 || (c >= 1 && c <= g.nown));
}

template <bool RELAX, int MATH, bool FAST>
__global__ void __launch_bounds__(SLOT_CHUNK, FLK_SLOT_MINBLOCKS)
k_slot_step(const __grid_constant__ Geom g, const __grid_constant__ Params p, const __grid_constant__ SlotArgs S,
            const __grid_constant__ Store s, const __grid_constant__ StepArgs a, int col_lo) {
    constexpr int NR = SLOT_NR;
    __shared__ __align__(128) float4 tile_s[FAST ? 3 * TILE_CAP + 1 : 1];
    __shared__ uint32_t idtile[FAST ? TILE_CAP : 1];
    __shared__ uint32_t offs[FAST ? 3 * NR + 2 : 1];   // exclusive prefix over the staged cell counts: own column first,
                                                       // then the left and the right one (+ total)
    __shared__ uint8_t rowof[FAST ? TILE_CAP : 1];     // own column: staged row of every tile entry
    __shared__ uint32_t wtot[SLOT_CHUNK / 32];
    __shared__ int sh_flag;
    const int t = threadIdx.x;
    const uint32_t dh = (uint32_t)g.dh;
    const int c = col_lo + (int)blockIdx.y;
    const bool split = FAST && tile_column(g, c);      // this column's seam-row birds are stepped by its blockIdx.x = 0 CTA

    if (blockIdx.x == 0) {
        // ---- seam rows of a tile column: one warp per cell (rows 0, my-1, my)
        if (!split) return;
        const int k3 = t >> 5;
        if (k3 >= 3) return;
        const int row = k3 == 0 ? 0 : (k3 == 1 ? g.my - 1 : g.my);
        const uint32_t n = __ldg(S.rcnt + (uint32_t)c * dh + (uint32_t)row);
        for (uint32_t kk = (uint32_t)t & 31u; kk < n; kk += 32) slot_step_bird_general<RELAX, MATH>(g, p, S, s, a, c, row, n, kk);
        return;
    }
    const uint32_t coln = __ldg(S.colcount + c);
    const uint32_t nchunk = (coln + SLOT_CHUNK - 1) / SLOT_CHUNK;
    // the last CTA of the column takes the chunks the grid does not cover (the host's bound follows the population with
    // a delay): correctness never depends on the bound
    for (uint32_t m = blockIdx.x - 1; m < nchunk; m += (blockIdx.x == gridDim.x - 1 ? 1u : 0x40000000u)) {
        if (m >= S.stride) break;                      // flagged by k_slot_colscan
        const uint32_t tab = __ldg(S.coltab + (size_t)c * S.stride + m);
        const int rowA = (int)(tab & 0xffffu);         // row of the chunk's first bird
        const uint32_t o0 = tab >> 16;                 // ... which is bird number o0 of that cell
        const uint32_t nb = min((uint32_t)SLOT_CHUNK, coln - m * SLOT_CHUNK);
        bool done = false;
        if (split) {   // CTA-uniform
            if (m != blockIdx.x - 1) __syncthreads();  // a second chunk in this CTA: the tile is being reused
            // ---- counts (and id orders) of the staged cells: rows rS .. rS+NR-1 of columns c, c-1, c+1 (in that order)
            const int rS = rowA >= 1 ? rowA - 1 : 0;
            const int rA = rowA - rS;                                // staged row of the chunk's first bird (0 or 1)
            const int wc = t / NR, wr = t - wc * NR;                 // wc: 0 own column, 1 left, 2 right
            const int row = rS + wr;
            uint32_t v = 0, cell = 0;
            unsigned long long perm = 0;
            if (t < 3 * NR && row < (int)dh) {
                cell = (uint32_t)(c + (wc == 0 ? 0 : (wc == 1 ? -1 : 1))) * dh + (uint32_t)row;
                v = __ldg(S.rcnt + cell);
                perm = __ldg(reinterpret_cast<const unsigned long long *>(S.rperm) + cell);
            }
            if (t == 0) sh_flag = 0;
            const uint32_t inc = warp_incl_scan(v);
            if ((t & 31) == 31) wtot[t >> 5] = inc;
            __syncthreads();
            uint32_t ex = inc - v;
#pragma unroll
            for (int w = 0; w < SLOT_CHUNK / 32 - 1; ++w) ex += (w < (t >> 5)) ? wtot[w] : 0u;
            if (t <= 3 * NR) offs[t] = ex;                      // offs[3*NR] = total (threads >= 3*NR hold v = 0)
            if (v > (uint32_t)S.C) atomicOr(&sh_flag, 1);       // a spilled cell in reach
            if (wc == 0) {
                for (uint32_t k = 0; k < v; ++k)
                    if (ex + k < (uint32_t)TILE_CAP) rowof[ex + k] = (uint8_t)wr;
            }
            __syncthreads();
            // ---- rows the chunk's birds span; is everything they need staged?
            const uint32_t j0 = offs[rA] + o0;                  // tile index of the chunk's first bird
            const uint32_t jl = j0 + nb - 1;
            // the last bird sits in a staged row <= NR-2 (its row+1 is staged), or the staged rows reach the column's end
            bool ok = sh_flag == 0 && jl < (uint32_t)TILE_CAP && (jl < offs[NR - 1] || rS + NR >= (int)dh);
            int rB = 0;
            if (ok) {
                rB = min((int)rowof[jl] + 1, NR - 1);           // last staged row any of the chunk's birds reads
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) ok = ok && (offs[cc * NR + rB + 1] - offs[cc * NR] <= (uint32_t)TILE_CAP);
            }
            if (ok) {   // CTA-uniform
                // tile order: left column, own column, right column (the reference's query order is x-outer)
                float4 *const tileL = tile_s, *const tileO = tile_s + TILE_CAP, *const tileR = tile_s + 2 * TILE_CAP;
                if (t < 3 * NR && wr <= rB && v) {
                    const uint32_t dst0 = ex - offs[wc * NR];
                    float4 *dst = (wc == 0 ? tileO : (wc == 1 ? tileL : tileR)) + dst0;
                    const size_t src = (size_t)cell * (uint32_t)S.C;
                    for (uint32_t k = 0; k < v; ++k) {          // asynchronous copies: none of them waits for another
                        const uint32_t slot = (uint32_t)(perm >> (4 * k)) & 15u;
                        cp_async16(dst + k, S.rrec + src + slot);
                        if (wc == 0) cp_async4(idtile + dst0 + k, S.rid + src + slot);
                    }
                }
                cp_async_wait_all();
                __syncthreads();
                done = true;
                if ((uint32_t)t < nb) {
                    const uint32_t j = j0 + (uint32_t)t;
                    const int r = rowof[j];                         // staged row of this bird
                    const int grow = rS + r;
                    if (grow != 0 && grow < g.my - 1) {             // seam rows: stepped by the column's seam CTA
                        const uint32_t baseL = offs[NR], baseR = offs[2 * NR];
                        const uint32_t a0 = offs[NR + r - 1] - baseL, b0 = offs[NR + r + 2] - baseL;
                        const uint32_t a1 = offs[r - 1], b1 = offs[r + 2];
                        const uint32_t a2 = offs[2 * NR + r - 1] - baseR, b2 = offs[2 * NR + r + 2] - baseR;
                        const float4 mer = tileO[j];
                        float xa = 0.0f, ya = 0.0f, xc = 0.0f, yc = 0.0f, xs = 0.0f, ys = 0.0f;
                        if (MATH == 0 && FLK_PACKED) {
                            const f32x2 ME = pk(mer.x, mer.y);
                            f32x2 A = 0ull, Cc = 0ull, Ss = 0ull;
                            accumulate_run_smem_packed<false>(tileL, a0, b0, 0, ME, A, Cc, Ss);
                            accumulate_run_smem_packed<true>(tileO, a1, b1, j, ME, A, Cc, Ss);
                            accumulate_run_smem_packed<false>(tileR, a2, b2, 0, ME, A, Cc, Ss);
                            unpk(A, xa, ya); unpk(Cc, xc, yc); unpk(Ss, xs, ys);
                        } else {
                            accumulate_run_smem<MATH, false>(tileL, a0, b0, 0, mer.x, mer.y, xa, ya, xc, yc, xs, ys);
                            accumulate_run_smem<MATH, true>(tileO, a1, b1, j, mer.x, mer.y, xa, ya, xc, yc, xs, ys);
                            accumulate_run_smem<MATH, false>(tileR, a2, b2, 0, mer.x, mer.y, xa, ya, xc, yc, xs, ys);
                        }
                        Acc acc;
                        acc.xa = xa; acc.ya = ya; acc.xc = xc; acc.yc = yc; acc.xs = xs; acc.ys = ys;
                        acc.nvec = (b0 - a0) + (b1 - a1) + (b2 - a2);
                        acc.count = (int)acc.nvec - 1;
                        const Rec me = Rec{mer.x, mer.y, mer.z, mer.w};
                        const uint32_t my_id = idtile[j];
                        const Rec out = bird_finish<MATH>(g, p, a, s.step_dev, me, my_id, acc);
                        slot_emit(g, S, s, out, my_id, true, make_recip(g.d));
                    }
                }
            }
        }
        if (!done && (uint32_t)t < nb) {
            // ---- general path: find this thread's bird by walking the column from the chunk's first cell
            uint32_t q = o0 + (uint32_t)t, n = 0;
            int row = rowA;
            bool found = false;
            while (row < (int)dh) {
                n = __ldg(S.rcnt + (uint32_t)c * dh + (uint32_t)row);
                if (q < n) { found = true; break; }
                q -= n;
                ++row;
            }
            if (found && !(split && (row == 0 || row >= g.my - 1)))   // a tile column's seam rows: the seam CTA
                slot_step_bird_general<RELAX, MATH>(g, p, S, s, a, c, row, n, q);
        }
    }
}

void launch_slot_step(const Geom &g, const Params &p, const SlotArgs &S, const Store &s, const StepArgs &a, int col_lo,
                      int col_hi, uint32_t max_col_birds, cudaStream_t st) {
    if (col_hi <= col_lo) return;
    const bool fast = p.relax && g.toroidal && p.k == 1 && g.mx >= 6 && g.my >= 6 && g.d >= 0x1p-20f && g.d <= 4096.0f &&
                      !a.identity && p.radius > 0.0f;
    // blockIdx.x = 0: the column's seam cells; 1 + m: chunk m.  max_col_birds bounds the largest column loosely (the last
    // CTA of a column loops over whatever the bound misses)
    uint32_t chunks = (max_col_birds + SLOT_CHUNK - 1) / SLOT_CHUNK;
    if (chunks < 1) chunks = 1;
    if (chunks > 65000) chunks = 65000;
    const dim3 grid(1 + chunks, (unsigned)(col_hi - col_lo)), block(SLOT_CHUNK);
#define FLK_SL(R, M, F) k_slot_step<R, M, F><<<grid, block, 0, st>>>(g, p, S, s, a, col_lo)
    if (p.relax) {
        if (a.math_mode == 1) { if (fast) FLK_SL(true, 1, true); else FLK_SL(true, 1, false); }
        else { if (fast) FLK_SL(true, 0, true); else FLK_SL(true, 0, false); }
    } else {
        if (a.math_mode == 1) FLK_SL(false, 1, false);
        else FLK_SL(false, 0, false);
    }
#undef FLK_SL
}

// ------------------------------------------------------------------------------------------------
// k_slot_colscan — Field2D::lazy_update on the slotted store
// ------------------------------------------------------------------------------------------------
// slots [0, n) of one cell in ascending-id order, 4 bits per slot.  Every pair is compared once (ranks of both sides
// updated); equal ids (not expected in one field) fall in descending slot order, so the result is always a permutation.
template <int NN>
__device__ __forceinline__ unsigned long long rank_perm(const uint32_t (&a)[SLOT_C_MAX], int n) {
    uint32_t rk[NN];
#pragma unroll
    for (int k = 0; k < NN; ++k) rk[k] = 0;
#pragma unroll
    for (int k = 1; k < NN; ++k) {
        if (k < n) {
#pragma unroll
            for (int j = 0; j < k; ++j) {
                const uint32_t c = a[j] < a[k] ? 1u : 0u;
                rk[k] += c;
                rk[j] += 1u - c;
            }
        }
    }
    unsigned long long perm = 0;
#pragma unroll
    for (int k = 0; k < NN; ++k)
        if (k < n) perm |= (unsigned long long)k << (4 * rk[k]);
    return perm;
}

__device__ __forceinline__ unsigned long long sort_perm(const uint32_t *ids, int n, int C) {
    uint32_t a[SLOT_C_MAX];
    const uint4 *p4 = reinterpret_cast<const uint4 *>(ids);   // C is a multiple of 4: 16-byte aligned
#pragma unroll
    for (int q = 0; q < SLOT_C_MAX / 4; ++q) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (4 * q < n) v = p4[q];
        a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
    (void)C;
    if (n <= 8) return rank_perm<8>(a, n);
    return rank_perm<16>(a, n);
}

struct ColscanArgs {
    const uint32_t *cnt;        // counters of the buffer that becomes READ
    const uint32_t *id;         // its ids
    uint32_t *zero_cnt;         // counters of the buffer that becomes WRITE (cleared: lazy_update empties the bags)
    unsigned long long *perm;
    uint32_t *colcount, *colmax, *colchunk, *coltab, *totals, *ticket;
    const uint32_t *spill_new_read;
    uint32_t *spill_new_write;
    uint32_t *n_read_dev;
    uint64_t *step_dev;
    uint32_t *err_flag;
    uint32_t stride, spill_cap, ncols;
    int C, bump_step;
};

constexpr int CS_T = 256;            // threads per CTA
constexpr int CS_E = 4;              // consecutive cells per thread
constexpr int CS_SEG = CS_T * CS_E;  // rows of one column a CTA sweeps

// grid = ncols x nseg CTAs (blockIdx.x = column * nseg + segment).  A CTA first sums the counters of the rows that precede
// its segment in its column (they come out of L2), then sweeps its own 1024 rows once: no CTA waits for another one.
__global__ void __launch_bounds__(CS_T) k_slot_colscan(const Geom g, const ColscanArgs A, uint32_t nseg) {
    __shared__ uint32_t sh_last, sh_sum, sh_maxcol, sh_maxcell;
    const uint32_t c = blockIdx.x / nseg, sg = blockIdx.x - c * nseg, dh = (uint32_t)g.dh;
    const size_t base = (size_t)c * dh;
    const int t = threadIdx.x;
    if (t == 0) { sh_maxcell = 0; sh_last = 0; }
    const uint32_t r0 = sg * CS_SEG;
    uint32_t before = 0;
    for (uint32_t r = (uint32_t)t; r < r0; r += CS_T) before += A.cnt[base + r];
    const uint32_t r = r0 + (uint32_t)t * CS_E;
    uint32_t n[CS_E], v = 0;
#pragma unroll
    for (int q = 0; q < CS_E; ++q) {
        n[q] = r + q < dh ? A.cnt[base + r + q] : 0u;
        v += n[q];
    }
    uint32_t run, total;
    block_excl_scan<CS_T>(before, &run);                 // run = birds of this column before the segment
    uint32_t pfx = block_excl_scan<CS_T>(v, &total) + run;
    uint32_t lmax = 0;
#pragma unroll
    for (int q = 0; q < CS_E; ++q) {
        if (r + q >= dh) break;
        const uint32_t nq = n[q];
        if (nq) {
            if (nq > 0xffffu) atomicOr(A.err_flag, ERR_CAPACITY);
            // chunks of the step kernel that start inside this cell
            for (uint32_t m = (pfx + SLOT_CHUNK - 1) / SLOT_CHUNK; m * SLOT_CHUNK < pfx + nq; ++m) {
                if (m < A.stride) A.coltab[(size_t)c * A.stride + m] = (r + q) | ((m * SLOT_CHUNK - pfx) << 16);
                else atomicOr(A.err_flag, ERR_CAPACITY);
            }
        }
        const int ns = (int)min(nq, (uint32_t)A.C);
        A.perm[base + r + q] = ns >= 2 ? sort_perm(A.id + (base + r + q) * (uint32_t)A.C, ns, A.C) : 0ull;
        A.zero_cnt[base + r + q] = 0u;
        lmax = max(lmax, nq);
        pfx += nq;
    }
    if (lmax) atomicMax(&sh_maxcell, lmax);
    __syncthreads();
    if (t == 0) {
        if (sg == nseg - 1) A.colcount[c] = run + total;
        if (sh_maxcell) atomicMax(A.colmax + c, sh_maxcell);
        __threadfence();
        sh_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!sh_last) return;
    // ---- the last CTA to finish: chunk prefix over the columns, totals of the new READ buffer
    __threadfence();
    if (t == 0) { sh_sum = 0; sh_maxcol = 0; sh_maxcell = 0; }
    __syncthreads();
    uint32_t carry = 0;
    for (uint32_t c0 = 0; c0 < A.ncols; c0 += CS_T) {
        const uint32_t cc = c0 + (uint32_t)t;
        const uint32_t nn = cc < A.ncols ? __ldcg(A.colcount + cc) : 0u;
        const uint32_t mc = cc < A.ncols ? __ldcg(A.colmax + cc) : 0u;
        if (cc < A.ncols) A.colmax[cc] = 0u;            // ready for the next lazy_update
        const uint32_t ch = (nn + SLOT_CHUNK - 1) / SLOT_CHUNK;
        uint32_t tot;
        const uint32_t ex = block_excl_scan<CS_T>(ch, &tot);
        if (cc < A.ncols) A.colchunk[cc] = carry + ex;
        carry += tot;
        if (nn) { atomicAdd(&sh_sum, nn); atomicMax(&sh_maxcol, nn); atomicMax(&sh_maxcell, mc); }
    }
    __syncthreads();
    if (t == 0) {
        A.colchunk[A.ncols] = carry;
        A.totals[0] = sh_sum; A.totals[1] = sh_maxcol; A.totals[2] = sh_maxcell;
        A.totals[3] = min(*A.spill_new_read, A.spill_cap);
        *A.spill_new_write = 0u;
        *A.ticket = 0u;
        *A.n_read_dev = sh_sum;
        if (A.bump_step) *A.step_dev += 1;
    }
}

void launch_slot_colscan(const Geom &g, const SlotStore &S, const Store &s, uint32_t ncols, int bump_step, cudaStream_t st) {
    const int nr = 1 - S.rd;   // the WRITE buffer becomes READ
    ColscanArgs A;
    A.cnt = S.cnt[nr]; A.id = S.id[nr]; A.zero_cnt = S.cnt[S.rd];
    A.perm = reinterpret_cast<unsigned long long *>(S.perm);
    A.colcount = S.colcount; A.colmax = S.colmax; A.colchunk = S.colchunk; A.coltab = S.coltab;
    A.totals = S.totals; A.ticket = S.ticket;
    A.spill_new_read = S.spill_cnt + nr; A.spill_new_write = S.spill_cnt + S.rd;
    A.n_read_dev = s.n_read_dev; A.step_dev = s.step_dev; A.err_flag = s.err_flag;
    A.stride = S.stride; A.spill_cap = S.spill_cap; A.ncols = ncols;
    A.C = S.C; A.bump_step = bump_step;
    const uint32_t nseg = ((uint32_t)g.dh + CS_SEG - 1) / CS_SEG;
    k_slot_colscan<<<ncols * nseg, CS_T, 0, st>>>(g, A, nseg);
}

// ------------------------------------------------------------------------------------------------
// uploads: bulk Field2D::set_object_location (state.rs:49) into the WRITE buffer
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slot_ingest(const Geom g, const SlotArgs S, const Store s, uint32_t n,
                                                     const uint32_t *__restrict__ id, const float2 *__restrict__ loc,
                                                     const float2 *__restrict__ ld, int owned_only) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float2 l = loc[j], d = ld[j];
    const int cx = cell_coord(l.x, g.d, g.mx + 1);
    const int cy = cell_coord(l.y, g.d, g.dh);
    int lc = local_col(g, cx);
    // strips: only birds this rank owns are accepted (flockers_mpi/src/model/state.rs:75-85); ghosts arrive through the halo
    if (owned_only && g.mode != 0 && (lc == 0 || lc == g.nown + 1)) lc = FLK_DROP;
    if (lc == FLK_DROP) return;
    slot_insert(S, s.err_flag, (uint32_t)lc * (uint32_t)g.dh + (uint32_t)cy, make_float4(l.x, l.y, d.x, d.y), id[j]);
}

void launch_slot_ingest(const Geom &g, const SlotArgs &S, const Store &s, uint32_t n, const uint32_t *id, const float2 *loc,
                        const float2 *ld, int owned_only, cudaStream_t st) {
    if (n == 0) return;
    k_slot_ingest<<<(n + 255) / 256, 256, 0, st>>>(g, S, s, n, id, loc, ld, owned_only);
}

// dense form (ids id0 .. id0+n-1 implied by the position) and the records of a sorted-runs READ buffer (ids given)
__global__ void __launch_bounds__(256) k_slot_ingest_dense(const Geom g, const SlotArgs S, const Store s, uint32_t n,
                                                           uint32_t id0, const float4 *__restrict__ rec4,
                                                           const uint32_t *__restrict__ ids) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 r = rec4[j];
    const int cx = cell_coord(r.x, g.d, g.mx + 1);
    const int cy = cell_coord(r.y, g.d, g.dh);
    const int lc = local_col(g, cx);
    if (lc == FLK_DROP) return;
    slot_insert(S, s.err_flag, (uint32_t)lc * (uint32_t)g.dh + (uint32_t)cy, r, ids ? ids[j] : id0 + j);
}

void launch_slot_ingest_dense(const Geom &g, const SlotArgs &S, const Store &s, uint32_t n, uint32_t id0, const float4 *rec4,
                              const uint32_t *ids, cudaStream_t st) {
    if (n == 0) return;
    k_slot_ingest_dense<<<(n + 255) / 256, 256, 0, st>>>(g, S, s, n, id0, rec4, ids);
}

// strips: birds received from the two neighbouring strips (migrants + next tick's ghosts) join the WRITE buffer
// (flockers_mpi/src/model/state.rs:128-130,167-174)
__global__ void __launch_bounds__(256) k_slot_unpack(const Geom g, const SlotArgs S, const Store s) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * s.capx) return;
    const int side = t >= s.capx;
    const uint32_t j = side ? t - s.capx : t;
    uint8_t *buf = side ? s.recvR : s.recvL;
    const uint32_t n = min(reinterpret_cast<const XHdr *>(buf)->count, s.capx);
    if (j >= n) return;
    const float4 r = reinterpret_cast<const float4 *>(xbuf_rec(buf))[j];
    const uint32_t id = xbuf_id(buf, s.capx)[j];
    const int lc = local_col(g, cell_coord(r.x, g.d, g.mx + 1));
    const int cy = cell_coord(r.y, g.d, g.dh);
    if (lc == FLK_DROP) return;
    slot_insert(S, s.err_flag, (uint32_t)lc * (uint32_t)g.dh + (uint32_t)cy, r, id);
}

void launch_slot_unpack(const Geom &g, const SlotArgs &S, const Store &s, cudaStream_t st) {
    k_slot_unpack<<<(2 * s.capx + 255) / 256, 256, 0, st>>>(g, S, s);
}

// ------------------------------------------------------------------------------------------------
// read side
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_slot_materialize(const __grid_constant__ Geom g, const __grid_constant__ SlotArgs S,
                                                          const __grid_constant__ Store s, uint32_t ncell) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell == 0) *s.n_read_dev = min(s.cell_start[ncell], s.cap);
    if (cell >= ncell) return;
    uint32_t w = s.cell_start[cell];
    for_each_entry(S, cell, [&](const float4 &r, uint32_t id) {
        if (w < s.cap) {
            reinterpret_cast<float4 *>(s.rec)[w] = r;
            s.ids[w] = id;
        } else {
            atomicOr(s.err_flag, ERR_CAPACITY);
        }
        ++w;
    });
}

void launch_slot_materialize(const Geom &g, const SlotArgs &S, const Store &s, uint32_t ncell, cudaStream_t st) {
    k_slot_materialize<<<(ncell + 255) / 256, 256, 0, st>>>(g, S, s, ncell);
}

__global__ void __launch_bounds__(256) k_slot_gather_dense(const SlotArgs S, uint32_t ncell, uint32_t id0, uint32_t n,
                                                           float4 *__restrict__ out, uint32_t *__restrict__ outside) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    const uint32_t cnt = S.rcnt[cell];
    if (cnt == 0) return;
    const uint32_t C = (uint32_t)S.C;
    const size_t base = (size_t)cell * C;
    for (uint32_t q = 0; q < min(cnt, C); ++q) {
        const uint32_t k = S.rid[base + q] - id0;
        if (k < n) out[k] = S.rrec[base + q];
        else atomicAdd(outside, 1u);
    }
    if (cnt > C) {
        const uint32_t nsp = min(*S.rspill_cnt, S.spill_cap);
        for (uint32_t j = 0; j < nsp; ++j) {
            if (S.rspill[j].cell != cell) continue;
            const uint32_t k = S.rspill[j].id - id0;
            if (k < n) out[k] = S.rspill[j].rec;
            else atomicAdd(outside, 1u);
        }
    }
}

void launch_slot_gather_dense(const SlotArgs &S, uint32_t ncell, uint32_t id0, uint32_t n, float4 *out, uint32_t *outside,
                              cudaStream_t st) {
    if (ncell == 0) return;
    k_slot_gather_dense<<<(ncell + 255) / 256, 256, 0, st>>>(S, ncell, id0, n, out, outside);
}

// statistics of a sorted-runs READ buffer for the layout decision
__global__ void __launch_bounds__(256) k_cell_stats(const Store s, uint32_t ncell, int C, uint32_t *out) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 0;
    if (cell < ncell) n = s.cell_start[cell + 1] - s.cell_start[cell];
    uint32_t mx = n, over = n > (uint32_t)C ? n - (uint32_t)C : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        over += __shfl_xor_sync(0xffffffffu, over, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mx) atomicMax(out, mx);
        if (over) atomicAdd(out + 1, over);
    }
}

void launch_cell_stats(const Store &s, uint32_t ncell, int C, uint32_t *out, cudaStream_t st) {
    if (ncell == 0) return;
    k_cell_stats<<<(ncell + 255) / 256, 256, 0, st>>>(s, ncell, C, out);
}

}  // namespace flk
