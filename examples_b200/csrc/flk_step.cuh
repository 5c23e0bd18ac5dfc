// flk_step.cuh — the parts of Bird::step (flockers/src/model/bird.rs:27-145) shared by the step kernels of both READ
// layouts (flk_kernels.cu: cells as runs of one sorted array; flk_slots.cu: fixed-capacity cells): the neighbour
// accumulators, the shared-memory candidate loops and everything after the loop.
#pragma once
#include "flk_kernels.cuh"

namespace flk {

// ------------------------------------------------------------------------------------------------
// neighbour accumulation (bird.rs:44-71), candidates visited in query order
// ------------------------------------------------------------------------------------------------
struct Acc {
    float xa = 0.0f, ya = 0.0f, xc = 0.0f, yc = 0.0f, xs = 0.0f, ys = 0.0f;   // bird.rs:44-49
    uint32_t nvec = 0;   // size of the query result (bird.rs:42 tests emptiness)
    int count = 0;       // bird.rs:50
};

// ---- TMA staging helpers (1-D bulk copies global -> shared, completion on an mbarrier) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// append one bird to a neighbour's message (flockers_mpi/src/model/bird.rs:194: agents_to_send[owner].push)
__device__ __forceinline__ void xbuf_append(const Store &s, uint8_t *buf, const Rec &r, uint32_t id,
                                            const uint32_t *payload = nullptr) {
    const uint32_t slot = atomicAdd(&reinterpret_cast<XHdr *>(buf)->count, 1u);
    if (slot < s.capx) {
        reinterpret_cast<float4 *>(xbuf_rec(buf))[slot] = make_float4(r.x, r.y, r.ldx, r.ldy);
        xbuf_id(buf, s.capx)[slot] = id;
        for (uint32_t k = 0; k < s.pw; ++k) xbuf_pay(buf, s.capx)[(size_t)slot * s.pw + k] = payload ? payload[k] : 0u;
    } else {
        atomicOr(s.err_flag, ERR_EXCHANGE);
    }
}

#ifndef FLK_STEP_BS
#define FLK_STEP_BS 128   // birds (threads) per k_step CTA
#endif
constexpr int TILE_CAP = 2 * FLK_STEP_BS;   // records staged per window column (4 KB); 3 columns = 12 KB per CTA

// everything of Bird::step after the neighbour loop (bird.rs:73-140): combine, draw, normalise, move -> the successor
template <int MATH>
__device__ __forceinline__ Rec bird_finish(const Geom &g, const Params &p, const StepArgs &a, const uint64_t *step_dev,
                                           const Rec &me, uint32_t my_id, const Acc &acc) {
    const float W = g.W, H = g.H;
    float xa = acc.xa, ya = acc.ya, xc = acc.xc, yc = acc.yc, xs = acc.xs, ys = acc.ys;
    const uint32_t nvec = acc.nvec;
    const int count = acc.count;
    float avo_x = 0.0f, avo_y = 0.0f, coh_x = 0.0f, coh_y = 0.0f;   // bird.rs:36-40
    float rnd_x = 0.0f, rnd_y = 0.0f, con_x = 0.0f, con_y = 0.0f;
    if (nvec != 0) {                                                  // :42
        if (count > 0) {                                              // :73
            const float c = (float)count;
            // eight divisions by `count` (:74-84): one shared reciprocal when all six sums are in the safe range
            const float lo = fminf(fminf(fminf(fabsf(xa), fabsf(ya)), fminf(fabsf(xc), fabsf(yc))), fminf(fabsf(xs), fabsf(ys)));
            const float hi = fmaxf(fmaxf(fmaxf(fabsf(xa), fabsf(ya)), fmaxf(fabsf(xc), fabsf(yc))), fmaxf(fabsf(xs), fabsf(ys)));
            if (lo >= 0x1p-60f && hi <= 0x1p60f) {
                const Recip R = make_recip(c);                        // c in [1, 2^24]
                xa = div_rn_fast(xa, R); ya = div_rn_fast(ya, R);     // :74-75
                xc = div_rn_fast(xc, R); yc = div_rn_fast(yc, R);     // :76-77
                xs = div_rn_fast(xs, R); ys = div_rn_fast(ys, R);     // :78-79
                con_x = div_rn_fast(xs, R); con_y = div_rn_fast(ys, R);   // :81-84 (second division, kept)
            } else {
                xa = __fdiv_rn(xa, c); ya = __fdiv_rn(ya, c);
                xc = __fdiv_rn(xc, c); yc = __fdiv_rn(yc, c);
                xs = __fdiv_rn(xs, c); ys = __fdiv_rn(ys, c);
                con_x = __fdiv_rn(xs, c); con_y = __fdiv_rn(ys, c);
            }
        } else {
            con_x = xs; con_y = ys;                                   // :85-90
        }
        avo_x = __fmul_rn(400.0f, xa); avo_y = __fmul_rn(400.0f, ya); // :92-95
        div2_rn_checked(-xc, -yc, 10.0f, coh_x, coh_y);               // :97-100
        float r1, r2;                                                 // :103-107
        if (a.injected != nullptr && my_id < a.n_injected) {
            const float2 r = __ldg(a.injected + my_id);
            r1 = r.x; r2 = r.y;
        } else {
            const uint64_t step = a.use_step_dev ? *step_dev : a.step;
            uint32_t o0, o1;
            philox4x32_10(my_id, (uint32_t)step, (uint32_t)(step >> 32), 0u, (uint32_t)a.seed,
                          (uint32_t)(a.seed >> 32), o0, o1);
            r1 = u32_to_unit(o0); r2 = u32_to_unit(o1);
        }
        const float xr = __fadd_rn(__fmul_rn(r1, 2.0f), -1.0f);
        const float yr = __fadd_rn(__fmul_rn(r2, 2.0f), -1.0f);
        const float sq = __fsqrt_rn(__fadd_rn(__fmul_rn(xr, xr), __fmul_rn(yr, yr)));   // :109
        div2_rn_checked(__fmul_rn(0.05f, xr), __fmul_rn(0.05f, yr), sq, rnd_x, rnd_y);  // :110-113
    }
    // :116-127, summed left to right
    float dx = __fmul_rn(p.cohesion, coh_x);
    dx = __fadd_rn(dx, __fmul_rn(p.avoidance, avo_x));
    dx = __fadd_rn(dx, __fmul_rn(p.consistency, con_x));
    dx = __fadd_rn(dx, __fmul_rn(p.randomness, rnd_x));
    dx = __fadd_rn(dx, __fmul_rn(p.momentum, me.ldx));
    float dy = __fmul_rn(p.cohesion, coh_y);
    dy = __fadd_rn(dy, __fmul_rn(p.avoidance, avo_y));
    dy = __fadd_rn(dy, __fmul_rn(p.consistency, con_y));
    dy = __fadd_rn(dy, __fmul_rn(p.randomness, rnd_y));
    dy = __fadd_rn(dy, __fmul_rn(p.momentum, me.ldy));
    const float dis = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // :129
    if (dis > 0.0f) {                                                                // :130
        float ux, uy;
        div2_rn_checked(dx, dy, dis, ux, uy);
        dx = __fmul_rn(ux, p.jump);                                                  // :131
        dy = __fmul_rn(uy, p.jump);                                                  // :132
    }
    Rec out;
    out.ldx = dx; out.ldy = dy;                                                      // :135
    out.x = toroidal_transform(__fadd_rn(me.x, dx), W);                              // :137
    out.y = toroidal_transform(__fadd_rn(me.y, dy), H);                              // :138
    return out;
}

// interior accumulation out of the CTA's shared-memory tile (same order, same arithmetic as accumulate_interior)
#ifndef FLK_SMEM_UNROLL
#define FLK_SMEM_UNROLL 2   // packed loop, k_step ms: without prefetch 1: 0.877, 2: 0.909, 4: 0.912; with prefetch 1: 0.901, 2: 0.856, 3: 0.892, 4: 0.896
#endif
constexpr int kSmemUnroll = FLK_SMEM_UNROLL;
// SELF: the run contains the bird itself at index `self` (bird.rs:53 skips it by id).  Its own record contributes
// exactly nothing to the avoidance and cohesion sums (dx = dy = +0, 0/den = +0, and an accumulator that started at +0
// is unchanged by adding +0), so only the consistency terms need a select — cheaper than splitting the run in two.
// Packed form of the same loop for exact arithmetic: Blackwell's two-lane f32 instructions (add/mul/fma.rn.f32x2 ->
// FADD2/FMUL2/FFMA2) carry the x and y halves of every step in one issue slot.  Each lane is the same IEEE
// round-to-nearest operation as the scalar form, so results are bit-identical:
//   (dx,dy) = fma(o.xy, -1, me.xy)   == me - o      (the product by -1 is exact, one rounding)
//   nden    = (-sq*sq) + (-1)        == -(sq*sq+1)  (round-to-nearest is symmetric)
#ifndef FLK_PACKED
#define FLK_PACKED 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float a, float b) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ void unpk(f32x2 d, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// FLK_PACKMASK selects which groups of the loop use the two-lane form (bit 0: difference and squares, bit 1: the two
// quotients, bit 2: the three accumulator pairs).  A two-lane instruction occupies the FMA pipe for ~2.3 cycles against
// 2 x 1.05 for its scalar pair (tools/micro/x2_throughput.cu), so packing only pays where issue slots are the limit.
#ifndef FLK_PACKMASK
#define FLK_PACKMASK 7
#endif
#ifndef FLK_PREFETCH
#define FLK_PREFETCH 1   // next candidate's LDS.128 issued before the current one is consumed (with unroll 2: 0.867 -> 0.856 ms)
#endif
template <bool SELF>
__device__ __forceinline__ void accumulate_run_smem_packed(const float4 *tile, uint32_t j, uint32_t e, uint32_t self,
                                                           f32x2 ME, f32x2 &A, f32x2 &Cc, f32x2 &Ss) {
    const ulonglong2 *t2 = reinterpret_cast<const ulonglong2 *>(tile);
    const f32x2 NEG1 = pk(-1.0f, -1.0f);
    float mx_, my_;
    unpk(ME, mx_, my_);
#if FLK_PREFETCH
    ulonglong2 nxt = t2[j];   // one slot past a run is still inside the tile array (padded by one record)
#endif
#pragma unroll kSmemUnroll
    for (; j < e; ++j) {
#if FLK_PREFETCH
        ulonglong2 o = nxt;
        nxt = t2[j + 1];
#else
        ulonglong2 o = t2[j];
#endif
        if (SELF && j == self) o.y = 0ull;                     // (last_d) := (+0,+0), see accumulate_run_smem
        f32x2 Dv;
        float dx, dy, s0, s1;
        if (FLK_PACKMASK & 1) {
            Dv = fma2(o.x, NEG1, ME);                           // bird.rs:54-55 (no seam in reach)
            unpk(Dv, dx, dy);
            unpk(mul2(Dv, Dv), s0, s1);
        } else {
            float ox, oy;
            unpk(o.x, ox, oy);
            dx = __fadd_rn(mx_, -ox); dy = __fadd_rn(my_, -oy);
            Dv = pk(dx, dy);
            s0 = __fmul_rn(dx, dx); s1 = __fmul_rn(dy, dy);
        }
        const float sq = __fadd_rn(s0, s1);                    // :59
        const float nden = __fadd_rn(-__fmul_rn(sq, sq), -1.0f);
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-nden));
        const float er = __fmaf_rn(nden, r, 1.0f);             // div2_rn_interior, two lanes at a time
        r = __fmaf_rn(r, er, r);
        f32x2 Q;
        if (FLK_PACKMASK & 2) {
            const f32x2 RR = pk(r, r), ND = pk(nden, nden);
            const f32x2 Q0 = mul2(Dv, RR);
            const f32x2 REM = fma2(ND, Q0, Dv);
            Q = fma2(RR, REM, Q0);                              // :60-61
        } else {
            const float q0 = __fmul_rn(dx, r), q1 = __fmul_rn(dy, r);
            const float rem0 = __fmaf_rn(nden, q0, dx), rem1 = __fmaf_rn(nden, q1, dy);
            Q = pk(__fmaf_rn(r, rem0, q0), __fmaf_rn(r, rem1, q1));
        }
        if (FLK_PACKMASK & 4) {
            A = add2(A, Q);
            Cc = add2(Cc, Dv);                                  // :64-65
            Ss = add2(Ss, o.y);                                 // :68-69
        } else {
            float a0, a1, q0, q1;
            unpk(A, a0, a1); unpk(Q, q0, q1); A = pk(__fadd_rn(a0, q0), __fadd_rn(a1, q1));
            unpk(Cc, a0, a1); Cc = pk(__fadd_rn(a0, dx), __fadd_rn(a1, dy));
            unpk(Ss, a0, a1); unpk(o.y, q0, q1); Ss = pk(__fadd_rn(a0, q0), __fadd_rn(a1, q1));
        }
    }
}

template <int MATH, bool SELF>
__device__ __forceinline__ void accumulate_run_smem(const float4 *tile, uint32_t j, uint32_t e, uint32_t self, float mx_,
                                                    float my_, float &xa, float &ya, float &xc, float &yc, float &xs,
                                                    float &ys) {
#pragma unroll kSmemUnroll
    for (; j < e; ++j) {
        float4 o = tile[j];
        if (SELF && j == self) { o.z = 0.0f; o.w = 0.0f; }
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

}  // namespace flk
