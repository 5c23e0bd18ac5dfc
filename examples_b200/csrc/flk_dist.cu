// flk_dist.cu — multi-GPU Flockers: 1-D periodic strip decomposition over cell columns, one strip per GPU.
//
// Follows the flockers_mpi protocol (flockers_mpi/src/model/state.rs:105-175, bird.rs:185-195):
//   before_step  ghosts of the neighbouring strips are present in the READ buffer          (halo)
//   step         every OWNED bird steps; its successor is routed by the owner of its new location
//   after_step   leavers go to their new owner, arrivals are inserted                       (migration)
//   update       lazy_update
// JUMP (0.7) < discretisation, so a bird crosses at most one column per tick: both the migrants and next
// tick's ghosts of a neighbour are exactly "successors that landed in one of the two columns next to the
// shared boundary".  One message per neighbour per tick therefore carries halo AND migration.
//
// Two transports move the fixed-capacity messages:
//   NCCL   one process per GPU (torchrun); ncclSend/ncclRecv grouped on the handle's stream.  libnccl is
//          dlopen'ed so that single-GPU users need no NCCL at all.
//   LOCAL  all strips driven by one host thread (flk_mgpu_*): cudaMemcpyPeerAsync between the handles'
//          buffers, ordered with events.  Also how the strip logic is tested on a single GPU.
#include <dlfcn.h>

#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "../../include/flk.h"
#include "flk_internal.cuh"

using namespace flk;

// ---- minimal NCCL surface (ABI-stable subset of nccl.h; the library is resolved at run time) ----------------
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
}
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclUint8 = 1;  // ncclUint8

int nccl_load() {
    if (g_nccl.lib) return FLK_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(FLK_E_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                            \
    if (!g_nccl.field) return fail(FLK_E_NCCL, "libnccl lacks %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return FLK_OK;
}
#define NCCL_TRY(x)                                                                                      \
    do {                                                                                                 \
        ncclResult_t r_ = (x);                                                                           \
        if (r_ != 0) return fail(FLK_E_NCCL, "%s: %s", #x, g_nccl.GetErrorString(r_));                   \
    } while (0)
}  // namespace

struct flk_dist_state {
    int rank = 0, world = 1;
    int transport = 0;  // 0 NCCL, 1 LOCAL
    ncclComm_t comm = nullptr;
    int left = 0, right = 0;
    uint8_t *xbuf = nullptr;           // sendL | sendR | recvL | recvR
    flk_field **group = nullptr;       // LOCAL: all handles of the group (owned by rank 0's state)
    cudaStream_t comm_stream = nullptr;   // high-priority stream the exchange runs on, beside the interior step
    cudaEvent_t ev_tick = nullptr;     // main stream: the previous tick's rebin and this tick's header clears are done
    cudaEvent_t ev_sent = nullptr;     // my send buffers are complete (border columns stepped)
    cudaEvent_t ev_copied = nullptr;   // the exchange is complete: neighbours' messages are in my recv buffers
    // host-side upper bound of the READ population (grid sizing): the device count is copied back asynchronously
    // every tick and used once it has landed, so no tick ever waits for it
    uint32_t *h_count = nullptr;       // pinned
    cudaEvent_t ev_count = nullptr;
    bool count_pending = false;
    uint64_t known = 0, known_tick = 0, copy_tick = 0, tick_no = 0;   // tick_no = ticks enqueued so far
    bool have_known = false;
    bool overlap = true;               // FLK_DIST_OVERLAP=0 steps all columns before the exchange (A/B measurements)
    bool committed = false;
    std::vector<int> bounds;           // column boundaries: strip / tile column px owns [bounds[px], bounds[px+1])
    // tiles (mode 2): ranks form a Px x Py grid, rank = py*Px + px; rows are split like the columns
    int Px = 1, Py = 1, px = 0, py = 0;
    int down = 0, up = 0;
    std::vector<int> rbounds;          // row boundaries (Py+1)
    bool bands = false;                // tiles: this tick stepped the border row bands ahead of the interior
    cudaEvent_t ev_stepped = nullptr;  // tiles: every owned bird of this tick is stepped (main stream)
    cudaEvent_t ev_fwd = nullptr;      // my lower/upper messages are complete (own successors + forwarded arrivals)
    cudaEvent_t ev_copied_y = nullptr; // the row-direction exchange is complete
};

namespace flk {

void dist_free(flk_field *f) {
    flk_dist_state *d = f->dist;
    if (!d) return;
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    if (d->xbuf) cudaFree(d->xbuf);
    if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
    if (d->h_count) cudaFreeHost(d->h_count);
    if (d->ev_count) cudaEventDestroy(d->ev_count);
    if (d->ev_sent) cudaEventDestroy(d->ev_sent);
    if (d->ev_tick) cudaEventDestroy(d->ev_tick);
    if (d->ev_copied) cudaEventDestroy(d->ev_copied);
    if (d->ev_fwd) cudaEventDestroy(d->ev_fwd);
    if (d->ev_stepped) cudaEventDestroy(d->ev_stepped);
    if (d->ev_copied_y) cudaEventDestroy(d->ev_copied_y);
    if (d->rank == 0 && d->group) delete[] d->group;
    delete d;
    f->dist = nullptr;
}

static void strip_of(int mx, int world, int rank, int *c0, int *c1) {
    *c0 = (int)((long long)rank * mx / world);
    *c1 = (int)((long long)(rank + 1) * mx / world);
}

// strip boundaries: the equal-width default or the caller's (Kdtree-style partition by population, flk_balance_columns)
static int make_bounds(int mx, int world, const int32_t *col_bounds, std::vector<int> &b) {
    b.resize(world + 1);
    for (int r = 0; r <= world; ++r) {
        int c0, c1;
        if (col_bounds) b[r] = col_bounds[r];
        else if (r == world) b[r] = mx;
        else { strip_of(mx, world, r, &c0, &c1); b[r] = c0; }
    }
    if (b[0] != 0 || b[world] != mx) return fail(FLK_E_INVALID, "column boundaries must start at 0 and end at %d", mx);
    for (int r = 0; r < world; ++r)
        if (b[r + 1] - b[r] < 2) return fail(FLK_E_UNSUPPORTED, "strip %d would be %d columns wide (minimum 2)", r, b[r + 1] - b[r]);
    return FLK_OK;
}

// common part of flk_dist_create / flk_mgpu_create
static int dist_field_new(float w, float h, float disc, uint64_t capacity, int device, int rank, int world,
                          int transport, flk_field **out, const int32_t *col_bounds = nullptr, int Py = 1,
                          const int32_t *row_bounds = nullptr, uint32_t payload_bytes = 0) {
    *out = nullptr;
    if (payload_bytes % 4 != 0 || payload_bytes > 256) return fail(FLK_E_INVALID, "payload_bytes must be a multiple of 4, <= 256");
    if (world < 2) return fail(FLK_E_INVALID, "strip decomposition needs world >= 2 (use flk_field_create for one GPU)");
    if (rank < 0 || rank >= world) return fail(FLK_E_INVALID, "rank %d out of range", rank);
    if (capacity == 0 || capacity > 0x7fffffffull) return fail(FLK_E_INVALID, "capacity must be in [1, 2^31)");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(FLK_E_CUDA, "no CUDA device (%s): libflk has no CPU fallback", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(FLK_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    flk_field *f = new flk_field();
    f->device = device;
    int rc = setup_geometry(f, w, h, disc, 1);
    if (rc) { delete f; return rc; }
    Geom &g = f->g;
    if (Py < 1 || world % Py != 0) { delete f; return fail(FLK_E_INVALID, "%d ranks do not form a grid with %d rows", world, Py); }
    const int Px = world / Py, px = rank % Px, py = rank / Px;
    if (Py > 1 && Px < 2) { delete f; return fail(FLK_E_UNSUPPORTED, "a tile grid needs at least 2 x 2 ranks (use strips for one row or column)"); }
    if (g.mx < 2 * Px || (Py > 1 && g.my < 2 * Py)) {
        delete f;
        return fail(FLK_E_UNSUPPORTED, "%d x %d cells cannot be split into %d x %d tiles of >= 2 x 2 cells", g.mx, g.my, Px, Py);
    }
    std::vector<int> bounds, rbounds;
    rc = make_bounds(g.mx, Px, col_bounds, bounds);
    if (!rc && Py > 1) rc = make_bounds(g.my, Py, row_bounds, rbounds);
    if (rc) { delete f; return rc; }
    const int c0 = bounds[px], c1 = bounds[px + 1];
    int narrowest = g.mx;
    for (int r = 0; r < Px; ++r) narrowest = std::min(narrowest, bounds[r + 1] - bounds[r]);
    for (int r = 0; r < Py && Py > 1; ++r) narrowest = std::min(narrowest, rbounds[r + 1] - rbounds[r]);
    g.mode = Py > 1 ? 2 : 1; g.col0 = c0; g.nown = c1 - c0; g.has_slack = (c0 == 0);
    g.ncols = g.nown + 3;
    if (Py > 1) {   // rows laid out like the columns: lower ghost, owned, upper ghost, slack
        g.row0 = rbounds[py]; g.rown = rbounds[py + 1] - rbounds[py]; g.has_slack_row = (g.row0 == 0);
        g.dh = g.rown + 3;
    }
    f->p = Params{0.8f, 1.0f, 1.1f, 0.7f, 1.0f, 0.7f, 10.0f, 1, 1};
    rc = update_window(f);
    if (rc) { delete f; return rc; }
    f->cap = capacity;
    f->ncell = (uint32_t)g.ncols * (uint32_t)g.dh;
    // exchange capacity: the two boundary columns hold ~2/nown of the strip; allow 8x that (clustered flocks).
    // Must be identical on every rank (it fixes the message layout), hence the narrowest strip width.
    uint64_t capx = 8 * (capacity / (uint64_t)narrowest) + 4096;
    if (capx > capacity) capx = capacity;
    f->s.capx = (uint32_t)capx;
    f->s.pw = payload_bytes / 4;      // Field2D<O>: the payload words travel in the halo / migration messages too
    rc = [&]() -> int {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&f->ev0));
        CUDA_TRY(cudaEventCreate(&f->ev1));
        return field_alloc(f);
    }();
    if (rc) { field_free(f); delete f; return rc; }
    flk_dist_state *d = new flk_dist_state();
    f->dist = d;
    d->rank = rank; d->world = world; d->transport = transport;
    d->bounds = bounds; d->rbounds = rbounds;
    d->Px = Px; d->Py = Py; d->px = px; d->py = py;
    if (const char *e = getenv("FLK_DIST_OVERLAP")) d->overlap = atoi(e) != 0;
    d->left = py * Px + (px + Px - 1) % Px;
    d->right = py * Px + (px + 1) % Px;
    d->down = ((py + Py - 1) % Py) * Px + px;
    d->up = ((py + 1) % Py) * Px + px;
    const size_t xb = (xbuf_bytes(f->s.capx, f->s.pw) + 255) & ~(size_t)255;
    const int nbuf = Py > 1 ? 8 : 4;
    cudaError_t e = cudaMalloc(&d->xbuf, nbuf * xb);
    if (e == cudaSuccess) e = cudaMemsetAsync(d->xbuf, 0, nbuf * xb, f->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_fwd, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_stepped, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_copied_y, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        e = cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, hi_p);
    }
    if (e == cudaSuccess) e = cudaHostAlloc(&d->h_count, 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_count, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_sent, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_tick, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_copied, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        dist_free(f); field_free(f); delete f;
        return fail(FLK_E_CUDA, "dist alloc: %s", cudaGetErrorString(e));
    }
    f->s.sendL = d->xbuf; f->s.sendR = d->xbuf + xb; f->s.recvL = d->xbuf + 2 * xb; f->s.recvR = d->xbuf + 3 * xb;
    if (Py > 1) {
        f->s.sendD = d->xbuf + 4 * xb; f->s.sendU = d->xbuf + 5 * xb; f->s.recvD = d->xbuf + 6 * xb; f->s.recvU = d->xbuf + 7 * xb;
    }
    f->n_read = 0;
    *out = f;
    return FLK_OK;
}

// upper bound of the READ population for grid sizing: the last count that has landed on the host plus what the
// exchanges enqueued since then can have added (2*capx per tick), never above the capacity
static uint32_t read_upper_bound(flk_field *f) {
    flk_dist_state *d = f->dist;
    if (d->count_pending && cudaEventQuery(d->ev_count) == cudaSuccess) {
        d->known = *d->h_count;          // population after tick `copy_tick`
        d->known_tick = d->copy_tick;
        d->have_known = true;
        d->count_pending = false;
    }
    if (!d->have_known) return (uint32_t)f->cap;
    const uint64_t b = d->known + (d->tick_no - d->known_tick) * (uint64_t)f->s.nrecv * f->s.capx;
    return (uint32_t)(b < f->cap ? b : f->cap);
}

// phase 1 of a tick on one strip: clear the outgoing messages, step (or re-stage) every owned bird.
// Successors that go into a message can only come from the two owned columns next to each boundary (a bird moves
// less than one column per tick) and from the slack column, so those columns are stepped on the comm stream, which
// then runs the exchange, while the interior columns are stepped on the main stream (north star: halo exchange overlapped with
// interior-cell compute on separate streams).
static int dist_phase_step(flk_field *f, uint64_t step, uint64_t seed, const float2 *d_inj, uint32_t n_inj, int identity) {
    flk_dist_state *d = f->dist;
    CUDA_TRY(cudaMemsetAsync(f->s.sendL, 0, sizeof(XHdr), f->stream));
    CUDA_TRY(cudaMemsetAsync(f->s.sendR, 0, sizeof(XHdr), f->stream));
    if (f->g.mode == 2) {
        CUDA_TRY(cudaMemsetAsync(f->s.sendD, 0, sizeof(XHdr), f->stream));
        CUDA_TRY(cudaMemsetAsync(f->s.sendU, 0, sizeof(XHdr), f->stream));
    }
    StepArgs a;
    a.step = step; a.use_step_dev = 0; a.seed = seed;
    a.injected = d_inj; a.n_injected = n_inj;
    a.out_base = 0;
    a.math_mode = f->math_mode;
    a.identity = identity;
    a.skip_bands = 0;
    a.nseam = 0; a.gen_threads = 0;
    const bool prof = f->profile && !identity;
    if (prof) {
        cudaEvent_t e0, e1;
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        f->prof_events.push_back(e0); f->prof_events.push_back(e1);
        CUDA_TRY(cudaEventRecord(e0, f->stream));
    }
    const uint32_t dh = (uint32_t)f->g.dh, nown = (uint32_t)f->g.nown;
    // tiles: the border COLUMNS are stepped first too, so the left/right exchange and its unpack run beside the interior
    // step; the lower/upper messages are complete only when every column is done (border rows are not an index range)
    const bool split = nown >= 5 && d->overlap;
    if (split) {
        // border: columns {ghost, 1, 2} and {nown-1, nown, ghost, slack} (ghost threads only invalidate their WRITE
        // slot), on the high-priority comm stream so that they run BESIDE the interior launch instead of ahead of it
        // (two launches of ~500 CTAs would leave most of the GPU idle).  The grid covers capx birds per range and the
        // kernel flags ERR_EXCHANGE if a range holds more (capx is 8x the mean population of two columns).
        CUDA_TRY(cudaEventRecord(d->ev_tick, f->stream));            // previous rebin + header clears are done
        CUDA_TRY(cudaStreamWaitEvent(d->comm_stream, d->ev_tick, 0));
        const uint32_t ranges[2][2] = {{0, 3 * dh}, {(nown - 1) * dh, (nown + 3) * dh}};
        for (int r = 0; r < 2; ++r) {
            a.cell_lo = ranges[r][0]; a.cell_hi = ranges[r][1];
            launch_step(f->g, f->p, f->s, a, f->s.capx, d->comm_stream);
            f->launches += 1;
        }
        // tiles: the two owned rows next to each row boundary (and the slack row) of the interior columns are the only
        // other birds whose successors enter a message: step them now as well, so that the lower/upper exchange does
        // not have to wait for the interior step either
        d->bands = f->g.mode == 2 && f->g.rown >= 5 && !identity;
        if (d->bands) {
            launch_step_bands(f->g, f->p, f->s, a, 3, (int)nown - 1, d->comm_stream);
            f->launches += 1;
        }
        CUDA_TRY(cudaEventRecord(d->ev_sent, d->comm_stream));
        a.cell_lo = 3 * dh; a.cell_hi = (nown - 1) * dh;
        a.skip_bands = d->bands ? 1 : 0;
        launch_step(f->g, f->p, f->s, a, read_upper_bound(f), f->stream);
        a.skip_bands = 0;
    a.nseam = 0; a.gen_threads = 0;
        f->launches += 1;
    } else {
        d->bands = false;
        a.cell_lo = 0; a.cell_hi = f->ncell;
        launch_step(f->g, f->p, f->s, a, read_upper_bound(f), f->stream);
        f->launches += 1;
        CUDA_TRY(cudaEventRecord(d->ev_sent, f->stream));
    }
    if (f->g.mode == 2) CUDA_TRY(cudaEventRecord(d->ev_stepped, f->stream));   // every owned bird of this tick is stepped
    if (prof) CUDA_TRY(cudaEventRecord(f->prof_events.back(), f->stream));
    CUDA_TRY(cudaGetLastError());
    return FLK_OK;
}

// phase 2: move the messages (on the comm stream, as soon as the border columns are done)
static int dist_phase_exchange(flk_field *f) {
    flk_dist_state *d = f->dist;
    const size_t nbytes = xbuf_bytes(f->s.capx, f->s.pw);
    cudaStream_t cs = d->comm_stream;
    // my own border step of THIS tick is behind my unpack/scatter of the previous tick on the main stream, so waiting
    // for it also protects the receive buffers the previous tick may still be reading
    CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_sent, 0));
    if (d->transport == 0) {
        NCCL_TRY(g_nccl.GroupStart());
        NCCL_TRY(g_nccl.Send(f->s.sendL, nbytes, kNcclUint8, d->left, d->comm, cs));
        NCCL_TRY(g_nccl.Send(f->s.sendR, nbytes, kNcclUint8, d->right, d->comm, cs));
        // order matters only for world==2 (both messages come from the same peer): its sendL is my recvR
        NCCL_TRY(g_nccl.Recv(f->s.recvR, nbytes, kNcclUint8, d->right, d->comm, cs));
        NCCL_TRY(g_nccl.Recv(f->s.recvL, nbytes, kNcclUint8, d->left, d->comm, cs));
        NCCL_TRY(g_nccl.GroupEnd());
    } else {
        flk_field *L = d->group[d->left], *R = d->group[d->right];
        CUDA_TRY(cudaStreamWaitEvent(cs, L->dist->ev_sent, 0));
        CUDA_TRY(cudaStreamWaitEvent(cs, R->dist->ev_sent, 0));
        CUDA_TRY(cudaMemcpyPeerAsync(f->s.recvL, f->device, L->s.sendR, L->device, nbytes, cs));
        CUDA_TRY(cudaMemcpyPeerAsync(f->s.recvR, f->device, R->s.sendL, R->device, nbytes, cs));
    }
    CUDA_TRY(cudaEventRecord(d->ev_copied, cs));
    return FLK_OK;
}

// tiles, phase 2b: the left/right arrivals join the WRITE buffer; those in a border row are forwarded into the
// lower/upper messages (the diagonal neighbours' birds travel in two hops, so every tile talks to four peers only)
static int dist_phase_unpack_x(flk_field *f) {
    flk_dist_state *d = f->dist;
    cudaStream_t cs = d->comm_stream;           // behind the left/right exchange in stream order, beside the interior step
    launch_unpack(f->g, f->s, cs, 0);
    f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    // my own successors must all be in the lower/upper messages: with the border bands stepped first they already are
    if (!d->bands) CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_stepped, 0));
    CUDA_TRY(cudaEventRecord(d->ev_fwd, cs));
    return FLK_OK;
}

// tiles, phase 2c: the row-direction exchange
static int dist_phase_exchange_y(flk_field *f) {
    flk_dist_state *d = f->dist;
    const size_t nbytes = xbuf_bytes(f->s.capx, f->s.pw);
    cudaStream_t cs = d->comm_stream;
    CUDA_TRY(cudaStreamWaitEvent(cs, d->ev_fwd, 0));
    if (d->transport == 0) {
        NCCL_TRY(g_nccl.GroupStart());
        NCCL_TRY(g_nccl.Send(f->s.sendD, nbytes, kNcclUint8, d->down, d->comm, cs));
        NCCL_TRY(g_nccl.Send(f->s.sendU, nbytes, kNcclUint8, d->up, d->comm, cs));
        // order matters only for two tile rows (both messages come from the same peer): its sendD is my recvU
        NCCL_TRY(g_nccl.Recv(f->s.recvU, nbytes, kNcclUint8, d->up, d->comm, cs));
        NCCL_TRY(g_nccl.Recv(f->s.recvD, nbytes, kNcclUint8, d->down, d->comm, cs));
        NCCL_TRY(g_nccl.GroupEnd());
    } else {
        flk_field *Dn = d->group[d->down], *Up = d->group[d->up];
        CUDA_TRY(cudaStreamWaitEvent(cs, Dn->dist->ev_fwd, 0));
        CUDA_TRY(cudaStreamWaitEvent(cs, Up->dist->ev_fwd, 0));
        CUDA_TRY(cudaMemcpyPeerAsync(f->s.recvD, f->device, Dn->s.sendU, Dn->device, nbytes, cs));
        CUDA_TRY(cudaMemcpyPeerAsync(f->s.recvU, f->device, Up->s.sendD, Up->device, nbytes, cs));
    }
    CUDA_TRY(cudaEventRecord(d->ev_copied_y, cs));
    return FLK_OK;
}

// phase 3: arrivals join the WRITE buffer, then lazy_update
static int dist_phase_finish(flk_field *f) {
    flk_dist_state *d = f->dist;
    const bool tiles = f->g.mode == 2;
    CUDA_TRY(cudaStreamWaitEvent(f->stream, tiles ? d->ev_copied_y : d->ev_copied, 0));
    if (d->transport == 1) {
        // LOCAL: my neighbours read MY send buffers with their own copies; the next tick may not clear them before
        CUDA_TRY(cudaStreamWaitEvent(f->stream, d->group[d->left]->dist->ev_copied, 0));
        CUDA_TRY(cudaStreamWaitEvent(f->stream, d->group[d->right]->dist->ev_copied, 0));
        if (tiles) {
            CUDA_TRY(cudaStreamWaitEvent(f->stream, d->group[d->down]->dist->ev_copied_y, 0));
            CUDA_TRY(cudaStreamWaitEvent(f->stream, d->group[d->up]->dist->ev_copied_y, 0));
        }
    }
    launch_unpack(f->g, f->s, f->stream, tiles ? 1 : 0);
    f->launches += 1;
    CUDA_TRY(cudaGetLastError());
    // grid bounds only; validity is decided on the device.  The new READ population is at most the old bound plus
    // the messages.
    const uint64_t ub = (uint64_t)read_upper_bound(f) + (uint64_t)f->s.nrecv * f->s.capx;
    f->w_count = ub < f->cap ? ub : f->cap;
    int rc = enqueue_lazy_update(f, 0);
    f->n_read = f->cap;   // the host only keeps an upper bound (the exact count lives on the device)
    f->w_count = 0;
    if (rc) return rc;
    d->tick_no += 1;
    if (!d->count_pending) {
        CUDA_TRY(cudaMemcpyAsync(d->h_count, f->s.n_read_dev, 4, cudaMemcpyDeviceToHost, f->stream));
        CUDA_TRY(cudaEventRecord(d->ev_count, f->stream));
        d->count_pending = true;
        d->copy_tick = d->tick_no;
    }
    return FLK_OK;
}

static int dist_tick(flk_field *f, uint64_t step, uint64_t seed, const float2 *d_inj, uint32_t n_inj, int identity) {
    int rc = dist_phase_step(f, step, seed, d_inj, n_inj, identity);
    if (!rc) rc = dist_phase_exchange(f);
    if (!rc && f->g.mode == 2) rc = dist_phase_unpack_x(f);
    if (!rc && f->g.mode == 2) rc = dist_phase_exchange_y(f);
    if (!rc) rc = dist_phase_finish(f);
    return rc;
}

// lazy_update of the birds uploaded so far (owned only), before the first halo exchange
static int dist_publish_uploads(flk_field *f) {
    const uint64_t w = f->w_count;
    launch_scan(f->s, f->ncell + 1, f->stream);
    if (w) {
        // uploads sit at WRITE [0,w): plain (non-dist) scatter
        launch_scatter(f->s, (uint32_t)w, 0, f->stream);
    }
    launch_rank(f->s, (uint32_t)w, f->ncell, 0, f->stream);
    f->launches += 2 + (w ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    f->n_read = w;
    f->w_count = 0;
    return FLK_OK;
}

// the population changes outside the tick loop (upload, clear): forget what the host knew about it.  A count copy
// enqueued before the change may still land later; it must not become `known` (it would under-size the interior grid)
void dist_reset_counts(flk_field *f) {
    flk_dist_state *d = f->dist;
    d->committed = false;
    d->have_known = false;
    d->count_pending = false;
    d->known = 0; d->known_tick = 0; d->copy_tick = 0; d->tick_no = 0;
}

int dist_set_objects(flk_field *f, uint64_t n, const uint32_t *id, const float *loc, const float *ld, const uint32_t *payload) {
    dist_reset_counts(f);
    return write_append(f, n, id, loc, ld, payload);
}

static int dist_check_errors(flk_field *f) {
    CUDA_TRY(cudaMemcpyAsync(f->h_err, f->s.err_flag, 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    const uint32_t flag = *f->h_err;
    if (flag & ERR_CAPACITY) return fail(FLK_E_CAPACITY, "rank %d holds more birds than its capacity %llu", f->dist->rank, (unsigned long long)f->cap);
    if (flag & ERR_EXCHANGE) return fail(FLK_E_CAPACITY, "rank %d: halo/migration message exceeded %u birds", f->dist->rank, f->s.capx);
    return FLK_OK;
}

int dist_refresh_counts(flk_field *f) {
    int rc = dist_check_errors(f);
    if (rc) return rc;
    uint32_t n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, f->s.n_read_dev, 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    f->n_read = n;
    return FLK_OK;
}

int dist_run(flk_field *f, uint64_t first_step, uint64_t n_steps, uint64_t seed) {
    flk_dist_state *d = f->dist;
    if (d->transport != 0) return fail(FLK_E_STATE, "use flk_mgpu_run on an in-process group");
    if (!d->committed) return fail(FLK_E_STATE, "call flk_dist_commit after uploading birds");
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    for (uint64_t t = 0; t < n_steps; ++t) {
        int rc = dist_tick(f, first_step + t, seed, nullptr, 0, 0);
        if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    return FLK_OK;
}

}  // namespace flk

extern "C" {

int flk_dist_unique_id(uint8_t *id128) {
    if (!id128) return fail(FLK_E_INVALID, "id128 is null");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return FLK_OK;
}

int flk_dist_create(float width, float height, float discretization, uint64_t capacity, int device, int rank, int world,
                    const uint8_t *id128, flk_field **out) {
    return flk_dist_create_cols(width, height, discretization, capacity, device, rank, world, id128, nullptr, out);
}

int flk_dist_create_cols(float width, float height, float discretization, uint64_t capacity, int device, int rank,
                         int world, const uint8_t *id128, const int32_t *col_bounds, flk_field **out) {
    return flk_dist_create_tiles(width, height, discretization, capacity, device, rank, world, id128, 1, col_bounds,
                                 nullptr, out);
}

int flk_dist_create_tiles(float width, float height, float discretization, uint64_t capacity, int device, int rank,
                          int world, const uint8_t *id128, int tile_rows, const int32_t *col_bounds,
                          const int32_t *row_bounds, flk_field **out) {
    return flk_dist_create_ex(width, height, discretization, capacity, device, rank, world, id128, tile_rows, col_bounds,
                              row_bounds, 0, out);
}

int flk_dist_create_ex(float width, float height, float discretization, uint64_t capacity, int device, int rank,
                       int world, const uint8_t *id128, int tile_rows, const int32_t *col_bounds,
                       const int32_t *row_bounds, uint32_t payload_bytes, flk_field **out) {
    if (!out || !id128) return fail(FLK_E_INVALID, "null argument");
    int rc = nccl_load();
    if (rc) return rc;
    rc = dist_field_new(width, height, discretization, capacity, device, rank, world, 0, out, col_bounds, tile_rows,
                        row_bounds, payload_bytes);
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&(*out)->dist->comm, world, id, rank);
    if (r != 0) {
        int code = fail(FLK_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
        flk_field_destroy(*out);
        *out = nullptr;
        return code;
    }
    return FLK_OK;
}

int flk_mgpu_create(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                    const int *devices, flk_field **out) {
    return flk_mgpu_create_cols(width, height, discretization, capacity_per_rank, n, devices, nullptr, out);
}

// Kdtree-style partition along x: equal-population strip boundaries from a sample of the x coordinates
int flk_balance_columns(const float *x, uint64_t n, uint64_t stride, float width, float discretization, int world,
                        int32_t *col_bounds) {
    if (!col_bounds || (n && !x) || world < 1 || stride == 0) return fail(FLK_E_INVALID, "bad argument");
    if (!(width > 0.0f) || !(discretization > 0.0f)) return fail(FLK_E_INVALID, "width, discretization must be > 0");
    const int mx = (int)std::ceil(width / discretization);
    if (mx < 2 * world) return fail(FLK_E_UNSUPPORTED, "%d cell columns cannot be split into %d strips of >= 2 columns", mx, world);
    std::vector<uint64_t> hist(mx, 0);
    for (uint64_t i = 0; i < n; ++i) {
        const float q = std::floor(x[i * stride] / discretization);   // f32, like the device binning
        int c = (q != q) ? 0 : (q >= (float)mx ? 0 : (q < 0.0f ? 0 : (int)q));   // the slack column belongs to strip 0
        hist[c] += 1;
    }
    // greedy quantile cut: boundary r sits where the running count reaches r/world of the sample, every strip keeps
    // at least 2 columns and leaves at least 2 for each strip still to come
    col_bounds[0] = 0;
    uint64_t run = 0;
    int c = 0;
    for (int r = 1; r < world; ++r) {
        const uint64_t target = (uint64_t)((double)n * r / world);
        const int lo = col_bounds[r - 1] + 2, hi = mx - 2 * (world - r);
        while (c < hi && (c < lo || run + hist[c] / 2 < target)) run += hist[c++];
        col_bounds[r] = c;
    }
    col_bounds[world] = mx;
    return FLK_OK;
}

int flk_mgpu_create_cols(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                         const int *devices, const int32_t *col_bounds, flk_field **out) {
    return flk_mgpu_create_tiles(width, height, discretization, capacity_per_rank, n, devices, 1, col_bounds, nullptr, out);
}

int flk_mgpu_create_tiles(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                          const int *devices, int tile_rows, const int32_t *col_bounds, const int32_t *row_bounds,
                          flk_field **out) {
    return flk_mgpu_create_ex(width, height, discretization, capacity_per_rank, n, devices, tile_rows, col_bounds, row_bounds,
                              0, out);
}

int flk_mgpu_create_ex(float width, float height, float discretization, uint64_t capacity_per_rank, int n,
                       const int *devices, int tile_rows, const int32_t *col_bounds, const int32_t *row_bounds,
                       uint32_t payload_bytes, flk_field **out) {
    if (!out || !devices || n < 2) return fail(FLK_E_INVALID, "need n >= 2 devices");
    flk_field **grp = new flk_field *[n];
    for (int r = 0; r < n; ++r) {
        int rc = dist_field_new(width, height, discretization, capacity_per_rank, devices[r], r, n, 1, &grp[r], col_bounds,
                                tile_rows, row_bounds, payload_bytes);
        if (rc) {
            for (int q = 0; q < r; ++q) { grp[q]->dist->group = nullptr; flk_field_destroy(grp[q]); }
            delete[] grp;
            return rc;
        }
    }
    for (int r = 0; r < n; ++r) {
        grp[r]->dist->group = grp;
        out[r] = grp[r];
        // direct peer copies over NVLink where the topology allows it
        for (int q = 0; q < n; ++q) {
            if (devices[q] == devices[r]) continue;
            int can = 0;
            cudaSetDevice(devices[r]);
            if (cudaDeviceCanAccessPeer(&can, devices[r], devices[q]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
                if (e != cudaSuccess) cudaGetLastError();  // already enabled
            }
        }
    }
    return FLK_OK;
}

int flk_dist_strip(const flk_field *f, int32_t *col0, int32_t *col1) {
    if (!f) return fail(FLK_E_INVALID, "null handle");
    if (col0) *col0 = f->g.mode != 0 ? f->g.col0 : 0;
    if (col1) *col1 = f->g.mode != 0 ? f->g.col0 + f->g.nown : f->g.mx;
    return FLK_OK;
}

int flk_dist_tile(const flk_field *f, int32_t *col0, int32_t *col1, int32_t *row0, int32_t *row1) {
    if (!f) return fail(FLK_E_INVALID, "null handle");
    flk_dist_strip(f, col0, col1);
    if (row0) *row0 = f->g.mode == 2 ? f->g.row0 : 0;
    if (row1) *row1 = f->g.mode == 2 ? f->g.row0 + f->g.rown : f->g.my;
    return FLK_OK;
}

static int cell_of(float v, float d) {   // same f32 arithmetic as the device binning
    float q = std::floor(v / d);
    int c = (q != q) ? 0 : (q >= 2147483648.0f ? 0x7fffffff : (q <= -2147483648.0f ? -0x7fffffff - 1 : (int)q));
    return c < 0 ? 0 : c;
}

int flk_dist_owner(const flk_field *f, float x, float y, int32_t *rank) {
    if (!f || !rank) return fail(FLK_E_INVALID, "null argument");
    if (f->g.mode == 2) {   // get_block_by_location on the tile grid; the slack column / row belong to column / row 0
        const flk_dist_state *d = f->dist;
        const int gc = cell_of(x, f->g.d), gr = cell_of(y, f->g.d);
        int px = 0, py = 0;
        if (gc < f->g.mx) for (int r = 0; r < d->Px; ++r) if (gc >= d->bounds[r] && gc < d->bounds[r + 1]) px = r;
        if (gr < f->g.my) for (int r = 0; r < d->Py; ++r) if (gr >= d->rbounds[r] && gr < d->rbounds[r + 1]) py = r;
        *rank = py * d->Px + px;
        return FLK_OK;
    }
    if (f->g.mode == 0) { *rank = 0; return FLK_OK; }
    float q = std::floor(x / f->g.d);   // same f32 arithmetic as the device binning
    int gc = (q != q) ? 0 : (q >= 2147483648.0f ? 0x7fffffff : (q <= -2147483648.0f ? -0x7fffffff - 1 : (int)q));
    if (gc < 0) gc = 0;
    if (gc >= f->g.mx) { *rank = 0; return FLK_OK; }   // the slack column belongs to the strip holding column 0
    const int world = f->dist->world;
    const std::vector<int> &b = f->dist->bounds;
    for (int r = 0; r < world; ++r)
        if (gc >= b[r] && gc < b[r + 1]) { *rank = r; return FLK_OK; }
    return fail(FLK_E_INVALID, "no owner for column %d", gc);
}

// publish the uploaded birds and build the first halo (state.rs:113-133)
int flk_dist_commit(flk_field *f) {
    CHECK_HANDLE(f);
    if (f->g.mode == 0) return flk_lazy_update(f);
    if (f->dist->transport != 0) return fail(FLK_E_STATE, "use flk_mgpu_commit on an in-process group");
    int rc = flush_pending(f);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    rc = dist_publish_uploads(f);
    if (!rc) rc = dist_tick(f, 0, 0, nullptr, 0, 1);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    f->dist->committed = true;
    return FLK_OK;
}

int flk_dist_step(flk_field *f, uint64_t step, uint64_t seed, const float *injected_r, uint64_t n_injected) {
    CHECK_HANDLE(f);
    if (f->g.mode == 0) {
        int rc = flk_step(f, step, seed, injected_r, n_injected);
        return rc ? rc : flk_lazy_update(f);
    }
    if (f->dist->transport != 0) return fail(FLK_E_STATE, "use flk_mgpu_step on an in-process group");
    if (!f->dist->committed) return fail(FLK_E_STATE, "call flk_dist_commit after uploading birds");
    const float2 *d_inj = nullptr;
    if (injected_r) {
        int rc = upload_injected(f, injected_r, n_injected);
        if (rc) return rc;
        d_inj = f->d_injected;
    }
    CUDA_TRY(cudaEventRecord(f->ev0, f->stream));
    int rc = dist_tick(f, step, seed, d_inj, (uint32_t)n_injected, 0);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(f->ev1, f->stream));
    return FLK_OK;
}

static int mgpu_check(flk_field **hs, int n) {
    if (!hs || n < 2) return fail(FLK_E_INVALID, "bad group");
    for (int r = 0; r < n; ++r)
        if (!hs[r] || !hs[r]->dist || hs[r]->dist->transport != 1 || hs[r]->dist->world != n || hs[r]->dist->rank != r)
            return fail(FLK_E_INVALID, "handle %d is not rank %d of an in-process group of %d", r, r, n);
    return FLK_OK;
}

static int mgpu_tick(flk_field **hs, int n, uint64_t step, uint64_t seed, const float *injected_r, uint64_t n_injected,
                     int identity) {
    for (int r = 0; r < n; ++r) {
        flk_field *f = hs[r];
        CUDA_TRY(cudaSetDevice(f->device));
        const float2 *d_inj = nullptr;
        if (injected_r) {
            int rc = upload_injected(f, injected_r, n_injected);
            if (rc) return rc;
            d_inj = f->d_injected;
        }
        int rc = dist_phase_step(f, step, seed, d_inj, (uint32_t)n_injected, identity);
        if (rc) return rc;
    }
    for (int r = 0; r < n; ++r) {
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        int rc = dist_phase_exchange(hs[r]);
        if (rc) return rc;
    }
    if (hs[0]->g.mode == 2) {
        for (int r = 0; r < n; ++r) {
            CUDA_TRY(cudaSetDevice(hs[r]->device));
            int rc = dist_phase_unpack_x(hs[r]);
            if (rc) return rc;
        }
        for (int r = 0; r < n; ++r) {
            CUDA_TRY(cudaSetDevice(hs[r]->device));
            int rc = dist_phase_exchange_y(hs[r]);
            if (rc) return rc;
        }
    }
    for (int r = 0; r < n; ++r) {
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        int rc = dist_phase_finish(hs[r]);
        if (rc) return rc;
    }
    return FLK_OK;
}

int flk_mgpu_commit(flk_field **hs, int n) {
    int rc = mgpu_check(hs, n);
    if (rc) return rc;
    for (int r = 0; r < n; ++r) {
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        rc = flush_pending(hs[r]);
        if (!rc) rc = dist_publish_uploads(hs[r]);
        if (rc) return rc;
    }
    rc = mgpu_tick(hs, n, 0, 0, nullptr, 0, 1);
    if (rc) return rc;
    for (int r = 0; r < n; ++r) hs[r]->dist->committed = true;
    return FLK_OK;
}

int flk_mgpu_step(flk_field **hs, int n, uint64_t step, uint64_t seed, const float *injected_r, uint64_t n_injected) {
    int rc = mgpu_check(hs, n);
    if (rc) return rc;
    for (int r = 0; r < n; ++r)
        if (!hs[r]->dist->committed) return fail(FLK_E_STATE, "call flk_mgpu_commit after uploading birds");
    return mgpu_tick(hs, n, step, seed, injected_r, n_injected, 0);
}

int flk_mgpu_run(flk_field **hs, int n, uint64_t first_step, uint64_t n_steps, uint64_t seed) {
    int rc = mgpu_check(hs, n);
    if (rc) return rc;
    for (int r = 0; r < n; ++r) {
        if (!hs[r]->dist->committed) return fail(FLK_E_STATE, "call flk_mgpu_commit after uploading birds");
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        CUDA_TRY(cudaEventRecord(hs[r]->ev0, hs[r]->stream));
    }
    for (uint64_t t = 0; t < n_steps; ++t) {
        rc = mgpu_tick(hs, n, first_step + t, seed, nullptr, 0, 0);
        if (rc) return rc;
    }
    for (int r = 0; r < n; ++r) {
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        CUDA_TRY(cudaEventRecord(hs[r]->ev1, hs[r]->stream));
    }
    return FLK_OK;
}

int flk_dist_num_owned(flk_field *f, uint64_t *n) {
    CHECK_HANDLE(f);
    if (!n) return fail(FLK_E_INVALID, "n is null");
    if (f->g.mode == 0) { *n = f->n_read; return FLK_OK; }
    int rc = dist_check_errors(f);
    if (rc) return rc;
    if (f->g.mode == 2) {   // tiles: count by compaction (owned cells are not an index range)
        uint32_t *d_cnt = reinterpret_cast<uint32_t *>(f->scratch + ((f->cap * 20 + 15) & ~(uint64_t)15) + 96);
        uint32_t h = 0;
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 4, f->stream));
        launch_compact_owned(f->g, f->s, (uint32_t)f->cap, nullptr, nullptr, nullptr, d_cnt, f->stream);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&h, d_cnt, 4, cudaMemcpyDeviceToHost, f->stream));
        CUDA_TRY(cudaStreamSynchronize(f->stream));
        *n = h;
        return FLK_OK;
    }
    // owned birds = cells of local columns [1, nown] plus the slack column
    uint32_t b[4];
    const uint32_t dh = (uint32_t)f->g.dh, nown = (uint32_t)f->g.nown;
    const uint32_t idx[4] = {1 * dh, (nown + 1) * dh, (nown + 2) * dh, (nown + 3) * dh};
    for (int i = 0; i < 4; ++i)
        CUDA_TRY(cudaMemcpyAsync(&b[i], f->s.cell_start + idx[i], 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    *n = (uint64_t)(b[1] - b[0]) + (b[3] - b[2]);
    return FLK_OK;
}

// snapshot of the birds this rank OWNS (ghost copies excluded), (cell, id) order
static int dist_snapshot_owned(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n, bool sync,
                               void *payload = nullptr) {
    CHECK_HANDLE(f);
    if (f->g.mode == 0) {
        if (payload) return flk_snapshot_ex(f, id, loc, last_d, payload, n);
        return sync ? flk_snapshot(f, id, loc, last_d, n) : flk_snapshot_async(f, id, loc, last_d, n);
    }
    int rc = dist_check_errors(f);
    if (rc) return rc;
    if (!f->s.pw) payload = nullptr;
    if (f->g.mode == 2) {   // tiles: compact the owned birds into the scratch block, then copy
        const uint64_t cap = f->cap;
        uint32_t *d_id = reinterpret_cast<uint32_t *>(f->scratch);
        float2 *d_loc = reinterpret_cast<float2 *>(f->scratch + ((cap * 4 + 7) & ~(uint64_t)7));
        float2 *d_ld = d_loc + cap;
        uint32_t *d_cnt = reinterpret_cast<uint32_t *>(f->scratch + ((cap * 20 + 15) & ~(uint64_t)15) + 96);
        uint32_t h = 0;
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 4, f->stream));
        // the owned payload rows are compacted into the WRITE payload array (free between ticks)
        launch_compact_owned(f->g, f->s, (uint32_t)cap, id ? d_id : nullptr, loc ? d_loc : nullptr, last_d ? d_ld : nullptr,
                             d_cnt, f->stream, payload ? f->s.tpay : nullptr);
        f->launches += 1;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&h, d_cnt, 4, cudaMemcpyDeviceToHost, f->stream));
        CUDA_TRY(cudaStreamSynchronize(f->stream));
        if (n) *n = h;
        if (h && payload) CUDA_TRY(cudaMemcpyAsync(payload, f->s.tpay, (uint64_t)h * f->s.pw * 4, cudaMemcpyDeviceToHost, f->stream));
        if (h && id) CUDA_TRY(cudaMemcpyAsync(id, d_id, (uint64_t)h * 4, cudaMemcpyDeviceToHost, f->stream));
        if (h && loc) CUDA_TRY(cudaMemcpyAsync(loc, d_loc, (uint64_t)h * 8, cudaMemcpyDeviceToHost, f->stream));
        if (h && last_d) CUDA_TRY(cudaMemcpyAsync(last_d, d_ld, (uint64_t)h * 8, cudaMemcpyDeviceToHost, f->stream));
        if (sync) CUDA_TRY(cudaStreamSynchronize(f->stream));
        return FLK_OK;
    }
    uint32_t b[4];
    const uint32_t dh = (uint32_t)f->g.dh, nown = (uint32_t)f->g.nown;
    const uint32_t idx[4] = {1 * dh, (nown + 1) * dh, (nown + 2) * dh, (nown + 3) * dh};
    for (int i = 0; i < 4; ++i)
        CUDA_TRY(cudaMemcpyAsync(&b[i], f->s.cell_start + idx[i], 4, cudaMemcpyDeviceToHost, f->stream));
    CUDA_TRY(cudaStreamSynchronize(f->stream));
    const uint64_t n0 = b[1] - b[0], n1 = b[3] - b[2];
    const uint64_t tot = (uint64_t)b[3];
    if (n) *n = n0 + n1;
    if (tot && (loc || last_d)) {
        float2 *d_loc = reinterpret_cast<float2 *>(f->scratch);
        float2 *d_ld = d_loc + tot;
        launch_split(f->s, (uint32_t)tot, d_loc, d_ld, f->stream);
        f->launches += 1;
        CUDA_TRY(cudaGetLastError());
        const uint64_t off[2] = {b[0], b[2]}, len[2] = {n0, n1}, dst[2] = {0, n0};
        for (int k = 0; k < 2; ++k) {
            if (!len[k]) continue;
            if (loc) CUDA_TRY(cudaMemcpyAsync(loc + 2 * dst[k], d_loc + off[k], len[k] * 8, cudaMemcpyDeviceToHost, f->stream));
            if (last_d) CUDA_TRY(cudaMemcpyAsync(last_d + 2 * dst[k], d_ld + off[k], len[k] * 8, cudaMemcpyDeviceToHost, f->stream));
        }
    }
    if (id) {
        if (n0) CUDA_TRY(cudaMemcpyAsync(id, f->s.ids + b[0], n0 * 4, cudaMemcpyDeviceToHost, f->stream));
        if (n1) CUDA_TRY(cudaMemcpyAsync(id + n0, f->s.ids + b[2], n1 * 4, cudaMemcpyDeviceToHost, f->stream));
    }
    if (payload) {   // payload rows follow the READ order: the same two index ranges
        const uint64_t row = (uint64_t)f->s.pw * 4;
        uint8_t *dst = static_cast<uint8_t *>(payload);
        if (n0) CUDA_TRY(cudaMemcpyAsync(dst, f->s.pay + (uint64_t)b[0] * f->s.pw, n0 * row, cudaMemcpyDeviceToHost, f->stream));
        if (n1) CUDA_TRY(cudaMemcpyAsync(dst + n0 * row, f->s.pay + (uint64_t)b[2] * f->s.pw, n1 * row, cudaMemcpyDeviceToHost, f->stream));
    }
    if (sync) CUDA_TRY(cudaStreamSynchronize(f->stream));
    return FLK_OK;
}

int flk_dist_snapshot_owned(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n) {
    return dist_snapshot_owned(f, id, loc, last_d, n, true);
}

int flk_dist_snapshot_owned_async(flk_field *f, uint32_t *id, float *loc, float *last_d, uint64_t *n) {
    return dist_snapshot_owned(f, id, loc, last_d, n, false);
}

int flk_dist_snapshot_owned_ex(flk_field *f, uint32_t *id, float *loc, float *last_d, void *payload, uint64_t *n) {
    return dist_snapshot_owned(f, id, loc, last_d, n, true, payload);
}

}  // extern "C"
