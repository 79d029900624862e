// sph_multi.cu -- slab-decomposed multi-GPU step: one context (= one process, one GPU) per z-slab.
//
// The reference is single-process; this is the north-star's new capability.  Design:
//  * The global GRID table geometry (cell = floor(pred/r) inside the box) is identical on all ranks.
//    Rank k OWNS the particles whose PREDICTED position falls in global z layers [L_k, L_k+1); its
//    local table spans those layers plus ONE ghost layer per side, so every owned particle finds all
//    27 cells locally and the neighbour sets equal the single-GPU ones exactly.
//  * Per step: (1) predict + classify every owned row (stay / migrate lo|hi / also-a-ghost-for lo|hi),
//    (2) order-preserving pack of migrants (raw state) and ghosts (predicted positions), (3) ONE
//    exchange round with both neighbours carrying migrants + ghosts (counts first), (4) one sort of
//    owned + ghost rows; z is the slowest key digit so the sorted array is
//    [ghost-lo layer | owned rows | ghost-hi layer | departed rows] and every later halo is a
//    CONTIGUOUS row range, (5) density on owned rows -> halo of densities -> pressure -> halo of v'
//    -> viscosity -> integrate, which writes the owned rows back compacted.
//  * Halos 2 and 3 are plain ncclSend/ncclRecv of contiguous ranges of the sorted arrays straight
//    into the neighbour's ghost rows: both sides hold the same particles in the same order because
//    the sort is stable and both build the layer from (the owner's resident rows in owner order,
//    then the rows that migrated in this step in sender order).  The range lengths are cross-checked
//    every step; a mismatch is an error, never a hang.
#include <nccl.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "sph_context.h"
#include "sph_device.cuh"

using namespace sphb200;

// NCCL is bound at run time, not link time: a host process that already carries an NCCL (PyTorch
// bundles its own libnccl.so.2) must keep exactly one copy, so we adopt the loaded one if there is
// one, else $SPH_NCCL_LIB, else the system libnccl.so.2.  Only the long-stable point-to-point API is used.
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl()
{
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) { const char* p = getenv("SPH_NCCL_LIB"); if (p && *p) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define SPH_SYM(field, name)                                                                  \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                  \
    if (!g_nccl.field) { g_nccl.error = std::string("libnccl.so.2 lacks ") + name; return; }
    SPH_SYM(GetUniqueId, "ncclGetUniqueId") SPH_SYM(CommInitRank, "ncclCommInitRank") SPH_SYM(CommDestroy, "ncclCommDestroy")
    SPH_SYM(Send, "ncclSend") SPH_SYM(Recv, "ncclRecv") SPH_SYM(GroupStart, "ncclGroupStart") SPH_SYM(GroupEnd, "ncclGroupEnd")
    SPH_SYM(GetErrorString, "ncclGetErrorString") SPH_SYM(AllReduce, "ncclAllReduce")
#undef SPH_SYM
    g_nccl.ok = true;
}
const NcclApi& nccl()
{
    std::call_once(g_nccl_once, load_nccl);
    return g_nccl;
}
}  // namespace
#define ncclGetUniqueId nccl().GetUniqueId
#define ncclCommInitRank nccl().CommInitRank
#define ncclCommDestroy nccl().CommDestroy
#define ncclSend nccl().Send
#define ncclRecv nccl().Recv
#define ncclAllReduce nccl().AllReduce
#define ncclGroupStart nccl().GroupStart
#define ncclGroupEnd nccl().GroupEnd
#define ncclGetErrorString nccl().GetErrorString

struct SlabState {
    uint32_t xcap = 0;          // rows per exchange buffer
    uint32_t gcap = 0;          // rows of the ghost buffer
    uint8_t* cls = nullptr;     // [cap] classification of owned rows
    float4* mig_send[2] = {nullptr, nullptr};     // [2*xcap] pos rows then vel rows, lo / hi
    float4* ghost_send[2] = {nullptr, nullptr};   // [xcap] predicted positions, lo / hi
    float4* keep[2] = {nullptr, nullptr};         // [xcap] migrants that stay visible as ghosts, lo / hi
    float4* ghost_pred = nullptr;                 // [gcap] recv_lo | keep_lo | recv_hi | keep_hi
    uint32_t* block_counts = nullptr;             // [6][nblocks_cap] counts (k_predict_key), then [6][nblocks_cap] offsets (k_slab_scan)
    uint32_t* dev_small = nullptr;                // 64 u32: totals[6], picks[8] at 16, peer lengths[2] at 24, count messages out (lo, hi: 8 words each) at 32, in at 48
    uint32_t* host_small = nullptr;               // pinned mirror
    uint32_t nblocks_cap = 0;
    std::vector<int> layers;                      // nranks + 1 global layer indices
    uint32_t o0 = 0, o1 = 0;                      // owned rows of the sorted arrays in the last step
    uint32_t expect[5] = {0, 0, 0, 0, 0};         // row ranges derived on the host, checked against the table one step later
    bool verify_pending = false;
    uint32_t stats[5] = {0, 0, 0, 0, 0};
    bool have_planes = false;
    // A capacity or consistency error of one rank must not leave its neighbours blocked in a matched exchange: every rank
    // keeps the communication pattern of the step, links whose two ends cannot both go ahead are skipped BY BOTH ENDS
    // (the verdict is computed from the same eight words on either side), the error is returned when the step has been
    // enqueued, and it is sticky: a failed rank says so in its next count message, so the failure spreads one hop per
    // step until every rank has returned it.  sph_upload_owned clears it.
    bool failed = false;
    std::string failure;
    cudaStream_t halo_stream = nullptr;           // halos 2 and 3 travel here, overlapped with interior compute
    cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr, ev_counts = nullptr, ev_msg = nullptr;
    cudaStream_t bnd_stream = nullptr;            // the boundary layers of a gather pass run here, concurrently with the interior
    cudaEvent_t ev_sorted = nullptr, ev_b[2] = {nullptr, nullptr}, ev_i[2] = {nullptr, nullptr};
    // SPH_SLAB_TIMING=1: finer timers of the spatial stage (events on the stream + host clock around the syncs)
    bool prof = false;
    cudaEvent_t pe[8] = {};
    double pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t psteps = 0, pseen = 0;
    // re-balancing (sph_comm_rebalance): the layers the table of the last step was built for, histogram buffers
    bool table_valid = false;
    int t_zlo = 0, t_own_lo = 0, t_own_hi = 0;
    uint32_t* hist_dev = nullptr;                 // [hist_cap] per-layer counts (+ 1 word: the smallest exchange buffer of any rank)
    uint32_t* hist_host = nullptr;                // pinned mirror
    uint32_t hist_cap = 0;
};

namespace {

// (the six lists -- L_MIG_LO ... -- and the pack block geometry live in sph_internal.h: k_predict_key counts into them;
// the three lists of a side are contiguous, so a side's counts travel as one message)

#define SPH_NCCL(c, call)                                                                   \
    do {                                                                                    \
        ncclResult_t r__ = (call);                                                          \
        if (r__ != ncclSuccess)                                                             \
            return fail((c), SPH_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

__device__ __forceinline__ bool in_list(uint8_t k, int list)
{
    switch (list) {
    case L_MIG_LO: return k & CLS_MIG_LO;
    case L_MIG_HI: return k & CLS_MIG_HI;
    case L_GHOST_LO: return k & CLS_GHOST_LO;
    case L_GHOST_HI: return k & CLS_GHOST_HI;
    case L_KEEP_LO: return (k & CLS_MIG_LO) && (k & CLS_KEEP);
    default: return (k & CLS_MIG_HI) && (k & CLS_KEEP);
    }
}

// the predicted position a ghost row carries: the owner's own expression (predict, sph_device.cuh)
__device__ __forceinline__ float4 ghost_pred_of(const float4 p, const float4 v0, const DevParams& P, float dt)
{
    float4 v = v0;
    float3 pr;
    predict(p, v, pr, P, dt);
    return make_float4(pr.x, pr.y, pr.z, 0.0f);
}

// The per-block list counts come from k_predict_key (it classifies the rows anyway).  One block per list: exclusive scan
// of its row of block counts into `offsets`, total to totals[list]; the block that finishes last writes the two count
// messages of the step.
// The count message of a side: (migrants, ghosts, kept migrants) towards that neighbour, then what the neighbour needs to
// reach the same verdict about the link as this rank: status (0 = fine), rows an exchange buffer holds, free rows and free
// ghost rows this rank can take FROM that side (half of what is left once its own kept migrants are in).
__global__ void __launch_bounds__(kPackThreads)
k_slab_scan(const uint32_t* __restrict__ block_counts, uint32_t* __restrict__ offsets, uint32_t* __restrict__ totals,
            const uint32_t nblocks, uint32_t* __restrict__ done, uint32_t* __restrict__ msg, const uint32_t status,
            const uint32_t xcap, const uint32_t cap, const uint32_t gcap, const uint32_t n_old)
{
    __shared__ uint32_t wsum[kPackThreads / 32];
    __shared__ uint32_t carry_s;
    const uint32_t* row = block_counts + (size_t)blockIdx.x * nblocks;
    uint32_t* out = offsets + (size_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += kPackThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? row[i] : 0u;
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; w++) wbase += wsum[w];
        const uint32_t carry = carry_s;
        if (i < nblocks) out[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == kPackThreads - 1) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    totals[blockIdx.x] = carry_s;
    __threadfence();
    if (atomicAdd(done, 1u) != NLISTS - 1) return;
    __threadfence();
    volatile const uint32_t* T = totals;
    const uint32_t keep = T[L_KEEP_LO] + T[L_KEEP_HI];
    const uint32_t used = n_old + keep;
    for (int side = 0; side < 2; side++) {                   // 0: to lo, 1: to hi
        uint32_t* m = msg + 8 * side;
        m[0] = T[3 * side]; m[1] = T[3 * side + 1]; m[2] = T[3 * side + 2];
        m[3] = status; m[4] = xcap;
        m[5] = cap > used ? (cap - used) / 2u : 0u;
        m[6] = gcap > keep ? (gcap - keep) / 2u : 0u;
        m[7] = 0u;
    }
}

// order-preserving pack: list l, entry rank = (#members in earlier blocks) + (#members before me in this block).
// A block covers kPackSpan rows as kPackIters slices of kPackThreads; one pass ballots every (slice, warp) into shared
// memory, one warp per list turns them into running offsets, a second pass writes -- three barriers per block.
__global__ void __launch_bounds__(kPackThreads)
k_slab_pack(const uint8_t* __restrict__ cls, const float4* __restrict__ pos, const float4* __restrict__ vel,
            const uint32_t* __restrict__ block_counts, const uint32_t* __restrict__ block_offsets,
            float4* __restrict__ mig_lo, float4* __restrict__ mig_hi,
            float4* __restrict__ ghost_lo, float4* __restrict__ ghost_hi, float4* __restrict__ keep_lo,
            float4* __restrict__ keep_hi, const uint32_t rows, const uint32_t nblocks, const uint32_t xcap,
            const DevParams P, const float dt)
{
    constexpr int kWarps = kPackThreads / 32, kCells = kPackIters * kWarps;     // (slice, warp) pairs in row order
    __shared__ uint32_t off[NLISTS][kCells];
    __shared__ uint32_t any_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        uint32_t a = 0;
        for (int l = 0; l < NLISTS; l++) a |= block_counts[(size_t)l * nblocks + blockIdx.x];
        any_s = a;
    }
    __syncthreads();
    if (!any_s) return;                              // interior blocks: nothing leaves, nothing is a ghost
    uint8_t k[kPackIters];
    #pragma unroll
    for (int it = 0; it < kPackIters; it++) {
        const uint32_t s = blockIdx.x * kPackSpan + it * kPackThreads + threadIdx.x;
        k[it] = s < rows ? cls[s] : (uint8_t)0;
    }
    #pragma unroll
    for (int it = 0; it < kPackIters; it++) {
        #pragma unroll
        for (int l = 0; l < NLISTS; l++) {
            const uint32_t bal = __ballot_sync(0xffffffffu, in_list(k[it], l));
            if (lane == 0) off[l][it * kWarps + warp] = __popc(bal);
        }
    }
    __syncthreads();
    if (warp < NLISTS) {                             // warp l: counts of list l -> offsets, kCells / 32 consecutive cells per lane
        constexpr int kPer = kCells / 32;
        uint32_t v[kPer], sum = 0;
        #pragma unroll
        for (int j = 0; j < kPer; j++) { v[j] = off[warp][lane * kPer + j]; sum += v[j]; }
        uint32_t inc = sum;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        uint32_t run = block_offsets[(size_t)warp * nblocks + blockIdx.x] + inc - sum;
        #pragma unroll
        for (int j = 0; j < kPer; j++) { off[warp][lane * kPer + j] = run; run += v[j]; }
    }
    __syncthreads();
    #pragma unroll 1
    for (int it = 0; it < kPackIters; it++) {
        const uint8_t kk = k[it];
        if (!__any_sync(0xffffffffu, kk != 0)) continue;
        const uint32_t s = blockIdx.x * kPackSpan + it * kPackThreads + threadIdx.x;
        float4 p = make_float4(0, 0, 0, 0), v = p, pr = p;
        if (kk) { p = pos[s]; v = vel[s]; pr = ghost_pred_of(p, v, P, dt); }
        #pragma unroll
        for (int l = 0; l < NLISTS; l++) {
            const bool in = in_list(kk, l);
            const uint32_t bal = __ballot_sync(0xffffffffu, in);
            if (!in) continue;
            const uint32_t o = off[l][it * kWarps + warp] + __popc(bal & ((1u << lane) - 1u));
            if (o >= xcap) continue;                         // overflow is detected on the host from the totals
            switch (l) {
            case L_MIG_LO: mig_lo[o] = p; mig_lo[xcap + o] = v; break;
            case L_MIG_HI: mig_hi[o] = p; mig_hi[xcap + o] = v; break;
            case L_GHOST_LO: ghost_lo[o] = pr; break;
            case L_GHOST_HI: ghost_hi[o] = pr; break;
            case L_KEEP_LO: keep_lo[o] = pr; break;
            default: keep_hi[o] = pr; break;
            }
        }
    }
}

__global__ void k_slab_pick(const uint32_t* __restrict__ table, uint32_t* __restrict__ out, uint32_t i0, uint32_t i1,
                            uint32_t i2, uint32_t i3, uint32_t i4, const DevParams P)
{
    if (threadIdx.x == 0) {
        const uint32_t t0 = tbl(table, P, i0), t1 = tbl(table, P, i1), t2 = tbl(table, P, i2), t3 = tbl(table, P, i3);
        out[0] = t0; out[1] = t1; out[2] = t2; out[3] = t3; out[4] = tbl(table, P, i4);
        // boundary-layer lengths this rank will SEND in the later halos: lo, hi
        out[5] = t1 - t0;
        out[6] = t3 - t2;
        // (boundary layer I send, ghost layer I expect) per side, at words 8..11 of the small buffer: the pair that
        // crosses a link in the synchronous cross-check
        uint32_t* pair = out - 8;
        pair[0] = t1 - t0; pair[1] = t0;
        pair[2] = t3 - t2; pair[3] = out[4] - t3;
    }
}

int ceil_log2_u64(uint64_t v) { int b = 0; while ((1ull << b) < v && b < 63) b++; return b; }

// re-balancing: rows per OWNED global z layer, read off the prefix table of the last step (layer l of the local table
// spans the entries [l * plane, (l + 1) * plane)); the other layers of the global histogram stay zero on this rank
__global__ void __launch_bounds__(256)
k_layer_hist(const uint32_t* __restrict__ table, uint32_t* __restrict__ hist, const uint32_t plane, const int zlo,
             const int own_lo, const int own_hi, const uint32_t seg_off)
{
    const int g = own_lo + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (g >= own_hi) return;
    const uint32_t l = (uint32_t)(g - zlo);
    hist[g] = tbl_at(table, seg_off, (l + 1u) * plane) - tbl_at(table, seg_off, l * plane);
}

}  // namespace

namespace sphb200 {

void multi_adopt_upload(SphContext* c, uint32_t n)
{
    if (c->slab) { c->slab->o0 = 0; c->slab->o1 = n; c->slab->table_valid = false; c->slab->failed = false; c->slab->failure.clear(); c->slab->verify_pending = false; }
}

int multi_params_changed(SphContext* c)
{
    if (!c->slab || !c->slab->have_planes) return SPH_OK;
    const std::vector<float> planes = c->planes;         // sph_comm_set_planes assigns c->planes
    c->slab->table_valid = false;                        // the last table was laid out for the old layer grid
    return sph_comm_set_planes(c, planes.data());
}

void multi_teardown(SphContext* c)
{
    if (c->comm) { ncclCommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
    SlabState* s = c->slab;
    if (!s) return;
    if (s->prof && s->psteps) {
        fprintf(stderr, "[slab rank %d] spatial sub-stages, ms/step over %llu steps: pack %.3f | count xchg %.3f | (host wait %.3f) | payload xchg %.3f | keys+sort %.3f | table+reorder %.3f | pick+check xchg %.3f | (host wait %.3f)\n",
                c->rank, (unsigned long long)s->psteps, s->pacc[0] / s->psteps, s->pacc[1] / s->psteps, s->pacc[6] / s->psteps,
                s->pacc[2] / s->psteps, s->pacc[3] / s->psteps, s->pacc[4] / s->psteps, s->pacc[5] / s->psteps, s->pacc[7] / s->psteps);
        for (auto& e : s->pe) if (e) cudaEventDestroy(e);
    }
    void* ptrs[] = {s->cls, s->mig_send[0], s->mig_send[1], s->ghost_send[0], s->ghost_send[1], s->keep[0], s->keep[1],
                    s->ghost_pred, s->block_counts, s->dev_small};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (s->host_small) cudaFreeHost(s->host_small);
    if (s->hist_dev) cudaFree(s->hist_dev);
    if (s->hist_host) cudaFreeHost(s->hist_host);
    if (s->halo_stream) { cudaStreamSynchronize(s->halo_stream); cudaStreamDestroy(s->halo_stream); }
    if (s->ev_boundary) cudaEventDestroy(s->ev_boundary);
    if (s->ev_halo) cudaEventDestroy(s->ev_halo);
    if (s->ev_counts) cudaEventDestroy(s->ev_counts);
    for (cudaEvent_t e : {s->ev_sorted, s->ev_b[0], s->ev_b[1], s->ev_i[0], s->ev_i[1]}) if (e) cudaEventDestroy(e);
    if (s->bnd_stream) { cudaStreamSynchronize(s->bnd_stream); cudaStreamDestroy(s->bnd_stream); }
    if (s->ev_msg) cudaEventDestroy(s->ev_msg);
    delete s;
    c->slab = nullptr;
}

static int slab_params(SphContext* c, DevParams* P, uint32_t n_rows)
{
    SlabState* s = c->slab;
    make_dev_params(c, n_rows, P);
    const int GZ = c->gdim[2];
    const int own_lo = s->layers[c->rank], own_hi = s->layers[c->rank + 1];
    const int has_lo = c->rank > 0, has_hi = c->rank < c->nranks - 1;
    const int zlo = own_lo - (has_lo ? 1 : 0), zhi = own_hi + (has_hi ? 1 : 0);
    P->slab = 1; P->zlo = zlo; P->gz_global = GZ; P->own_lo = own_lo; P->own_hi = own_hi;
    P->has_lo = has_lo; P->has_hi = has_hi;
    P->gdim[2] = zhi - zlo;
    P->ncell = (uint32_t)((uint64_t)P->gdim[0] * P->gdim[1] * P->gdim[2]);
    P->mode = SPH_TABLE_GRID;
    P->seg_off = (uint32_t)table_layout(P->ncell).cells_pad;      // the slab's own table, not the whole grid's
    P->cnt_off = P->seg_off + (uint32_t)table_layout(P->ncell).nseg_pad;
    return SPH_OK;
}

int multi_step(SphContext* c, float dt)
{
    SlabState* s = c->slab;
    if (!s || !c->comm) return fail(c, SPH_ERR_INVALID, "slab mode: call sph_comm_init first");
    if (!s->have_planes) return fail(c, SPH_ERR_INVALID, "slab mode: call sph_comm_set_planes first");
    SPH_CUDA(c, cudaSetDevice(c->device));
    s->table_valid = false;                     // until this step has rebuilt it (sph_comm_rebalance reads it)
    ncclComm_t comm = (ncclComm_t)c->comm;
    cudaStream_t st = c->st;
    const uint32_t n_old = c->n;
    const int lo = c->rank - 1, hi = c->rank + 1;
    const bool has_lo = c->rank > 0, has_hi = c->rank < c->nranks - 1;
    DevParams P;
    slab_params(c, &P, n_old);
    int rc = ensure_tables(c, P);
    if (rc != SPH_OK) return rc;
    const bool timing = c->timing;
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[0], st));

    // (1) predict + classify + key of the resident rows
    const bool binned = counting_sort_enabled();
    const size_t padded = scan_pad((size_t)P.ncell + 3);
    const TableLayout TL = table_layout(P.ncell);
    if (c->table_ncell != P.ncell) { c->table_two_level = false; c->table_ncell = P.ncell; }   // the planes moved: another layout
    c->table_seg_off = P.seg_off;
    const uint32_t nblocks = n_old ? (n_old + kPackSpan - 1) / kPackSpan : 1;
    bool zeroed = false;                        // the step's small counters: dev_small and the per-pack-block list counts
    if (binned) {
        if (!c->table_two_level) { SPH_CUDA(c, cudaMemsetAsync(c->tstart, 0, TL.total * sizeof(uint32_t), st)); c->table_two_level = true; }
        else { launch_table_clear(st, c->tstart, TL, &c->launches, s->dev_small, 64u, s->block_counts, NLISTS * nblocks); zeroed = true; }
    }
    if (!zeroed) {                              // (otherwise k_table_clear zeroed them on its way)
        SPH_CUDA(c, cudaMemsetAsync(s->dev_small, 0, 64 * sizeof(uint32_t), st));
        SPH_CUDA(c, cudaMemsetAsync(s->block_counts, 0, (size_t)NLISTS * nblocks * sizeof(uint32_t), st));
    }
    // ... which also counts, per pack block, the rows of each of the six lists (k_slab_scan / k_slab_pack below)
    launch_predict_key(st, c->A_pos, c->A_vel, c->key_a, s->cls, n_old, true, P, dt, binned ? c->tstart : nullptr, c->perm_b,
                       &c->launches, s->block_counts, nblocks);
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[1], st));

    #define SLAB_MARK(i) do { if (s->prof) cudaEventRecord(s->pe[i], st); } while (0)
    auto host_now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    SLAB_MARK(0);
    // (2) order-preserving pack of the six lists
    uint32_t* totals = s->dev_small;            // [6] (word 7: k_slab_scan's finished-block counter)
    uint32_t* picks = s->dev_small + 16;        // [8]
    uint32_t* peer = s->dev_small + 24;         // [2] boundary lengths of the neighbours' layers
    uint32_t* msg_out = s->dev_small + 32;
    uint32_t* msg_in = s->dev_small + 48;
    // (the per-block counts of the six lists were left by k_predict_key)
    uint32_t* offsets = s->block_counts + (size_t)NLISTS * s->nblocks_cap;
    k_slab_scan<<<NLISTS, kPackThreads, 0, st>>>(s->block_counts, offsets, totals, n_old ? nblocks : 0u, s->dev_small + 7, msg_out,
                                                 s->failed ? 1u : 0u, s->xcap, c->cap, s->gcap, n_old);
    ++c->launches;
    // (3a) count messages to / from the neighbours (written by the last block of k_slab_scan).  They travel on the halo stream, with the copy of the
    // counts to the host behind them, WHILE the solver's stream packs the six lists: the host's round trip (the one
    // synchronisation of the step) is hidden behind the pack kernel instead of leaving the GPU idle.
    cudaStream_t cs = s->halo_stream;
    SPH_CUDA(c, cudaEventRecord(s->ev_counts, st));
    SPH_CUDA(c, cudaStreamWaitEvent(cs, s->ev_counts, 0));
    SPH_NCCL(c, ncclGroupStart());
    if (has_lo) { SPH_NCCL(c, ncclSend(msg_out, 8, ncclUint32, lo, comm, cs)); SPH_NCCL(c, ncclRecv(msg_in, 8, ncclUint32, lo, comm, cs)); }
    if (has_hi) { SPH_NCCL(c, ncclSend(msg_out + 8, 8, ncclUint32, hi, comm, cs)); SPH_NCCL(c, ncclRecv(msg_in + 8, 8, ncclUint32, hi, comm, cs)); }
    SPH_NCCL(c, ncclGroupEnd());
    // (words 16..31 of the pinned mirror still hold the previous step's table picks: they are compared below)
    SPH_CUDA(c, cudaMemcpyAsync(s->host_small, s->dev_small, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
    SPH_CUDA(c, cudaMemcpyAsync(s->host_small + 32, s->dev_small + 32, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
    SPH_CUDA(c, cudaEventRecord(s->ev_msg, cs));
    // (2) order-preserving pack of the six lists
    if (n_old) {
        k_slab_pack<<<nblocks, kPackThreads, 0, st>>>(s->cls, c->A_pos, c->A_vel, s->block_counts, offsets, s->mig_send[0],
                                                      s->mig_send[1], s->ghost_send[0], s->ghost_send[1], s->keep[0],
                                                      s->keep[1], n_old, nblocks, s->xcap, P, dt);
        ++c->launches;
    }
    SLAB_MARK(1);
    SLAB_MARK(2);
    const double h0 = host_now();
    SPH_CUDA(c, cudaEventSynchronize(s->ev_msg));
    const double h1 = host_now();
    // The verdict about a link, from the two messages that crossed it: identical on both of its ends (sph_slab_link_ok).
    auto link_ok = [](const uint32_t* a, const uint32_t* b) { return sph_slab_link_ok(a, b) == 1; };
    const uint32_t* MO = s->host_small + 32;
    const uint32_t* MI = s->host_small + 48;
    const bool was_failed = s->failed;
    const bool link_lo = has_lo && link_ok(MO, MI), link_hi = has_hi && link_ok(MO + 8, MI + 8);
    if ((has_lo && !link_lo) || (has_hi && !link_hi)) {
        if (!s->failed) {
            auto why = [&](const uint32_t* mo, const uint32_t* mi, const char* side) -> std::string {
                if (mi[3]) return std::string("the ") + side + " neighbour reported a failure";
                if (mo[0] > mo[4] || mo[1] > mo[4] || mo[2] > mo[4]) return std::string("exchange buffer towards the ") + side + " neighbour too small (raise capacity)";
                if (mi[0] > mi[4] || mi[1] > mi[4] || mi[2] > mi[4]) return std::string("the ") + side + " neighbour's exchange buffer is too small";
                if ((uint64_t)mi[0] + mi[1] > mo[5] || mi[1] > mo[6]) return std::string("capacity too small for the arrivals + ghosts from the ") + side + " neighbour";
                return std::string("the ") + side + " neighbour has no room for this rank's migrants + ghosts";
            };
            s->failure = "slab mode: " + ((has_lo && !link_lo) ? why(MO, MI, "lower") : why(MO + 8, MI + 8, "upper"));
        }
        s->failed = true;
    }
    // a link that does not go ahead carries nothing this step, in either direction
    uint32_t Teff[NLISTS], Reff[NLISTS];
    for (int l = 0; l < 3; l++) {
        Teff[l] = link_lo ? MO[l] : 0u;       Reff[l] = link_lo ? MI[l] : 0u;
        Teff[3 + l] = link_hi ? MO[8 + l] : 0u; Reff[3 + l] = link_hi ? MI[8 + l] : 0u;
    }
    // kept migrants are this rank's own rows (ghost copies for its own boundary layers): they do not depend on the link
    Teff[L_KEEP_LO] = s->host_small[L_KEEP_LO] <= s->xcap ? s->host_small[L_KEEP_LO] : 0u;
    Teff[L_KEEP_HI] = s->host_small[L_KEEP_HI] <= s->xcap ? s->host_small[L_KEEP_HI] : 0u;
    if (!link_lo) Teff[L_KEEP_LO] = 0u;
    if (!link_hi) Teff[L_KEEP_HI] = 0u;
    const uint32_t* T = Teff;
    const uint32_t* R = Reff;
    const bool has_lo_x = link_lo, has_hi_x = link_hi;               // the links that exchange payloads and halos this step
    const uint32_t mig_in_lo = R[0], ghost_in_lo = R[1], mig_in_hi = R[3], ghost_in_hi = R[4];
    const uint32_t kept_by_lo = R[2], kept_by_hi = R[5];   // arrivals the neighbour still mirrors: they lie in my boundary layers
    // deferred check of the previous step's row ranges (computed on the host, see below) against the table
    if (s->verify_pending) {
        s->verify_pending = false;
        const uint32_t* V = s->host_small + 16;
        for (int q = 0; q < 5 && !s->failed; q++)
            if (V[q] != s->expect[q] && !(q == 1 && !has_lo) && !(q == 2 && !has_hi)) {    // no neighbour: no boundary layer on that side
                s->failed = true;
                s->failure = "slab mode: row ranges derived from the exchanged counts disagree with the table (range " +
                             std::to_string(q) + ": " + std::to_string(s->expect[q]) + " vs " + std::to_string(V[q]) + ")";
            }
    }
    const uint32_t n_a = n_old + mig_in_lo + mig_in_hi;                 // resident + arrived (departed rows still inside)
    const uint32_t n_ghost = ghost_in_lo + T[L_KEEP_LO] + ghost_in_hi + T[L_KEEP_HI];
    if ((uint64_t)n_a + n_ghost > c->cap || n_ghost > s->gcap)          // cannot happen for links that passed the verdict
        return fail(c, SPH_ERR_CAPACITY, "slab mode: internal error: accepted links exceed the capacity");
    // (3b) payloads: migrants land straight behind the resident rows, ghosts in the ghost buffer
    float4* g_recv_lo = s->ghost_pred;
    float4* g_keep_lo = g_recv_lo + ghost_in_lo;
    float4* g_recv_hi = g_keep_lo + T[L_KEEP_LO];
    float4* g_keep_hi = g_recv_hi + ghost_in_hi;
    SPH_NCCL(c, ncclGroupStart());
    if (has_lo_x) {
        if (T[L_MIG_LO]) {
            SPH_NCCL(c, ncclSend(s->mig_send[0], (size_t)T[L_MIG_LO] * 4, ncclFloat, lo, comm, st));
            SPH_NCCL(c, ncclSend(s->mig_send[0] + s->xcap, (size_t)T[L_MIG_LO] * 4, ncclFloat, lo, comm, st));
        }
        if (T[L_GHOST_LO]) SPH_NCCL(c, ncclSend(s->ghost_send[0], (size_t)T[L_GHOST_LO] * 4, ncclFloat, lo, comm, st));
        if (mig_in_lo) {
            SPH_NCCL(c, ncclRecv(c->A_pos + n_old, (size_t)mig_in_lo * 4, ncclFloat, lo, comm, st));
            SPH_NCCL(c, ncclRecv(c->A_vel + n_old, (size_t)mig_in_lo * 4, ncclFloat, lo, comm, st));
        }
        if (ghost_in_lo) SPH_NCCL(c, ncclRecv(g_recv_lo, (size_t)ghost_in_lo * 4, ncclFloat, lo, comm, st));
    }
    if (has_hi_x) {
        if (T[L_MIG_HI]) {
            SPH_NCCL(c, ncclSend(s->mig_send[1], (size_t)T[L_MIG_HI] * 4, ncclFloat, hi, comm, st));
            SPH_NCCL(c, ncclSend(s->mig_send[1] + s->xcap, (size_t)T[L_MIG_HI] * 4, ncclFloat, hi, comm, st));
        }
        if (T[L_GHOST_HI]) SPH_NCCL(c, ncclSend(s->ghost_send[1], (size_t)T[L_GHOST_HI] * 4, ncclFloat, hi, comm, st));
        if (mig_in_hi) {
            SPH_NCCL(c, ncclRecv(c->A_pos + n_old + mig_in_lo, (size_t)mig_in_hi * 4, ncclFloat, hi, comm, st));
            SPH_NCCL(c, ncclRecv(c->A_vel + n_old + mig_in_lo, (size_t)mig_in_hi * 4, ncclFloat, hi, comm, st));
        }
        if (ghost_in_hi) SPH_NCCL(c, ncclRecv(g_recv_hi, (size_t)ghost_in_hi * 4, ncclFloat, hi, comm, st));
    }
    SPH_NCCL(c, ncclGroupEnd());
    if (T[L_KEEP_LO]) SPH_CUDA(c, cudaMemcpyAsync(g_keep_lo, s->keep[0], (size_t)T[L_KEEP_LO] * 16, cudaMemcpyDeviceToDevice, st));
    if (T[L_KEEP_HI]) SPH_CUDA(c, cudaMemcpyAsync(g_keep_hi, s->keep[1], (size_t)T[L_KEEP_HI] * 16, cudaMemcpyDeviceToDevice, st));

    SLAB_MARK(3);
    // (4) keys of the arrivals (they may not migrate again this step) and of the ghosts; one sort of everything
    const uint32_t n_all = n_a + n_ghost;
    slab_params(c, &P, n_all);
    P.n_a = n_a;
    launch_predict_key(st, c->A_pos + n_old, c->A_vel + n_old, c->key_a + n_old, nullptr, n_a - n_old, false, P, dt,
                       binned ? c->tstart : nullptr, c->perm_b + n_old, &c->launches);
    launch_ghost_key(st, s->ghost_pred, c->key_a + n_a, n_ghost, P, binned ? c->tstart : nullptr, c->perm_b + n_a, &c->launches);
    if (binned) {
        // table over ncell + 1 "cells": the extra one collects the departed rows (key == ncell)
        exclusive_scan_u32(st, c->tstart + TL.cells_pad + TL.nseg_pad, c->tstart + TL.cells_pad, TL.nseg_pad, c->scan_tmp, &c->launches);
        launch_inseg_scan(st, c->tstart, TL, &c->launches);
        launch_place(st, c->key_a, c->perm_b, c->tstart, c->perm_a, n_all, P, &c->launches);
        SLAB_MARK(4);
        launch_reorder(st, c->perm_a, c->key_a, c->tstart, c->key_b, c->A_pos, c->A_vel, s->ghost_pred, c->S_pos, c->S_vel, c->pred, c->predpk, P, dt, &c->launches);
        c->sorted_where = 1;
    } else {
        const int bits = ceil_log2_u64((uint64_t)P.ncell + 1);
        c->sorted_where = radix_sort_pairs(st, c->key_a, c->key_b, c->perm_a, c->perm_b, true, n_all, bits, c->counts, &c->launches);
        SLAB_MARK(4);
        const uint32_t* keys = c->sorted_where ? c->key_b : c->key_a;
        const uint32_t* perm = c->sorted_where ? c->perm_b : c->perm_a;
        DevParams PT = P;
        PT.ncell = P.ncell + 1;
        launch_build_table(st, keys, c->tstart, c->tend, c->gap_list, PT, &c->launches);
        c->table_two_level = false;
        launch_reorder(st, perm, nullptr, nullptr, nullptr, c->A_pos, c->A_vel, s->ghost_pred, c->S_pos, c->S_vel, c->pred, c->predpk, P, dt,
                       &c->launches);
    }
    SLAB_MARK(5);
    // owned rows and boundary layers of the sorted arrays
    const uint32_t plane = (uint32_t)P.gdim[0] * (uint32_t)P.gdim[1];
    const uint32_t l_own_lo = (uint32_t)(P.own_lo - P.zlo), l_own_hi = (uint32_t)(P.own_hi - P.zlo);
    k_slab_pick<<<1, 32, 0, st>>>(c->tstart, picks, l_own_lo * plane, (l_own_lo + 1) * plane, (l_own_hi - 1) * plane,
                                  l_own_hi * plane, P.ncell, P);
    ++c->launches;
    // Row ranges of the sorted arrays.  With at least three owned layers they follow from the exchanged counts --
    //   ghost-lo layer  = ghosts received from lo + my migrants to lo that I keep mirroring
    //   lo boundary     = my rows flagged ghost-for-lo + the arrivals lo still mirrors          (same on the hi side)
    //   departed        = my migrants
    // -- so the step needs no second host synchronisation; the table's own answer (k_slab_pick) is copied back
    // asynchronously and compared at the next step's synchronisation.  Thin slabs (boundary layers may coincide) and
    // SPH_SLAB_CHECK=1 read the table now and cross-check the layer lengths with the neighbours, as before.
    static const bool force_check = [] { const char* e = getenv("SPH_SLAB_CHECK"); return e && e[0] == '1'; }();
    uint32_t o0, b_lo_end, b_hi_begin, o1, live_end;
    bool halo_lo = has_lo_x, halo_hi = has_hi_x;          // the links whose boundary / ghost layers travel in the gather stage
    const double h2 = host_now();
    if (!force_check && P.own_hi - P.own_lo >= 3) {
        o0 = ghost_in_lo + T[L_KEEP_LO];
        live_end = n_all - (s->host_small[L_MIG_LO] + s->host_small[L_MIG_HI]);     // every row classified as a migrant sorted behind the table
        o1 = live_end - (ghost_in_hi + T[L_KEEP_HI]);
        b_lo_end = o0 + T[L_GHOST_LO] + kept_by_lo;
        b_hi_begin = o1 - (T[L_GHOST_HI] + kept_by_hi);
        SPH_CUDA(c, cudaMemcpyAsync(s->host_small + 16, s->dev_small + 16, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        s->expect[0] = o0; s->expect[1] = b_lo_end; s->expect[2] = b_hi_begin; s->expect[3] = o1; s->expect[4] = live_end;
        s->verify_pending = true;
        SLAB_MARK(6);
    } else {
        // cross-check: the neighbour's boundary layer must be exactly as long as my ghost layer, and the other way
        // round; both numbers cross the link, so its two ends reach the same verdict
        uint32_t* pair_out = s->dev_small + 8;      // written by k_slab_pick: (my boundary, my ghost layer) lo, then hi
        uint32_t* pair_in = s->dev_small + 12;
        SPH_NCCL(c, ncclGroupStart());
        if (has_lo_x) { SPH_NCCL(c, ncclSend(pair_out, 2, ncclUint32, lo, comm, st)); SPH_NCCL(c, ncclRecv(pair_in, 2, ncclUint32, lo, comm, st)); }
        if (has_hi_x) { SPH_NCCL(c, ncclSend(pair_out + 2, 2, ncclUint32, hi, comm, st)); SPH_NCCL(c, ncclRecv(pair_in + 2, 2, ncclUint32, hi, comm, st)); }
        SPH_NCCL(c, ncclGroupEnd());
        SPH_CUDA(c, cudaMemcpyAsync(s->host_small + 8, s->dev_small + 8, 24 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SLAB_MARK(6);
        SPH_CUDA(c, cudaStreamSynchronize(st));
        const uint32_t* K = s->host_small + 16;
        o0 = K[0]; b_lo_end = K[1]; b_hi_begin = K[2]; o1 = K[3]; live_end = K[4];
        const uint32_t* PO = s->host_small + 8;
        const uint32_t* PI = s->host_small + 12;
        const bool agree_lo = !has_lo_x || (PI[0] == PO[1] && PI[1] == PO[0]);
        const bool agree_hi = !has_hi_x || (PI[2] == PO[3] && PI[3] == PO[2]);
        if (!agree_lo || !agree_hi) {
            // that link carries no halo this step -- on either end -- the step is completed and the error returned at its end
            if (!s->failed)
                s->failure = "slab mode: ghost layer and neighbour boundary layer disagree (" + std::to_string(PO[1]) + " vs " +
                             std::to_string(PI[0]) + ", " + std::to_string(PO[3]) + " vs " + std::to_string(PI[2]) + ")";
            s->failed = true;
            if (!agree_lo) halo_lo = false;
            if (!agree_hi) halo_hi = false;
        }
    }
    const double h3 = host_now();
    if (s->prof && ++s->pseen > 3) {
        SPH_CUDA(c, cudaEventSynchronize(s->pe[6]));
        float ms;
        for (int i = 0; i < 6; i++) { cudaEventElapsedTime(&ms, s->pe[i], s->pe[i + 1]); s->pacc[i] += ms; }
        s->pacc[6] += h1 - h0; s->pacc[7] += h3 - h2;
        s->psteps++;
    }
    s->o0 = o0; s->o1 = o1;
    P.row0 = o0; P.row1 = o1;
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[2], st));

    // (5) gather passes.  Each pass runs its two BOUNDARY layers first; their results go to the neighbours
    // on the halo stream while the interior rows (which never touch a ghost row) are computed, so the halo
    // the NEXT pass needs arrives during this pass's interior work.
    NbrList L;
    rc = ensure_list(c, &L);
    if (rc != SPH_OK) return rc;
    const uint32_t hi_begin = b_hi_begin > b_lo_end ? b_hi_begin : b_lo_end;      // thin slab: the layers may coincide
    const uint32_t seg[3][2] = {{o0, b_lo_end}, {hi_begin, o1}, {b_lo_end, hi_begin}};   // lo layer, hi layer, interior
    cudaStream_t hs = s->halo_stream;
    auto halo = [&](void* base, const size_t fpr, cudaEvent_t boundary_done) -> int {   // boundary layers out, ghost layers in (contiguous ranges of rows of `fpr` floats)
        float* rows = static_cast<float*>(base);
        SPH_CUDA(c, cudaStreamWaitEvent(hs, boundary_done, 0));
        SPH_NCCL(c, ncclGroupStart());
        if (halo_lo) {
            if (b_lo_end > o0) SPH_NCCL(c, ncclSend(rows + o0 * fpr, (size_t)(b_lo_end - o0) * fpr, ncclFloat, lo, comm, hs));
            if (o0) SPH_NCCL(c, ncclRecv(rows, (size_t)o0 * fpr, ncclFloat, lo, comm, hs));
        }
        if (halo_hi) {
            if (o1 > b_hi_begin) SPH_NCCL(c, ncclSend(rows + b_hi_begin * fpr, (size_t)(o1 - b_hi_begin) * fpr, ncclFloat, hi, comm, hs));
            if (live_end > o1) SPH_NCCL(c, ncclRecv(rows + o1 * fpr, (size_t)(live_end - o1) * fpr, ncclFloat, hi, comm, hs));
        }
        SPH_NCCL(c, ncclGroupEnd());
        SPH_CUDA(c, cudaEventRecord(s->ev_halo, hs));
        return SPH_OK;
    };
    // Boundary layers and interior of a pass are independent of each other, so they run CONCURRENTLY: the two boundary
    // launches on a high-priority stream of their own (their blocks get the SMs first, their results leave on the halo
    // stream as soon as they exist), the interior launch on the solver's stream, filling the machine around them --
    // no tail of a small launch is waited for.  What a pass needs from the previous one crosses streams by events:
    //   boundary rows of pass k+1  <-  ghost rows (halo of pass k) + interior rows of pass k (their inward neighbours)
    //   interior rows of pass k+1  <-  boundary rows of pass k (their outward neighbours)
    cudaStream_t bs = s->bnd_stream;
    DevParams Q = P;
    auto boundary = [&](auto&& launch_rows) {
        for (int g = 0; g < 2; g++) { Q.row0 = seg[g][0]; Q.row1 = seg[g][1]; launch_rows(bs); }
    };
    auto interior = [&](auto&& launch_rows) { Q.row0 = seg[2][0]; Q.row1 = seg[2][1]; launch_rows(st); };
    SPH_CUDA(c, cudaEventRecord(s->ev_sorted, st));
    SPH_CUDA(c, cudaStreamWaitEvent(bs, s->ev_sorted, 0));
    // density
    auto dens_rows = [&](cudaStream_t q) { launch_density(q, c->pred, c->predpk, c->tstart, c->tend, c->dens, L, Q, &c->launches); };
    boundary(dens_rows);
    SPH_CUDA(c, cudaEventRecord(s->ev_b[0], bs));
    rc = halo(c->dens, 8, s->ev_b[0]); if (rc != SPH_OK) return rc;
    interior(dens_rows);
    SPH_CUDA(c, cudaEventRecord(s->ev_i[0], st));
    SPH_CUDA(c, cudaStreamWaitEvent(st, s->ev_b[0], 0));            // the solver's stream has every owned density from here on
    if (L.idx) { rc = copy_list_words(c, st); if (rc != SPH_OK) return rc; }        // on a side stream: nothing waits for it
    c->ncount_valid = true;
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[3], st));

    // pressure: the boundary layers need the ghost densities that were travelling during the interior density work
    auto pres_rows = [&](cudaStream_t q) { launch_pressure(q, c->pred, c->dens, c->S_vel, c->tstart, c->tend, c->velp, L, Q, dt, &c->launches); };
    SPH_CUDA(c, cudaStreamWaitEvent(bs, s->ev_halo, 0));
    SPH_CUDA(c, cudaStreamWaitEvent(bs, s->ev_i[0], 0));
    boundary(pres_rows);
    SPH_CUDA(c, cudaEventRecord(s->ev_b[1], bs));
    rc = halo(c->velp, 4, s->ev_b[1]); if (rc != SPH_OK) return rc;
    interior(pres_rows);
    SPH_CUDA(c, cudaStreamWaitEvent(st, s->ev_b[1], 0));
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[4], st));

    SPH_CUDA(c, cudaStreamWaitEvent(st, s->ev_halo, 0));    // ghost post-pressure velocities are in
    launch_viscosity(st, c->pred, c->velp, c->tstart, c->tend, c->S_vel, L, P, dt, &c->launches);
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[5], st));
    launch_integrate(st, c->S_pos, c->S_vel, c->A_pos, c->A_vel, P, dt, &c->launches);
    if (timing) SPH_CUDA(c, cudaEventRecord(c->ev[6], st));
    c->ev_recorded = timing;
    SPH_CUDA(c, cudaGetLastError());
    c->n = o1 - o0;
    c->step_valid = true;
    s->stats[0] = c->n; s->stats[1] = o0; s->stats[2] = live_end - o1; s->stats[3] = T[L_MIG_LO]; s->stats[4] = T[L_MIG_HI];
    s->t_zlo = P.zlo; s->t_own_lo = P.own_lo; s->t_own_hi = P.own_hi; s->table_valid = true;
    if (s->failed) {                      // the step was enqueued in full (no neighbour is left waiting); its result is not valid
        (void)was_failed;
        return fail(c, SPH_ERR_CAPACITY, s->failure.empty() ? std::string("slab mode: a neighbour rank failed") : s->failure);
    }
    return SPH_OK;
}

}  // namespace sphb200

extern "C" {

// Pure host function (no device, no communicator): may the link between two slab neighbours carry its payloads this
// step?  `mine` is the 8-word count message this rank sent over the link, `theirs` the one it received (k_slab_scan:
// migrants, ghosts, kept migrants, status, exchange-buffer rows, free rows, free ghost rows, 0).  The expression is
// symmetric -- sph_slab_link_ok(a, b) == sph_slab_link_ok(b, a) -- so both ends decide alike without another exchange.
int sph_slab_link_ok(const uint32_t* mine, const uint32_t* theirs)
{
    if (!mine || !theirs) return -1;
    auto one_way = [](const uint32_t* from, const uint32_t* to) {
        return from[0] <= from[4] && from[1] <= from[4] && from[2] <= from[4] &&      // the sender's lists fit its exchange buffers
               (uint64_t)from[0] + from[1] <= to[5] && from[1] <= to[6];              // ... and the receiver has the room
    };
    return (mine[3] == 0u && theirs[3] == 0u && one_way(mine, theirs) && one_way(theirs, mine)) ? 1 : 0;
}

size_t sph_comm_id_bytes(void) { return sizeof(ncclUniqueId); }

int sph_comm_get_id(void* id_out, size_t id_bytes)
{
    if (!id_out || id_bytes < sizeof(ncclUniqueId)) return SPH_ERR_INVALID;
    if (!nccl().ok) return fail(nullptr, SPH_ERR_NCCL, nccl().error);
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, SPH_ERR_NCCL, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
    memcpy(id_out, &id, sizeof(id));
    return SPH_OK;
}

int sph_comm_init(SphContext* c, int rank, int nranks, const void* id, size_t id_bytes)
{
    if (!c || !id || id_bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks) return SPH_ERR_INVALID;
    if (c->comm) return fail(c, SPH_ERR_INVALID, "sph_comm_init: already initialised");
    if (!nccl().ok) return fail(c, SPH_ERR_NCCL, nccl().error);
    SPH_CUDA(c, cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    SPH_NCCL(c, ncclCommInitRank(&comm, nranks, uid, rank));
    c->comm = (ncclComm*)comm;
    c->rank = rank; c->nranks = nranks; c->mode = SPH_TABLE_GRID;
    SlabState* s = new SlabState();
    c->slab = s;
    s->xcap = c->cap / 8 > 65536 ? c->cap / 8 : (c->cap < 65536 ? c->cap : 65536);
    s->gcap = 4 * s->xcap;
    s->nblocks_cap = (c->cap + kPackSpan - 1) / kPackSpan + 1;
    SPH_CUDA(c, cudaMalloc(&s->cls, c->cap));
    for (int d = 0; d < 2; d++) {
        SPH_CUDA(c, cudaMalloc(&s->mig_send[d], (size_t)s->xcap * 32));
        SPH_CUDA(c, cudaMalloc(&s->ghost_send[d], (size_t)s->xcap * 16));
        SPH_CUDA(c, cudaMalloc(&s->keep[d], (size_t)s->xcap * 16));
    }
    SPH_CUDA(c, cudaMalloc(&s->ghost_pred, (size_t)s->gcap * 16));
    SPH_CUDA(c, cudaMalloc(&s->block_counts, (size_t)2 * NLISTS * s->nblocks_cap * 4));   // counts, then offsets
    SPH_CUDA(c, cudaMalloc(&s->dev_small, 64 * sizeof(uint32_t)));
    SPH_CUDA(c, cudaMallocHost(&s->host_small, 64 * sizeof(uint32_t)));
    {   // highest priority: the halo's NCCL kernel must get SM slots ahead of the queued blocks of the interior pass
        int lo_p = 0, hi_p = 0;
        SPH_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
        const char* e = getenv("SPH_HALO_PRIO");
        SPH_CUDA(c, cudaStreamCreateWithPriority(&s->halo_stream, cudaStreamNonBlocking, (e && e[0] == '0') ? lo_p : hi_p));
        // the boundary layers of a pass: ahead of the interior blocks too (their results are what the neighbours wait for)
        SPH_CUDA(c, cudaStreamCreateWithPriority(&s->bnd_stream, cudaStreamNonBlocking, hi_p));
    }
    SPH_CUDA(c, cudaEventCreateWithFlags(&s->ev_boundary, cudaEventDisableTiming));
    SPH_CUDA(c, cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming));
    SPH_CUDA(c, cudaEventCreateWithFlags(&s->ev_counts, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&s->ev_sorted, &s->ev_b[0], &s->ev_b[1], &s->ev_i[0], &s->ev_i[1]}) SPH_CUDA(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    SPH_CUDA(c, cudaEventCreateWithFlags(&s->ev_msg, cudaEventDisableTiming));
    if (const char* e = getenv("SPH_SLAB_TIMING")) {
        s->prof = e[0] == '1';
        if (s->prof) for (auto& ev : s->pe) SPH_CUDA(c, cudaEventCreate(&ev));
    }
    return SPH_OK;
}

int sph_comm_set_planes(SphContext* c, const float* planes)
{
    if (!c || !planes) return SPH_ERR_INVALID;
    SlabState* s = c->slab;
    if (!s) return fail(c, SPH_ERR_INVALID, "sph_comm_set_planes: call sph_comm_init first");
    const int GZ = c->gdim[2];
    std::vector<int> L(c->nranks + 1);
    for (int k = 0; k <= c->nranks; k++) {
        int l = (int)floorf(planes[k] / c->params.interaction_radius) - c->gmin[2];
        L[k] = l < 0 ? 0 : (l > GZ ? GZ : l);
    }
    L[0] = 0; L[c->nranks] = GZ;
    for (int k = 0; k < c->nranks; k++)
        if (L[k + 1] - L[k] < 3) return fail(c, SPH_ERR_INVALID, "sph_comm_set_planes: every slab needs at least three cell layers");
    s->layers = L;
    s->have_planes = true;
    c->planes.assign(planes, planes + c->nranks + 1);
    return SPH_OK;
}

int sph_comm_get_layers(const SphContext* c, int32_t* layers_out)
{
    if (!c || !layers_out || !c->slab || !c->slab->have_planes) return SPH_ERR_INVALID;
    for (int k = 0; k <= c->nranks; k++) layers_out[k] = c->slab->layers[k];
    return SPH_OK;
}

int sph_slab_balance_layers(const uint32_t* hist, int32_t gz, int32_t nranks, const int32_t* layers_old, uint32_t max_shift,
                            uint64_t row_budget, int32_t* layers_new)
{
    constexpr int kMinLayers = 3;                   // sph_comm_set_planes: every slab needs three cell layers
    if (!hist || !layers_new || nranks < 1 || gz < kMinLayers * nranks) return SPH_ERR_INVALID;
    const int R = nranks;
    std::vector<uint64_t> cum((size_t)gz + 1, 0);
    for (int l = 0; l < gz; l++) cum[l + 1] = cum[l] + hist[l];
    const uint64_t total = cum[gz];
    if (layers_old) {
        if (layers_old[0] != 0 || layers_old[R] != gz) return SPH_ERR_INVALID;
        for (int k = 0; k < R; k++) if (layers_old[k + 1] - layers_old[k] < kMinLayers) return SPH_ERR_INVALID;
    }
    const int s = (int)(max_shift > 3u ? 3u : max_shift);       // a plane never crosses its old neighbours (single-hop migration)
    std::vector<int> L((size_t)R + 1, 0);
    L[R] = gz;
    for (int k = 1; k < R; k++) {
        // the layer boundary whose cumulative count is nearest to k/R of the total (exact integer arithmetic; the
        // lower boundary on a tie)
        const unsigned __int128 target = (unsigned __int128)total * (unsigned)k;
        int lo = 0, hi = gz;                        // first boundary with cum * R >= target
        while (lo < hi) { const int m = (lo + hi) / 2; if ((unsigned __int128)cum[m] * (unsigned)R >= target) hi = m; else lo = m + 1; }
        int cut = lo;
        if (cut > 0 && target - (unsigned __int128)cum[cut - 1] * (unsigned)R <= (unsigned __int128)cum[cut] * (unsigned)R - target) cut--;
        if (layers_old) {
            const int old = layers_old[k];
            if (cut > old + s) cut = old + s;
            if (cut < old - s) cut = old - s;
            if (row_budget) {                        // rows that change owner through this plane
                while (cut > old && cum[cut] - cum[old] > row_budget) cut--;
                while (cut < old && cum[old] - cum[cut] > row_budget) cut++;
            }
            // hysteresis: a quantile that sits in the middle of a layer makes both of its boundaries equally good, and
            // the plane would flip between them (a whole layer migrating each time) on the smallest change of the
            // histogram.  A move must bring the plane nearer to its quantile by at least a quarter of the rows it moves.
            while (cut != old) {
                const unsigned __int128 c_old = (unsigned __int128)cum[old] * (unsigned)R, c_new = (unsigned __int128)cum[cut] * (unsigned)R;
                const unsigned __int128 d_old = c_old > target ? c_old - target : target - c_old;
                const unsigned __int128 d_new = c_new > target ? c_new - target : target - c_new;
                const unsigned __int128 moved = (c_new > c_old ? c_new - c_old : c_old - c_new);
                if (d_old > d_new && (d_old - d_new) * 4u >= moved) break;
                cut += cut > old ? -1 : 1;           // not worth it: try the next nearer boundary, down to staying put
            }
        }
        L[k] = cut;
    }
    for (int k = 1; k < R; k++) if (L[k] < L[k - 1] + kMinLayers) L[k] = L[k - 1] + kMinLayers;
    for (int k = R - 1; k >= 1; k--) if (L[k] > L[k + 1] - kMinLayers) L[k] = L[k + 1] - kMinLayers;
    if (layers_old) {
        // the thickness passes can push a plane past what the limits allowed; then nothing moves this time
        bool ok = true;
        for (int k = 1; k < R && ok; k++) {
            const int old = layers_old[k], d = L[k] > old ? L[k] - old : old - L[k];
            const uint64_t moved = L[k] > old ? cum[L[k]] - cum[old] : cum[old] - cum[L[k]];
            if (d > s || (row_budget && moved > row_budget)) ok = false;
        }
        if (!ok) for (int k = 1; k < R; k++) L[k] = layers_old[k];
    }
    for (int k = 0; k <= R; k++) layers_new[k] = L[k];
    return SPH_OK;
}

int sph_comm_rebalance(SphContext* c, uint32_t max_shift, int32_t* layers_out, uint32_t* hist_out, size_t hist_entries, int* changed_out)
{
    if (!c) return SPH_ERR_INVALID;
    SlabState* s = c->slab;
    if (!s || !c->comm) return fail(c, SPH_ERR_INVALID, "sph_comm_rebalance: call sph_comm_init first");
    if (!s->have_planes) return fail(c, SPH_ERR_INVALID, "sph_comm_rebalance: call sph_comm_set_planes first");
    const int GZ = c->gdim[2];
    // the layers the last step's table was built for (a single-rank context steps through the plain path: whole grid)
    bool table_valid = s->table_valid;
    int t_zlo = s->t_zlo, t_own_lo = s->t_own_lo, t_own_hi = s->t_own_hi;
    if (c->nranks == 1) { table_valid = c->step_valid && c->mode == SPH_TABLE_GRID && !c->grid_too_large; t_zlo = 0; t_own_lo = 0; t_own_hi = GZ; }
    if (!table_valid) return fail(c, SPH_ERR_INVALID, "sph_comm_rebalance: needs a completed sph_step (the histogram is read off its table)");
    if (hist_out && hist_entries < (size_t)GZ) return fail(c, SPH_ERR_INVALID, "sph_comm_rebalance: hist_out too small");
    SPH_CUDA(c, cudaSetDevice(c->device));
    ncclComm_t comm = (ncclComm_t)c->comm;
    cudaStream_t st = c->st;
    if (s->hist_cap < (uint32_t)GZ + 1u) {
        SPH_CUDA(c, cudaStreamSynchronize(st));
        if (s->hist_dev) cudaFree(s->hist_dev);
        if (s->hist_host) cudaFreeHost(s->hist_host);
        s->hist_dev = nullptr; s->hist_host = nullptr; s->hist_cap = 0;
        SPH_CUDA(c, cudaMalloc(&s->hist_dev, ((size_t)GZ + 1) * sizeof(uint32_t)));
        SPH_CUDA(c, cudaMallocHost(&s->hist_host, ((size_t)GZ + 1) * sizeof(uint32_t)));
        s->hist_cap = (uint32_t)GZ + 1u;
    }
    // global histogram: every rank fills its owned layers, the sum over ranks is the whole column.  One extra word,
    // reduced with ncclMin, carries the smallest exchange buffer of any rank (it bounds the rows a plane may move).
    SPH_CUDA(c, cudaMemsetAsync(s->hist_dev, 0, ((size_t)GZ + 1) * sizeof(uint32_t), st));
    const int own = t_own_hi - t_own_lo;
    const uint32_t plane = (uint32_t)c->gdim[0] * (uint32_t)c->gdim[1];
    if (own > 0) {
        k_layer_hist<<<(own + 255) / 256, 256, 0, st>>>(c->tstart, s->hist_dev, plane, t_zlo, t_own_lo, t_own_hi, c->table_seg_off);
        ++c->launches;
    }
    SPH_CUDA(c, cudaMemcpyAsync(s->hist_dev + GZ, &s->xcap, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SPH_NCCL(c, ncclAllReduce(s->hist_dev, s->hist_dev, (size_t)GZ, ncclUint32, ncclSum, comm, st));
    SPH_NCCL(c, ncclAllReduce(s->hist_dev + GZ, s->hist_dev + GZ, 1, ncclUint32, ncclMin, comm, st));
    SPH_CUDA(c, cudaMemcpyAsync(s->hist_host, s->hist_dev, ((size_t)GZ + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SPH_CUDA(c, cudaStreamSynchronize(st));
    if (hist_out) memcpy(hist_out, s->hist_host, (size_t)GZ * sizeof(uint32_t));
    std::vector<int32_t> fresh((size_t)c->nranks + 1);
    bool changed = false;
    if (max_shift && c->nranks > 1) {
        // rows changing owner travel as migrants (+ one layer of them as kept ghosts): half an exchange buffer at most
        const uint64_t budget = s->hist_host[GZ] / 2u;
        const int rc = sph_slab_balance_layers(s->hist_host, GZ, c->nranks, s->layers.data(), max_shift, budget ? budget : 1u, fresh.data());
        if (rc != SPH_OK) return fail(c, rc, "sph_comm_rebalance: the layer histogram cannot be cut (fewer than three layers per rank?)");
        for (int k = 0; k <= c->nranks; k++) changed |= fresh[k] != s->layers[k];
        if (changed) {
            for (int k = 0; k <= c->nranks; k++) {
                s->layers[k] = fresh[k];
                c->planes[k] = ((float)(fresh[k] + c->gmin[2]) + 0.5f) * c->params.interaction_radius;   // a z inside the slab's first layer
            }
        }
    }
    if (layers_out) for (int k = 0; k <= c->nranks; k++) layers_out[k] = s->layers[k];
    if (changed_out) *changed_out = changed ? 1 : 0;
    return SPH_OK;
}

int sph_upload_owned(SphContext* c, uint32_t n, const uint32_t* global_id, const float* pos3, const float* vel3)
{
    if (!c) return SPH_ERR_INVALID;
    if (n > c->cap) return fail(c, SPH_ERR_CAPACITY, "sph_upload_owned: n exceeds capacity");
    if (n && (!pos3 || !global_id)) return fail(c, SPH_ERR_INVALID, "sph_upload_owned: NULL input");
    SPH_CUDA(c, cudaSetDevice(c->device));
    float* dpos = (float*)c->stage;
    float* dvel = (float*)(c->stage + (size_t)c->cap * 12);
    uint32_t* dids = (uint32_t*)(c->stage + (size_t)c->cap * 24);
    if (n) {
        SPH_CUDA(c, cudaMemcpyAsync(dpos, pos3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st));
        if (vel3) SPH_CUDA(c, cudaMemcpyAsync(dvel, vel3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st));
        SPH_CUDA(c, cudaMemcpyAsync(dids, global_id, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
        launch_pack_state(c->st, dpos, vel3 ? dvel : nullptr, dids, c->A_pos, c->A_vel, n, &c->launches);
        SPH_CUDA(c, cudaGetLastError());
    }
    SPH_CUDA(c, cudaStreamSynchronize(c->st));
    c->n = n;
    c->step_valid = false;
    c->ncount_valid = false;
    if (c->slab) { c->slab->o0 = 0; c->slab->o1 = n; c->slab->table_valid = false; c->slab->failed = false; c->slab->failure.clear(); c->slab->verify_pending = false; }
    return SPH_OK;
}

static size_t owned_field_bytes(int field)
{
    switch (field) {
    case SPH_FIELD_POSITIONS: case SPH_FIELD_VELOCITIES: case SPH_FIELD_PREDICTED: case SPH_FIELD_VEL_AFTER_PRESSURE:
    case SPH_FIELD_VEL_AFTER_VISCOSITY: return 12;
    case SPH_FIELD_OUT_POSITIONS: case SPH_FIELD_COLORS: return 16;
    case SPH_FIELD_DENSITIES: return 8;
    case SPH_FIELD_HASH: case SPH_FIELD_KEY: case SPH_FIELD_NEIGHBOUR_COUNT: case SPH_FIELD_SPEED_NORMALIZED: return 4;
    default: return 0;
    }
}

// export `field` of the owned rows, device order, into dev_out (enqueued on the solver's stream)
static int owned_export(SphContext* c, int field, void* dev_out, uint32_t n, bool by_id = false)
{
    // per-step arrays live at sorted rows [o0, o1); the state arrays were compacted to [0, n)
    const uint32_t off = c->slab ? c->slab->o0 : 0;
    const void* src = nullptr;
    bool needs_step = true;
    switch (field) {
    case SPH_FIELD_POSITIONS: case SPH_FIELD_OUT_POSITIONS: src = c->A_pos; needs_step = false; break;
    case SPH_FIELD_VELOCITIES: case SPH_FIELD_SPEED_NORMALIZED: case SPH_FIELD_COLORS: src = c->A_vel; needs_step = false; break;
    case SPH_FIELD_DENSITIES: src = c->dens + off; break;
    case SPH_FIELD_PREDICTED: case SPH_FIELD_HASH: case SPH_FIELD_KEY: src = c->pred + off; break;
    case SPH_FIELD_VEL_AFTER_PRESSURE: src = c->velp + off; break;
    case SPH_FIELD_VEL_AFTER_VISCOSITY: src = c->S_vel + off; break;
    case SPH_FIELD_NEIGHBOUR_COUNT:
        if (!c->ncount_valid) return fail(c, SPH_ERR_INVALID, "neighbour counts not recorded");
        src = c->ncount + off; break;
    default: return fail(c, SPH_ERR_INVALID, "unknown field");
    }
    if (needs_step && !c->step_valid) return fail(c, SPH_ERR_INVALID, "field needs a step first");
    DevParams P;
    make_dev_params(c, n, &P);
    launch_export(c->st, field, c->A_pos, src, nullptr, dev_out, n, P, by_id, &c->launches);
    SPH_CUDA(c, cudaGetLastError());
    return SPH_OK;
}

int sph_download_owned_scatter(SphContext* c, int field, void* host_base, size_t host_elems, uint32_t* out_n)
{
    if (!c) return SPH_ERR_INVALID;
    SPH_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n = c->n;
    if (out_n) *out_n = n;
    if (!owned_field_bytes(field)) return fail(c, SPH_ERR_INVALID, "unknown field");
    if (!n) return SPH_OK;
    if (!host_base || host_elems < n) return fail(c, SPH_ERR_INVALID, "sph_download_owned_scatter: host array missing or smaller than the owned count");
    // the export kernel writes row s to host_base[id(s)] itself, over PCIe: the array must be page-locked and mapped
    // into this device's address space (sph_host_register: portable + mapped under unified addressing)
    void* dev = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dev, host_base, 0);
    if (e != cudaSuccess || !dev) {
        cudaGetLastError();
        return fail(c, SPH_ERR_INVALID, "sph_download_owned_scatter: the host array is not page-locked / device-mapped (sph_host_register it)");
    }
    int rc = owned_export(c, field, dev, n, true);
    if (rc != SPH_OK) return rc;
    SPH_CUDA(c, cudaStreamSynchronize(c->st));
    return SPH_OK;
}

int sph_download_owned(SphContext* c, int field, uint32_t* global_id, void* host, size_t host_bytes, uint32_t* out_n)
{
    if (!c) return SPH_ERR_INVALID;
    SPH_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n = c->n;
    if (out_n) *out_n = n;
    if (!n) return SPH_OK;
    const size_t per = owned_field_bytes(field);
    if (!per) return fail(c, SPH_ERR_INVALID, "unknown field");
    if (host && host_bytes < per * n) return fail(c, SPH_ERR_INVALID, "sph_download_owned: host buffer too small");
    if (global_id) {
        launch_export_ids(c->st, c->A_pos, (uint32_t*)c->stage, n, &c->launches);
        SPH_CUDA(c, cudaMemcpyAsync(global_id, c->stage, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
        SPH_CUDA(c, cudaStreamSynchronize(c->st));
    }
    if (host) {
        int rc = owned_export(c, field, c->stage, n);
        if (rc != SPH_OK) return rc;
        SPH_CUDA(c, cudaMemcpyAsync(host, c->stage, per * n, cudaMemcpyDeviceToHost, c->st));
        SPH_CUDA(c, cudaStreamSynchronize(c->st));
    }
    return SPH_OK;
}

// ---- pipelined transfers, slab mode (see sph_api.cu: same streams, events and staging buffers) -------------------
int sph_upload_owned_begin(SphContext* c, uint32_t n, const uint32_t* global_id, const float* pos3, const float* vel3)
{
    if (!c) return SPH_ERR_INVALID;
    if (n > c->cap) return fail(c, SPH_ERR_CAPACITY, "sph_upload_owned_begin: n exceeds capacity");
    if (n && (!pos3 || !global_id)) return fail(c, SPH_ERR_INVALID, "sph_upload_owned_begin: NULL input");
    if (c->upload_pending) return fail(c, SPH_ERR_INVALID, "sph_upload_owned_begin: an upload is already pending (commit it first)");
    SPH_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_pipeline(c);
    if (rc != SPH_OK) return rc;
    if (!c->stage_in) SPH_CUDA(c, cudaMalloc((void**)&c->stage_in, (size_t)c->cap * 28));
    if (c->pack_recorded) SPH_CUDA(c, cudaStreamWaitEvent(c->st_in, c->ev_pack, 0));
    if (n) {
        SPH_CUDA(c, cudaMemcpyAsync(c->stage_in, pos3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st_in));
        if (vel3) SPH_CUDA(c, cudaMemcpyAsync(c->stage_in + (size_t)c->cap * 12, vel3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st_in));
        SPH_CUDA(c, cudaMemcpyAsync(c->stage_in + (size_t)c->cap * 24, global_id, (size_t)n * 4, cudaMemcpyHostToDevice, c->st_in));
    }
    SPH_CUDA(c, cudaEventRecord(c->ev_h2d, c->st_in));
    c->upload_pending = true;
    c->upload_has_vel = vel3 != nullptr;
    c->upload_has_ids = true;
    c->upload_n = n;
    return SPH_OK;
}

int sph_download_owned_begin(SphContext* c, int field, uint32_t* global_id, void* host, size_t host_bytes, uint32_t* out_n)
{
    if (!c) return SPH_ERR_INVALID;
    if (c->download_pending) return fail(c, SPH_ERR_INVALID, "sph_download_owned_begin: a download is already pending (wait for it first)");
    SPH_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n = c->n;
    if (out_n) *out_n = n;
    const size_t per = owned_field_bytes(field);
    if (!per) return fail(c, SPH_ERR_INVALID, "unknown field");
    if (n && !host) return fail(c, SPH_ERR_INVALID, "sph_download_owned_begin: host is NULL");
    if (host_bytes < per * n) return fail(c, SPH_ERR_INVALID, "sph_download_owned_begin: host buffer too small");
    int rc = ensure_pipeline(c);
    if (rc != SPH_OK) return rc;
    if (!c->stage_out) SPH_CUDA(c, cudaMalloc((void**)&c->stage_out, (size_t)c->cap * 20));
    if (n) {
        rc = owned_export(c, field, c->stage_out, n);
        if (rc != SPH_OK) return rc;
        uint32_t* dids = (uint32_t*)(c->stage_out + (size_t)c->cap * 16);
        if (global_id) launch_export_ids(c->st, c->A_pos, dids, n, &c->launches);
        SPH_CUDA(c, cudaEventRecord(c->ev_export, c->st));
        SPH_CUDA(c, cudaStreamWaitEvent(c->st_out, c->ev_export, 0));
        SPH_CUDA(c, cudaMemcpyAsync(host, c->stage_out, per * n, cudaMemcpyDeviceToHost, c->st_out));
        if (global_id) SPH_CUDA(c, cudaMemcpyAsync(global_id, dids, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st_out));
    }
    SPH_CUDA(c, cudaEventRecord(c->ev_d2h, c->st_out));
    c->download_pending = true;
    return SPH_OK;
}

int sph_comm_stats(const SphContext* c, uint32_t* out5)
{
    if (!c || !out5 || !c->slab) return SPH_ERR_INVALID;
    for (int i = 0; i < 5; i++) out5[i] = c->slab->stats[i];
    return SPH_OK;
}

}  // extern "C"
