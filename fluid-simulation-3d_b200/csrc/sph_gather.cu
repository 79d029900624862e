// sph_gather.cu -- the three neighbour-gather passes (density, pressure, viscosity) and their launchers.
//
//   density   S3  physicsWorld.cc:304-311, 325-365      pressure  S4  :367-422
//   viscosity S5  :424-464 (snapshot semantics)          kernels   engine/physics/kernels.h:25-82
//
// The physics of a (particle, neighbour) pair lives once, in sph_gather.cuh (load_self / fetch / eval / finish);
// the kernels here differ only in how they enumerate candidates.  What runs by default:
//
//   k_density_pk     GRID table: two-phase density over packed candidate pairs (see the kernel): phase A culls two
//                    candidates per pair record (a 128-bit + a 64-bit load) with FFMA2 math and pushes survivors (row, d^2) on a small
//                    shared-memory stack; phase B pops them converged, sums the kernels, writes the NEIGHBOUR LIST
//                    rows and the viscosity weight of every entry.
//   k_density_list   REFERENCE_HASH table (27 bucket walks, hash filter): the same two phases with scalar exact math.
//   k_gather_list    the pressure pass replays the list (same predicted positions + same predicate => same set):
//                    ~18 entries instead of ~100 candidates, one 256-bit record load per neighbour.
//                    A particle whose list overflowed walks the table instead (walk_particle).
//   k_viscosity_w    the viscosity pass over the recorded weights: per entry one 16-byte gather of v'_j.
//
// ncu (profiles/): every variant of these passes is bound by L1 wavefronts and instruction issue, not by HBM --
// the gather re-reads neighbours from L1/L2 by design.  Variants kept for A/B runs, all parity-tested
// (tests/test_variants_gpu.py), selected with SPH_GATHER / SPH_DENSITY:
//   k_gather_walk    generation 1: every pass walks the table (SPH_GATHER=v1; also the no-list configuration)
//   k_gather2        packed fp32x2 two-particles-per-thread two-phase cull in every pass (SPH_GATHER=v2)
//   k_density_pair / k_density_pair2   packed two-particle density feeding the list (SPH_DENSITY=pair / 2).
//                    FADD2/FMUL2/FFMA2 halve the cull instructions, but a warp then spans 64 particles, its loads
//                    touch ~1.7x more lines and the union windows waste slots: 190-205 us vs 140 us at 1 M particles.
#include <cstdlib>

#include "sph_gather.cuh"

namespace sphb200 {

namespace {

constexpr int GT = 128;    // threads per block (2 particles each)
constexpr int GK = 16;     // stack entries per particle (a small stack keeps shared memory from eating the L1)
constexpr int GCH = 4;     // candidates per chunk = loads in flight per thread between two overflow checks
// evaluate one candidate; the density pass also records it in the neighbour list
template <int PASS>
__device__ __forceinline__ void accept(const GatherArgs& A, const DevParams& P, const Self& s, const uint32_t j,
                                       const Fetched& f, Acc& acc)
{
    const uint32_t before = acc.cnt;
    const bool ok = eval<PASS>(P, s, j, f, acc);
    if (PASS == PASS_DENSITY && ok && A.list_idx && before < A.list_k)
        A.list_idx[(size_t)before * A.list_stride + s.i] = j;
}

// ---- generation 1: one thread per particle walks the table (both table modes) -------------------
// Also the fallback of the list kernels for a particle whose list overflowed.
template <int MODE, int PASS>
__device__ __forceinline__ void walk_particle(const GatherArgs& A, const DevParams& P, const Self& s, Acc& acc)
{
    for_each_candidate<MODE>(A.pred, A.table, A.tend, s.p, P, [&](const uint32_t j, const float4 q) {
        Fetched f;
        f.q = q;
        f.aux = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (PASS != PASS_DENSITY) {
            float ox, oy, oz;
            if (sqr_dist(q, s.p, ox, oy, oz) > P.sqr_r) return;      // do not fetch the aux row of a non-neighbour
            f = fetch<PASS>(A, j);
        }
        accept<PASS>(A, P, s, j, f, acc);
    });
}

#ifndef SPH_WALK_THREADS
#define SPH_WALK_THREADS 128          // block size of the one-thread-per-particle gather kernels (tools/sweep_variants.sh)
#endif
constexpr int kWalkThreads = SPH_WALK_THREADS;
// The neighbour list streams through once per pass (176 MB at 1 M particles, more than the L2): mark its traffic
// evict-first so that it does not push the particle records, which every pass re-reads, out of the L2.
#ifdef SPH_LIST_PLAIN
#define LIST_LD(p) __ldg(p)
#define LIST_ST(p, v) (*(p) = (v))
#else
#define LIST_LD(p) __ldcs(p)
#define LIST_ST(p, v) __stcs((p), (v))
#endif
#ifndef SPH_LIST_UNROLL
#define SPH_LIST_UNROLL 4
#endif
constexpr int kListUnroll = SPH_LIST_UNROLL;   // neighbour-list entries in flight per thread
#ifndef SPH_VISC_UNROLL
#define SPH_VISC_UNROLL 4
#endif

template <int MODE, int PASS>
__global__ void __launch_bounds__(kWalkThreads)
k_gather_walk(const GatherArgs A, const DevParams P, const float dt)
{
    const uint32_t i = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.row1) return;
    const Self s = load_self<PASS>(A, P, i);
    Acc acc = {0.0f, 0.0f, 0.0f, 0u};
    walk_particle<MODE, PASS>(A, P, s, acc);
    finish<PASS>(A, P, s, acc, dt);
    if (PASS == PASS_DENSITY && A.list_cnt) A.list_cnt[i] = acc.cnt;
}

// The longest list that did not fit, for the host's auto-grow: one atomic per WARP at most, and only while the
// value still rises (a plain read first), so a scene where every particle overflows costs nothing.
// Must be called by all 32 lanes.
__device__ __forceinline__ void report_overflow(const GatherArgs& A, const uint32_t n_over)
{
    const uint32_t m = __reduce_max_sync(0xffffffffu, n_over);
    if (m && A.list_overflow && (threadIdx.x & 31) == 0 && m > *(volatile uint32_t*)A.list_overflow) atomicMax(A.list_overflow, m);
}

// ---- density pass, two-phase, one thread per particle (default) ------------------------------------
// Phase A walks the table and applies ONLY the reference's exact predicate; survivors are pushed on a
// small per-thread shared-memory stack with a predicated store (no divergent branch, 4 candidate loads
// in flight).  Phase B (flush) pops the stack, evaluates the smoothing kernels -- the expensive code runs
// for the ~16 % of candidates that are neighbours, converged -- and writes the neighbour-list column the
// pressure / viscosity passes replay.  Rows and segments are warp-uniform loop levels (redux.sync), so the
// flush is collective and happens at most once per SEG candidates whatever the density.
__device__ __forceinline__ void sts_if(uint32_t* p, const uint32_t v, const bool c)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p st.shared.b32 [%0], %1; }" ::"r"(a), "r"(v), "r"((int)c) : "memory");
}

constexpr int KS = 48;     // stack entries per thread
constexpr int SEG = 16;    // candidates between two flush checks

template <int MODE>
__device__ __forceinline__ void row_bounds(const GatherArgs& A, const DevParams& P, const int3 c, const int3 g,
                                           const int x0, const int x1, const uint32_t rows, const int r,
                                           uint32_t& b, uint32_t& e, float& hf)
{
    b = e = 0;
    hf = 0.0f;
    if (MODE == SPH_TABLE_GRID) {                          // 9 rows (dy, dz) x one contiguous x window
        const int z = g.z + r / 3 - 1, y = g.y + r % 3 - 1;
        if (!((rows >> r) & 1u) || z < 0 || z >= P.gdim[2] || y < 0 || y >= P.gdim[1]) return;
        const uint32_t row = ((uint32_t)z * (uint32_t)P.gdim[1] + (uint32_t)y) * (uint32_t)P.gdim[0];
        b = tbl(A.table, P, row + x0);
        e = tbl(A.table, P, row + x1 + 1);
    } else {                                               // 27 buckets, offsets[27] order (physicsWorld.h:131-143)
        const uint32_t h = hash_cell(c.x + r / 9 - 1, c.y + (r / 3) % 3 - 1, c.z + r % 3 - 1);
        const uint32_t key = key_of_hash(h, P);
        const uint32_t s0 = __ldg(&A.table[key]);
        if (s0 >= P.n) return;                             // 0x7FFFFFFF: empty bucket (:339)
        b = s0;
        e = __ldg(&A.tend[key]);
        hf = __uint2float_rn(h);                           // `index.y != hash` is compared in float (:346)
    }
}

template <int MODE>
__global__ void __launch_bounds__(kWalkThreads)
k_density_list(const GatherArgs A, const DevParams P)
{
    __shared__ uint32_t stk[KS][kWalkThreads];
    const int tid = threadIdx.x;
    const uint32_t iraw = P.row0 + blockIdx.x * blockDim.x + tid;
    const bool valid = iraw < P.row1;
    const uint32_t i = valid ? iraw : P.row1 - 1;          // idle tail threads shadow the last row (they join the collectives)
    const Self s = load_self<PASS_DENSITY>(A, P, i);
    const uint32_t K = A.list_idx ? A.list_k : 0u;
    const size_t stride = A.list_stride;
    uint32_t* col = A.list_idx + i;
    uint32_t n = 0, ns = 0;
    Acc acc = {0.0f, 0.0f, 0.0f, 0u};

    auto flush = [&]() {
        for (uint32_t k0 = 0; k0 < ns; k0 += 4) {
            uint32_t j[4];
            Fetched f[4];
            #pragma unroll
            for (int u = 0; u < 4; u++) j[u] = (k0 + u < ns) ? stk[k0 + u][tid] : i;
            #pragma unroll
            for (int u = 0; u < 4; u++) f[u] = fetch<PASS_DENSITY>(A, j[u]);
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k0 + u < ns) {
                    (void)eval<PASS_DENSITY>(P, s, j[u], f[u], acc);
                    if (valid && n + k0 + u < K) col[(size_t)(n + k0 + u) * stride] = j[u];
                }
            }
        }
        n += ns;
        ns = 0;
    };

    const int3 c = cell_of(s.p.x, s.p.y, s.p.z, P.r);
    Win W = {};
    if (MODE == SPH_TABLE_GRID) W = window_of(s.p.x, s.p.y, s.p.z, P);
    const int3 g = W.g;
    const int x0 = W.x0, x1 = W.x1;
    constexpr int ROWS = (MODE == SPH_TABLE_GRID) ? 9 : 27;
    #pragma unroll 1
    for (int r = 0; r < ROWS; r++) {
        uint32_t b, e;
        float hf;
        row_bounds<MODE>(A, P, c, g, x0, x1, W.rows, r, b, e, hf);
        const uint32_t segs = (__reduce_max_sync(0xffffffffu, e - b) + SEG - 1) / SEG;
        #pragma unroll 1
        for (uint32_t sg = 0; sg < segs; sg++) {
            if (__any_sync(0xffffffffu, ns > KS - SEG)) flush();
            const uint32_t j0 = b + sg * SEG;
            const uint32_t je = min(j0 + SEG, e);
            #pragma unroll 1
            for (uint32_t jb = j0; jb < je; jb += 4) {
                const float4* qp = A.pred + jb;            // the array is padded: reading up to 3 rows past `e` is safe
                float4 q[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) q[u] = __ldg(qp + u);
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    float ox, oy, oz;
                    // branch-free: bitwise &, predicated store
                    bool ok = (jb + u < je) & !(sqr_dist(q[u], s.p, ox, oy, oz) > P.sqr_r);    // :357, exact (Q8)
                    if (MODE == SPH_TABLE_REFERENCE_HASH) ok = ok & (q[u].w == hf);
                    sts_if(&stk[ns][tid], jb + u, ok);
                    ns += ok;
                }
            }
        }
    }
    flush();
    if (valid) {
        finish<PASS_DENSITY>(A, P, s, acc, 0.0f);
        if (A.list_cnt) A.list_cnt[i] = n;
    }
    report_overflow(A, (valid && n > K) ? n : 0u);
}

// ---- neighbour-list passes: pressure and viscosity replay the exact neighbour set recorded by the
// density pass (same predicted positions, same predicate => same set), 4 entries in flight per thread.
template <int MODE, int PASS>
__global__ void __launch_bounds__(kWalkThreads)
k_gather_list(const GatherArgs A, const DevParams P, const float dt)
{
    chain_prologue();
    const uint32_t i = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.row1) return;
    const Self s = load_self<PASS>(A, P, i);
    Acc acc = {0.0f, 0.0f, 0.0f, 0u};
    const uint32_t cnt = A.list_cnt[i];
    if (cnt > A.list_k) {                                  // overflowed list: walk the table like generation 1
        walk_particle<MODE, PASS>(A, P, s, acc);
    } else {
        const uint32_t* __restrict__ col = A.list_idx + i;
        constexpr int U = kListUnroll;
        // software pipeline: the list entries of chunk c+1 are requested before chunk c's rows are fetched,
        // so the (cold, HBM-resident) list read overlaps the (L1/L2-resident) particle rows of the chunk before
        uint32_t jn[U];
        #pragma unroll
        for (int u = 0; u < U; u++) jn[u] = (u < cnt) ? LIST_LD(&col[(size_t)u * A.list_stride]) : s.i;
        for (uint32_t k0 = 0; k0 < cnt; k0 += U) {
            uint32_t j[U];
            Fetched f[U];
            #pragma unroll
            for (int u = 0; u < U; u++) j[u] = jn[u];
            #pragma unroll
            for (int u = 0; u < U; u++) jn[u] = (k0 + U + u < cnt) ? LIST_LD(&col[(size_t)(k0 + U + u) * A.list_stride]) : s.i;
            #pragma unroll
            for (int u = 0; u < U; u++) {            // the particle itself (its own entry, and the padding) is not fetched
                f[u].q = s.p;
                if (j[u] != s.i) f[u] = fetch<PASS>(A, j[u]);
            }
            #pragma unroll
            for (int u = 0; u < U; u++) (void)eval<PASS>(P, s, j[u], f[u], acc);   // ... and skipped
        }
    }
    finish<PASS>(A, P, s, acc, dt);
}

// ---- viscosity pass over the list WITH weights (default on the GRID table) --------------------------------------
// k_density_pk leaves, next to every list entry j, the viscosity kernel value of the pair (exact neighbour set:
// band candidates were re-tested; padding entries and entries at d >= r carry weight 0).  The pass then needs no
// positions at all: per entry one coalesced index, one coalesced weight and ONE 16-byte gather of v'_j.  The
// particle's own entry contributes (v'_i - v'_i) * w = 0, exactly like the reference's `continue` (:450).
// (12 blocks per SM = 40 registers, no spills: 61.5-63.5 us against 64.6 at the 42 registers ptxas picks unprompted; 16 blocks
// spill and lose (71.7 us); the pressure pass gains nothing from 48 registers / 10 blocks)
template <int MODE>
__global__ void __launch_bounds__(kWalkThreads, 12)
k_viscosity_w(const GatherArgs A, const DevParams P, const float dt)
{
    chain_prologue();
    const uint32_t i = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.row1) return;
    const uint32_t cnt = A.list_cnt[i];
    if (cnt > A.list_k) {                                  // overflowed list: walk the table like generation 1
        const Self s = load_self<PASS_VISCOSITY>(A, P, i);
        Acc acc = {0.0f, 0.0f, 0.0f, 0u};
        walk_particle<MODE, PASS_VISCOSITY>(A, P, s, acc);
        finish<PASS_VISCOSITY>(A, P, s, acc, dt);
        return;
    }
    const float4 vi = A.velp[i];
    const uint32_t* __restrict__ col = A.list_idx + i;
    const float* __restrict__ colw = A.list_w + i;
    const size_t stride = A.list_stride;
    constexpr int U = SPH_VISC_UNROLL;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    uint32_t jn[U];
    float wn[U];
    #pragma unroll
    for (int u = 0; u < U; u++) {
        jn[u] = (u < cnt) ? LIST_LD(&col[(size_t)u * stride]) : i;
        wn[u] = (u < cnt) ? LIST_LD(&colw[(size_t)u * stride]) : 0.0f;
    }
    for (uint32_t k0 = 0; k0 < cnt; k0 += U) {
        uint32_t j[U];
        float w[U];
        float4 v[U];
        #pragma unroll
        for (int u = 0; u < U; u++) { j[u] = jn[u]; w[u] = wn[u]; }
        #pragma unroll
        for (int u = 0; u < U; u++) {
            jn[u] = (k0 + U + u < cnt) ? LIST_LD(&col[(size_t)(k0 + U + u) * stride]) : i;
            wn[u] = (k0 + U + u < cnt) ? LIST_LD(&colw[(size_t)(k0 + U + u) * stride]) : 0.0f;
        }
        #pragma unroll
        for (int u = 0; u < U; u++) { v[u] = vi; if (j[u] != i) v[u] = __ldg(&A.velp[j[u]]); }
        #pragma unroll
        for (int u = 0; u < U; u++) {
            ax = fmaf(v[u].x - vi.x, w[u], ax);
            ay = fmaf(v[u].y - vi.y, w[u], ay);
            az = fmaf(v[u].z - vi.z, w[u], az);
        }
    }
    const float k = P.mu * dt;                             // :463
    A.velv_out[i] = make_float4(fmaf(ax, k, vi.x), fmaf(ay, k, vi.y), fmaf(az, k, vi.z), 0.0f);
}

// 9 row windows of a cell: rows (dy, dz), x window [xa, xb]
__device__ __forceinline__ void row_range(const uint32_t* __restrict__ table, const DevParams& P, const int3 g,
                                          const int xa, const int xb, const uint32_t rows, const int r9,
                                          uint32_t& b, uint32_t& e)
{
    const int dz = r9 / 3 - 1, dy = r9 % 3 - 1;
    const int z = g.z + dz, y = g.y + dy;
    b = e = 0;
    if (r9 >= 9 || !((rows >> r9) & 1u) || z < 0 || z >= P.gdim[2] || y < 0 || y >= P.gdim[1]) return;
    const uint32_t row = ((uint32_t)z * (uint32_t)P.gdim[1] + (uint32_t)y) * (uint32_t)P.gdim[0];
    b = tbl(table, P, row + xa);
    e = tbl(table, P, row + xb + 1);
}

template <int PASS>
__global__ void __launch_bounds__(GT)
k_gather2(const GatherArgs A, const DevParams P, const float dt)
{
    __shared__ uint32_t stk[2][GK][GT];
    const int tid = threadIdx.x;
    const uint32_t last = P.row1 - 1;
    const uint32_t i0r = P.row0 + 2u * (blockIdx.x * GT + tid), i1r = i0r + 1u;
    const bool valid0 = i0r <= last, valid1 = i1r <= last;
    const Self s0 = load_self<PASS>(A, P, valid0 ? i0r : last);
    Self s1 = load_self<PASS>(A, P, valid1 ? i1r : last);
    const Win W0 = window_of(s0.p.x, s0.p.y, s0.p.z, P), W1 = window_of(s1.p.x, s1.p.y, s1.p.z, P);
    const int3 g0 = W0.g, g1 = W1.g;
    const bool straddle = valid1 && (g0.y != g1.y || g0.z != g1.z);
    const bool pair = valid1 && !straddle;
    const uint32_t rows01 = pair ? (W0.rows | W1.rows) : W0.rows;
    // a particle that is not processed in the packed loop is parked far away: it never passes the cull
    const float far = 1.0e18f;
    const uint64_t px = pk(s0.p.x, pair ? s1.p.x : far), py = pk(s0.p.y, pair ? s1.p.y : far), pz = pk(s0.p.z, pair ? s1.p.z : far);
    const int xa = pair ? min(W0.x0, W1.x0) : W0.x0;
    const int xb = pair ? max(W0.x1, W1.x1) : W0.x1;
    Acc a0 = {0.0f, 0.0f, 0.0f, 0u}, a1 = {0.0f, 0.0f, 0.0f, 0u};
    uint32_t n0 = 0, n1 = 0;
    const uint32_t nlast = P.n - 1;
    const float cull_hi = P.cull_hi;

    // phase B: pop both stacks in lockstep; the fetches of entry k+1 are issued before entry k is evaluated
    auto flush = [&]() {
        const uint32_t m = __reduce_max_sync(0xffffffffu, max(n0, n1));
        uint32_t j0 = n0 ? stk[0][0][tid] : s0.i, j1 = n1 ? stk[1][0][tid] : s1.i;
        Fetched f0 = fetch<PASS>(A, j0), f1 = fetch<PASS>(A, j1);
        for (uint32_t k = 0; k < m; k++) {
            const uint32_t nj0 = (k + 1 < n0) ? stk[0][k + 1][tid] : s0.i, nj1 = (k + 1 < n1) ? stk[1][k + 1][tid] : s1.i;
            const Fetched nf0 = fetch<PASS>(A, nj0), nf1 = fetch<PASS>(A, nj1);
            if (k < n0) accept<PASS>(A, P, s0, j0, f0, a0);
            if (k < n1) accept<PASS>(A, P, s1, j1, f1, a1);
            j0 = nj0; j1 = nj1; f0 = nf0; f1 = nf1;
        }
        n0 = n1 = 0;
    };

    // phase A: every thread walks its own 9 row windows as ONE flat candidate sequence, GCH candidates per
    // iteration; the warp iterates until its slowest thread is done (trip = max of the TOTAL candidate
    // counts, not the sum of per-row maxima).  The next row's range is fetched one row ahead.
    int r = 0;
    uint32_t j, e, bn, en;
    row_range(A.table, P, g0, xa, xb, rows01, 0, j, e);
    row_range(A.table, P, g0, xa, xb, rows01, 1, bn, en);
    auto advance = [&]() {
        do {
            r++; j = bn; e = en;
            row_range(A.table, P, g0, xa, xb, rows01, r + 1, bn, en);
        } while (r < 9 && j >= e);
    };
    if (j >= e) advance();
    while (__any_sync(0xffffffffu, r < 9)) {
        const bool act = r < 9;
        const uint32_t jb = act ? j : 0u, ee = act ? e : 0u;
        float4 q[GCH];
        #pragma unroll
        for (int u = 0; u < GCH; u++) q[u] = __ldg(&A.pred[min(jb + u, nlast)]);
        #pragma unroll
        for (int u = 0; u < GCH; u++) {
            const uint32_t jj = jb + u;
            const uint64_t ox = sub2(pk(q[u].x, q[u].x), px), oy = sub2(pk(q[u].y, q[u].y), py), oz = sub2(pk(q[u].z, q[u].z), pz);
            const uint64_t d2 = fma2(oz, oz, fma2(oy, oy, mul2(ox, ox)));
            float d0, d1;
            upk(d2, d0, d1);
            const bool in = jj < ee;
            if (in && !(d0 > cull_hi)) { stk[0][n0][tid] = jj; n0++; }
            if (in && !(d1 > cull_hi)) { stk[1][n1][tid] = jj; n1++; }
        }
        if (act) { j += GCH; if (j >= e) advance(); }
        if (__any_sync(0xffffffffu, max(n0, n1) > GK - GCH)) flush();
    }
    flush();
    if (valid0) finish<PASS>(A, P, s0, a0, dt);
    if (pair) finish<PASS>(A, P, s1, a1, dt);

    // second particles whose row differs from the first's: one at a time, whole warp on the candidates
    uint32_t todo = __ballot_sync(0xffffffffu, straddle);
    const int lane = tid & 31;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t ic = __shfl_sync(0xffffffffu, s1.i, src);
        const Self sc = load_self<PASS>(A, P, ic);
        const Win Wc = window_of(sc.p.x, sc.p.y, sc.p.z, P);
        const int3 gc = Wc.g;
        const int ca = Wc.x0, cb = Wc.x1;
        Acc ac = {0.0f, 0.0f, 0.0f, 0u};
        #pragma unroll 1
        for (int r9 = 0; r9 < 9; r9++) {
            uint32_t b, e;
            row_range(A.table, P, gc, ca, cb, Wc.rows, r9, b, e);
            for (uint32_t j = b + lane; j < e; j += 32) term<PASS>(A, P, sc, j, ac);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ac.a += __shfl_xor_sync(0xffffffffu, ac.a, o);
            ac.b += __shfl_xor_sync(0xffffffffu, ac.b, o);
            ac.c += __shfl_xor_sync(0xffffffffu, ac.c, o);
            ac.cnt += __shfl_xor_sync(0xffffffffu, ac.cnt, o);
        }
        if (lane == src) finish<PASS>(A, P, sc, ac, dt);
    }
}

// ---- density pass, generation 3 (GRID table, neighbour list on) ------------------------------------
// Phase A: two particles per thread cull every candidate of their shared 9 row windows with packed
// fp32x2 math (FMA-fused d^2 against the conservatively widened cull_hi) and append the survivors to
// the particles' NEIGHBOUR LIST COLUMNS in global memory -- pure predication, no divergent branch.
// Phase B: each particle replays its own column, applies the reference's exact predicate and sums the
// smoothing kernels.  The list therefore doubles as the compaction buffer of this pass and as the
// input of the pressure and viscosity passes (which re-apply the exact predicate themselves, so the
// <= 1e-6 fraction of borderline extras in the list is harmless).
__global__ void __launch_bounds__(GT)
k_density_pair(const GatherArgs A, const DevParams P)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t last = P.row1 - 1;
    const uint32_t i0r = P.row0 + 2u * (blockIdx.x * GT + tid), i1r = i0r + 1u;
    const bool valid0 = i0r <= last, valid1 = i1r <= last;
    const Self s0 = load_self<PASS_DENSITY>(A, P, valid0 ? i0r : last);
    const Self s1 = load_self<PASS_DENSITY>(A, P, valid1 ? i1r : last);
    const Win W0 = window_of(s0.p.x, s0.p.y, s0.p.z, P), W1 = window_of(s1.p.x, s1.p.y, s1.p.z, P);
    const int3 g0 = W0.g, g1 = W1.g;
    const bool straddle = valid1 && (g0.y != g1.y || g0.z != g1.z);
    const bool pair = valid1 && !straddle;
    const uint32_t rows01 = pair ? (W0.rows | W1.rows) : W0.rows;
    const float far = 1.0e18f;       // a particle that is not processed here never passes the cull
    const uint64_t px = pk(s0.p.x, pair ? s1.p.x : far), py = pk(s0.p.y, pair ? s1.p.y : far), pz = pk(s0.p.z, pair ? s1.p.z : far);
    const int xa = pair ? min(W0.x0, W1.x0) : W0.x0;
    const int xb = pair ? max(W0.x1, W1.x1) : W0.x1;
    const uint32_t K = A.list_k;
    const size_t stride = A.list_stride;
    uint32_t* col0 = A.list_idx + s0.i;
    uint32_t* col1 = A.list_idx + s1.i;
    uint32_t n0 = 0, n1 = 0;
    const uint32_t nlast = P.n - 1;
    const float cull_hi = P.cull_hi;

    #pragma unroll 1
    for (int r9 = 0; r9 < 9; r9++) {
        uint32_t b, e;
        row_range(A.table, P, g0, xa, xb, rows01, r9, b, e);
        const uint32_t chunks = (__reduce_max_sync(0xffffffffu, e - b) + GCH - 1) / GCH;
        #pragma unroll 1
        for (uint32_t c = 0; c < chunks; c++) {
            const uint32_t jb = b + c * GCH;
            float4 q[GCH];
            #pragma unroll
            for (int u = 0; u < GCH; u++) q[u] = __ldg(&A.pred[min(jb + u, nlast)]);
            #pragma unroll
            for (int u = 0; u < GCH; u++) {
                const uint32_t jj = jb + u;
                const uint64_t ox = sub2(pk(q[u].x, q[u].x), px), oy = sub2(pk(q[u].y, q[u].y), py), oz = sub2(pk(q[u].z, q[u].z), pz);
                const uint64_t d2 = fma2(oz, oz, fma2(oy, oy, mul2(ox, ox)));
                float d0, d1;
                upk(d2, d0, d1);
                const bool in = jj < e;
                if (in && !(d0 > cull_hi)) { if (n0 < K) col0[(size_t)n0 * stride] = jj; n0++; }
                if (in && !(d1 > cull_hi)) { if (n1 < K) col1[(size_t)n1 * stride] = jj; n1++; }
            }
        }
    }

    // phase B: replay both columns in lockstep, 4 entries in flight per particle
    Acc a0 = {0.0f, 0.0f, 0.0f, 0u}, a1 = {0.0f, 0.0f, 0.0f, 0u};
    const uint32_t m0 = min(n0, K), m1 = pair ? min(n1, K) : 0u;
    const uint32_t m = __reduce_max_sync(0xffffffffu, max(m0, m1));
    for (uint32_t k0 = 0; k0 < m; k0 += 4) {
        uint32_t j0[4], j1[4];
        Fetched f0[4], f1[4];
        #pragma unroll
        for (int u = 0; u < 4; u++) {
            j0[u] = (k0 + u < m0) ? col0[(size_t)(k0 + u) * stride] : s0.i;   // plain loads: written by this thread above
            j1[u] = (k0 + u < m1) ? col1[(size_t)(k0 + u) * stride] : s1.i;
        }
        #pragma unroll
        for (int u = 0; u < 4; u++) { f0[u] = fetch<PASS_DENSITY>(A, j0[u]); f1[u] = fetch<PASS_DENSITY>(A, j1[u]); }
        #pragma unroll
        for (int u = 0; u < 4; u++) {
            if (k0 + u < m0) (void)eval<PASS_DENSITY>(P, s0, j0[u], f0[u], a0);
            if (k0 + u < m1) (void)eval<PASS_DENSITY>(P, s1, j1[u], f1[u], a1);
        }
    }
    GatherArgs W = A;            // the fallback walk must not record a second time
    W.list_idx = nullptr;
    if (n0 > K) { a0 = {0.0f, 0.0f, 0.0f, 0u}; walk_particle<SPH_TABLE_GRID, PASS_DENSITY>(W, P, s0, a0); }
    if (pair && n1 > K) { a1 = {0.0f, 0.0f, 0.0f, 0u}; walk_particle<SPH_TABLE_GRID, PASS_DENSITY>(W, P, s1, a1); }
    if (valid0) { finish<PASS_DENSITY>(A, P, s0, a0, 0.0f); A.list_cnt[s0.i] = n0; }
    if (pair) { finish<PASS_DENSITY>(A, P, s1, a1, 0.0f); A.list_cnt[s1.i] = n1; }
    report_overflow(A, max((valid0 && n0 > K) ? n0 : 0u, (pair && n1 > K) ? n1 : 0u));

    // second particles that live in another (y,z) row than their partner: whole warp on one particle
    uint32_t todo = __ballot_sync(0xffffffffu, straddle);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t ic = __shfl_sync(0xffffffffu, s1.i, src);
        const Self sc = load_self<PASS_DENSITY>(A, P, ic);
        const Win Wc = window_of(sc.p.x, sc.p.y, sc.p.z, P);
        const int3 gc = Wc.g;
        const int ca = Wc.x0, cb = Wc.x1;
        uint32_t* colc = A.list_idx + ic;
        Acc ac = {0.0f, 0.0f, 0.0f, 0u};
        uint32_t nc = 0;
        #pragma unroll 1
        for (int r9 = 0; r9 < 9; r9++) {
            uint32_t b, e;
            row_range(A.table, P, gc, ca, cb, Wc.rows, r9, b, e);
            for (uint32_t jb = b; jb < e; jb += 32) {
                const uint32_t j = jb + lane;
                bool ok = false;
                if (j < e) { const Fetched f = fetch<PASS_DENSITY>(A, j); ok = eval<PASS_DENSITY>(P, sc, j, f, ac); }
                const uint32_t mask = __ballot_sync(0xffffffffu, ok);
                const uint32_t pos = nc + __popc(mask & ((1u << lane) - 1u));
                if (ok && pos < K) colc[(size_t)pos * stride] = j;
                nc += __popc(mask);
            }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ac.a += __shfl_xor_sync(0xffffffffu, ac.a, o);
            ac.b += __shfl_xor_sync(0xffffffffu, ac.b, o);
            ac.cnt += __shfl_xor_sync(0xffffffffu, ac.cnt, o);
        }
        if (lane == src) { finish<PASS_DENSITY>(A, P, sc, ac, 0.0f); A.list_cnt[ic] = nc; }
    }
}

// ---- density pass, two particles per thread + packed fp32x2 cull + shared-memory stacks (SPH_DENSITY=pair2) ----
// Same two-phase structure as k_density_list, but every candidate row is loaded ONCE for two consecutive
// particles and culled with 6 packed instructions (FADD2 x3, FMUL2, FFMA2 x2) against the conservatively widened
// cull_hi; phase B applies the reference's exact predicate.  Halves the candidate loads per test (the single-
// particle kernel is bound by L1 wavefronts) and the cull instructions per test.
constexpr int KS2 = 24;    // stack entries per particle
constexpr int SEG2 = 8;    // candidates between two flush checks

__global__ void __launch_bounds__(GT)
k_density_pair2(const GatherArgs A, const DevParams P)
{
    __shared__ uint32_t stk[2][KS2][GT];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t last = P.row1 - 1;
    const uint32_t i0r = P.row0 + 2u * (blockIdx.x * GT + tid), i1r = i0r + 1u;
    const bool valid0 = i0r <= last, valid1 = i1r <= last;
    const Self s0 = load_self<PASS_DENSITY>(A, P, valid0 ? i0r : last);
    const Self s1 = load_self<PASS_DENSITY>(A, P, valid1 ? i1r : last);
    const Win W0 = window_of(s0.p.x, s0.p.y, s0.p.z, P), W1 = window_of(s1.p.x, s1.p.y, s1.p.z, P);
    const bool straddle = valid1 && (W0.g.y != W1.g.y || W0.g.z != W1.g.z);
    const bool pair = valid1 && !straddle;
    const uint32_t rows01 = pair ? (W0.rows | W1.rows) : W0.rows;
    const float far = 1.0e18f;       // a particle that is not processed here never passes the cull
    const uint64_t px = pk(s0.p.x, pair ? s1.p.x : far), py = pk(s0.p.y, pair ? s1.p.y : far), pz = pk(s0.p.z, pair ? s1.p.z : far);
    const int xa = pair ? min(W0.x0, W1.x0) : W0.x0;
    const int xb = pair ? max(W0.x1, W1.x1) : W0.x1;
    const uint32_t K = A.list_k;
    const size_t stride = A.list_stride;
    uint32_t* col0 = A.list_idx + s0.i;
    uint32_t* col1 = A.list_idx + s1.i;
    uint32_t n0 = 0, n1 = 0, ns0 = 0, ns1 = 0;
    Acc a0 = {0.0f, 0.0f, 0.0f, 0u}, a1 = {0.0f, 0.0f, 0.0f, 0u};
    const float cull_hi = P.cull_hi;

    auto flush = [&]() {
        const uint32_t m = max(ns0, ns1);
        for (uint32_t k0 = 0; k0 < m; k0 += 2) {
            uint32_t j0[2], j1[2];
            Fetched f0[2], f1[2];
            #pragma unroll
            for (int u = 0; u < 2; u++) {
                j0[u] = (k0 + u < ns0) ? stk[0][k0 + u][tid] : s0.i;
                j1[u] = (k0 + u < ns1) ? stk[1][k0 + u][tid] : s1.i;
            }
            #pragma unroll
            for (int u = 0; u < 2; u++) { f0[u] = fetch<PASS_DENSITY>(A, j0[u]); f1[u] = fetch<PASS_DENSITY>(A, j1[u]); }
            #pragma unroll
            for (int u = 0; u < 2; u++) {
                // the list records the survivors of the conservative cull; every later pass re-applies the exact predicate
                if (k0 + u < ns0) { (void)eval<PASS_DENSITY>(P, s0, j0[u], f0[u], a0); if (valid0 && n0 + k0 + u < K) col0[(size_t)(n0 + k0 + u) * stride] = j0[u]; }
                if (k0 + u < ns1) { (void)eval<PASS_DENSITY>(P, s1, j1[u], f1[u], a1); if (pair && n1 + k0 + u < K) col1[(size_t)(n1 + k0 + u) * stride] = j1[u]; }
            }
        }
        n0 += ns0; n1 += ns1;
        ns0 = ns1 = 0;
    };

    #pragma unroll 1
    for (int r9 = 0; r9 < 9; r9++) {
        uint32_t b, e;
        row_range(A.table, P, W0.g, xa, xb, rows01, r9, b, e);
        const uint32_t segs = (__reduce_max_sync(0xffffffffu, e - b) + SEG2 - 1) / SEG2;
        #pragma unroll 1
        for (uint32_t sg = 0; sg < segs; sg++) {
            if (__any_sync(0xffffffffu, max(ns0, ns1) > KS2 - SEG2)) flush();
            const uint32_t j0 = b + sg * SEG2;
            const uint32_t je = min(j0 + SEG2, e);
            #pragma unroll 1
            for (uint32_t jb = j0; jb < je; jb += 4) {
                const float4* qp = A.pred + jb;            // padded array: up to 3 rows past `e` are readable
                float4 q[4];
                #pragma unroll
                for (int u = 0; u < 4; u++) q[u] = __ldg(qp + u);
                #pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint64_t ox = sub2(pk(q[u].x, q[u].x), px), oy = sub2(pk(q[u].y, q[u].y), py), oz = sub2(pk(q[u].z, q[u].z), pz);
                    const uint64_t d2 = fma2(oz, oz, fma2(oy, oy, mul2(ox, ox)));
                    float d0, d1;
                    upk(d2, d0, d1);
                    const bool in = jb + u < je;
                    const bool ok0 = in & !(d0 > cull_hi), ok1 = in & !(d1 > cull_hi);
                    sts_if(&stk[0][ns0][tid], jb + u, ok0);
                    sts_if(&stk[1][ns1][tid], jb + u, ok1);
                    ns0 += ok0; ns1 += ok1;
                }
            }
        }
    }
    flush();
    if (valid0) { finish<PASS_DENSITY>(A, P, s0, a0, 0.0f); A.list_cnt[s0.i] = n0; }
    if (pair) { finish<PASS_DENSITY>(A, P, s1, a1, 0.0f); A.list_cnt[s1.i] = n1; }
    report_overflow(A, max((valid0 && n0 > K) ? n0 : 0u, (pair && n1 > K) ? n1 : 0u));

    // second particles that live in another (y,z) row than their partner: whole warp on one particle
    uint32_t todo = __ballot_sync(0xffffffffu, straddle);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t ic = __shfl_sync(0xffffffffu, s1.i, src);
        const Self sc = load_self<PASS_DENSITY>(A, P, ic);
        const Win Wc = window_of(sc.p.x, sc.p.y, sc.p.z, P);
        uint32_t* colc = A.list_idx + ic;
        Acc ac = {0.0f, 0.0f, 0.0f, 0u};
        uint32_t nc = 0;
        #pragma unroll 1
        for (int r9 = 0; r9 < 9; r9++) {
            uint32_t b, e;
            row_range(A.table, P, Wc.g, Wc.x0, Wc.x1, Wc.rows, r9, b, e);
            for (uint32_t jb = b; jb < e; jb += 32) {
                const uint32_t j = jb + lane;
                bool ok = false;
                if (j < e) { const Fetched f = fetch<PASS_DENSITY>(A, j); ok = eval<PASS_DENSITY>(P, sc, j, f, ac); }
                const uint32_t mask = __ballot_sync(0xffffffffu, ok);
                const uint32_t pos = nc + __popc(mask & ((1u << lane) - 1u));
                if (ok && pos < K) colc[(size_t)pos * stride] = j;
                nc += __popc(mask);
            }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ac.a += __shfl_xor_sync(0xffffffffu, ac.a, o);
            ac.b += __shfl_xor_sync(0xffffffffu, ac.b, o);
            ac.cnt += __shfl_xor_sync(0xffffffffu, ac.cnt, o);
        }
        if (lane == src) { finish<PASS_DENSITY>(A, P, sc, ac, 0.0f); A.list_cnt[ic] = nc; }
        report_overflow(A, (lane == src && nc > K) ? nc : 0u);
    }
}

// ---- density pass, packed candidate pairs (default for the GRID table) ---------------------------------------
// One thread per particle, as k_density_list, but
//  * the candidates come from the PAIR-INTERLEAVED copy of the predicted positions (PredPair, sph_internal.h): one
//    128-bit + one 64-bit load bring two candidates as three aligned register pairs, and their d^2 costs 6 packed instructions
//    (3 FADD2, FMUL2, 2 FFMA2) instead of 16 scalar ones.  The FMA-fused d^2 decides everything outside a 2e-6 wide
//    band around sqrRadius; inside the band the reference's exact predicate is evaluated on the plain rows
//    (phase B, ~1e-6 of the candidates);
//  * rows and 2-pair chunks are warp-uniform loop levels; the bounds of the next row window are requested while
//    the current one is culled;
//  * the stack keeps (row, d^2): phase B needs no second look at the candidate -- no gather loads at all -- and it
//    also leaves the viscosity kernel value of every entry next to the index (k_viscosity_w);
//  * phase B writes list row k for the whole warp at once (lanes that run short pad with their own index, which
//    the pressure / viscosity passes skip), so list writes stay coalesced and list lengths are warp-uniform.
// ncu (round 1, one 32-byte record per pair): bound by the L1 data pipe (l1tex__data_pipe_lsu_wavefronts 85 % at 1 M, 94 % at
// 8 M particles): a 256-bit load is served in 8 passes of 4 lanes, >= 1 wavefront each, i.e. >= 4 wavefronts per
// warp-candidate (measured 5.25).  Round 2 (24 bytes per pair): issue 76-79 %, L1 data pipe 71-78 %.
#ifndef SPH_PKS
#define SPH_PKS 24
#endif
// The survivor pushes are inline PTX; SPH_PK_NOCLOB=1 (default) gives them no "memory" clobber and fences the stack
// only where C++ code reads it (flush), so nothing forces the compiler to re-read state around every push.
#ifndef SPH_PK_NOCLOB
#define SPH_PK_NOCLOB 1
#endif
#if SPH_PK_NOCLOB
#define SPH_PK_CLOBBER
#define SPH_PK_FENCE() asm volatile("" ::: "memory")
#else
#define SPH_PK_CLOBBER : "memory"
#define SPH_PK_FENCE() do {} while (0)
#endif
constexpr int PKS = SPH_PKS;    // stack entries per thread: sparse scenes (one or two flushes per particle)
constexpr int PKS_DENSE = 72;   // ... when the lists are long throughout (measured mean above 56 rows per warp, ensure_list): fewer, fuller flushes

// ncu: bound by the L1 data pipe; candidates are read as a 16-byte (x0, x1, y0, y1) and an 8-byte (z0, z1) load per pair:
// 24 bytes per pair instead of the former 32-byte record with its two dead w words (-26 % of the cull's wavefronts).
// (occupancy sweep at the end of round 2, profiles/r02_density_occupancy_sweep.txt: a 20- or 16-row stack, or 40 / 48 registers
// by launch bounds, fill more warp slots and are no faster -- 135-152 us against 135-136 for 24 rows at 47 registers)
#ifndef SPH_PK_BLOCKS
#define SPH_PK_BLOCKS 1
#endif
#if SPH_PK_BLOCKS > 1
__global__ void __launch_bounds__(kWalkThreads, SPH_PK_BLOCKS)
#else
__global__ void __launch_bounds__(kWalkThreads)
#endif
k_density_pk(const GatherArgs A, const DevParams P, const uint32_t stack_rows)
{
    chain_prologue();
    extern __shared__ uint2 stk_raw[];             // [stack_rows][kWalkThreads] survivors: (row, bits of the FMA-fused d^2)
    uint2 (*stk)[kWalkThreads] = reinterpret_cast<uint2 (*)[kWalkThreads]>(stk_raw);
    const uint32_t full_mark = (stack_rows - 4u) * (kWalkThreads * 8u);
    constexpr uint32_t kRow = kWalkThreads * 8;    // bytes between two stack rows of a thread
    const int tid = threadIdx.x;
    const uint32_t iraw = P.row0 + blockIdx.x * blockDim.x + tid;
    const bool valid = iraw < P.row1;
    const uint32_t i = valid ? iraw : P.row1 - 1;  // idle tail threads shadow the last row (they join the collectives)
    const Self s = load_self<PASS_DENSITY>(A, P, i);
    const uint32_t K = A.list_k;
    const size_t stride = A.list_stride;
    uint32_t* col = A.list_idx + i;
    uint32_t kbase = 0;                            // list rows written so far (warp-uniform)
    (void)stride;
    Acc acc = {0.0f, 0.0f, 0.0f, 0u};
    const uint32_t sa0 = (uint32_t)__cvta_generic_to_shared(&stk[0][tid]);
    uint32_t sa = sa0;

    // phase B: list row k of the whole warp at once (coalesced).  A flush in mid-walk writes only as many rows as
    // EVERY lane can fill -- or as many as it takes to get the fullest lane down to half a stack, lanes that run
    // short then pad with their own index -- and the entries that stay are moved to the bottom of the stack.  The
    // last flush writes everything, so a warp's list length is the longest true list of its lanes unless the
    // lanes' counts drift apart by more than half a stack.  The sums are scaled by the kernel volumes in the end.
    uint32_t* lp = col;
    float* lw = A.list_w + i;                      // viscosity weight of the entry, same row / column as the index
    const uint32_t klim = valid ? K : 0u;
    const uint32_t keep = stack_rows / 2u;
    auto flush = [&](const bool last) {
        SPH_PK_FENCE();
        const uint32_t ns = (sa - sa0) / kRow;
        const uint32_t mx = __reduce_max_sync(0xffffffffu, ns);
        uint32_t rows = mx;
        if (!last) rows = max(__reduce_min_sync(0xffffffffu, ns), mx > keep ? mx - keep : 0u);
        #pragma unroll 2
        for (uint32_t k = 0; k < rows; k++) {
            const uint2 en = stk[k][tid];
            float d2 = __uint_as_float(en.y);
            bool nb = k < ns;
            if (nb && d2 >= P.cull_lo) {           // inside the band: the reference's exact predicate (:357, Q8)
                float ox, oy, oz;
                d2 = sqr_dist(__ldg(&A.pred[en.x]), s.p, ox, oy, oz);
                nb = !(d2 > P.sqr_r);
            }
            const float w = nb ? fmaxf(P.r - sqrt_approx(d2), 0.0f) : 0.0f;   // kernels.h:27,39: zero unless d < r
            acc.cnt += nb;
            acc.a = fmaf(w, w, acc.a);
            acc.b = fmaf(w * w, w, acc.b);
            // SmoothingViscoPoly6 of the pair (kernels.h:73-82): (r^2 - d^2)^3 * scale where d < r, else 0; 0 for padding
            const float u = nb ? fmaxf(P.rr - d2, 0.0f) : 0.0f;
            if (kbase < klim) { LIST_ST(lp, nb ? en.x : i); LIST_ST(lw, u * u * (u * P.sv)); }
            lp += stride;
            lw += stride;
            kbase++;
        }
        if (last) return;
        for (uint32_t k = rows; k < mx; k++) stk[k - rows][tid] = stk[k][tid];
        sa = sa0 + (ns > rows ? ns - rows : 0u) * kRow;
        SPH_PK_FENCE();
    };

    const Win W = window_of(s.p.x, s.p.y, s.p.z, P);
    // candidates: pair m = rows (2m, 2m+1) as (x0, x1, y0, y1) in one array and (z0, z1) in another (sph_internal.h)
    const float4* __restrict__ pxy = A.predpk;
    const float2* __restrict__ pzz = reinterpret_cast<const float2*>(A.predpk + P.pair_cap);
    const uint32_t last_pair = (P.n - 1u) >> 1;
    const uint64_t px = pk(s.p.x, s.p.x), py = pk(s.p.y, s.p.y), pz = pk(s.p.z, s.p.z);
    const float cull_hi = P.cull_hi;

    // cull one pair: rows (cj, cj + 1), the first at position t of a window of `len` rows (t = -1 when the window
    // starts on an odd row).  Row cj is inside iff 0 <= t < len (one unsigned compare), row cj + 1 iff t < len - 1.
    auto cull = [&](const float4 cxy, const float2 cz, const int t, const int len, const int len1, const uint32_t cj) {
        const uint64_t ox = sub2(pk(cxy.x, cxy.y), px), oy = sub2(pk(cxy.z, cxy.w), py), oz = sub2(pk(cz.x, cz.y), pz);
        const uint64_t d2 = fma2(oz, oz, fma2(oy, oy, mul2(ox, ox)));
        float d0, d1;
        upk(d2, d0, d1);
        asm volatile("{ .reg .pred p, q;\n"
                     " setp.gt.f32 p, %3, %5;\n"
                     " setp.lt.and.u32 q, %1, %2, !p;\n"
                     " @q st.shared.v2.b32 [%0], {%7, %3};\n"
                     " @q add.u32 %0, %0, %9;\n"
                     " setp.gt.f32 p, %4, %5;\n"
                     " setp.lt.and.s32 q, %1, %6, !p;\n"
                     " @q st.shared.v2.b32 [%0], {%8, %4};\n"
                     " @q add.u32 %0, %0, %9; }"
                     : "+r"(sa)
                     : "r"(t), "r"(len), "f"(d0), "f"(d1), "f"(cull_hi), "r"(len1), "r"(cj), "r"(cj + 1u), "n"(kRow)
                     SPH_PK_CLOBBER);
    };
    auto ldxy = [](const float4* p) {
        float4 r;
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        return r;
    };
    auto ldz = [](const float2* p) {
        float2 r;
        asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
        return r;
    };

    // 9 row windows; the bounds of the next row are requested while this one is culled.  Rows and 2-pair chunks
    // are warp-uniform loop levels.
    uint32_t vm = W.rows;                          // rows that exist and can hold a neighbour
    #pragma unroll
    for (int a3 = 0; a3 < 3; a3++) {
        if ((uint32_t)(W.g.y + a3 - 1) >= (uint32_t)P.gdim[1]) vm &= ~(0x49u << a3);
        if ((uint32_t)(W.g.z + a3 - 1) >= (uint32_t)P.gdim[2]) vm &= ~(0x7u << (3 * a3));
    }
    const uint32_t gd0 = (uint32_t)P.gdim[0];
    const uint32_t zstep = (uint32_t)P.gdim[1] * gd0 - 3u * gd0;
    uint32_t tc = (uint32_t)(((int64_t)(W.g.z - 1) * P.gdim[1] + (W.g.y - 1)) * (int64_t)gd0 + W.x0);   // only read where the row exists
    const uint32_t xspan = (uint32_t)(W.x1 - W.x0) + 1u;
    int dyc = 0;
    auto bounds = [&](const int r9, uint32_t& b, uint32_t& e) {
        b = e = 0;
        if ((vm >> r9) & 1u) { b = tbl(A.table, P, tc); e = tbl(A.table, P, tc + xspan); }
        tc += gd0;
        if (++dyc == 3) { dyc = 0; tc += zstep; }
    };

    uint32_t bn, en;
    bounds(0, bn, en);
    #pragma unroll 1
    for (int r9 = 0; r9 < 9; r9++) {
        const uint32_t b = bn, e = en;
        bounds(r9 + 1, bn, en);
        const int len = (int)(e - b);
        const uint32_t p0 = b >> 1;                                    // first pair of the window
        const uint32_t np = len ? ((e + 1u) >> 1) - p0 : 0u;
        const uint32_t iters = __reduce_max_sync(0xffffffffu, np);
        int t = -(int)(b & 1u);
        const int len1 = len - 1;
        if (iters <= kPairPad) {                   // every lane's reads end inside the padded allocation: running pointers
            const float4* qxy = pxy + p0;
            const float2* qz = pzz + p0;
            uint32_t cj = 2u * p0;
            #pragma unroll 1
            for (uint32_t it = 0; it < iters; it += 2, t += 4, qxy += 2, qz += 2, cj += 4u) {
                if (__any_sync(0xffffffffu, sa - sa0 > full_mark)) flush(false);
                const float4 a0 = ldxy(qxy), a1 = ldxy(qxy + 1);
                const float2 z0 = ldz(qz), z1 = ldz(qz + 1);
                cull(a0, z0, t, len, len1, cj);
                cull(a1, z1, t + 2, len, len1, cj + 2u);
            }
            continue;
        }
        #pragma unroll 1
        for (uint32_t it = 0; it < iters; it += 2, t += 4) {
            if (__any_sync(0xffffffffu, sa - sa0 > full_mark)) flush(false);
            const uint32_t q0 = min(p0 + it, last_pair), q1 = min(p0 + it + 1u, last_pair);
            const float4 a0 = ldxy(pxy + q0), a1 = ldxy(pxy + q1);
            const float2 z0 = ldz(pzz + q0), z1 = ldz(pzz + q1);
            cull(a0, z0, t, len, len1, 2u * (p0 + it));
            cull(a1, z1, t + 2, len, len1, 2u * (p0 + it) + 2u);
        }
    }
    flush(true);
    if (valid) {
        acc.a *= P.vol2;
        acc.b *= P.vol3;
        finish<PASS_DENSITY>(A, P, s, acc, 0.0f);
        A.list_cnt[i] = kbase;
    }
    report_overflow(A, (valid && kbase > K) ? kbase : 0u);
    // list rows and warps so far, one 64-bit counter (warps << 32 | rows, the sums carry into each other consistently and the
    // host takes 64-bit differences): kbase is warp-uniform
    if ((tid & 31) == 0 && A.rows_sum) atomicAdd(reinterpret_cast<unsigned long long*>(A.rows_sum), (1ull << 32) | (unsigned long long)kbase);
}

template <int PASS>
__global__ void __launch_bounds__(kWalkThreads)
k_rim_fix(const GatherArgs A, const DevParams P, const float dt)
{
    chain_prologue();
    const uint32_t i = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.row1) return;
    const float4 p = A.pred[i];
    if (!near_table_rim(grid_cell(p.x, p.y, p.z, P), P)) return;
    const Self s = load_self<PASS>(A, P, i);
    Acc acc = {0.0f, 0.0f, 0.0f, 0u};
    for_each_candidate<SPH_TABLE_GRID, true>(A.pred, A.table, A.tend, s.p, P, [&](const uint32_t j, const float4 q) {
        Fetched f;
        f.q = q;
        f.aux = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (PASS != PASS_DENSITY) {
            float ox, oy, oz;
            if (sqr_dist(q, s.p, ox, oy, oz) > P.sqr_r) return;
            f = fetch<PASS>(A, j);
        }
        (void)eval<PASS>(P, s, j, f, acc);
    });
    finish<PASS>(A, P, s, acc, dt);
    // the list the main density kernel recorded holds the extras: mark it overflowed, the later passes walk the
    // table for this particle and their own fix-up overwrites the result
    if (PASS == PASS_DENSITY && A.list_cnt) A.list_cnt[i] = 0xFFFFFFFFu;
}

template <int PASS>
static void launch_rim_fix(cudaStream_t st, const GatherArgs& A, const DevParams& P, float dt, uint64_t* launches)
{
    if (!P.rim_check || P.mode != SPH_TABLE_GRID || P.row1 <= P.row0) return;
    launch_chained(k_rim_fix<PASS>, dim3((P.row1 - P.row0 + kWalkThreads - 1) / kWalkThreads), dim3(kWalkThreads), 0, st, A, P, dt);
    ++*launches;
}

template <int PASS>
void launch(cudaStream_t st, const GatherArgs& A, const DevParams& P, float dt, uint64_t* launches)
{
    if (P.row1 <= P.row0) return;
    const uint32_t rows = P.row1 - P.row0;
    const uint32_t threads = (rows + 1) / 2;
    k_gather2<PASS><<<(threads + GT - 1) / GT, GT, 0, st>>>(A, P, dt);
    ++*launches;
}

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------
// SPH_GATHER selects the enumeration: "list" (default: density walks and records the neighbour list,
// pressure + viscosity replay it), "v1" (every pass walks the table), "v2" (GRID table only: packed
// two-phase cull in every pass).
static int gather_variant()
{
    static const int v = [] {
        const char* e = getenv("SPH_GATHER");
        if (e && e[0] == 'v' && e[1] == '1') return 1;
        if (e && e[0] == 'v' && e[1] == '2') return 2;
        return 0;                            // "list" (lane per particle + list) or "tile" (sph_tile.cu, the default)
    }();
    return v;
}

// SPH_DENSITY selects the density kernel when the neighbour list is on: default k_density_pk on the GRID table and
// k_density_list on the REFERENCE_HASH table; "list" = k_density_list on both, "pair" / "2" = packed
// two-particles-per-thread culls (GRID only), "walk" = single-phase walk.
static int density_variant()
{
    static const int v = [] {
        const char* e = getenv("SPH_DENSITY");
        if (e && e[0] == 'p') return 1;
        if (e && e[0] == 'w') return 2;
        if (e && e[0] == '2') return 3;
        if (e && e[0] == 'l') return 4;      // "list": the scalar two-phase kernel on the GRID table too
        return 0;
    }();
    return v;
}

template <int PASS>
static void launch_walk_or_list(cudaStream_t st, GatherArgs A, const DevParams& P, float dt, bool use_list, uint64_t* launches)
{
    if (P.row1 <= P.row0) return;
    const uint32_t blocks = (P.row1 - P.row0 + kWalkThreads - 1) / kWalkThreads;
    if (!use_list || PASS == PASS_DENSITY) {
        if (P.mode == SPH_TABLE_REFERENCE_HASH) k_gather_walk<SPH_TABLE_REFERENCE_HASH, PASS><<<blocks, kWalkThreads, 0, st>>>(A, P, dt);
        else k_gather_walk<SPH_TABLE_GRID, PASS><<<blocks, kWalkThreads, 0, st>>>(A, P, dt);
    } else {
        if (P.mode == SPH_TABLE_REFERENCE_HASH) launch_chained(k_gather_list<SPH_TABLE_REFERENCE_HASH, PASS>, dim3(blocks), dim3(kWalkThreads), 0, st, A, P, dt);
        else launch_chained(k_gather_list<SPH_TABLE_GRID, PASS>, dim3(blocks), dim3(kWalkThreads), 0, st, A, P, dt);
    }
    ++*launches;
}

// the packed density kernel (and only it) leaves the viscosity weights next to the list entries
static bool density_is_pk(const NbrList& L, const DevParams& P)
{
    return gather_variant() == 0 && density_variant() == 0 && P.mode == SPH_TABLE_GRID && L.idx && L.k && L.w;
}

static GatherArgs base_args(const float4* pred_s, const uint32_t* tstart, const uint32_t* tend, const NbrList& L)
{
    GatherArgs A = {};
    A.pred = pred_s; A.table = tstart; A.tend = tend;
    if (gather_variant() == 0 && L.idx && L.k) { A.list_idx = L.idx; A.list_k = L.k; A.list_stride = L.stride; }
    A.list_cnt = L.cnt;
    A.list_overflow = L.overflow;
    A.key_sorted = L.keys;
    A.list16 = reinterpret_cast<uint16_t*>(L.idx);
    A.list_w = L.w;
    A.tile_need = L.tile_need;
    A.rows_sum = L.rows_sum;
    return A;
}

// the tile generation (sph_tile.cu) runs a pass when the GRID table, the list and the sorted keys are all there
static bool use_tile(const NbrList& L, const DevParams& P)
{
    return tile_enabled() && P.mode == SPH_TABLE_GRID && L.idx && L.k && L.w && L.keys && L.capn;
}

static void launch_density_main(cudaStream_t st, const float4* pred_s, const float4* pred_pk, const uint32_t* tstart, const uint32_t* tend,
                                Rec8* dens, const NbrList& L, const DevParams& P, uint64_t* launches)
{
    GatherArgs A = base_args(pred_s, tstart, tend, L);
    A.predpk = pred_pk;
    A.dens_out = dens; A.ncount = L.ncount;
    if (use_tile(L, P) && launch_tile(st, PASS_DENSITY, A, P, L.capn, 0.0f, launches) == 0) return;
    A.list_w = nullptr;
    if (gather_variant() == 2 && P.mode == SPH_TABLE_GRID) launch<PASS_DENSITY>(st, A, P, 0.0f, launches);
    else if (A.list_idx && density_variant() == 3 && P.mode == SPH_TABLE_GRID) {       // SPH_DENSITY=2: pair2
        if (P.row1 <= P.row0) return;
        const uint32_t threads = (P.row1 - P.row0 + 1) / 2;
        k_density_pair2<<<(threads + GT - 1) / GT, GT, 0, st>>>(A, P);
        ++*launches;
    } else if (A.list_idx && density_variant() == 1 && P.mode == SPH_TABLE_GRID) {     // SPH_DENSITY=pair
        if (P.row1 <= P.row0) return;
        const uint32_t threads = (P.row1 - P.row0 + 1) / 2;
        k_density_pair<<<(threads + GT - 1) / GT, GT, 0, st>>>(A, P);
        ++*launches;
    } else if (A.list_idx && (density_variant() == 0 || density_variant() == 4)) {     // default: two-phase list density
        if (P.row1 <= P.row0) return;
        const uint32_t blocks = (P.row1 - P.row0 + kWalkThreads - 1) / kWalkThreads;
        if (density_is_pk(L, P) && pred_pk) {
            A.list_w = L.w;
            // the deep stack needs the opt-in above 48 KB; the attribute is per device, so it is (re)set whenever used
            uint32_t rows = PKS;
            // (a capacity that auto-grew a little past 64 because ONE pile-up in a corner needed it must not cost every block
            // three quarters of its occupancy: the deep stack is for scenes whose lists are long throughout, like C5)
            if (L.deep) {
                if (cudaFuncSetAttribute(k_density_pk, cudaFuncAttributeMaxDynamicSharedMemorySize, PKS_DENSE * kWalkThreads * 8) == cudaSuccess) rows = PKS_DENSE;
                else cudaGetLastError();
            }
            launch_chained(k_density_pk, dim3(blocks), dim3(kWalkThreads), (size_t)rows * kWalkThreads * 8, st, A, P, rows);
        }
        else if (P.mode == SPH_TABLE_REFERENCE_HASH) k_density_list<SPH_TABLE_REFERENCE_HASH><<<blocks, kWalkThreads, 0, st>>>(A, P);
        else k_density_list<SPH_TABLE_GRID><<<blocks, kWalkThreads, 0, st>>>(A, P);
        ++*launches;
    } else launch_walk_or_list<PASS_DENSITY>(st, A, P, 0.0f, false, launches);           // SPH_DENSITY=walk / no list
}

static void launch_pressure_main(cudaStream_t st, const float4* pred_s, const Rec8* dens, const float4* vel_s,
                                 const uint32_t* tstart, const uint32_t* tend, float4* vel_p, const NbrList& L, const DevParams& P,
                                 float dt, uint64_t* launches)
{
    GatherArgs A = base_args(pred_s, tstart, tend, L);
    A.dens = dens; A.vel_s = vel_s; A.velp_out = vel_p;
    if (use_tile(L, P) && launch_tile(st, PASS_PRESSURE, A, P, L.capn, dt, launches) == 0) return;
    if (gather_variant() == 2 && P.mode == SPH_TABLE_GRID) launch<PASS_PRESSURE>(st, A, P, dt, launches);
    else launch_walk_or_list<PASS_PRESSURE>(st, A, P, dt, A.list_idx != nullptr, launches);
}

static void launch_viscosity_main(cudaStream_t st, const float4* pred_s, const float4* vel_p, const uint32_t* tstart,
                                  const uint32_t* tend, float4* vel_v, const NbrList& L, const DevParams& P, float dt,
                                  uint64_t* launches)
{
    GatherArgs A = base_args(pred_s, tstart, tend, L);
    A.velp = vel_p; A.velv_out = vel_v;
    if (use_tile(L, P) && launch_tile(st, PASS_VISCOSITY, A, P, L.capn, dt, launches) == 0) return;
    if (density_is_pk(L, P) && !getenv("SPH_VISC_NOW")) {        // weights recorded by k_density_pk
        if (P.row1 <= P.row0) return;
        A.list_w = L.w;
        launch_chained(k_viscosity_w<SPH_TABLE_GRID>, dim3((P.row1 - P.row0 + kWalkThreads - 1) / kWalkThreads), dim3(kWalkThreads), 0, st, A, P, dt);
        ++*launches;
    } else if (gather_variant() == 2 && P.mode == SPH_TABLE_GRID) launch<PASS_VISCOSITY>(st, A, P, dt, launches);
    else launch_walk_or_list<PASS_VISCOSITY>(st, A, P, dt, A.list_idx != nullptr, launches);
}

// the public launchers: the pass itself, then (Q2 regime only, never for the tile generation, which compares true cells
// itself) the fix-up of the particles next to the table's rim
void launch_density(cudaStream_t st, const float4* pred_s, const float4* pred_pk, const uint32_t* tstart, const uint32_t* tend,
                    Rec8* dens, const NbrList& L, const DevParams& P, uint64_t* launches)
{
    launch_density_main(st, pred_s, pred_pk, tstart, tend, dens, L, P, launches);
    if (!P.rim_check || use_tile(L, P)) return;
    GatherArgs A = base_args(pred_s, tstart, tend, L);
    A.dens_out = dens; A.ncount = L.ncount;
    launch_rim_fix<PASS_DENSITY>(st, A, P, 0.0f, launches);
}

void launch_pressure(cudaStream_t st, const float4* pred_s, const Rec8* dens, const float4* vel_s,
                     const uint32_t* tstart, const uint32_t* tend, float4* vel_p, const NbrList& L, const DevParams& P,
                     float dt, uint64_t* launches)
{
    launch_pressure_main(st, pred_s, dens, vel_s, tstart, tend, vel_p, L, P, dt, launches);
    if (!P.rim_check || use_tile(L, P)) return;
    GatherArgs A = base_args(pred_s, tstart, tend, L);
    A.dens = dens; A.vel_s = vel_s; A.velp_out = vel_p;
    launch_rim_fix<PASS_PRESSURE>(st, A, P, dt, launches);
}

void launch_viscosity(cudaStream_t st, const float4* pred_s, const float4* vel_p, const uint32_t* tstart,
                      const uint32_t* tend, float4* vel_v, const NbrList& L, const DevParams& P, float dt,
                      uint64_t* launches)
{
    launch_viscosity_main(st, pred_s, vel_p, tstart, tend, vel_v, L, P, dt, launches);
    if (!P.rim_check || use_tile(L, P)) return;
    GatherArgs A = base_args(pred_s, tstart, tend, L);
    A.velp = vel_p; A.velv_out = vel_v;
    launch_rim_fix<PASS_VISCOSITY>(st, A, P, dt, launches);
}

}  // namespace sphb200
