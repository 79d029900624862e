// sph_tile.cu -- the three neighbour-gather passes, "tile" generation (SPH_GATHER=tile; an A/B variant, not the default).
//
//   density   S3  physicsWorld.cc:304-311, 325-365      pressure  S4  :367-422
//   viscosity S5  :424-464 (snapshot semantics)          kernels   engine/physics/kernels.h:25-82
//
// One WARP owns 32 consecutive sorted rows (its "targets").  Sorted order is (z, y, x), so the targets of a warp lie
// in one grid row (y, z) -- or a few: the warp then works segment by segment -- and cover a short run of reference
// cells [c0, c1] along x.  Everything the 27-cell walks of those targets can touch is nine contiguous runs of the
// sorted arrays: rows (y+dy, z+dz), cells [c0-1, c1+1].
//
//  * STAGING.  Lanes 0..8 each issue ONE bulk copy (cp.async.bulk, the TMA engine; SASS UBLKCP) of their run into the
//    warp's own shared-memory buffer and the warp waits on its own mbarrier: no LSU wavefront, no register staging,
//    no block-wide barrier anywhere in these kernels.  The density pass stages predicted positions, the pressure
//    pass the 32-byte density records, the viscosity pass the post-pressure velocities.
//  * DENSITY, phase 1 -- lanes = CANDIDATES.  For every reference cell of the run the warp flattens the cell's 27-cell
//    neighbourhood (nine pieces of the staged runs) over its lanes, four candidates per lane held in registers as
//    fp32x2 pairs, and then loops over the cell's own targets: target position by one broadcast LDS, 12 packed
//    instructions for four d^2, four votes -> a 128-bit accept mask per target in shared memory.  A candidate is
//    loaded once per CELL, not once per target per lane: the L1 data pipe, the roof of the lane-per-particle
//    kernels (sph_gather.cu, 85-95 % busy), carries ~1/3 of the wavefronts.
//  * DENSITY, phase 2 -- lanes = TARGETS.  Every lane walks the set bits of its own mask: staged candidate by LDS,
//    the reference's exact predicate and d^2, the two density kernels, the viscosity kernel value, and the
//    neighbour-list entry -- a 16-bit index into the staged runs -- written row by row for the whole warp (coalesced).
//  * PRESSURE / VISCOSITY replay the list: the same staging (same rounds, same offsets: a pure function of the
//    sorted keys, the table and the staging capacity), then one LDS gather per entry instead of a global gather.
//
// Exactness: phase 1 culls with the FMA-fused d^2 against the widened cull_hi (a superset); phase 2 applies
// (x*x + y*y) + z*z <= sqrRadius in the reference's rounding (:357, Q8).  The candidate set is exactly the
// reference's 27 cells: whole cells, no distance windows, so an interaction radius below the cut-off (Q2) needs no
// special case; targets within one cell of the table border, where cells hold clamped outliers, also compare true
// cell coordinates.  Anything that does not fit the staging buffer (a single cell whose neighbourhood exceeds it)
// or the list falls back to the table walk for the affected targets, and reports the size it needed so that the
// host grows the buffer for the next step.
#include <cstdlib>

#include "sph_gather.cuh"

namespace sphb200 {

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int kTileWarps = 4;                   // warps per block; every warp works alone
constexpr uint32_t kPad16 = 0xFFFFu;             // list padding: "no neighbour in this row of the list"

// ---- mbarrier / bulk copy (PTX ISA: mbarrier, cp.async.bulk) --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(const uint32_t dst, const void* src, const uint32_t bytes, const uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- geometry of a warp's work ------------------------------------------------------------------------------
struct TileCfg {
    uint32_t capn;          // staged candidates per warp (multiple of 128)
    uint32_t warp_bytes;    // shared memory per warp
    uint32_t off_mask, off_map, off_bar;   // byte offsets inside the warp's region (density: masks, slot map)
};

// one staged round: rows (lane k < 9) of cells [c0 - 1, c1 + 1]
struct Round {
    uint32_t rb;            // lane k: table index of the first fine cell of neighbour row k
    bool rv;                // lane k: the row exists
    uint32_t runb, len, so; // lane k: first sorted row of the staged run, its length, its offset in the buffer
    uint32_t total;         // staged candidates; 0xFFFFFFFF: cell c0 alone does not fit
    int c0, c1;
};

template <int PASS> struct PassTraits;
template <> struct PassTraits<PASS_DENSITY>   { static constexpr uint32_t kElem = 16; };
template <> struct PassTraits<PASS_PRESSURE>  { static constexpr uint32_t kElem = 32; };
template <> struct PassTraits<PASS_VISCOSITY> { static constexpr uint32_t kElem = 16; };

__device__ __forceinline__ uint32_t warp_incl_scan16(uint32_t v, const int lane)
{   // inclusive scan over lanes 0..15 (the nine row lanes live there)
    #pragma unroll
    for (int o = 1; o < 16; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, v, o); if (lane >= o) v += u; }
    return v;
}

// S4 term with the reference's own distance: sqrtf is correctly rounded there, and (r - d) loses every bit that an
// approximate root gets wrong when d is within ulps of r (a lone neighbour at the rim of the kernel)
__device__ __forceinline__ void pressure_term(const DevParams& P, const float4 p, const float c0, const float c1s,
                                              const float4 qlo, const float4 qhi, float& ax, float& ay, float& az)
{
    float ox, oy, oz;
    const float d2 = sqr_dist(qlo, p, ox, oy, oz);
    if (d2 > P.sqr_r) return;                              // :402
    const float rs = rsqrt_approx(d2);
    const bool zero = !(d2 > 0.0f);
    const float d0 = d2 * rs;
    const float dn = fmaf(fmaf(-d0, d0, d2), 0.5f * rs, d0);   // one Newton step on the approximate root: sqrtf's value
    const float d = zero ? 0.0f : dn;
    if (d <= P.r) {                                        // kernels.h:51,63
        const float inv = zero ? 0.0f : rs;
        const float v = P.r - d;
        const float c1 = fmaf(c0, qhi.y, P.k);             // (P_i + P_j)/rho_j = (P_i - k rho0)/rho_j + k
        const float c2 = fmaf(c1s, qhi.z, P.kn);           // (nP_i + nP_j)/nrho_j = nP_i/nrho_j + kn
        const float coef = -0.5f * v * fmaf(v * P.s3, c2, P.s2 * c1);
        const float ci = coef * inv;
        ax = fmaf(ox, ci, ax);
        ay = fmaf(oy, ci, ay);
        az = fmaf(oz, ci, az);
        if (zero) ay += coef;                              // dist == 0: direction (0,1,0)  (:414)
    }
}

// ---- table-walk fallback of one target (oversized neighbourhood / overflowed list) -----------------------------
template <int PASS>
__device__ __noinline__ void walk_target(const GatherArgs& A, const DevParams& P, const uint32_t t, const bool border,
                                         const float dt)
{
    const Self s = load_self<PASS>(A, P, t);
    Acc acc = {0.0f, 0.0f, 0.0f, 0u};
    for_each_candidate<SPH_TABLE_GRID, true>(A.pred, A.table, A.tend, s.p, P, [&](const uint32_t j, const float4 q) {
        float ox, oy, oz;
        if (sqr_dist(q, s.p, ox, oy, oz) > P.sqr_r) return;
        if (PASS == PASS_PRESSURE) {
            if (j == t) return;
            const Rec8 r = ld256(&A.dens[j]);
            pressure_term(P, s.p, s.c0, s.c1, r.lo, r.hi, acc.a, acc.b, acc.c);
        } else {
            Fetched f;
            f.q = q;
            f.aux = (PASS == PASS_VISCOSITY) ? __ldg(&A.velp[j]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            (void)eval<PASS>(P, s, j, f, acc);
        }
    });
    finish<PASS>(A, P, s, acc, dt);
}

// ---- the kernel ---------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(kTileWarps * 32)
k_tile(const __grid_constant__ GatherArgs A, const __grid_constant__ DevParams P, const TileCfg T, const float dt)
{
    extern __shared__ __align__(128) unsigned char tile_smem[];
    constexpr uint32_t ESZ = PassTraits<PASS>::kElem;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t wA = P.row0 + (blockIdx.x * kTileWarps + wib) * 32u;
    if (wA >= P.row1) return;                              // warps are independent: no block-wide barrier below
    const uint32_t wB = min(wA + 32u, P.row1);
    unsigned char* const wsm = tile_smem + (size_t)wib * T.warp_bytes;
    const uint32_t cand_s = smem_u32(wsm);                 // staged elements, capn + 1 of them (the last one is "far away")
    const uint32_t bar = smem_u32(wsm + T.off_bar);
    if (lane == 0) {
        mbar_init(bar, 1);
        // element capn, behind everything a copy can reach: a candidate that never passes the cull (idle slots read it)
        if (PASS == PASS_DENSITY) reinterpret_cast<float4*>(wsm)[T.capn] = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.0f);
    }
    __syncwarp();
    uint32_t parity = 0;

    const uint32_t gx = (uint32_t)P.gdim[0], gyd = (uint32_t)P.gdim[1];
    const int xs = __ffs(P.xsub) - 1;
    const int ncx = P.gdim[0] >> xs;                       // reference cells per grid row
    const uint32_t t = wA + lane;
    const bool tv = t < wB;
    const uint32_t key = tv ? __ldg(&A.key_sorted[t]) : 0u;
    const uint32_t R = key / gx;
    const int cx = (int)((key - R * gx) >> xs);
    const int gyy = (int)(R % gyd), gzz = (int)(R / gyd);
    const int gzg = gzz + P.zlo;                           // global z layer (slab mode: the table is a window of it)
    const bool border = P.rim_check && (cx <= 1 || cx >= ncx - 2 || gyy <= 1 || gyy >= (int)gyd - 2 || gzg <= 1 || gzg >= P.gz_global - 2);   // near_table_rim
    const uint32_t K = A.list_k;
    const size_t stride = A.list_stride;

    // per-target state of the pass
    const float4 p = tv ? __ldg(&A.pred[t]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const uint32_t lcnt = (PASS != PASS_DENSITY && tv) ? A.list_cnt[t] : 0u;
    float pc0 = 0.0f, pc1 = 0.0f, prho = 1.0f;
    float4 sv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (PASS == PASS_PRESSURE && tv) {
        const Rec8 d = A.dens[t];
        pc0 = (d.lo.w - P.rho0) * P.k - P.k * P.rho0;      // :371, folded with the neighbour's share (pressure_term)
        pc1 = d.hi.x * P.kn;                               // :372
        prho = d.lo.w;
        sv = A.vel_s[t];
    } else if (PASS == PASS_VISCOSITY && tv) {
        sv = A.velp[t];
    }

    uint32_t pend = __ballot_sync(FULL, tv);
    while (pend) {
        // ---- next segment: the pending targets of one grid row
        const int first = __ffs(pend) - 1;
        const uint32_t Rf = __shfl_sync(FULL, R, first);
        const uint32_t seg = __ballot_sync(FULL, tv && R == Rf) & pend;
        pend &= ~seg;
        const int lastl = 31 - __clz(seg);
        const int cA = __shfl_sync(FULL, cx, first), cB = __shfl_sync(FULL, cx, lastl);
        const int gy0 = __shfl_sync(FULL, gyy, first), gz0 = __shfl_sync(FULL, gzz, first);
        Round Rd;
        {
            const int y = gy0 + (lane % 3) - 1, z = gz0 + (lane / 3) - 1;
            Rd.rv = lane < 9 && (uint32_t)y < gyd && (uint32_t)z < (uint32_t)P.gdim[2];
            Rd.rb = Rd.rv ? ((uint32_t)z * gyd + (uint32_t)y) * gx : 0u;
        }
        int c0 = cA;
        while (c0 <= cB) {
            // ---- next round: as many cells of the segment as the staging buffer holds
            int c1 = cB;
            for (;;) {
                const uint32_t xlo = (uint32_t)max(c0 - 1, 0) << xs, xhi = (uint32_t)min(c1 + 2, ncx) << xs;
                Rd.runb = Rd.rv ? tbl(A.table, P, Rd.rb + xlo) : 0u;
                const uint32_t rune = Rd.rv ? tbl(A.table, P, Rd.rb + xhi) : 0u;
                Rd.len = rune - Rd.runb;
                const uint32_t inc = warp_incl_scan16(Rd.len, lane);
                Rd.total = __shfl_sync(FULL, inc, 8);
                Rd.so = inc - Rd.len;
                if (Rd.total <= T.capn) break;
                if (c1 == c0) {
                    if (lane == 0 && A.tile_need && Rd.total > *(volatile uint32_t*)A.tile_need) atomicMax(A.tile_need, Rd.total);
                    Rd.total = 0xFFFFFFFFu;
                    break;
                }
                c1 = c0 + ((c1 - c0) >> 1);
            }
            Rd.c0 = c0; Rd.c1 = c1;
            const bool mine = ((seg >> lane) & 1u) && cx >= c0 && cx <= c1;
            const uint32_t act = __ballot_sync(FULL, mine);
            c0 = c1 + 1;
            if (!act) continue;                            // cells without a target of this warp
            if (Rd.total == 0xFFFFFFFFu) {                 // this cell's neighbourhood does not fit: walk the table
                if (mine) {
                    walk_target<PASS>(A, P, t, border, dt);
                    if (PASS == PASS_DENSITY) A.list_cnt[t] = 0xFFFFFFFFu;      // "overflowed": the later passes walk too
                }
                __syncwarp();
                continue;
            }
            // ---- stage the nine runs (TMA bulk copies into this warp's buffer)
            const unsigned char* src = (PASS == PASS_DENSITY)    ? reinterpret_cast<const unsigned char*>(A.pred)
                                       : (PASS == PASS_PRESSURE) ? reinterpret_cast<const unsigned char*>(A.dens)
                                                                 : reinterpret_cast<const unsigned char*>(A.velp);
            fence_proxy_async();                           // the buffer's previous readers (generic proxy) are done
            if (lane == 0) mbar_expect_tx(bar, Rd.total * ESZ);
            __syncwarp();
            if (Rd.len) bulk_g2s(cand_s + Rd.so * ESZ, src + (size_t)Rd.runb * ESZ, Rd.len * ESZ, bar);
            const uint32_t so4 = __shfl_sync(FULL, Rd.so, 4), runb4 = __shfl_sync(FULL, Rd.runb, 4);

            if (PASS == PASS_DENSITY) {
                uint32_t* const masks = reinterpret_cast<uint32_t*>(wsm + T.off_mask);     // [32][MW]
                uint16_t* const smap = reinterpret_cast<uint16_t*>(wsm + T.off_map);       // flat position -> staged index
                const float4* const cand = reinterpret_cast<const float4*>(wsm);
                const uint32_t MW = T.capn >> 5;
                uint32_t my_off = 0, my_L = 0;
                uint32_t smoff = 0;
                bool waited = false;
                // ---- phase 1: cell by cell, lanes = candidates
                uint32_t cpend = act;
                while (cpend) {
                    const int fl = __ffs(cpend) - 1;
                    const int c = __shfl_sync(FULL, cx, fl);
                    const uint32_t cl = __ballot_sync(FULL, mine && cx == c);               // this cell's targets (lanes)
                    cpend &= ~cl;
                    const uint32_t plo = (uint32_t)max(c - 1, 0) << xs, phi = (uint32_t)min(c + 2, ncx) << xs;
                    const uint32_t cb = Rd.rv ? tbl(A.table, P, Rd.rb + plo) : 0u;
                    const uint32_t ce = Rd.rv ? tbl(A.table, P, Rd.rb + phi) : 0u;
                    const uint32_t plen = ce - cb;
                    const uint32_t pinc = warp_incl_scan16(plen, lane);
                    const uint32_t L = __shfl_sync(FULL, pinc, 8);
                    const uint32_t ppre = pinc - plen;
                    const uint32_t pss = Rd.so + (cb - Rd.runb);                            // staged start of the piece
                    // slot map of the cell: flat position -> staged index, piece by piece
                    #pragma unroll
                    for (int k = 0; k < 9; k++) {
                        const uint32_t kl = __shfl_sync(FULL, plen, k);
                        if (kl == 0) continue;
                        const uint32_t kp = __shfl_sync(FULL, ppre, k), ks = __shfl_sync(FULL, pss, k);
                        for (uint32_t e = lane; e < kl; e += 32) smap[smoff + kp + e] = (uint16_t)(ks + e);
                    }
                    __syncwarp();
                    if ((cl >> lane) & 1u) { my_off = smoff; my_L = L; }
                    if (!waited) { mbar_wait(bar, parity); parity ^= 1u; waited = true; }
                    for (uint32_t q0 = 0; q0 < L; q0 += 128) {
                        // four candidates per lane, as fp32x2 pairs (slots 0|1 and 2|3)
                        float4 cq[4];
                        #pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const uint32_t pos = q0 + 32u * s + lane;
                            const uint32_t ci = pos < L ? (uint32_t)smap[smoff + pos] : T.capn;   // capn: the far-away element
                            cq[s] = cand[ci];
                        }
                        const uint64_t x01 = pk(cq[0].x, cq[1].x), y01 = pk(cq[0].y, cq[1].y), z01 = pk(cq[0].z, cq[1].z);
                        const uint64_t x23 = pk(cq[2].x, cq[3].x), y23 = pk(cq[2].y, cq[3].y), z23 = pk(cq[2].z, cq[3].z);
                        uint32_t tl = cl;
                        while (tl) {
                            const int tlane = __ffs(tl) - 1;
                            tl &= tl - 1;
                            const float4 tp = cand[so4 + (wA + (uint32_t)tlane - runb4)];   // the target itself, broadcast
                            const uint64_t px = pk(tp.x, tp.x), py = pk(tp.y, tp.y), pz = pk(tp.z, tp.z);
                            uint64_t ox = sub2(x01, px), oy = sub2(y01, py), oz = sub2(z01, pz);
                            const uint64_t da = fma2(oz, oz, fma2(oy, oy, mul2(ox, ox)));
                            ox = sub2(x23, px); oy = sub2(y23, py); oz = sub2(z23, pz);
                            const uint64_t db = fma2(oz, oz, fma2(oy, oy, mul2(ox, ox)));
                            float d0, d1, d2, d3;
                            upk(da, d0, d1);
                            upk(db, d2, d3);
                            const uint32_t b0 = __ballot_sync(FULL, !(d0 > P.cull_hi)), b1 = __ballot_sync(FULL, !(d1 > P.cull_hi));
                            const uint32_t b2 = __ballot_sync(FULL, !(d2 > P.cull_hi)), b3 = __ballot_sync(FULL, !(d3 > P.cull_hi));
                            if (lane == 0) *reinterpret_cast<uint4*>(&masks[(uint32_t)tlane * MW + (q0 >> 5)]) = make_uint4(b0, b1, b2, b3);
                        }
                    }
                    smoff += L;
                }
                __syncwarp();
                // ---- phase 2: lanes = targets, walk the accept masks
                float rho = 0.0f, rhon = 0.0f;
                uint32_t cnt = 0, nl = 0;
                const uint32_t nw = mine ? (my_L + 31u) >> 5 : 0u;
                const uint32_t* const mrow = masks + (uint32_t)lane * MW;
                uint32_t w = 0, m = 0, base = 0;
                uint16_t* lp = A.list16 + t;
                float* lw = A.list_w + t;
                for (;;) {
                    while (m == 0u && w < nw) { m = mrow[w]; base = w << 5; ++w; }
                    if (!__any_sync(FULL, m != 0u)) break;
                    uint32_t ent = kPad16;
                    float wv = 0.0f;
                    if (m) {
                        const uint32_t pos = base + (uint32_t)__ffs(m) - 1u;
                        m &= m - 1u;
                        const uint32_t ci = smap[my_off + pos];
                        const float4 q = cand[ci];
                        float ox, oy, oz;
                        const float d2 = sqr_dist(q, p, ox, oy, oz);                       // the reference's rounding (Q8)
                        bool nb = !(d2 > P.sqr_r);                                          // :357
                        if (nb && border) nb = within_27(q, p, P.r);
                        if (nb) {
                            ++cnt;
                            const float v = fmaxf(P.r - sqrt_approx(d2), 0.0f);            // kernels.h:27,39: zero unless d < r
                            rho = fmaf(v, v, rho);
                            rhon = fmaf(v * v, v, rhon);
                            const float u = fmaxf(P.rr - d2, 0.0f);                        // kernels.h:78
                            wv = u * u * (u * P.sv);
                            ent = ci;
                        }
                    }
                    if (mine && nl < K) { __stcs(lp, (uint16_t)ent); __stcs(lw, wv); }
                    lp += stride;
                    lw += stride;
                    ++nl;
                }
                if (mine) {
                    Rec8 r;
                    const float a = rho * P.vol2, b = rhon * P.vol3;
                    r.lo = make_float4(p.x, p.y, p.z, a);
                    r.hi = make_float4(b, __fdiv_rn(1.0f, a), __fdiv_rn(1.0f, b), 0.0f);
                    A.dens_out[t] = r;
                    if (A.ncount) A.ncount[t] = cnt;
                    A.list_cnt[t] = nl;
                    if (nl > K && A.list_overflow && nl > *(volatile uint32_t*)A.list_overflow) atomicMax(A.list_overflow, nl);
                }
                __syncwarp();
            } else {
                // ---- replay: lanes = targets, entries are indices into the staged runs
                mbar_wait(bar, parity);
                parity ^= 1u;
                const bool walk = mine && lcnt > K;        // overflowed (or never recorded): walk the table
                const uint32_t myn = (mine && !walk) ? lcnt : 0u;
                const uint32_t nmax = __reduce_max_sync(FULL, myn);
                const uint32_t self = so4 + (t - runb4);
                const uint16_t* lp = A.list16 + t;
                float ax = 0.0f, ay = 0.0f, az = 0.0f;
                if (PASS == PASS_PRESSURE) {
                    const float4* const rec = reinterpret_cast<const float4*>(wsm);
                    #pragma unroll 2
                    for (uint32_t k = 0; k < nmax; k++) {
                        const uint32_t ci = k < myn ? (uint32_t)__ldcs(lp + (size_t)k * stride) : kPad16;
                        if (ci != kPad16 && ci != self) {
                            const float4 lo = rec[2u * ci], hi = rec[2u * ci + 1u];
                            pressure_term(P, p, pc0, pc1, lo, hi, ax, ay, az);
                        }
                    }
                    if (mine && !walk) {
                        const float kk = dt / prho;                                         // :421
                        A.velp_out[t] = make_float4(fmaf(ax, kk, sv.x), fmaf(ay, kk, sv.y), fmaf(az, kk, sv.z), 0.0f);
                    }
                } else {
                    const float4* const vel = reinterpret_cast<const float4*>(wsm);
                    const float* lw = A.list_w + t;
                    #pragma unroll 2
                    for (uint32_t k = 0; k < nmax; k++) {
                        const bool in = k < myn;
                        const uint32_t ci = in ? (uint32_t)__ldcs(lp + (size_t)k * stride) : kPad16;
                        const float wv = in ? __ldcs(lw + (size_t)k * stride) : 0.0f;
                        if (ci != kPad16) {                 // the particle's own entry contributes (v_i - v_i) * w = 0 (:450)
                            const float4 vj = vel[ci];
                            ax = fmaf(vj.x - sv.x, wv, ax);
                            ay = fmaf(vj.y - sv.y, wv, ay);
                            az = fmaf(vj.z - sv.z, wv, az);
                        }
                    }
                    if (mine && !walk) {
                        const float kk = P.mu * dt;                                         // :463
                        A.velv_out[t] = make_float4(fmaf(ax, kk, sv.x), fmaf(ay, kk, sv.y), fmaf(az, kk, sv.z), 0.0f);
                    }
                }
                if (walk) walk_target<PASS>(A, P, t, border, dt);
                __syncwarp();
            }
        }
    }
}

template <int PASS>
TileCfg tile_cfg(const uint32_t capn)
{
    TileCfg T;
    T.capn = capn;
    uint32_t off = (capn + 1u) * PassTraits<PASS>::kElem;          // + the far-away element
    off = (off + 15u) & ~15u;
    T.off_mask = off;
    if (PASS == PASS_DENSITY) off += capn * 4u;                    // [32][capn / 32] mask words
    T.off_map = off;
    if (PASS == PASS_DENSITY) off += 3u * capn * 2u;               // every staged candidate is in at most three cells' neighbourhoods
    off = (off + 15u) & ~15u;
    T.off_bar = off;
    off += 16u;
    T.warp_bytes = (off + 127u) & ~127u;
    return T;
}

}  // namespace

uint32_t tile_default_capn()
{
    static const uint32_t v = [] {
        const char* e = getenv("SPH_TILE_CAPN");
        uint32_t c = e ? (uint32_t)atoi(e) : 512u;
        c = (c + 127u) & ~127u;
        return c < 128u ? 128u : (c > 1664u ? 1664u : c);
    }();
    return v;
}

bool tile_enabled()
{   // an A/B generation (measured slower than the lane-per-particle kernels): SPH_GATHER=tile
    static const bool on = [] { const char* e = getenv("SPH_GATHER"); return e && e[0] == 't'; }();
    return on;
}

template <int PASS>
static int launch_tile_pass(cudaStream_t st, const GatherArgs& A, const DevParams& P, const uint32_t capn, const float dt,
                            uint64_t* launches)
{
    if (P.row1 <= P.row0) return 0;
    const TileCfg T = tile_cfg<PASS>(capn);
    const size_t bytes = (size_t)T.warp_bytes * kTileWarps;
    static size_t opted[3] = {0, 0, 0};
    if (bytes > 48u * 1024u && bytes > opted[PASS]) {
        if (cudaFuncSetAttribute(k_tile<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        opted[PASS] = bytes;
    }
    const uint32_t rows = P.row1 - P.row0;
    const uint32_t blocks = (rows + kTileWarps * 32u - 1u) / (kTileWarps * 32u);
    k_tile<PASS><<<blocks, kTileWarps * 32, bytes, st>>>(A, P, T, dt);
    ++*launches;
    return 0;
}

int launch_tile(cudaStream_t st, const int pass, const GatherArgs& A, const DevParams& P, const uint32_t capn, const float dt,
                uint64_t* launches)
{
    if (pass == PASS_DENSITY) return launch_tile_pass<PASS_DENSITY>(st, A, P, capn, dt, launches);
    if (pass == PASS_PRESSURE) return launch_tile_pass<PASS_PRESSURE>(st, A, P, capn, dt, launches);
    return launch_tile_pass<PASS_VISCOSITY>(st, A, P, capn, dt, launches);
}

}  // namespace sphb200
