// sph_gather.cuh -- the physics of the three gather passes, shared by every kernel variant.
//
//   density   S3  physicsWorld.cc:304-311, 325-365      pressure  S4  :367-422
//   viscosity S5  :424-464 (snapshot semantics)          kernels   engine/physics/kernels.h:25-82
//
// load_self / fetch / eval / finish are the ONLY implementation of the per-pair arithmetic; the kernels
// in sph_gather.cu differ only in how they enumerate candidates (27-cell walk of either table, packed
// two-phase cull, or the neighbour list recorded by the density pass).
#pragma once
#include "sph_device.cuh"

namespace sphb200 {

enum { PASS_DENSITY = 0, PASS_PRESSURE = 1, PASS_VISCOSITY = 2 };

struct GatherArgs {
    const float4* pred;        // sorted predicted positions (+ float(hash) in w)
    const float4* predpk;      // the same rows pair-interleaved (PredPair, sph_internal.h): density pass, GRID table
    const uint32_t* table;     // GRID prefix table / REFERENCE_HASH startIndices
    const uint32_t* tend;      // REFERENCE_HASH bucket ends
    // neighbour list recorded by the density pass: entry k of row i at list_idx[k * list_stride + i],
    // list_cnt[i] entries (a count above list_k means "overflowed: walk the table instead")
    uint32_t* list_idx;
    float* list_w;             // viscosity weights of the entries (k_density_pk); nullptr: the viscosity pass computes them
    uint32_t* list_cnt;
    uint32_t list_k, list_stride;
    uint32_t* list_overflow;   // device word: largest list length seen above list_k (drives the host's auto-grow)
    Rec8* dens_out;            // density pass: density records
    uint32_t* ncount;          // density pass, optional
    const Rec8* dens;          // pressure pass
    const float4* vel_s;       // pressure pass: own velocity after S1
    float4* velp_out;          // pressure pass: v' (velocity after pressure)
    const float4* velp;        // viscosity pass: post-pressure snapshot
    float4* velv_out;          // viscosity pass
    // tile generation (sph_tile.cu)
    const uint32_t* key_sorted; // GRID keys of the sorted rows
    uint16_t* list16;          // neighbour list as 16-bit indices into the warp's staged runs (same geometry as list_idx)
    uint32_t* tile_need;       // device word: largest single-cell neighbourhood that did not fit the staging buffer
    uint32_t* rows_sum;        // [2] running sums of list rows / warps (k_density_pk): the host picks the stack depth from their mean
};

// ---- packed fp32x2 (sm_100a) -------------------------------------------------
__device__ __forceinline__ uint64_t pk(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- per-particle state ---------------------------------------------------------
__device__ __forceinline__ Rec8 ld256(const Rec8* p)
{
    Rec8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
}

struct Self {
    float4 p;        // predicted position
    uint32_t i;      // sorted row
    float c0, c1;    // pressure: P_i - k*rho0, near P_i
    float4 v;        // pressure: own velocity (S1);  viscosity: own post-pressure velocity
    float rho;       // pressure: own density
};
struct Acc {
    float a, b, c;   // density: (rho, near rho, -); pressure / viscosity: force
    uint32_t cnt;    // density: accepted neighbours incl. self
};

template <int PASS>
__device__ __forceinline__ Self load_self(const GatherArgs& A, const DevParams& P, uint32_t i)
{
    Self s;
    s.i = i;
    s.p = A.pred[i];
    s.c0 = s.c1 = s.rho = 0.0f;
    s.v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (PASS == PASS_PRESSURE) {
        const Rec8 d = A.dens[i];
        const float pressure = (d.lo.w - P.rho0) * P.k;   // :371
        s.c0 = pressure - P.k * P.rho0;
        s.c1 = d.hi.x * P.kn;                               // :372
        s.rho = d.lo.w;
        s.v = A.vel_s[i];
    } else if (PASS == PASS_VISCOSITY) {
        s.v = A.velp[i];
    }
    return s;
}

// One candidate that survived the cull: fetch (issued one iteration ahead), then evaluate with the
// reference's exact predicate.
struct Fetched {
    float4 q;      // predicted position of the candidate
    float4 aux;    // pressure: (rho, near rho, 1/rho, 1/near rho) of the candidate; viscosity: its post-pressure velocity
};

template <int PASS>
__device__ __forceinline__ Fetched fetch(const GatherArgs& A, const uint32_t j)
{
    Fetched f;
    if (PASS == PASS_PRESSURE) {            // (rho, near rho, 1/rho, 1/near rho)
        const Rec8 r = ld256(&A.dens[j]);
        f.q = r.lo;
        f.aux = make_float4(r.lo.w, r.hi.x, r.hi.y, r.hi.z);
    } else if (PASS == PASS_VISCOSITY) {    // post-pressure velocity
        f.q = __ldg(&A.pred[j]);
        f.aux = __ldg(&A.velp[j]);
    } else {
        f.q = __ldg(&A.pred[j]);
        f.aux = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    return f;
}

template <int PASS>
__device__ __forceinline__ bool eval(const DevParams& P, const Self& s, const uint32_t j, const Fetched& f, Acc& acc)
{   // returns whether j is a neighbour by the reference's predicate (and, for S4/S5, not the particle itself)
    float ox, oy, oz;
    const float d2 = sqr_dist(f.q, s.p, ox, oy, oz);       // (ox*ox + oy*oy) + oz*oz, no FMA (Q8)
    if (d2 > P.sqr_r) return false;                           // :357 / :402 / :456
    if (PASS == PASS_DENSITY) {
        acc.cnt++;
        const float d = sqrt_approx(d2);
        if (d < P.r) {                                     // kernels.h:27,39
            const float v = P.r - d;
            const float v2 = v * v;
            acc.a = fmaf(v2, P.vol2, acc.a);
            acc.b = fmaf(v2 * v, P.vol3, acc.b);
        }
    } else if (PASS == PASS_PRESSURE) {
        if (j == s.i) return false;                        // :396
        // the reference's own distance (sqrtf is correctly rounded): r - d keeps no bit that an approximate root gets
        // wrong when d is within ulps of r -- a lone neighbour at the rim of the kernel (tools/gpu_fuzz.py)
        const float rs = rsqrt_approx(d2);
        const bool zero = !(d2 > 0.0f);
        const float d0 = d2 * rs;
        const float d = zero ? 0.0f : fmaf(fmaf(-d0, d0, d2), 0.5f * rs, d0);     // one Newton step: sqrtf's value
        const float inv = zero ? 0.0f : rs;
        if (d <= P.r) {                                    // kernels.h:51,63
            const float v = P.r - d;
            // (P_i + P_j)/rho_j = (P_i - k rho0)/rho_j + k ;  (nP_i + nP_j)/nrho_j = nP_i/nrho_j + kn
            const float c1 = fmaf(s.c0, f.aux.z, P.k);
            const float c2 = fmaf(s.c1, f.aux.w, P.kn);
            const float coef = -0.5f * v * fmaf(v * P.s3, c2, P.s2 * c1);
            const float ci = coef * inv;
            acc.a = fmaf(ox, ci, acc.a);
            acc.b = fmaf(oy, ci, acc.b);
            acc.c = fmaf(oz, ci, acc.c);
            if (zero) acc.b += coef;                       // dist == 0: direction (0,1,0)  (:414)
        }
    } else {
        if (j == s.i) return false;                        // :450
        const float w = P.rr - d2;                         // kernels.h:78
        if (w > 0.0f) {
            const float w3 = w * w * w * P.sv;
            acc.a = fmaf(f.aux.x - s.v.x, w3, acc.a);
            acc.b = fmaf(f.aux.y - s.v.y, w3, acc.b);
            acc.c = fmaf(f.aux.z - s.v.z, w3, acc.c);
        }
    }
    return true;
}

template <int PASS>
__device__ __forceinline__ void term(const GatherArgs& A, const DevParams& P, const Self& s, const uint32_t j, Acc& acc)
{
    const Fetched f = fetch<PASS>(A, j);
    (void)eval<PASS>(P, s, j, f, acc);
}

template <int PASS>
__device__ __forceinline__ void finish(const GatherArgs& A, const DevParams& P, const Self& s, const Acc& acc, const float dt)
{
    if (PASS == PASS_DENSITY) {
        Rec8 r;
        r.lo = make_float4(s.p.x, s.p.y, s.p.z, acc.a);
        r.hi = make_float4(acc.b, __fdiv_rn(1.0f, acc.a), __fdiv_rn(1.0f, acc.b), 0.0f);
        A.dens_out[s.i] = r;
        if (A.ncount) A.ncount[s.i] = acc.cnt;
    } else if (PASS == PASS_PRESSURE) {
        const float k = dt / s.rho;                        // :421
        A.velp_out[s.i] = make_float4(fmaf(acc.a, k, s.v.x), fmaf(acc.b, k, s.v.y), fmaf(acc.c, k, s.v.z), 0.0f);
    } else {
        const float k = P.mu * dt;                         // :463
        A.velv_out[s.i] = make_float4(fmaf(acc.a, k, s.v.x), fmaf(acc.b, k, s.v.y), fmaf(acc.c, k, s.v.z), 0.0f);
    }
}

// ---- neighbour walk ---------------------------------------------------------
// Calls f(j, pred_j) for every candidate row the reference's walk would reach for a particle at pi.
template <int MODE, bool RIM = false, class F>
__device__ __forceinline__ void for_each_candidate(const float4* __restrict__ pred_s,
                                                   const uint32_t* __restrict__ tstart,
                                                   const uint32_t* __restrict__ tend,
                                                   const float4 pi, const DevParams& P, F&& f)
{
    const int3 c = cell_of(pi.x, pi.y, pi.z, P.r);
    if (MODE == SPH_TABLE_REFERENCE_HASH) {
        #pragma unroll 1
        for (int i = 0; i < 27; i++) {               // offsets[27]: x outer, y, z inner (physicsWorld.h:131-143)
            const int dx = i / 9 - 1, dy = (i / 3) % 3 - 1, dz = i % 3 - 1;
            const uint32_t h = hash_cell(c.x + dx, c.y + dy, c.z + dz);
            const uint32_t key = key_of_hash(h, P);
            const uint32_t b = __ldg(&tstart[key]);
            if (b >= P.n) continue;                  // 0x7FFFFFFF: empty bucket (:339)
            const uint32_t e = __ldg(&tend[key]);
            const float hf = __uint2float_rn(h);
            for (uint32_t j = b; j < e; j++) {
                const float4 q = __ldg(&pred_s[j]);
                if (q.w != hf) continue;             // `index.y != hash`, compared in float (:346)
                f(j, q);
            }
        }
    } else {
        const Win W = window_of(pi.x, pi.y, pi.z, P);
        const int3 g = W.g;
        const int x0 = W.x0, x1 = W.x1;
        const bool rim = RIM && near_table_rim(g, P);
        #pragma unroll 1
        for (int dz = -1; dz <= 1; dz++) {
            const int z = g.z + dz;
            if (z < 0 || z >= P.gdim[2]) continue;
            #pragma unroll 1
            for (int dy = -1; dy <= 1; dy++) {
                const int y = g.y + dy;
                if (y < 0 || y >= P.gdim[1] || !((W.rows >> ((dz + 1) * 3 + dy + 1)) & 1u)) continue;
                const uint32_t row = ((uint32_t)z * (uint32_t)P.gdim[1] + (uint32_t)y) * (uint32_t)P.gdim[0];
                const uint32_t b = tbl(tstart, P, row + x0);
                const uint32_t e = tbl(tstart, P, row + x1 + 1);
                for (uint32_t j = b; j < e; j++) {
                    const float4 q = __ldg(&pred_s[j]);
                    if (RIM && rim && !within_27(q, pi, P.r)) continue;      // clamped outliers in the rim cells (Q2)
                    f(j, q);
                }
            }
        }
    }
}


}  // namespace sphb200
