// sph_device.cuh -- device helpers shared by the kernels (exact cell / hash / key / predict arithmetic).
#pragma once
#include "sph_internal.h"

namespace sphb200 {

// Programmatic dependent launch (launch_chained, sph_internal.h): first statement of every kernel of the step's chain.
// The kernel may be scheduled as soon as its predecessor's blocks have exited; the wait holds it until that grid has
// completed and its writes are visible.  Every kernel of the chain waits before it touches memory, so completion stays
// transitive down the chain.  A no-op under an ordinary launch.
// (Measured, C2 replayed step: wait only 0.298 ms, ordinary launches 0.3005 ms; with an early
// `griddepcontrol.launch_dependents` in the small kernels 0.300 ms, and in the density kernel 0.377 ms -- the pressure
// pass's blocks then take registers and warp slots next to the density pass's for its whole run, and only wait.)
__device__ __forceinline__ void chain_prologue()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// PositionToCellCoord (:499-503): floor(pos / r) with true IEEE division, C-cast to int.
__device__ __forceinline__ int3 cell_of(float x, float y, float z, float r)
{
    int3 c;
    c.x = __float2int_rz(floorf(__fdiv_rn(x, r)));
    c.y = __float2int_rz(floorf(__fdiv_rn(y, r)));
    c.z = __float2int_rz(floorf(__fdiv_rn(z, r)));
    return c;
}
// HashCell (:505-511): (uint32_t)float on the reference's platform = two's-complement wrap (Q5).
__device__ __forceinline__ uint32_t hash_cell(int cx, int cy, int cz)
{
    return (uint32_t)cx * 15823u + (uint32_t)cy * 9737333u + (uint32_t)cz * 440817757u;
}
// GetKeyFromHash (:513-516): hash % n, exact for every 32-bit operand pair (Lemire fastmod).
__device__ __forceinline__ uint32_t key_of_hash(uint32_t h, const DevParams& P)
{
    const uint64_t low = P.modM * (uint64_t)h;
    return (uint32_t)__umul64hi(low, (uint64_t)P.n);
}
// GRID prefix table entry c = number of rows with key < c, from its two levels (sph_internal.h: kSegShift).  A table
// built flat (radix path) has an all-zero segment array.
__device__ __forceinline__ uint32_t tbl_at(const uint32_t* __restrict__ t, const uint32_t seg_off, const uint32_t c)
{
    return __ldg(&t[c]) + __ldg(&t[seg_off + (c >> kSegShift)]);
}
__device__ __forceinline__ uint32_t tbl(const uint32_t* __restrict__ t, const DevParams& P, const uint32_t c)
{
    return tbl_at(t, P.seg_off, c);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// GRID table cell of a predicted position: (fine x, y, z).  y and z are the reference's cells
// floor(p / r) (:499-503); x is subdivided xsub = 2^k times -- floor((p.x / r) * xsub), exact nesting because
// the scale is a power of two -- so that a row's x window can be cut to the particle's own reach instead of
// three whole cells.  Everything is clamped into the table (border cells collect what lies beyond).
__device__ __forceinline__ int3 grid_cell(const float px, const float py, const float pz, const DevParams& P)
{
    const float tx = __fdiv_rn(px, P.r);
    const int xf = __float2int_rz(floorf(tx * (float)P.xsub));
    const int cy = __float2int_rz(floorf(__fdiv_rn(py, P.r)));
    const int cz = __float2int_rz(floorf(__fdiv_rn(pz, P.r)));
    int3 g;
    g.x = clampi(xf - P.gmin[0] * P.xsub, 0, P.gdim[0] - 1);
    g.y = clampi(cy - P.gmin[1], 0, P.gdim[1] - 1);
    g.z = clampi(clampi(cz - P.gmin[2], 0, P.gz_global - 1) - P.zlo, 0, P.gdim[2] - 1);
    return g;
}

// The candidate window of a particle: its cell, the fine-x range [x0, x1] to read in each (y,z) row, and a
// 9-bit mask of the rows (dz+1)*3 + (dy+1) that can hold a neighbour at all.
//  * x: [p.x - w, p.x + w] with w = sqrt(sqrRadius) widened past any rounding, intersected with the
//    reference's three cells (so a radius smaller than the cull still sees exactly the reference's 27 cells, Q2);
//  * rows: a row is skipped when the particle is farther than w from the row's (y,z) slab (the corner rows,
//    21 % of the time) -- no particle there can pass d^2 <= sqrRadius.
struct Win { int3 g; int x0, x1; uint32_t rows; };

__device__ __forceinline__ Win window_of(const float px, const float py, const float pz, const DevParams& P)
{
    Win W;
    W.g = grid_cell(px, py, pz, P);
    const float S = (float)P.xsub;
    const float slack = fabsf(px) * 3e-7f;
    const int cx = __float2int_rz(floorf(__fdiv_rn(px, P.r)));
    int lo = __float2int_rz(floorf(__fdiv_rn(px - P.xwin - slack, P.r) * S));
    int hi = __float2int_rz(floorf(__fdiv_rn(px + P.xwin + slack, P.r) * S));
    lo = max(lo, (cx - 1) * P.xsub);
    hi = min(hi, (cx + 2) * P.xsub - 1);
    W.x0 = clampi(lo - P.gmin[0] * P.xsub, 0, P.gdim[0] - 1);
    W.x1 = clampi(hi - P.gmin[0] * P.xsub, 0, P.gdim[0] - 1);
    // distance of the particle to the neighbouring slabs of its own cell, in y and z
    const int cy = __float2int_rz(floorf(__fdiv_rn(py, P.r)));
    const int cz = __float2int_rz(floorf(__fdiv_rn(pz, P.r)));
    const float w = P.xwin + (fabsf(py) + fabsf(pz)) * 3e-7f + 1e-6f;
    const float w2 = w * w;
    const float ylo = fmaxf(py - (float)cy * P.r, 0.0f), yhi = fmaxf((float)(cy + 1) * P.r - py, 0.0f);
    const float zlo = fmaxf(pz - (float)cz * P.r, 0.0f), zhi = fmaxf((float)(cz + 1) * P.r - pz, 0.0f);
    uint32_t rows = 0x1FFu;
    // only when the cell was not clamped into the table (then the geometry above does not describe the cell)
    const bool inside = (cy - P.gmin[1] == W.g.y) && (clampi(cz - P.gmin[2], 0, P.gz_global - 1) - P.zlo == W.g.z) &&
                        (cz - P.gmin[2] >= 0) && (cz - P.gmin[2] < P.gz_global);
    if (inside) {
        if (ylo * ylo + zlo * zlo > w2) rows &= ~(1u << 0);     // (dy,dz) = (-1,-1)
        if (yhi * yhi + zlo * zlo > w2) rows &= ~(1u << 2);     // (+1,-1)
        if (ylo * ylo + zhi * zhi > w2) rows &= ~(1u << 6);     // (-1,+1)
        if (yhi * yhi + zhi * zhi > w2) rows &= ~(1u << 8);     // (+1,+1)
    }
    W.rows = rows;
    return W;
}
// Cells on the rim of the GRID table collect what lies beyond it (grid_cell clamps), so there the table cell is not the
// reference's cell.  A particle whose table cell is within one cell of the rim can meet such outliers among its 27
// table cells; for it the walk also compares TRUE cell coordinates: j is in one of the reference's 27 cells of i iff
// they differ by at most one per axis (:331-337).  It matters when the cut-off exceeds the cell size (Q2): with
// sqrt(sqrRadius) <= r the distance test alone excludes everything outside the 27 true cells.
__device__ __forceinline__ bool near_table_rim(const int3 g, const DevParams& P)
{
    if (!P.rim_check) return false;
    const int cx = g.x / P.xsub, ncx = P.gdim[0] / P.xsub, gzg = g.z + P.zlo;
    return cx <= 1 || cx >= ncx - 2 || g.y <= 1 || g.y >= P.gdim[1] - 2 || gzg <= 1 || gzg >= P.gz_global - 2;
}
static __device__ __noinline__ bool within_27(const float4 q, const float4 p, const float r)
{
    const int3 a = cell_of(q.x, q.y, q.z, r), b = cell_of(p.x, p.y, p.z, r);
    return abs(a.x - b.x) <= 1 && abs(a.y - b.y) <= 1 && abs(a.z - b.z) <= 1;
}

__device__ __forceinline__ uint32_t grid_key(int3 g, const DevParams& P)
{
    return ((uint32_t)g.z * (uint32_t)P.gdim[1] + (uint32_t)g.y) * (uint32_t)P.gdim[0] + (uint32_t)g.x;
}

// S1: velocity += externalForce * dt ; predicted = position + velocity * (1/120)   (:45-47, Q1)
__device__ __forceinline__ void predict(const float4 p, float4& v, float3& pred, const DevParams& P, float dt)
{
    const float gy = P.gravity ? -P.g : 0.0f;
    v.x = __fadd_rn(v.x, __fmul_rn(0.0f, dt));
    v.y = __fadd_rn(v.y, __fmul_rn(gy, dt));
    v.z = __fadd_rn(v.z, __fmul_rn(0.0f, dt));
    const float look = 1.0f / 120.0f;
    pred.x = __fadd_rn(p.x, __fmul_rn(v.x, look));
    pred.y = __fadd_rn(p.y, __fmul_rn(v.y, look));
    pred.z = __fadd_rn(p.z, __fmul_rn(v.z, look));
}

// glm::dot on the offset, no FMA: (x*x + y*y) + z*z   (Q8)
__device__ __forceinline__ float sqr_dist(const float4 q, const float4 pi, float& ox, float& oy, float& oz)
{
    ox = __fsub_rn(q.x, pi.x); oy = __fsub_rn(q.y, pi.y); oz = __fsub_rn(q.z, pi.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz));
}


}  // namespace sphb200
