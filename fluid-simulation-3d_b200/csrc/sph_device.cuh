// sph_device.cuh -- device helpers shared by the kernels (exact cell / hash / key / predict arithmetic).
#pragma once
#include "sph_internal.h"

namespace sphb200 {

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
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int3 grid_cell(int3 c, const DevParams& P)
{
    int3 g;
    g.x = clampi(c.x - P.gmin[0], 0, P.gdim[0] - 1);
    g.y = clampi(c.y - P.gmin[1], 0, P.gdim[1] - 1);
    g.z = clampi(clampi(c.z - P.gmin[2], 0, P.gz_global - 1) - P.zlo, 0, P.gdim[2] - 1);
    return g;
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
