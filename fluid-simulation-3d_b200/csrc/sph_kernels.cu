// sph_kernels.cu -- sm_100a kernels of the per-timestep SPH update.
//
// Stage map (reference: engine/physics/physicsWorld.cc, kernels: engine/physics/kernels.h):
//   k_predict_key   S1 :42-48 (+CalculateExternalFoce :313-323) and the key half of S2 :473-482
//   k_build_table*  S2 :486-496 (start table) -- GRID mode builds a gap-free prefix table instead
//   k_reorder       "particle reordering for locality": gathers the state into sorted order
//   k_density       S3 :304-311, 325-365         k_pressure  S4 :367-422
//   k_viscosity     S5 :424-464 (snapshot/Jacobi) k_integrate S6 :81-108
//
// Data layout: SoA of float4 rows in DEVICE ORDER (= sorted order of the latest step):
//   pos4  = (x, y, z, bits(global particle id))      vel4 = (vx, vy, vz, 0)
//   pred4 = (px, py, pz, float(hash))                dens = float2 (rho, near rho)
//
// Exactness rules (SURVEY App.A Q5/Q6/Q8): everything that feeds an integer result (cell, hash, key,
// the d^2 <= sqrRadius predicate) is computed with explicit round-to-nearest, non-fused
// __fmul_rn/__fadd_rn/__fdiv_rn in the reference's operation order.  Float-only results may use FMA
// and approximate sqrt/rcp (tolerance 1e-5 of the stage scale, tests/test_parity_gpu.py).
#include <cstdlib>

#include "sph_internal.h"
#include "sph_device.cuh"

namespace sphb200 {

namespace {


__device__ __forceinline__ void count_segment(uint32_t* __restrict__ count, const DevParams& P, const uint32_t k)
{
    const uint32_t sg = k >> kSegShift;
    const uint32_t peers = __match_any_sync(__activemask(), sg);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&count[P.cnt_off + sg], (uint32_t)__popc(peers));
    // (Keeping the sums of the scan's 4096-segment tiles here as well, to save the scan its reduce launch, was measured:
    // one more atomic per warp on a handful of addresses serialises -- k_predict_key 19 -> 46 us at 1 M rows, 0.11 -> 0.55 ms
    // at 8 M.  The reduce launch costs 3 us.)
}

__global__ void __launch_bounds__(256)
k_predict_key(const float4* __restrict__ pos, const float4* __restrict__ vel, uint32_t* __restrict__ key,
              uint8_t* __restrict__ cls, const uint32_t rows, const bool may_migrate, const DevParams P, const float dt,
              uint32_t* __restrict__ count, uint32_t* __restrict__ rank, uint32_t* __restrict__ pack_counts,
              const uint32_t pack_blocks)
{
    chain_prologue();
    // counting sort (GRID table): the row takes a ticket in its cell's counter; the tickets are made
    // canonical (ascending source row) later, in k_reorder_binned
    // ... and is counted in its segment (one atomic per distinct segment of the warp: rows arrive nearly sorted)
    #define SPH_EMIT_KEY(K) do { const uint32_t k__ = (K); key[s] = k__; if (count) { rank[s] = atomicAdd(&count[k__], 1u); count_segment(count, P, k__); } } while (0)
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= rows) return;
    const float4 p = pos[s];
    float4 v = vel[s];
    float3 pr;
    predict(p, v, pr, P, dt);
    const int3 c = cell_of(pr.x, pr.y, pr.z, P.r);
    if (P.mode == SPH_TABLE_REFERENCE_HASH) {
        SPH_EMIT_KEY(key_of_hash(hash_cell(c.x, c.y, c.z), P));
        return;
    }
    if (P.slab) {
        // slab mode: ownership follows the z layer of the PREDICTED position, so every owned row finds
        // all its 27 cells in the local table (own layers + one ghost layer per side)
        const int gz = clampi(c.z - P.gmin[2], 0, P.gz_global - 1);
        uint8_t k = CLS_STAY;
        if (may_migrate) {
            if (gz < P.own_lo && P.has_lo) k = CLS_MIG_LO;
            else if (gz >= P.own_hi && P.has_hi) k = CLS_MIG_HI;
            else {
                if (gz == P.own_lo && P.has_lo) k |= CLS_GHOST_LO;
                if (gz == P.own_hi - 1 && P.has_hi) k |= CLS_GHOST_HI;
            }
            if ((k & (CLS_MIG_LO | CLS_MIG_HI)) && (gz == P.own_lo - 1 || gz == P.own_hi)) k |= CLS_KEEP;
            cls[s] = k;
            // rows of this warp in each of the six pack lists, added to the counters of its pack block (sph_multi.cu:
            // the order-preserving pack needs them per block; only warps at a slab boundary have any)
            const uint32_t act = __activemask();
            if (pack_counts && __any_sync(act, k != 0)) {
                const bool keep = k & CLS_KEEP;
                const uint32_t b[NLISTS] = {__ballot_sync(act, k & CLS_MIG_LO), __ballot_sync(act, k & CLS_GHOST_LO),
                                            __ballot_sync(act, (k & CLS_MIG_LO) && keep), __ballot_sync(act, k & CLS_MIG_HI),
                                            __ballot_sync(act, k & CLS_GHOST_HI), __ballot_sync(act, (k & CLS_MIG_HI) && keep)};
                if ((int)(threadIdx.x & 31) == __ffs(act) - 1) {
                    #pragma unroll
                    for (int l = 0; l < NLISTS; l++)
                        if (b[l]) atomicAdd(&pack_counts[(size_t)l * pack_blocks + s / kPackSpan], (uint32_t)__popc(b[l]));
                }
            }
        }
        if (k & (CLS_MIG_LO | CLS_MIG_HI)) { SPH_EMIT_KEY(P.ncell); return; }  // leaves this rank: sorts past the table
        int3 g = grid_cell(pr.x, pr.y, pr.z, P);
        // A row that may not migrate (it arrived this step) but whose layer lies beyond this whole slab -- it moved
        // more than a slab in one step -- is parked one layer INSIDE the slab: the boundary layers must hold exactly
        // the rows the neighbour rank mirrors as ghosts.  It finds no neighbours there (the cull is by distance),
        // moves ballistically for this step and migrates onward in the next one.
        int gzc = gz;
        if (gz < P.own_lo) gzc = min(P.own_lo + 1, P.own_hi - 1);
        else if (gz >= P.own_hi) gzc = max(P.own_hi - 2, P.own_lo);
        g.z = gzc - P.zlo;
        SPH_EMIT_KEY(grid_key(g, P));
        return;
    }
    SPH_EMIT_KEY(grid_key(grid_cell(pr.x, pr.y, pr.z, P), P));
    #undef SPH_EMIT_KEY
}

// slab mode: keys of the ghost rows (predicted positions received from / kept for the neighbour ranks)
__global__ void __launch_bounds__(256)
k_ghost_key(const float4* __restrict__ ghost_pred, uint32_t* __restrict__ key, const uint32_t rows, const DevParams P,
            uint32_t* __restrict__ count, uint32_t* __restrict__ rank)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= rows) return;
    const float4 q = ghost_pred[s];
    const uint32_t k = grid_key(grid_cell(q.x, q.y, q.z, P), P);
    key[s] = k;
    if (count) { rank[s] = atomicAdd(&count[k], 1u); count_segment(count, P, k); }
}

// counting sort, placement: slot of row s = first slot of its cell + its ticket
__global__ void __launch_bounds__(256)
k_place(const uint32_t* __restrict__ key, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ table,
        uint32_t* __restrict__ slot_row, const uint32_t n, const DevParams P)
{
    chain_prologue();
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) slot_row[tbl(table, P, key[s]) + rank[s]] = s;
}

// ---- two-level GRID table (counting sort) -----------------------------------------------------------------------
// Layout of the allocation: [cells: ncell + 3, padded][segment bases, padded for the scan][rows per segment, padded].
// The counting kernels count into the third array; the scan turns it into the second (out of place), so the in-segment
// scan -- which only needs to know WHICH segments hold rows -- does not depend on the segment scan and the two run side
// by side in the recorded step (run_step).  A nonzero count also marks the segment as one the next step has to clear.
__global__ void __launch_bounds__(256)
k_table_clear(uint32_t* __restrict__ table, const uint32_t seg_off, const uint32_t dirty_off, const uint32_t nseg,
              const uint32_t nseg_pad, uint32_t* __restrict__ extra_a, const uint32_t words_a, uint32_t* __restrict__ extra_b,
              const uint32_t words_b)
{   // zero the cell counters of the segments the LAST step touched (row count != 0), and their row counts.  One THREAD looks at
    // one segment's flag (coalesced); the warp then zeroes its dirty segments together, 64 cells at a time.
    chain_prologue();
    const uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool in = sg < nseg;
    // slab mode: the step's small counters ride along (two memset nodes less on the solver's stream)
    for (uint32_t i = sg; i < words_a; i += gridDim.x * blockDim.x) extra_a[i] = 0u;
    for (uint32_t i = sg; i < words_b; i += gridDim.x * blockDim.x) extra_b[i] = 0u;
    const bool dirty = in && table[dirty_off + sg] != 0u;
    if (sg < nseg_pad && (dirty || !in)) table[dirty_off + sg] = 0u;      // (the bases are overwritten by the scan, padding included)
    uint32_t todo = __ballot_sync(0xffffffffu, dirty);
    const uint32_t sg0 = sg - lane;
    while (todo) {
        const uint32_t k = (uint32_t)__ffs(todo) - 1u;
        todo &= todo - 1u;
        reinterpret_cast<uint2*>(table + ((size_t)(sg0 + k) << kSegShift))[lane] = make_uint2(0u, 0u);
    }
}

__global__ void __launch_bounds__(256)
k_inseg_scan(uint32_t* __restrict__ table, const uint32_t seg_off, const uint32_t dirty_off, const uint32_t nseg)
{   // occupied segments: cell counters -> exclusive prefix inside the segment (the base comes from the segment scan).
    // One thread decides for one segment; the warp then scans its occupied segments together.
    chain_prologue();
    const uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool occ = sg < nseg && table[dirty_off + sg] != 0u;                     // empty: its cells are, and stay, zero
    uint32_t todo = __ballot_sync(0xffffffffu, occ);
    const uint32_t sg0 = sg - lane;
    // four occupied segments per round: their loads are issued together (one round trip to the L2 for all four)
    while (todo) {
        uint2* cells[4];
        uint2 v[4];
        #pragma unroll
        for (int u = 0; u < 4; u++) {
            cells[u] = nullptr;
            if (todo) {
                const uint32_t k = (uint32_t)__ffs(todo) - 1u;
                todo &= todo - 1u;
                cells[u] = reinterpret_cast<uint2*>(table + ((size_t)(sg0 + k) << kSegShift));
            }
        }
        #pragma unroll
        for (int u = 0; u < 4; u++) v[u] = cells[u] ? cells[u][lane] : make_uint2(0u, 0u);
        #pragma unroll
        for (int u = 0; u < 4; u++) {
            if (!cells[u]) continue;                                     // warp-uniform
            uint32_t inc = v[u].x + v[u].y;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t w = __shfl_up_sync(0xffffffffu, inc, o); if ((int)lane >= o) inc += w; }
            const uint32_t ex = inc - (v[u].x + v[u].y);
            cells[u][lane] = make_uint2(ex, ex + v[u].x);
        }
    }
}

__global__ void __launch_bounds__(256)
k_table_flatten(const uint32_t* __restrict__ table, uint32_t* __restrict__ flat, const uint32_t entries, const DevParams P)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < entries) flat[c] = tbl(table, P, c);
}

// ---- S2 tables -------------------------------------------------------------
// REFERENCE_HASH: startIndices[key] = first sorted row of the bucket (:486-496); we also keep the
// bucket end so the walk needs no per-row key compare (:343).
__global__ void __launch_bounds__(256)
k_fill_u32(uint32_t* __restrict__ p, const uint32_t v, const uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256)
k_build_table_hash(const uint32_t* __restrict__ key_sorted, uint32_t* __restrict__ tstart,
                   uint32_t* __restrict__ tend, const uint32_t n)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t k = key_sorted[s];
    if (s == 0 || key_sorted[s - 1] != k) tstart[k] = s;
    if (s == n - 1 || key_sorted[s + 1] != k) tend[k] = s + 1;
}

// GRID: prefix table t[c] = number of rows with key < c, for c in [0, ncell]; rows of cell c are
// [t[c], t[c+1]).  Built straight from the sorted keys: the row s at which the key steps from a to b
// owns the entries (a, b]; a warp fills its gaps cooperatively so the table is written exactly once,
// coalesced, with no memset and no scan.  Gaps wider than kBigGap go to a worklist for k_fill_gaps.
constexpr uint32_t kBigGap = 1024;
constexpr uint32_t kGapSegment = 16384;

__global__ void __launch_bounds__(256)
k_build_table_grid(const uint32_t* __restrict__ key_sorted, uint32_t* __restrict__ table,
                   uint32_t* __restrict__ gap_list, const uint32_t n, const uint32_t ncell)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;   // s in [0, n]: s == n closes the table
    const int lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = 0;                                     // this row owns table[lo..hi)
    if (s <= n) {
        const uint32_t b = (s < n) ? key_sorted[s] : ncell;      // virtual key ncell after the last row
        const uint32_t a1 = (s == 0) ? 0u : key_sorted[s - 1] + 1u;
        lo = a1; hi = b + 1u;
        if (s == n) hi = ncell + 1u;
        if (s < n && s > 0 && key_sorted[s - 1] == b) hi = lo;   // same cell as the previous row: nothing
    }
    uint32_t len = hi > lo ? hi - lo : 0u;
    if (len > kBigGap) {
        // wide gaps (air above the fluid, empty slabs) go to a worklist in kGapSegment pieces, one block each
        for (uint32_t o = lo; o < hi; o += kGapSegment) {
            const uint32_t slot = atomicAdd(&gap_list[0], 1u);
            gap_list[1 + 3 * slot] = o; gap_list[2 + 3 * slot] = min(o + kGapSegment, hi); gap_list[3 + 3 * slot] = s;
        }
        len = 0;
    } else if (len <= 8) {
        // the common case inside the fluid (the next occupied fine cell is a few entries away): write it here
        #pragma unroll
        for (uint32_t t = 0; t < 8; t++)
            if (t < len) table[lo + t] = s;
        len = 0;
    }
    uint32_t todo = __ballot_sync(0xffffffffu, len > 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t glo = __shfl_sync(0xffffffffu, lo, src);
        const uint32_t glen = __shfl_sync(0xffffffffu, len, src);
        const uint32_t gs = __shfl_sync(0xffffffffu, s, src);
        for (uint32_t i = lane; i < glen; i += 32) table[glo + i] = gs;
    }
}

__global__ void __launch_bounds__(256)
k_fill_gaps(uint32_t* __restrict__ table, const uint32_t* __restrict__ gap_list)
{
    const uint32_t ngaps = gap_list[0];
    for (uint32_t g = blockIdx.x; g < ngaps; g += gridDim.x) {       // one block per worklist entry
        const uint32_t lo = gap_list[1 + 3 * g], hi = gap_list[2 + 3 * g], s = gap_list[3 + 3 * g];
        for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) table[i] = s;
    }
}

// ---- reorder ----------------------------------------------------------------
// Source row of sorted slot s.  Radix path: perm[s].  Counting-sort path (key != nullptr): the tickets handed out by
// atomicAdd put a cell's rows into its slots in arbitrary order; the canonical (= stable sort) order is ascending
// source row, so slot number r of the cell takes the r-th smallest source row of the cell's slots.  Cells hold a
// handful of rows; a cell with more than kMaxCanonical rows (clamped border cells in a blow-up) keeps ticket order.
constexpr uint32_t kSmallCell = 32;         // up to here: rank every entry (m^2 compares per slot, m ~ 4)
constexpr uint32_t kMaxCanonical = 16384;   // up to here: bisection on the value (32 m compares per slot); beyond: ticket order

// Slot number r of a cell takes the r-th smallest source row of the cell's slots (the stable sort's order).  Cells hold a
// handful of rows, where ranking every entry is cheapest; crowded cells (a dense column, rows piled into a clamped rim
// cell) bisect on the value instead: the r-th smallest of m distinct values is the smallest x with #{v <= x} = r + 1.
// A cell above kMaxCanonical rows -- a blow-up that clamps a large part of the scene into one rim cell -- keeps its
// ticket order; each such cell is counted in *noncanonical (sph_noncanonical_cells): its neighbour sets are unaffected
// (the cull is by distance), only the summation order of floats, and with it the last bits, stop being reproducible.
__device__ __forceinline__ uint32_t source_row(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ key,
                                               const uint32_t* __restrict__ table, const uint32_t s, uint32_t* key_sorted,
                                               const DevParams& P, uint32_t* __restrict__ noncanonical)
{
    const uint32_t ncell = P.ncell;
    const uint32_t s0 = perm[s];
    if (!key) return s0;
    const uint32_t k = key[s0];
    if (key_sorted) key_sorted[s] = k;
    // slab mode: key == ncell collects the rows that leave this rank; they are dropped after the step, their order
    // is irrelevant -- and there can be hundreds of them
    if (k >= ncell) return s0;
    const uint32_t b = tbl(table, P, k), m = tbl(table, P, k + 1u) - b;
    if (m == 1u) return s0;
    const uint32_t r = s - b;
    if (m > kMaxCanonical) {
        if (r == 0u && noncanonical) atomicAdd(noncanonical, 1u);
        return s0;
    }
    if (m <= kSmallCell) {
        for (uint32_t t = 0; t < m; t++) {                   // the entry with exactly r smaller entries
            const uint32_t v = perm[b + t];
            uint32_t less = 0;
            for (uint32_t u = 0; u < m; u++) less += perm[b + u] < v;
            if (less == r) return v;
        }
        return s0;
    }
    uint32_t lo = 0u, hi = 0xFFFFFFFFu;                      // smallest x with #{v <= x} >= r + 1
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t le = 0;
        for (uint32_t u = 0; u < m; u++) le += perm[b + u] <= mid;
        if (le >= r + 1u) hi = mid; else lo = mid + 1u;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_reorder(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ key, const uint32_t* __restrict__ table,
          uint32_t* __restrict__ key_sorted, const float4* __restrict__ pos, const float4* __restrict__ vel,
          const float4* __restrict__ ghost_pred, float4* __restrict__ pos_s, float4* __restrict__ vel_s,
          float4* __restrict__ pred_s, float4* __restrict__ pred_pk, const DevParams P, const float dt)
{
    chain_prologue();
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < P.n;                    // no early return: the pair-interleaved copy is written with shuffles
    float4 q = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (valid) {
        const uint32_t src = source_row(perm, key, table, s, key_sorted, P, P.noncanonical);
        if (src >= P.n_a) {                 // slab mode ghost row: only its predicted position exists here
            const float4 gq = ghost_pred[src - P.n_a];
            const int3 c = cell_of(gq.x, gq.y, gq.z, P.r);
            q = make_float4(gq.x, gq.y, gq.z, __uint2float_rn(hash_cell(c.x, c.y, c.z)));
            pos_s[s] = make_float4(gq.x, gq.y, gq.z, __uint_as_float(0xFFFFFFFFu));
            vel_s[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        } else {
            const float4 p = pos[src];
            float4 v = vel[src];
            float3 pr;
            predict(p, v, pr, P, dt);          // bit-identical to k_predict_key: same inputs, same operations
            const int3 c = cell_of(pr.x, pr.y, pr.z, P.r);
            const uint32_t h = hash_cell(c.x, c.y, c.z);
            pos_s[s] = p;
            vel_s[s] = v;
            q = make_float4(pr.x, pr.y, pr.z, __uint2float_rn(h));   // spatialLookup[].y is float(hash) (:480)
        }
        pred_s[s] = q;
    }
    // candidate pair m = rows (2m, 2m+1): the even row's thread writes xy[m] = (x0, x1, y0, y1), the odd row's thread
    // z[m] = (z0, z1) (sph_internal.h).  The partner of the last row of an odd-sized array contributes zeros (that slot
    // is past every window, so it is never accepted).
    const float ax = __shfl_xor_sync(0xffffffffu, q.x, 1), ay = __shfl_xor_sync(0xffffffffu, q.y, 1);
    const float az = __shfl_xor_sync(0xffffffffu, q.z, 1);
    if (pred_pk && s < ((P.n + 1u) & ~1u)) {
        if (s & 1u) reinterpret_cast<float2*>(pred_pk + P.pair_cap)[s >> 1] = make_float2(az, q.z);
        else pred_pk[s >> 1] = make_float4(q.x, ax, q.y, ay);
    }
}

// S6 with the optional features of SphExtras: everything happens in the box's own axes (local = R^T * world).
//  1. stickiness: a particle within stick_d of a wall is pulled towards it, v -= dt * k * d * (1 - d / stick_d) * n
//     (n the wall's inward normal, d the distance to the wall);
//  2. the reference's move, clamp and -0.95 reflection (:84-107) on the local coordinates;
//  3. back to world axes.
__device__ __forceinline__ void integrate_extras(float4& p, float4& v, const DevParams& P, const float dt)
{
    const float* R = P.rot;
    float l[3] = {R[0] * p.x + R[3] * p.y + R[6] * p.z, R[1] * p.x + R[4] * p.y + R[7] * p.z, R[2] * p.x + R[5] * p.y + R[8] * p.z};
    float u[3] = {R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z};
    if (P.stick_k > 0.0f) {
        #pragma unroll
        for (int a = 0; a < 3; a++) {
            const float dlo = l[a] + P.half[a], dhi = P.half[a] - l[a];          // distances to the two walls of this axis
            if (dlo >= 0.0f && dlo < P.stick_d) u[a] -= dt * P.stick_k * dlo * (1.0f - dlo / P.stick_d);     // inward normal +a
            if (dhi >= 0.0f && dhi < P.stick_d) u[a] += dt * P.stick_k * dhi * (1.0f - dhi / P.stick_d);     // inward normal -a
        }
    }
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        l[a] = __fadd_rn(l[a], __fmul_rn(u[a], dt));
        if (__fsub_rn(P.half[a], fabsf(l[a])) <= 0.0f) {
            const float sg = (l[a] > 0.0f) ? 1.0f : ((l[a] < 0.0f) ? -1.0f : 0.0f);
            l[a] = __fmul_rn(P.half[a], sg);
            u[a] = __fmul_rn(u[a], -0.95f);
        }
    }
    p.x = R[0] * l[0] + R[1] * l[1] + R[2] * l[2]; p.y = R[3] * l[0] + R[4] * l[1] + R[5] * l[2]; p.z = R[6] * l[0] + R[7] * l[1] + R[8] * l[2];
    v.x = R[0] * u[0] + R[1] * u[1] + R[2] * u[2]; v.y = R[3] * u[0] + R[4] * u[1] + R[5] * u[2]; v.z = R[6] * u[0] + R[7] * u[1] + R[8] * u[2];
}

// S6 (:84-107)
__global__ void __launch_bounds__(256)
k_integrate(const float4* __restrict__ pos_s, const float4* __restrict__ vel_v, float4* __restrict__ pos_out,
            float4* __restrict__ vel_out, const DevParams P, const float dt)
{
    chain_prologue();
    const uint32_t s = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.row1) return;
    float4 p = pos_s[s];
    float4 v = vel_v[s];
    p.x = __fadd_rn(p.x, __fmul_rn(v.x, dt));
    p.y = __fadd_rn(p.y, __fmul_rn(v.y, dt));
    p.z = __fadd_rn(p.z, __fmul_rn(v.z, dt));
    const float damp = -0.95f;                       // -1 * dampFactor
    #define SPH_COLLIDE(c, h)                                                             \
        if (__fsub_rn(h, fabsf(p.c)) <= 0.0f) {                                           \
            const float sg = (p.c > 0.0f) ? 1.0f : ((p.c < 0.0f) ? -1.0f : 0.0f);         \
            p.c = __fmul_rn(h, sg);                                                       \
            v.c = __fmul_rn(v.c, damp);                                                   \
        }
    SPH_COLLIDE(x, P.half[0])
    SPH_COLLIDE(y, P.half[1])
    SPH_COLLIDE(z, P.half[2])
    #undef SPH_COLLIDE
    v.w = 0.0f;
    pos_out[s - P.row0] = p;      // .w still carries the particle id; owned rows compact to [0, row1-row0)
    vel_out[s - P.row0] = v;
}

// S6 with SphExtras on (launched instead of k_integrate, so the reference's configuration pays nothing for them)
__global__ void __launch_bounds__(256)
k_integrate_extras(const float4* __restrict__ pos_s, const float4* __restrict__ vel_v, float4* __restrict__ pos_out,
                   float4* __restrict__ vel_out, const DevParams P, const float dt)
{
    chain_prologue();
    const uint32_t s = P.row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.row1) return;
    float4 p = pos_s[s];
    float4 v = vel_v[s];
    integrate_extras(p, v, P, dt);
    v.w = 0.0f;
    pos_out[s - P.row0] = p;
    vel_out[s - P.row0] = v;
}

// ---- upload / export --------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_state(const float* __restrict__ pos3, const float* __restrict__ vel3, const uint32_t* __restrict__ ids,
             float4* __restrict__ pos, float4* __restrict__ vel, const uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = ids ? ids[i] : i;
    pos[i] = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], __uint_as_float(id));
    vel[i] = vel3 ? make_float4(vel3[3 * i], vel3[3 * i + 1], vel3[3 * i + 2], 0.0f) : make_float4(0, 0, 0, 0);
}

// ---- device-side scene spawn ---------------------------------------------------------------
// Jittered lattice block written straight into the state arrays: no host array, no upload.  The arithmetic is the
// synthetic-scene rule of SURVEY 8(d) (fluid-simulation-3d_b200/scenes.py: block) operation for operation -- lattice
// site in fp64 then rounded to fp32, jitter / velocity from the counter-based generator splitmix64(seed ^ (3*id +
// axis)) in fp32 -- so the device scene is bit-identical to the host one (tests/test_spawn_gpu.py).  Particle id <->
// lattice site follows the reference's fill order (GridArrangement, physicsWorld.cc:526-530): y outer from the top
// layer down, then x, then z.
__device__ __forceinline__ float spawn_uniform01(const uint64_t seed, const uint64_t id, const uint32_t axis)
{
    uint64_t x = (seed ^ (id * 3ull + axis)) + 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return __fmul_rn((float)(uint32_t)(x >> 40), 1.0f / 16777216.0f);   // top 24 bits -> [0, 1), exact
}

__global__ void __launch_bounds__(256)
k_spawn_block(float4* __restrict__ pos, float4* __restrict__ vel, const uint32_t nx, const uint32_t ny, const uint32_t nz,
              const double gap, const double lox, const double loy, const double loz, const float jitter_amp,
              const float vel_scale, const uint64_t seed, const uint32_t n)
{
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const uint32_t iz = id % nz, ix = (id / nz) % nx, iy = id / (nz * nx);
    float p[3];
    p[0] = (float)__dadd_rn(lox, __dmul_rn((double)ix, gap));
    p[1] = (float)__dadd_rn(loy, __dmul_rn((double)(ny - 1u - iy), gap));     // iy counts from the top layer down
    p[2] = (float)__dadd_rn(loz, __dmul_rn((double)iz, gap));
    float v[3] = {0.0f, 0.0f, 0.0f};
    #pragma unroll
    for (uint32_t a = 0; a < 3; a++) {
        if (jitter_amp != 0.0f) p[a] = __fadd_rn(p[a], __fmul_rn(__fsub_rn(spawn_uniform01(seed, id, a), 0.5f), jitter_amp));
        if (vel_scale != 0.0f) v[a] = __fmul_rn(__fsub_rn(spawn_uniform01(seed ^ 0x5EEDull, id, a), 0.5f), vel_scale);
    }
    pos[id] = make_float4(p[0], p[1], p[2], __uint_as_float(id));
    vel[id] = make_float4(v[0], v[1], v[2], 0.0f);
}

__device__ __forceinline__ float4 speed_color(float t)
{   // FluidSimCPU::updateColors (fluidSimCPU.cc:100-125)
    const float4 c1 = make_float4(0.0f, 0.75f, 1.0f, 1.0f), c2 = make_float4(0.0f, 1.0f, 0.0f, 1.0f);
    const float4 c3 = make_float4(1.0f, 1.0f, 0.0f, 1.0f), c4 = make_float4(1.0f, 0.0f, 0.0f, 1.0f);
    const float b1 = 0.33f, b2 = 0.66f;
    float a; float4 lo, hi;
    // the reference's operations one by one, unfused (oracle/colors.py restates them in numpy; bit-exact against it)
    if (t <= b1) { a = __fdiv_rn(t, b1); lo = c1; hi = c2; }
    else if (t <= b2) { a = __fdiv_rn(__fsub_rn(t, b1), __fsub_rn(b2, b1)); lo = c2; hi = c3; }
    else { a = __fdiv_rn(__fsub_rn(t, b2), __fsub_rn(1.0f, b2)); lo = c3; hi = c4; }
    const float ia = __fsub_rn(1.0f, a);
    return make_float4(__fadd_rn(__fmul_rn(ia, lo.x), __fmul_rn(a, hi.x)), __fadd_rn(__fmul_rn(ia, lo.y), __fmul_rn(a, hi.y)),
                       __fadd_rn(__fmul_rn(ia, lo.z), __fmul_rn(a, hi.z)), __fadd_rn(__fmul_rn(ia, lo.w), __fmul_rn(a, hi.w)));
}

// out[dst] = field of device row s, dst = particle id (by_id) or s.
__global__ void __launch_bounds__(256)
k_export(const int field, const float4* __restrict__ id_src, const void* __restrict__ src,
         void* __restrict__ out, const uint32_t n, const bool by_id, const DevParams P)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t dst = by_id ? __float_as_uint(id_src[s].w) : s;
    switch (field) {
    case SPH_FIELD_POSITIONS: case SPH_FIELD_VELOCITIES: case SPH_FIELD_PREDICTED: case SPH_FIELD_VEL_AFTER_PRESSURE:
    case SPH_FIELD_VEL_AFTER_VISCOSITY: {
        const float4 a = ((const float4*)src)[s];
        float* o = (float*)out + 3 * (size_t)dst;
        o[0] = a.x; o[1] = a.y; o[2] = a.z;
    } break;
    case SPH_FIELD_OUT_POSITIONS: {
        const float4 a = ((const float4*)src)[s];
        ((float4*)out)[dst] = make_float4(a.x, a.y, a.z, 0.34f);          // :107
    } break;
    case SPH_FIELD_DENSITIES:
        { const Rec8 d = ((const Rec8*)src)[s]; ((float2*)out)[dst] = make_float2(d.lo.w, d.hi.x); }   // density record
        break;
    case SPH_FIELD_HASH: case SPH_FIELD_KEY: {       // pure functions of the predicted position (:477-479)
        const float4 a = ((const float4*)src)[s];
        const int3 c = cell_of(a.x, a.y, a.z, P.r);
        const uint32_t h = hash_cell(c.x, c.y, c.z);
        ((uint32_t*)out)[dst] = (field == SPH_FIELD_HASH) ? h : key_of_hash(h, P);
    } break;
    case SPH_FIELD_NEIGHBOUR_COUNT:
        ((uint32_t*)out)[dst] = ((const uint32_t*)src)[s];
        break;
    case SPH_FIELD_SPEED_NORMALIZED: case SPH_FIELD_COLORS: {
        const float4 a = ((const float4*)src)[s];
        // glm::length = sqrt(dot): (x*x + y*y) + z*z, unfused; clamp; / 1.5f                       (:181)
        const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z)));
        const float t = __fdiv_rn(fminf(fmaxf(len, 0.0f), 1.5f), 1.5f);
        if (field == SPH_FIELD_SPEED_NORMALIZED) ((float*)out)[dst] = t;
        else ((float4*)out)[dst] = speed_color(t);
    } break;
    default: break;
    }
}

// getPosition/getVelocity/getDensity/getNearDensity/getSpeed/getSpeedNormalzied (:149-182) for one id
__global__ void __launch_bounds__(256)
k_row_of_id(const float4* __restrict__ pos, uint32_t* __restrict__ row_of, const uint32_t n)
{   // inverse of the device order: row_of[id] = row (ids are 0..n-1 on a single-GPU context)
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t id = __float_as_uint(pos[s].w);
    if (id < n) row_of[id] = s;
}

__global__ void k_read_particle(const float4* __restrict__ pos, const float4* __restrict__ vel, const Rec8* __restrict__ dens,
                                const uint32_t* __restrict__ row_of, const uint32_t id, float* __restrict__ out10)
{
    const uint32_t s = row_of[id];
    const float4 p = pos[s];
    const float4 v = vel[s];
    const float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    out10[0] = p.x; out10[1] = p.y; out10[2] = p.z; out10[3] = v.x; out10[4] = v.y; out10[5] = v.z;
    if (dens) { const Rec8 d = dens[s]; out10[6] = d.lo.w; out10[7] = d.hi.x; }
    out10[8] = len;
    out10[9] = fminf(fmaxf(len, 0.0f), 1.5f) / 1.5f;
}

__global__ void __launch_bounds__(256)
k_export_ids(const float4* __restrict__ id_src, uint32_t* __restrict__ out, const uint32_t n)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out[s] = __float_as_uint(id_src[s].w);
}

inline uint32_t blocks_for(uint32_t n, int threads) { return (n + threads - 1) / threads; }

}  // namespace

// ---- launchers ---------------------------------------------------------------
void launch_predict_key(cudaStream_t st, const float4* pos, const float4* vel, uint32_t* key, uint8_t* cls,
                        uint32_t rows, bool may_migrate, const DevParams& P, float dt, uint32_t* count, uint32_t* rank,
                        uint64_t* launches, uint32_t* pack_counts, uint32_t pack_blocks)
{
    if (rows == 0) return;
    launch_chained(k_predict_key, dim3(blocks_for(rows, 256)), dim3(256), 0, st, pos, vel, key, cls, rows, may_migrate, P, dt, count, rank,
                   pack_counts, pack_blocks);
    ++*launches;
}

void launch_ghost_key(cudaStream_t st, const float4* ghost_pred, uint32_t* key, uint32_t rows, const DevParams& P,
                      uint32_t* count, uint32_t* rank, uint64_t* launches)
{
    if (rows == 0) return;
    k_ghost_key<<<blocks_for(rows, 256), 256, 0, st>>>(ghost_pred, key, rows, P, count, rank);
    ++*launches;
}

void launch_build_table(cudaStream_t st, const uint32_t* key_sorted, uint32_t* tstart, uint32_t* tend,
                        uint32_t* gap_list, const DevParams& P, uint64_t* launches)
{
    if (P.mode == SPH_TABLE_REFERENCE_HASH) {
        if (P.n) { k_fill_u32<<<blocks_for(P.n, 256), 256, 0, st>>>(tstart, 0x7FFFFFFFu, P.n); ++*launches; }   // INT_MAX = empty (:481)
        if (P.n) { k_build_table_hash<<<blocks_for(P.n, 256), 256, 0, st>>>(key_sorted, tstart, tend, P.n); ++*launches; }
    } else {
        cudaMemsetAsync(gap_list, 0, sizeof(uint32_t), st);
        const TableLayout T = table_layout(P.ncell);     // a flat table: every segment base is zero (tbl(), sph_device.cuh)
        cudaMemsetAsync(tstart + P.seg_off, 0, T.nseg_pad * sizeof(uint32_t), st);
        k_build_table_grid<<<blocks_for(P.n + 1, 256), 256, 0, st>>>(key_sorted, tstart, gap_list, P.n, P.ncell);
        k_fill_gaps<<<148 * 4, 256, 0, st>>>(tstart, gap_list);
        *launches += 2;
    }
}

void launch_place(cudaStream_t st, const uint32_t* key, const uint32_t* rank, const uint32_t* table, uint32_t* slot_row,
                  uint32_t n, const DevParams& P, uint64_t* launches)
{
    if (n == 0) return;
    launch_chained(k_place, dim3(blocks_for(n, 256)), dim3(256), 0, st, key, rank, table, slot_row, n, P);
    ++*launches;
}

TableLayout table_layout(const uint32_t ncell)
{
    TableLayout T;
    T.cells_pad = scan_pad((size_t)ncell + 3);            // a multiple of 4096, hence of the segment size
    T.nseg = T.cells_pad >> kSegShift;
    T.nseg_pad = scan_pad(T.nseg + 1);                    // the exclusive scan runs in place over zero-padded entries
    T.total = T.cells_pad + 2 * T.nseg_pad;
    return T;
}

void launch_table_clear(cudaStream_t st, uint32_t* table, const TableLayout& T, uint64_t* launches, uint32_t* extra_a, uint32_t words_a,
                        uint32_t* extra_b, uint32_t words_b)
{
    launch_chained(k_table_clear, dim3(blocks_for((uint32_t)T.nseg_pad, 256)), dim3(256), 0, st, table, (uint32_t)T.cells_pad, (uint32_t)(T.cells_pad + T.nseg_pad), (uint32_t)T.nseg,
                   (uint32_t)T.nseg_pad, extra_a, words_a, extra_b, words_b);
    ++*launches;
}

void launch_inseg_scan(cudaStream_t st, uint32_t* table, const TableLayout& T, uint64_t* launches)
{
    launch_chained(k_inseg_scan, dim3(blocks_for((uint32_t)T.nseg, 256)), dim3(256), 0, st, table, (uint32_t)T.cells_pad, (uint32_t)(T.cells_pad + T.nseg_pad), (uint32_t)T.nseg);
    ++*launches;
}

void launch_table_flatten(cudaStream_t st, const uint32_t* table, const DevParams& P, uint32_t* flat, uint32_t entries, uint64_t* launches)
{
    if (entries == 0) return;
    k_table_flatten<<<blocks_for(entries, 256), 256, 0, st>>>(table, flat, entries, P);
    ++*launches;
}

void launch_reorder(cudaStream_t st, const uint32_t* perm, const uint32_t* key, const uint32_t* table, uint32_t* key_sorted,
                    const float4* pos, const float4* vel, const float4* ghost_pred, float4* pos_s, float4* vel_s,
                    float4* pred_s, float4* pred_pk, const DevParams& P, float dt, uint64_t* launches)
{
    if (P.n == 0) return;
    launch_chained(k_reorder, dim3(blocks_for(P.n, 256)), dim3(256), 0, st, perm, key, table, key_sorted, pos, vel, ghost_pred, pos_s, vel_s, pred_s, pred_pk, P, dt);
    ++*launches;
}

void launch_integrate(cudaStream_t st, const float4* pos_s, const float4* vel_v, float4* pos_out, float4* vel_out,
                      const DevParams& P, float dt, uint64_t* launches)
{
    if (P.row1 <= P.row0) return;
    if (P.extras) launch_chained(k_integrate_extras, dim3(blocks_for(P.row1 - P.row0, 256)), dim3(256), 0, st, pos_s, vel_v, pos_out, vel_out, P, dt);
    else launch_chained(k_integrate, dim3(blocks_for(P.row1 - P.row0, 256)), dim3(256), 0, st, pos_s, vel_v, pos_out, vel_out, P, dt);
    ++*launches;
}

void launch_spawn_block(cudaStream_t st, float4* pos, float4* vel, uint32_t nx, uint32_t ny, uint32_t nz, double gap,
                        const double lo[3], float jitter_amp, float vel_scale, uint64_t seed, uint32_t n, uint64_t* launches)
{
    if (n == 0) return;
    k_spawn_block<<<blocks_for(n, 256), 256, 0, st>>>(pos, vel, nx, ny, nz, gap, lo[0], lo[1], lo[2], jitter_amp, vel_scale, seed, n);
    ++*launches;
}

void launch_pack_state(cudaStream_t st, const float* pos3, const float* vel3, const uint32_t* ids,
                       float4* pos, float4* vel, uint32_t n, uint64_t* launches)
{
    if (n == 0) return;
    k_pack_state<<<blocks_for(n, 256), 256, 0, st>>>(pos3, vel3, ids, pos, vel, n);
    ++*launches;
}

void launch_export(cudaStream_t st, int field, const float4* id_src, const void* src, const void*, void* out,
                   uint32_t n, const DevParams& P, bool by_id, uint64_t* launches)
{
    if (n == 0) return;
    k_export<<<blocks_for(n, 256), 256, 0, st>>>(field, id_src, src, out, n, by_id, P);
    ++*launches;
}

void launch_find_particle(cudaStream_t st, const float4* pos, const float4* vel, const Rec8* dens, uint32_t n,
                          uint32_t id, float* out10, uint32_t* row_of, bool rebuild, uint64_t* launches)
{   // one particle by id: the id -> row map is rebuilt once per device order (a step, an upload), then a read is O(1)
    if (n == 0) return;
    if (rebuild) { k_row_of_id<<<blocks_for(n, 256), 256, 0, st>>>(pos, row_of, n); ++*launches; }
    k_read_particle<<<1, 1, 0, st>>>(pos, vel, dens, row_of, id, out10);
    ++*launches;
}

void launch_export_ids(cudaStream_t st, const float4* id_src, uint32_t* out, uint32_t n, uint64_t* launches)
{
    if (n == 0) return;
    k_export_ids<<<blocks_for(n, 256), 256, 0, st>>>(id_src, out, n);
    ++*launches;
}

}  // namespace sphb200
