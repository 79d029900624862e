// sph_internal.h -- shared declarations of the B200 SPH step (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sph_b200.h"

namespace sphb200 {

// Everything a kernel needs, passed by value (lives in the constant bank).
// Normalisation constants are computed on the HOST with the reference's own fp32
// expressions (kernels.h:29,41,53,65,77; SURVEY App.A Q17).
struct DevParams {
    uint32_t n;             // particles on this device (owned + ghosts in slab mode)
    uint32_t n_owned;       // slab mode: owned particles are rows [0, n_owned) BEFORE sorting; == n otherwise
    int      mode;          // SphTableMode
    int      gravity;
    float    r, sqr_r, rho0, k, kn, mu, g;
    float    half[3];       // BoundScale * 0.5f (physicsWorld.cc:88)
    float    vol2, vol3;    // SmoothingPow2 / Pow3 volumes
    float    s2, s3;        // SmoothingDerivativePow2 / Pow3 scales
    float    sv;            // SmoothingViscoPoly6 scale
    float    rr;            // r*r
    // The packed cull computes d^2 with FMAs (3 roundings) where the reference's (x*x + y*y) + z*z has 5: the two
    // differ by < 8 * 2^-24 = 4.8e-7 relative.  d^2' > cull_hi: certainly not a neighbour; d^2' < cull_lo: certainly
    // one; in between (a ~2e-6 wide band) the exact predicate is evaluated on the plain rows.
    float    cull_hi;       // sqr_r * (1 + 1e-6) rounded up
    float    cull_lo;       // sqr_r * (1 - 1e-6) rounded down
    // GRID table: cell g = clamp(floor(pred/r) - gmin, 0, gdim-1); key = (gz*gdim.y + gy)*gdim.x + gx
    int      gmin[3];       // first reference cell of the table per axis
    int      gdim[3];       // table extent: fine x cells (reference cells * xsub), y cells, z cells
    int      xsub;          // subdivision of the cell in x (power of two)
    float    xwin;          // half width of the x window: sqrt(sqr_r) * (1 + 1e-5)
    uint32_t ncell;
    // REFERENCE_HASH table: key = hash % n via Lemire fastmod, M = 2^64 / n + 1
    uint64_t modM;
    // rows the gather / integrate kernels process: [row0, row1)  (single GPU: [0, n))
    uint32_t row0, row1;
    // slab mode (multi-GPU): the local table holds global z layers [zlo, zlo + gdim[2]); this rank
    // owns global layers [own_lo, own_hi).  Single GPU: slab = 0, zlo = 0, gz_global = gdim[2].
    int      slab;
    int      zlo, gz_global, own_lo, own_hi;
    int      has_lo, has_hi;   // a neighbour rank exists below / above
    uint32_t n_a;              // reorder: perm values < n_a index the state arrays, the rest the ghost buffer
    uint32_t pair_cap;         // candidate pairs: float4 entries of the (x0,x1,y0,y1) array; the (z0,z1) array starts right behind it
    uint32_t cnt_off;          // GRID table: word offset of the per-segment row COUNTS (count_segment; nonzero = the segment is occupied / dirty)
    uint32_t seg_off;          // GRID table: word offset of the per-segment base array behind the per-cell array (tbl(), sph_device.cuh)
    uint32_t* noncanonical;    // device counter: cells too crowded for the canonical-order ranking of the counting sort (sph_kernels.cu)
    // optional features (SphExtras): box rotation (rows of R: world = R * local) and wall stickiness
    int      extras;           // 0: S6 is the reference's, bit for bit
    float    rot[9];
    float    stick_k, stick_d;
    int      rim_check;        // the cut-off exceeds the cell size (Q2): particles next to the table's rim compare true cells (sph_device.cuh)
};

// 32-byte per-particle record read with ONE 256-bit load (LDG.E.256 on sm_100a) by the pressure pass:
//   density record  lo = (pred.x, pred.y, pred.z, rho)   hi = (near rho, 1/rho, 1/near rho, 0)
// Position and per-pass payload of a neighbour sit in the same sector, so a neighbour costs one load
// instruction and one line instead of two of each.
struct __align__(32) Rec8 { float4 lo, hi; };
// pair records of padding behind `predpk`: a density-pass warp may read this far past the last row's pair without
// clamping its addresses (the slots are outside every window, so whatever they hold is rejected)
constexpr uint32_t kPairPad = 2048;
// GRID table, two levels: cells are grouped in segments of 64; t(c) = in-segment prefix[c] + segment base[c >> 6].
// Building it costs O(particles + occupied segments), not O(cells): the air above the fluid is never touched.
constexpr int kSegShift = 6;
constexpr uint32_t kSegCells = 1u << kSegShift;

// Pair-interleaved predicted positions (`predpk`), the density pass's candidate stream: rows 2m and 2m+1 form pair m,
//   xy[m] = (x0, x1, y0, y1)   at predpk[m]                          (16 bytes)
//   z[m]  = (z0, z1)           at ((float2*)(predpk + pair_cap))[m]   (8 bytes)
// so that a 128-bit and a 64-bit load hand the density pass two candidates as three aligned register pairs, ready for the
// packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2) -- 24 bytes per pair through the L1, not 32.  Written by k_reorder
// next to the plain `pred` rows.

// slab mode: classification of an owned row by the z layer of its predicted position
// (CLS_KEEP: a migrant that lands in the layer right across the plane -- it stays visible to this rank as a ghost)
enum : uint8_t { CLS_STAY = 0, CLS_MIG_LO = 1, CLS_MIG_HI = 2, CLS_GHOST_LO = 4, CLS_GHOST_HI = 8, CLS_KEEP = 16 };
// the six lists a slab rank packs for its neighbours; a pack block covers kPackSpan consecutive rows
enum { L_MIG_LO = 0, L_GHOST_LO = 1, L_KEEP_LO = 2, L_MIG_HI = 3, L_GHOST_HI = 4, L_KEEP_HI = 5, NLISTS = 6 };
constexpr int kPackThreads = 256;
constexpr int kPackIters = 16;
constexpr uint32_t kPackSpan = kPackIters * kPackThreads;

struct SortTemp {
    uint32_t* counts;       // [256][nblocks] digit counts -> exclusive offsets
    size_t    counts_len;
};

// ---- sph_sort.cu ----------------------------------------------------------
// Stable LSD radix sort of (key, value) pairs on `bits` low key bits.  vals_in == nullptr
// means "value = row index".  Returns 0 if the result is in (keys_a, vals_a), 1 if in (keys_b, vals_b).
size_t radix_sort_temp_entries(uint32_t n);
int radix_sort_pairs(cudaStream_t st, uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                     bool vals_identity, uint32_t n, int bits, uint32_t* counts, uint64_t* launches);

// ---- sph_kernels.cu -------------------------------------------------------
// count / rank non-null: counting sort of the GRID table (the row takes a ticket in its cell's counter)
void launch_predict_key(cudaStream_t st, const float4* pos, const float4* vel, uint32_t* key, uint8_t* cls,
                        uint32_t rows, bool may_migrate, const DevParams& P, float dt, uint32_t* count, uint32_t* rank,
                        uint64_t* launches, uint32_t* pack_counts = nullptr, uint32_t pack_blocks = 0);
void launch_ghost_key(cudaStream_t st, const float4* ghost_pred, uint32_t* key, uint32_t rows, const DevParams& P,
                      uint32_t* count, uint32_t* rank, uint64_t* launches);
void launch_place(cudaStream_t st, const uint32_t* key, const uint32_t* rank, const uint32_t* table, uint32_t* slot_row,
                  uint32_t n, const DevParams& P, uint64_t* launches);
// two-level GRID table (counting sort): clear what the last step touched, in-segment prefixes, flat copy for the taps
struct TableLayout { size_t cells_pad, nseg, nseg_pad, total; };   // [cells][segment bases][segment row counts], in words
TableLayout table_layout(uint32_t ncell);
void launch_table_clear(cudaStream_t st, uint32_t* table, const TableLayout& T, uint64_t* launches, uint32_t* extra_a = nullptr,
                        uint32_t words_a = 0, uint32_t* extra_b = nullptr, uint32_t words_b = 0);   // extra_*: words to zero along the way
void launch_inseg_scan(cudaStream_t st, uint32_t* table, const TableLayout& T, uint64_t* launches);
void launch_table_flatten(cudaStream_t st, const uint32_t* table, const DevParams& P, uint32_t* flat, uint32_t entries, uint64_t* launches);
// sph_sort.cu: in-place exclusive scan of a zero-padded array (multiple of 4096 entries)
size_t scan_pad(size_t entries);
size_t scan_temp_entries(size_t entries);
// (src == dst: in place)
void exclusive_scan_u32(cudaStream_t st, const uint32_t* src, uint32_t* dst, size_t padded_entries, uint32_t* blocksums, uint64_t* launches);
void launch_build_table(cudaStream_t st, const uint32_t* key_sorted, uint32_t* table_start, uint32_t* table_end,
                        uint32_t* gap_list, const DevParams& P, uint64_t* launches);
// key / table non-null: counting-sort path (perm = slot -> some row of the cell; canonical order restored here)
void launch_reorder(cudaStream_t st, const uint32_t* perm, const uint32_t* key, const uint32_t* table, uint32_t* key_sorted,
                    const float4* pos, const float4* vel, const float4* ghost_pred, float4* pos_s, float4* vel_s,
                    float4* pred_s, float4* pred_pk, const DevParams& P, float dt, uint64_t* launches);
// neighbour list recorded by the density pass (k-major: entry k of row i at idx[k*stride + i])
struct NbrList {
    uint32_t* idx;      // nullptr: no list, every pass walks the table
    float*    w;        // viscosity weight of every entry, same geometry as idx (filled by k_density_pk only)
    uint32_t* cnt;      // [rows] list length (may exceed k: overflowed; may include <= 1e-6 borderline extras)
    uint32_t* ncount;   // [rows] exact neighbour count incl. self (always written by the density pass)
    uint32_t* overflow; // host-mapped: largest list length that did not fit (0 = none)
    uint32_t* rows_sum; // [2] running sums: list rows written by the density pass, warps that wrote them
    bool      deep;     // density pass: deep survivor stack
    uint32_t  k;        // entries per row before a row counts as overflowed
    uint32_t  stride;
    // tile generation (sph_tile.cu): sorted GRID keys, staging capacity per warp, "did not fit" report
    const uint32_t* keys;
    uint32_t  capn;
    uint32_t* tile_need;
};
// Launch of a kernel of the step's chain: programmatic stream serialisation (PDL), so the kernel may be scheduled while
// its predecessor's last wave drains; the kernel itself orders its memory accesses with chain_prologue() (sph_device.cuh).
// Stream capture records the edge as a programmatic dependency.  SPH_PDL=0 launches them ordinarily.
bool chained_launches_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = chained_launches_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct GatherArgs;
bool tile_enabled();
uint32_t tile_default_capn();
int launch_tile(cudaStream_t st, int pass, const GatherArgs& A, const DevParams& P, uint32_t capn, float dt, uint64_t* launches);
void launch_density(cudaStream_t st, const float4* pred_s, const float4* pred_pk, const uint32_t* tstart, const uint32_t* tend,
                    Rec8* dens, const NbrList& L, const DevParams& P, uint64_t* launches);
void launch_pressure(cudaStream_t st, const float4* pred_s, const Rec8* dens, const float4* vel_s,
                     const uint32_t* tstart, const uint32_t* tend, float4* vel_p, const NbrList& L, const DevParams& P,
                     float dt, uint64_t* launches);
void launch_viscosity(cudaStream_t st, const float4* pred_s, const float4* vel_p, const uint32_t* tstart,
                      const uint32_t* tend, float4* vel_v, const NbrList& L, const DevParams& P, float dt,
                      uint64_t* launches);
void launch_integrate(cudaStream_t st, const float4* pos_s, const float4* vel_v, float4* pos_out, float4* vel_out,
                      const DevParams& P, float dt, uint64_t* launches);

// upload / export helpers (original particle index order <-> device order)
void launch_pack_state(cudaStream_t st, const float* pos3, const float* vel3, const uint32_t* ids,
                       float4* pos, float4* vel, uint32_t n, uint64_t* launches);
void launch_spawn_block(cudaStream_t st, float4* pos, float4* vel, uint32_t nx, uint32_t ny, uint32_t nz, double gap,
                        const double lo[3], float jitter_amp, float vel_scale, uint64_t seed, uint32_t n, uint64_t* launches);
void launch_export(cudaStream_t st, int field, const float4* id_src, const void* src, const void* src2, void* out,
                   uint32_t n, const DevParams& P, bool by_id, uint64_t* launches);
void launch_find_particle(cudaStream_t st, const float4* pos, const float4* vel, const Rec8* dens, uint32_t n,
                          uint32_t id, float* out10, uint32_t* row_of, bool rebuild, uint64_t* launches);
void launch_export_ids(cudaStream_t st, const float4* id_src, uint32_t* out, uint32_t n, uint64_t* launches);

}  // namespace sphb200
