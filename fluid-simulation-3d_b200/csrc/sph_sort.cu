// sph_sort.cu -- device radix sort of (cell key, row) pairs for the per-step neighbour table.
//
// Replaces the reference's serial std::sort over float rows (physicsWorld.cc:484) with a STABLE
// least-significant-digit radix sort, 8-bit digits, only as many passes as the key range needs
// (3 passes for up to 16 M cells / buckets).  Stable => equal keys keep their incoming order, which
// is the canonical tie order the parity tests compare against (SURVEY App.A Q14).
//
// Per pass: (1) per-block digit histogram, (2) per-digit row scan of the [digit][block]
// matrix (256 blocks), (3) scatter with a warp-synchronous stable rank (match.any + per-warp digit counters in
// shared memory).  Blocks own a contiguous span of tiles so the scanned matrix stays small
// (<= 256 x 1184 entries) at any N.  HBM traffic per pass: keys read twice, pairs written once
// = 20 B/pair (algorithmic 16 B/pair).
#include <cstdlib>
#include "sph_device.cuh"

namespace sphb200 {

namespace {
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 8;                          // keys per thread per tile
constexpr int kTile = kSortThreads * kItems;       // 2048 keys
constexpr uint32_t kMaxSortBlocks = 148 * 6;     // one resident wave of the scatter kernel (6 blocks per SM)

struct SortGeom { uint32_t nblocks, tiles_per_block; };

inline SortGeom sort_geom(uint32_t n)
{
    uint32_t tiles = (n + kTile - 1) / kTile;
    if (tiles == 0) tiles = 1;
    SortGeom g;
    g.tiles_per_block = (tiles + kMaxSortBlocks - 1) / kMaxSortBlocks;
    g.nblocks = (tiles + g.tiles_per_block - 1) / g.tiles_per_block;
    return g;
}

__global__ void __launch_bounds__(kSortThreads)
k_radix_hist(const uint32_t* __restrict__ keys, uint32_t* __restrict__ counts, uint32_t n, int shift,
             uint32_t tiles_per_block, uint32_t nblocks)
{
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t span = (uint64_t)tiles_per_block * kTile;
    const uint64_t begin = (uint64_t)blockIdx.x * span;
    uint64_t end = begin + span;
    if (end > n) end = n;
    for (uint64_t i = begin + threadIdx.x; i < end; i += kSortThreads) {
        uint32_t d = (keys[i] >> shift) & 0xFFu;
        atomicAdd(&hist[d], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// Row scan: block d turns row d of the [digit][block] count matrix into its exclusive prefix (in
// place, coalesced, carry across 256-wide chunks) and writes the row total to totals[d].  The
// scatter kernel adds the exclusive scan of the 256 totals itself.
__global__ void __launch_bounds__(kSortThreads)
k_radix_rowscan(uint32_t* __restrict__ counts, uint32_t* __restrict__ totals, uint32_t nblocks)
{
    __shared__ uint32_t warp_sums[kSortWarps];
    __shared__ uint32_t carry_s;
    uint32_t* row = counts + (size_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += kSortThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? row[i] : 0u;
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        #pragma unroll
        for (int w = 0; w < kSortWarps; w++) if (w < warp) wbase += warp_sums[w];
        const uint32_t carry = carry_s;
        if (i < nblocks) row[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == kSortThreads - 1) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

template <bool kIdentity>
__global__ void __launch_bounds__(kSortThreads, 6)
k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ totals, uint32_t n, int shift,
                uint32_t tiles_per_block, uint32_t nblocks)
{
    __shared__ uint32_t cnt[kSortWarps][256];
    __shared__ uint32_t base[256];
    __shared__ uint32_t texcl[256];
    __shared__ uint32_t wsum[kSortWarps];
    __shared__ uint32_t skey[kTile], sval[kTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    {   // digit base = exclusive scan of the 256 digit totals (thread d owns digit d)
        const uint32_t v = totals[threadIdx.x];
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        #pragma unroll
        for (int w = 0; w < kSortWarps; w++) if (w < warp) wbase += wsum[w];
        base[threadIdx.x] = wbase + inc - v + offsets[(size_t)threadIdx.x * nblocks + blockIdx.x];
    }

    for (uint32_t t = 0; t < tiles_per_block; t++) {
        const uint64_t tile_base = ((uint64_t)blockIdx.x * tiles_per_block + t) * kTile;
        if (tile_base >= n) break;
        #pragma unroll
        for (int w = 0; w < kSortWarps; w++) cnt[w][threadIdx.x] = 0;
        __syncthreads();

        uint32_t key[kItems], rank[kItems];
        #pragma unroll
        for (int r = 0; r < kItems; r++) {
            const uint64_t idx = tile_base + (uint64_t)warp * (32 * kItems) + r * 32 + lane;
            const bool valid = idx < n;
            key[r] = valid ? keys_in[idx] : 0u;
            const uint32_t d = (key[r] >> shift) & 0xFFu;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader && valid) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers); }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = old + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();
        // per digit d (thread d): exclusive scan across the warps of this tile; tile count; then an exclusive
        // scan of the 256 tile counts gives each digit's first slot inside the tile (texcl)
        uint32_t run = 0;
        {
            const int d = threadIdx.x;
            #pragma unroll
            for (int w = 0; w < kSortWarps; w++) { uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
            uint32_t inc = run;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t2 = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t2;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            uint32_t wbase = 0;
            #pragma unroll
            for (int w = 0; w < kSortWarps; w++) if (w < warp) wbase += wsum[w];
            texcl[d] = wbase + inc - run;
        }
        __syncthreads();
        // stage the tile in shared memory in digit order, then write it out linearly: consecutive threads
        // write consecutive slots of the same digit's run -> coalesced segments instead of 4-byte scatters
        #pragma unroll
        for (int r = 0; r < kItems; r++) {
            const uint64_t idx = tile_base + (uint64_t)warp * (32 * kItems) + r * 32 + lane;
            if (idx < n) {
                const uint32_t d = (key[r] >> shift) & 0xFFu;
                const uint32_t lp = texcl[d] + cnt[warp][d] + rank[r];
                skey[lp] = key[r];
                sval[lp] = kIdentity ? (uint32_t)idx : vals_in[idx];
            }
        }
        __syncthreads();
        const uint32_t tile_n = (uint32_t)min((uint64_t)kTile, (uint64_t)n - tile_base);
        for (uint32_t q = threadIdx.x; q < tile_n; q += kSortThreads) {
            const uint32_t k = skey[q];
            const uint32_t d = (k >> shift) & 0xFFu;
            const uint32_t dst = base[d] + (q - texcl[d]);
            keys_out[dst] = k;
            vals_out[dst] = sval[q];
        }
        __syncthreads();
        base[threadIdx.x] += run;   // advance the block's running global offset of digit d
        __syncthreads();
    }
}
}  // namespace

size_t radix_sort_temp_entries(uint32_t n)
{
    (void)n;                                   // nblocks is not monotonic in n: size for the maximum
    return (size_t)256 * kMaxSortBlocks + 256; // count matrix + digit totals
}

int radix_sort_pairs(cudaStream_t st, uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                     bool vals_identity, uint32_t n, int bits, uint32_t* counts, uint64_t* launches)
{
    if (n == 0) return 0;
    const SortGeom g = sort_geom(n);
    uint32_t* totals = counts + (size_t)256 * g.nblocks;
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    uint32_t *kin = keys_a, *kout = keys_b, *vin = vals_a, *vout = vals_b;
    int where = 0;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        k_radix_hist<<<g.nblocks, kSortThreads, 0, st>>>(kin, counts, n, shift, g.tiles_per_block, g.nblocks);
        k_radix_rowscan<<<256, kSortThreads, 0, st>>>(counts, totals, g.nblocks);
        if (p == 0 && vals_identity)
            k_radix_scatter<true><<<g.nblocks, kSortThreads, 0, st>>>(kin, nullptr, kout, vout, counts, totals, n, shift,
                                                                       g.tiles_per_block, g.nblocks);
        else
            k_radix_scatter<false><<<g.nblocks, kSortThreads, 0, st>>>(kin, vin, kout, vout, counts, totals, n, shift,
                                                                        g.tiles_per_block, g.nblocks);
        if (launches) *launches += 3;
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        where ^= 1;
    }
    return where;
}

}  // namespace sphb200

// ---- device-wide exclusive scan (in place) ---------------------------------------------------------
// Used by the counting sort of the GRID table: per-cell counts -> prefix table t[c] = rows with key < c.
// Reduce-then-scan in three launches over 4096-entry blocks; `data` must be padded to a multiple of 4096
// entries (zeros).  Traffic: 2 reads + 1 write of the table.
namespace sphb200 {
namespace {
constexpr int kScanThreads = 256;
constexpr int kScanPer = 16;                                 // entries per thread
constexpr int kScanBlock = kScanThreads * kScanPer;          // 4096

__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const uint4* __restrict__ data, uint32_t* __restrict__ blocksums)
{
    chain_prologue();
    __shared__ uint32_t wsum[kScanThreads / 32];
    const uint4* p = data + (size_t)blockIdx.x * (kScanBlock / 4);
    uint32_t s = 0;
    #pragma unroll
    for (int k = 0; k < kScanPer / 4; k++) { const uint4 v = p[k * kScanThreads + threadIdx.x]; s += v.x + v.y + v.z + v.w; }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; w++) t += wsum[w];
        blocksums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_scan_top(uint32_t* __restrict__ blocksums, const uint32_t nb)
{
    chain_prologue();
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? blocksums[i] : 0u;
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; w++) wbase += wsum[w];
        const uint32_t carry = carry_s;
        if (i < nb) blocksums[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
}

// FOLD: the block sums its predecessors' totals itself (raw block sums in, no k_scan_top launch) -- one launch and
// ~4 us less for tables of up to kFoldBlocks tiles, where the nb^2 / 2 extra reads are noise (small and 1 M-particle
// scenes); larger tables keep the three-kernel form.
template <bool FOLD>
__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const uint4* src, uint4* data, const uint32_t* __restrict__ blocksums)   // src == data: in place
{
    chain_prologue();
    __shared__ uint32_t wsum[kScanThreads / 32];
    __shared__ uint32_t fsum[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = 0;
    if (FOLD) {
        uint32_t a = 0;
        for (uint32_t b = threadIdx.x; b < blockIdx.x; b += kScanThreads) a += blocksums[b];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) fsum[warp] = a;
    }
    // thread t owns 16 CONSECUTIVE entries: uint4 words [4t, 4t+4) of the block
    const size_t at = (size_t)blockIdx.x * (kScanBlock / 4) + (size_t)threadIdx.x * (kScanPer / 4);
    const uint4* q = src + at;
    uint4* p = data + at;
    uint4 v[kScanPer / 4];
    uint32_t s = 0;
    #pragma unroll
    for (int k = 0; k < kScanPer / 4; k++) { v[k] = q[k]; s += v[k].x + v[k].y + v[k].z + v[k].w; }
    uint32_t inc = s;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (FOLD) { for (int w = 0; w < kScanThreads / 32; w++) base += fsum[w]; }
    else base = blocksums[blockIdx.x];
    uint32_t run = base + inc - s;
    for (int w = 0; w < warp; w++) run += wsum[w];
    #pragma unroll
    for (int k = 0; k < kScanPer / 4; k++) {
        uint4 o;
        o.x = run; run += v[k].x;
        o.y = run; run += v[k].y;
        o.z = run; run += v[k].z;
        o.w = run; run += v[k].w;
        p[k] = o;
    }
}
constexpr uint32_t kFoldBlocks = 2048;

// Small arrays (the segment counters of the reference's own scene sizes: a few thousand entries) in ONE launch of one
// block: thread t scans its kSmallPer consecutive entries, the block scans the thread totals.  At 10 k particles every
// launch of the step is latency, not work; this saves one.
constexpr int kSmallThreads = 1024;
constexpr uint32_t kSmallBlocks = 8;                          // up to 8 tiles = 32768 entries = 32 per thread
__global__ void __launch_bounds__(kSmallThreads)
k_scan_small(const uint4* src, uint4* data, const uint32_t per4)   // per4: uint4 words per thread (entries / 4096)
{
    chain_prologue();
    __shared__ uint32_t wsum[kSmallThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* q = src + (size_t)threadIdx.x * per4;
    uint4* p = data + (size_t)threadIdx.x * per4;
    uint4 v[kSmallBlocks];
    uint32_t s = 0;
    #pragma unroll
    for (uint32_t k = 0; k < kSmallBlocks; k++)
        if (k < per4) { v[k] = q[k]; s += v[k].x + v[k].y + v[k].z + v[k].w; }
    uint32_t inc = s;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = wsum[lane];
        uint32_t wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + inc - s;
    #pragma unroll
    for (uint32_t k = 0; k < kSmallBlocks; k++)
        if (k < per4) {
            uint4 o;
            o.x = run; run += v[k].x;
            o.y = run; run += v[k].y;
            o.z = run; run += v[k].z;
            o.w = run; run += v[k].w;
            p[k] = o;
        }
}
}  // namespace

size_t scan_pad(size_t entries) { return (entries + kScanBlock - 1) / kScanBlock * kScanBlock; }
size_t scan_temp_entries(size_t entries) { return scan_pad(entries) / kScanBlock + 1; }

void exclusive_scan_u32(cudaStream_t st, const uint32_t* src, uint32_t* data, size_t padded_entries, uint32_t* blocksums, uint64_t* launches)
{
    const uint32_t nb = (uint32_t)(padded_entries / kScanBlock);
    if (nb == 0) return;
    static const bool fold_ok = [] { const char* e = getenv("SPH_SCAN_FOLD"); return !(e && e[0] == '0'); }();
    if (fold_ok && nb <= kSmallBlocks) {                       // entries = nb * 4096 = 1024 threads * (nb uint4 words)
        launch_chained(k_scan_small, dim3(1), dim3(kSmallThreads), 0, st, (const uint4*)src, (uint4*)data, nb);
        if (launches) *launches += 1;
        return;
    }
    launch_chained(k_scan_reduce, dim3(nb), dim3(kScanThreads), 0, st, (const uint4*)src, blocksums);
    if (fold_ok && nb <= kFoldBlocks) {
        launch_chained(k_scan_apply<true>, dim3(nb), dim3(kScanThreads), 0, st, (const uint4*)src, (uint4*)data, blocksums);
        if (launches) *launches += 2;
        return;
    }
    launch_chained(k_scan_top, dim3(1), dim3(1024), 0, st, blocksums, nb);
    launch_chained(k_scan_apply<false>, dim3(nb), dim3(kScanThreads), 0, st, (const uint4*)src, (uint4*)data, blocksums);
    if (launches) *launches += 3;
}
}  // namespace sphb200
