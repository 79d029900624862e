// sph_api.cu -- the C ABI of include/sph_b200.h: context, buffers, step orchestration.
//
// One step = the reference's FluidSimulation::Update (engine/physics/physicsWorld.cc:39-111):
//   predict+key -> radix sort -> table -> reorder -> density -> pressure -> viscosity -> integrate
// all enqueued on the context's stream; six CUDA-event timers mirror getElapsedTime* (:184-212).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "sph_context.h"

using namespace sphb200;

namespace {
std::string g_create_error;
std::mutex g_create_mutex;
constexpr float kPi = (float)3.14159265358979323846264338327950288;   // glm::pi<float>()
}

namespace sphb200 {

int fail(SphContext* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    else { std::lock_guard<std::mutex> l(g_create_mutex); g_create_error = msg; }
    return code;
}
int cuda_fail(SphContext* c, cudaError_t e, const char* what)
{
    return fail(c, SPH_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static int ceil_log2(uint64_t v)
{
    int b = 0;
    while ((1ull << b) < v && b < 63) b++;
    return b;
}

// SPH_SORT=radix keeps the LSD radix sort + table build for the GRID table (the REFERENCE_HASH table always uses it)
// SPH_PDL=0: the chain's kernels are launched ordinarily (launch_chained, sph_internal.h)
bool chained_launches_enabled()
{
    static const bool on = [] { const char* e = getenv("SPH_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

bool counting_sort_enabled()
{
    static const bool on = [] { const char* e = getenv("SPH_SORT"); return !(e && e[0] == 'r'); }();
    return on;
}

int make_dev_params(SphContext* c, uint32_t n, DevParams* P)
{
    const SphParams& p = c->params;
    memset(P, 0, sizeof(*P));
    P->n = n; P->n_owned = n; P->mode = c->grid_too_large ? (int)SPH_TABLE_REFERENCE_HASH : c->mode; P->gravity = p.gravity ? 1 : 0;
    P->r = p.interaction_radius; P->sqr_r = p.sqr_radius; P->rho0 = p.target_density;
    P->k = p.pressure_multiplier; P->kn = p.near_pressure_multiplier; P->mu = p.viscosity_strength;
    P->g = p.gravity_scale;
    for (int a = 0; a < 3; a++) P->half[a] = p.bound[a] * 0.5f;                  // physicsWorld.cc:88
    const float r = p.interaction_radius;
    // the reference recomputes these per call in fp32 (kernels.h:29,41,53,65,77); same expressions, once
    P->vol2 = 15 / (2 * kPi * powf(r, 5));
    P->vol3 = 15.0f / (kPi * powf(r, 6));
    P->s2 = 15.0f / (powf(r, 5) * kPi);
    P->s3 = 45 / (powf(r, 6) * kPi);
    P->sv = 315 / (64 * kPi * powf(fabsf(r), 9));
    P->rr = r * r;
    P->cull_hi = nextafterf((float)((double)p.sqr_radius * (1.0 + 1e-6)), INFINITY);
    P->cull_lo = nextafterf((float)((double)p.sqr_radius * (1.0 - 1e-6)), -INFINITY);
    for (int a = 0; a < 3; a++) { P->gmin[a] = c->gmin[a]; P->gdim[a] = c->gdim[a]; }
    P->xsub = c->xsub;
    P->xwin = (float)(sqrt((double)p.sqr_radius) * (1.0 + 1e-5));
    P->ncell = c->ncell;
    P->modM = n ? (UINT64_MAX / n + 1) : 0;
    P->row0 = 0; P->row1 = n; P->n_a = n;
    P->pair_cap = (c->cap + 8u) / 2u + kPairPad;
    P->seg_off = (uint32_t)table_layout(c->ncell).cells_pad;
    P->cnt_off = P->seg_off + (uint32_t)table_layout(c->ncell).nseg_pad;
    P->noncanonical = c->d_noncanonical;
    P->extras = c->extras_on ? 1 : 0;
    for (int i = 0; i < 9; i++) P->rot[i] = c->rot[i];
    P->stick_k = c->extras.stick_strength; P->stick_d = c->extras.stick_distance;
    P->rim_check = (double)p.sqr_radius > (double)r * (double)r * (1.0 + 4e-7) ? 1 : 0;
    P->slab = 0; P->zlo = 0; P->gz_global = c->gdim[2]; P->own_lo = 0; P->own_hi = c->gdim[2];
    return SPH_OK;
}

// GRID geometry: cover the box plus two cells of margin per side for predicted positions that
// overshoot the walls; anything further out clamps into the border cells (still a superset walk).
static int update_grid_geometry(SphContext* c)
{
    const SphParams& p = c->params;
    if (!(p.interaction_radius > 0.0f)) return fail(c, SPH_ERR_INVALID, "interaction_radius must be > 0");
    // The UI's radius slider goes down to 0.01 (gameApp.cc:371) and the reference's setters cannot fail: a grid that
    // would get too large is first coarsened in x (the subdivision only trims candidate windows), and if even the
    // unsubdivided grid exceeds the limit the context steps on the reference's own table, whose size is the particle
    // count, instead of refusing the parameter.
    const uint64_t kMaxCells = 1ull << 30;          // 4 GB of cell counters (C4, 64 M particles in a 258 x 129 x 86 box: 0.56 G cells)
    int dims[3], los[3];
    uint64_t base_cells = 1;
    for (int a = 0; a < 3; a++) {
        const float half0 = p.bound[a] * 0.5f;
        if (!(half0 >= 0.0f) || !std::isfinite(half0)) return fail(c, SPH_ERR_INVALID, "bound must be finite and >= 0");
        // a rotated box (SphExtras): the table covers its bounding box, |R| * half
        double half = half0;
        if (c->extras_on) { half = 0.0; for (int b = 0; b < 3; b++) half += std::fabs((double)c->rot[3 * a + b]) * (double)(p.bound[b] * 0.5f); }
        double q = std::floor((double)half / (double)p.interaction_radius);
        if (q > 1e6) q = 1e6;                       // far beyond any table: falls through to grid_too_large
        const int hi = (int)q + 2, lo = -(int)q - 3;
        los[a] = lo;
        dims[a] = hi - lo + 1;
        base_cells *= (uint64_t)dims[a];
    }
    c->grid_too_large = base_cells > kMaxCells;
    int xs = c->xsub_pref;
    while (xs > 1 && base_cells * (uint64_t)xs > kMaxCells) xs >>= 1;
    if (c->grid_too_large) {                        // a token grid: nothing is built on it while the flag is set
        if (c->nranks > 1) return fail(c, SPH_ERR_INVALID, "grid table too large for slab mode (radius too small for this box)");
        for (int a = 0; a < 3; a++) { c->gmin[a] = 0; c->gdim[a] = 1; }
        c->xsub = 1;
        c->ncell = 1;
        return SPH_OK;
    }
    c->xsub = xs;
    uint64_t cells = 1;
    for (int a = 0; a < 3; a++) {
        c->gmin[a] = los[a];
        c->gdim[a] = dims[a] * (a == 0 ? c->xsub : 1);          // x is subdivided (sph_device.cuh: grid_cell)
        cells *= (uint64_t)c->gdim[a];
    }
    c->ncell = (uint32_t)cells;
    return SPH_OK;
}

int ensure_tables(SphContext* c, const DevParams& P)
{
    // GRID: two-level prefix table over ncell (+1 departed bucket in slab mode) + end: cells, segment bases, dirty flags
    const size_t need = (P.mode == SPH_TABLE_GRID) ? table_layout(P.ncell).total : (size_t)c->cap + 1;
    if (need > c->table_cap) {
        SPH_CUDA(c, cudaStreamSynchronize(c->st));
        if (c->tstart) cudaFree(c->tstart);
        c->tstart = nullptr; c->table_cap = 0;
        SPH_CUDA(c, cudaMalloc(&c->tstart, need * sizeof(uint32_t)));
        c->table_cap = need;
        c->table_two_level = false;
    }
    if (P.mode == SPH_TABLE_GRID) {
        const size_t sneed = scan_temp_entries(table_layout(P.ncell).nseg_pad);
        if (sneed > c->scan_cap) {
            SPH_CUDA(c, cudaStreamSynchronize(c->st));
            if (c->scan_tmp) cudaFree(c->scan_tmp);
            c->scan_tmp = nullptr; c->scan_cap = 0;
            SPH_CUDA(c, cudaMalloc(&c->scan_tmp, sneed * sizeof(uint32_t)));
            c->scan_cap = sneed;
        }
        const size_t gneed = 1 + 3 * ((size_t)P.ncell / 1024 + (size_t)P.ncell / 16384 + 8);   // worst case of k_build_table_grid's worklist
        if (gneed > c->gap_cap) {
            SPH_CUDA(c, cudaStreamSynchronize(c->st));
            if (c->gap_list) cudaFree(c->gap_list);
            c->gap_list = nullptr; c->gap_cap = 0;
            SPH_CUDA(c, cudaMalloc(&c->gap_list, gneed * sizeof(uint32_t)));
            c->gap_cap = gneed;
        }
    } else if (!c->tend) {
        SPH_CUDA(c, cudaMalloc(&c->tend, ((size_t)c->cap + 1) * sizeof(uint32_t)));
    }
    return SPH_OK;
}

// Depth of the density pass's survivor stack, from the list length the last steps actually produced (the running sums
// of list rows and of the warps that wrote them, read back with the overflow word): the deep stack -- a quarter of the
// occupancy -- pays only when lists are long THROUGHOUT (C5: ~110 rows per warp), not because one pile-up once pushed
// the CAPACITY up (an evolved C2 run sat at 280 instead of 125 us per density pass that way).  Hysteresis 40 / 56 rows.
// Called before every step, replayed ones too: a flipped decision changes the StepKey, so the step is launched plainly
// and recorded again.
static void update_stack_depth(SphContext* c)
{
    if (c->capturing || !c->h_overflow) return;
    // (warps << 32 | rows) in one aligned 64-bit word: read whole, differences taken in 64 bits
    const uint64_t now = *reinterpret_cast<const volatile uint64_t*>(c->h_overflow + 2);
    const uint64_t d = now - c->rows_warps_seen;
    const uint32_t rows = (uint32_t)d, warps = (uint32_t)(d >> 32);
    if (!warps) return;
    const double mean = (double)rows / (double)warps;
    if (mean > 56.0) c->deep_stack = true;
    else if (mean < 40.0) c->deep_stack = false;
    c->rows_warps_seen = now;
}

// List overflow / staging need / list-length sums of the density pass that `after` has just been given: one 16-byte copy to
// the pinned mirror.  It leaves on a side stream: as a node of the solver's stream a device-to-host copy costs the step
// ~9 us wherever it stands (C2 replayed: 0.2957 -> 0.2871 ms once it was off the chain).  The counters are monotonic and
// read without synchronising, so nothing waits for the copy -- except a recording, which must join its branches
// (run_step does, at the end of the step).
int copy_list_words(SphContext* c, cudaStream_t after)
{
    if (!c->st_fork) {
        SPH_CUDA(c, cudaMemcpyAsync(c->h_overflow, c->d_overflow, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, after));
        return SPH_OK;
    }
    SPH_CUDA(c, cudaEventRecord(c->ev_fork, after));
    SPH_CUDA(c, cudaStreamWaitEvent(c->st_fork, c->ev_fork, 0));
    SPH_CUDA(c, cudaMemcpyAsync(c->h_overflow, c->d_overflow, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st_fork));
    if (c->capturing) SPH_CUDA(c, cudaEventRecord(c->ev_join, c->st_fork));
    return SPH_OK;
}

int ensure_list(SphContext* c, NbrList* L)
{
    if (!c->h_overflow) {
        // (list overflow, staging need, list rows written so far, warps that wrote them): ONE copy per step
        SPH_CUDA(c, cudaMallocHost((void**)&c->h_overflow, 4 * sizeof(uint32_t)));
        memset(c->h_overflow, 0, 4 * sizeof(uint32_t));
        SPH_CUDA(c, cudaMalloc((void**)&c->d_overflow, 4 * sizeof(uint32_t)));
        SPH_CUDA(c, cudaMemsetAsync(c->d_overflow, 0, 4 * sizeof(uint32_t), c->st));
        c->h_tile_need = c->h_overflow + 1;
        c->d_tile_need = c->d_overflow + 1;
        c->tile_capn = tile_default_capn();
    }
    // tile generation: a cell whose 27-cell neighbourhood did not fit the staging buffer was walked instead (exact,
    // slower); grow the buffer for the next step, as far as four warps' worth of 32-byte records fit one SM
    if (!c->capturing && *c->h_tile_need > c->tile_capn) {
        const uint32_t want = (*c->h_tile_need * 9u / 8u + 127u) & ~127u;
        c->tile_capn = want > 1664u ? 1664u : want;
    }
    // auto-grow: the value may lag the kernels by a step or two (read without synchronising); overflowing
    // particles are exact meanwhile (the later passes walk the table for them), only slower
    if (c->list_auto && c->list_k && !c->capturing && *c->h_overflow > c->list_k) {
        const uint32_t want = (*c->h_overflow * 5u / 4u + 15u) & ~15u;
        c->list_k = want > 4096u ? 4096u : want;
    }
    if (c->list_k && c->list_k_alloc < c->list_k) {
        SPH_CUDA(c, cudaStreamSynchronize(c->st));
        if (c->nlist) cudaFree(c->nlist);
        c->nlist = nullptr; c->list_k_alloc = 0;
        // rows [0, k) hold the neighbour indices, rows [k, 2k) the viscosity weights of the same entries.  The allocation
        // takes up to twice the rows asked for while that stays within a quarter of the free memory: the capacity can then
        // grow (a splash, a pile-up in a corner) without another synchronise + free + allocate, a ~100 ms stall at 1 M rows
        const size_t per_row = 2 * (size_t)c->cap * sizeof(uint32_t);
        uint32_t k_alloc = c->list_k;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const size_t roomy = std::min<size_t>(2 * (size_t)c->list_k, (free_b / 4) / per_row);
            if (roomy > k_alloc) k_alloc = (uint32_t)roomy;
        } else cudaGetLastError();
        cudaError_t e = cudaMalloc(&c->nlist, (size_t)k_alloc * per_row);
        if (e != cudaSuccess && k_alloc > c->list_k) { cudaGetLastError(); k_alloc = c->list_k; e = cudaMalloc(&c->nlist, (size_t)k_alloc * per_row); }
        if (e != cudaSuccess) { cudaGetLastError(); c->list_k = 0; }        // no room: fall back to walking every pass
        else c->list_k_alloc = k_alloc;
    }
    L->idx = c->list_k ? c->nlist : nullptr;
    L->w = c->list_k ? reinterpret_cast<float*>(c->nlist + (size_t)c->list_k_alloc * c->cap) : nullptr;
    L->cnt = c->lcount;
    update_stack_depth(c);
    L->deep = c->deep_stack;
    L->rows_sum = c->d_overflow + 2;
    L->overflow = c->d_overflow;
    L->ncount = c->ncount;
    L->k = c->list_k;
    L->stride = c->cap;
    L->keys = c->sorted_where ? c->key_b : c->key_a;
    L->capn = c->tile_capn;
    L->tile_need = c->d_tile_need;
    return SPH_OK;
}

}  // namespace sphb200

static void free_all(SphContext* c)
{
    void* ptrs[] = {c->A_pos, c->A_vel, c->S_pos, c->S_vel, c->pred, c->predpk, c->velp, c->dens, c->key_a, c->key_b,
                    c->perm_a, c->perm_b, c->ncount, c->lcount, c->nlist, c->d_noncanonical, c->row_of, c->scan_tmp, c->tstart, c->tend, c->gap_list, c->counts, c->stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->h_overflow) cudaFreeHost(c->h_overflow);
    if (c->d_overflow) cudaFree(c->d_overflow);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->stage_in) cudaFree(c->stage_in);
    if (c->stage_out) cudaFree(c->stage_out);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {c->ev_h2d, c->ev_pack, c->ev_export, c->ev_d2h}) if (e) cudaEventDestroy(e);
    if (c->st_fork) cudaStreamDestroy(c->st_fork);
    for (cudaEvent_t e : {c->ev_fork, c->ev_join}) if (e) cudaEventDestroy(e);
    if (c->st_in) cudaStreamDestroy(c->st_in);
    if (c->st_out) cudaStreamDestroy(c->st_out);
    if (c->st) cudaStreamDestroy(c->st);
}

extern "C" {

int sph_abi_version(void) { return SPH_B200_ABI_VERSION; }

void sph_default_params(SphParams* p)
{   // physicsWorld.h:96-106,145
    p->interaction_radius = 0.35f;
    p->sqr_radius = 0.35f * 0.35f;
    p->target_density = 99.7f;
    p->pressure_multiplier = 300.0f;
    p->near_pressure_multiplier = 20.0f;
    p->viscosity_strength = 0.5f;
    p->gravity_scale = 10.0f;
    p->gravity = 0;
    p->bound[0] = p->bound[1] = p->bound[2] = 20.0f;
}

const char* sph_last_error(const SphContext* ctx)
{
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> l(g_create_mutex);
    return g_create_error.c_str();
}

int sph_create(SphContext** out, int device, uint32_t capacity)
{
    if (!out) return fail(nullptr, SPH_ERR_INVALID, "sph_create: out is NULL");
    *out = nullptr;
    if (capacity == 0) return fail(nullptr, SPH_ERR_INVALID, "sph_create: capacity must be > 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SPH_ERR_CUDA, std::string("sph_create: no CUDA device (there is no CPU fallback): ") +
                                               cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, SPH_ERR_INVALID, "sph_create: bad device ordinal");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");

    SphContext* c = new SphContext();
    if (const char* xs = getenv("SPH_XSUB")) { const int v = atoi(xs); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) c->xsub_pref = c->xsub = v; }
    c->device = device;
    c->cap = capacity;
    sph_default_params(&c->params);
    const size_t cap = capacity;
#define ALLOC(ptr, bytes)                                                        \
    do {                                                                         \
        e = cudaMalloc((void**)&(ptr), (bytes));                                 \
        if (e != cudaSuccess) { free_all(c); delete c; return cuda_fail(nullptr, e, "cudaMalloc " #ptr); } \
    } while (0)
    e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
    // the fork of the recorded step (run_step); without it the step simply stays serial
    if (cudaStreamCreateWithFlags(&c->st_fork, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); c->st_fork = nullptr; }
    ALLOC(c->A_pos, cap * 16); ALLOC(c->A_vel, cap * 16);
    ALLOC(c->S_pos, cap * 16); ALLOC(c->S_vel, cap * 16);
    ALLOC(c->pred, (cap + 8) * 16);   /* padded: the gather loads 4 rows at a time */  ALLOC(c->velp, cap * 16);
    ALLOC(c->dens, cap * 32);
    ALLOC(c->predpk, (cap + 8) * 16 + (size_t)kPairPad * 32);
    ALLOC(c->key_a, cap * 4);  ALLOC(c->key_b, cap * 4);
    ALLOC(c->perm_a, cap * 4); ALLOC(c->perm_b, cap * 4);
    ALLOC(c->ncount, cap * 4);
    ALLOC(c->lcount, cap * 4);
    ALLOC(c->stage, cap * 32);
    ALLOC(c->d_noncanonical, 4);
    cudaMemset(c->d_noncanonical, 0, 4);
    c->counts_cap = radix_sort_temp_entries(capacity);
    ALLOC(c->counts, c->counts_cap * 4);
#undef ALLOC
    for (auto& ev : c->ev) {
        e = cudaEventCreate(&ev);
        if (e != cudaSuccess) { free_all(c); delete c; return cuda_fail(nullptr, e, "cudaEventCreate"); }
    }
    int rc = update_grid_geometry(c);
    if (rc != SPH_OK) { std::string m = c->err; free_all(c); delete c; return fail(nullptr, rc, m); }
    *out = c;
    return SPH_OK;
}

int sph_destroy(SphContext* c)
{
    if (!c) return SPH_OK;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->st_fork) cudaStreamSynchronize(c->st_fork);      // the copy of the list words into the pinned mirror may still be on its way
    if (c->st_in) cudaStreamSynchronize(c->st_in);
    if (c->st_out) cudaStreamSynchronize(c->st_out);
    multi_teardown(c);
    free_all(c);
    delete c;
    return SPH_OK;
}

int sph_set_params(SphContext* c, const SphParams* p)
{
    if (!c || !p) return SPH_ERR_INVALID;
    const SphParams old = c->params;
    c->params = *p;
    int rc = update_grid_geometry(c);
    // slab mode: the planes are world-space z values; a new radius or bound changes the layer grid under them
    if (rc == SPH_OK) rc = multi_params_changed(c);
    if (rc != SPH_OK) { const std::string m = c->err; c->params = old; update_grid_geometry(c); multi_params_changed(c); c->err = m; return rc; }
    return SPH_OK;
}

int sph_set_extras(SphContext* c, const SphExtras* e)
{
    if (!c || !e) return SPH_ERR_INVALID;
    const float* q = e->bound_rotation;
    const double n2 = (double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3];
    if (!(n2 > 1e-12) || !std::isfinite(n2)) return fail(c, SPH_ERR_INVALID, "sph_set_extras: the rotation quaternion is zero or not finite");
    if (!(e->stick_strength >= 0.0f) || !std::isfinite(e->stick_strength)) return fail(c, SPH_ERR_INVALID, "sph_set_extras: stick_strength must be >= 0");
    if (e->stick_strength > 0.0f && !(e->stick_distance > 0.0f && std::isfinite(e->stick_distance)))
        return fail(c, SPH_ERR_INVALID, "sph_set_extras: stick_distance must be > 0 when stick_strength is");
    const SphExtras old = c->extras;
    float oldrot[9];
    memcpy(oldrot, c->rot, sizeof(oldrot));
    const bool old_on = c->extras_on;
    const double inv = 1.0 / std::sqrt(n2);
    const double x = q[0] * inv, y = q[1] * inv, z = q[2] * inv, w = q[3] * inv;
    c->extras = *e;
    c->extras.bound_rotation[0] = (float)x; c->extras.bound_rotation[1] = (float)y; c->extras.bound_rotation[2] = (float)z; c->extras.bound_rotation[3] = (float)w;
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int i = 0; i < 9; i++) c->rot[i] = (float)R[i];
    const bool rotated = std::fabs(w) < 1.0 - 1e-12 || x != 0.0 || y != 0.0 || z != 0.0;
    c->extras_on = rotated || e->stick_strength > 0.0f;
    if (!rotated) { const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; memcpy(c->rot, I, sizeof(I)); }
    int rc = update_grid_geometry(c);
    if (rc == SPH_OK) rc = multi_params_changed(c);
    if (rc != SPH_OK) {
        const std::string m = c->err;
        c->extras = old; memcpy(c->rot, oldrot, sizeof(oldrot)); c->extras_on = old_on;
        update_grid_geometry(c); multi_params_changed(c);
        c->err = m;
        return rc;
    }
    return SPH_OK;
}

int sph_get_extras(const SphContext* c, SphExtras* e)
{
    if (!c || !e) return SPH_ERR_INVALID;
    *e = c->extras;
    return SPH_OK;
}

int sph_get_params(const SphContext* c, SphParams* p)
{
    if (!c || !p) return SPH_ERR_INVALID;
    *p = c->params;
    return SPH_OK;
}

int sph_set_table_mode(SphContext* c, int mode)
{
    if (!c) return SPH_ERR_INVALID;
    if (mode != SPH_TABLE_GRID && mode != SPH_TABLE_REFERENCE_HASH) return fail(c, SPH_ERR_INVALID, "unknown table mode");
    if (mode == SPH_TABLE_REFERENCE_HASH && c->nranks > 1)
        return fail(c, SPH_ERR_UNSUPPORTED, "slab mode uses the grid table (the reference table is global)");
    c->mode = mode;
    return SPH_OK;
}
int sph_get_table_mode(const SphContext* c) { return c ? c->mode : -1; }
int sph_set_stage_timing(SphContext* c, int enabled) { if (!c) return SPH_ERR_INVALID; c->timing = enabled != 0; return SPH_OK; }
int sph_set_neighbour_count_tap(SphContext* c, int enabled) { if (!c) return SPH_ERR_INVALID; c->nc_tap = enabled != 0; return SPH_OK; }
int sph_set_neighbour_list_capacity(SphContext* c, uint32_t entries)
{
    if (!c) return SPH_ERR_INVALID;
    if (entries > 4096) return fail(c, SPH_ERR_INVALID, "neighbour list capacity above 4096 entries per particle");
    c->list_k = entries;
    c->list_auto = false;                                  // an explicit capacity is kept as is
    if (c->h_overflow) { *c->h_overflow = 0; cudaMemsetAsync(c->d_overflow, 0, sizeof(uint32_t), c->st); }
    return SPH_OK;
}
uint32_t sph_num_particles(const SphContext* c) { return c ? c->n : 0; }
uint64_t sph_launch_count(const SphContext* c) { return c ? c->launches : 0; }
uint64_t sph_graph_replays(const SphContext* c) { return c ? c->graph_replays : 0; }
int sph_set_graph_replay(SphContext* c, int enabled)
{
    if (!c) return SPH_ERR_INVALID;
    c->graph_user_off = enabled == 0;
    return SPH_OK;
}
uint64_t sph_noncanonical_cells(SphContext* c)
{   // cells (summed over all steps so far) too crowded for the counting sort's canonical-order ranking; synchronises
    if (!c || !c->d_noncanonical) return 0;
    uint32_t v = 0;
    cudaSetDevice(c->device);
    if (cudaStreamSynchronize(c->st) != cudaSuccess || cudaMemcpy(&v, c->d_noncanonical, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 0; }
    return v;
}
void* sph_stream(const SphContext* c) { return c ? (void*)c->st : nullptr; }
int sph_density_stack_rows(const SphContext* c) { return c ? (c->deep_stack ? 72 : 24) : 0; }

int sph_grid_x_subdivision(const SphContext* c) { return c ? c->xsub : 0; }

int sph_get_grid(const SphContext* c, int32_t* dims3, int32_t* origin3)
{
    if (!c) return SPH_ERR_INVALID;
    for (int a = 0; a < 3; a++) { if (dims3) dims3[a] = c->gdim[a]; if (origin3) origin3[a] = c->gmin[a]; }
    return SPH_OK;
}

int sph_upload_state(SphContext* c, uint32_t n, const float* pos3, const float* vel3)
{
    if (!c) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_INVALID, "slab mode: use sph_upload_owned");
    if (n > c->cap) return fail(c, SPH_ERR_CAPACITY, "sph_upload_state: n exceeds capacity");
    if (n && !pos3) return fail(c, SPH_ERR_INVALID, "sph_upload_state: pos3 is NULL");
    SPH_CUDA(c, cudaSetDevice(c->device));
    float* dpos = (float*)c->stage;
    float* dvel = (float*)(c->stage + (size_t)c->cap * 12);
    if (n) {
        SPH_CUDA(c, cudaMemcpyAsync(dpos, pos3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st));
        if (vel3) SPH_CUDA(c, cudaMemcpyAsync(dvel, vel3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st));
        launch_pack_state(c->st, dpos, vel3 ? dvel : nullptr, nullptr, c->A_pos, c->A_vel, n, &c->launches);
        SPH_CUDA(c, cudaGetLastError());
    }
    c->n = n;
    c->step_valid = false;
    c->ncount_valid = false;
    return SPH_OK;
}

// Stage timer i.  (Recording them INTO the graph as external event-record nodes was measured: seven such nodes cost a
// 1 M-particle step 0.066 ms, more than the replay saves; the recording therefore carries no timers, and a context with
// the timers on takes every 16th step through plain launches to refresh them -- step_once.)
static cudaError_t stage_event(SphContext* c, int i)
{
    return cudaEventRecord(c->ev[i], c->st);
}

// the whole step; dt == 0 with `advance == false` is InitializeData's tail (lookup + densities only)
static int run_step(SphContext* c, float dt, bool advance, bool allow_timing = true)
{
    SPH_CUDA(c, cudaSetDevice(c->device));
    DevParams P;
    int rc = make_dev_params(c, c->n, &P);
    if (rc != SPH_OK) return rc;
    if (c->n == 0) return SPH_OK;
    rc = ensure_tables(c, P);
    if (rc != SPH_OK) return rc;
    const bool timing = c->timing && advance && allow_timing;
    cudaStream_t st = c->st;
    if (timing) SPH_CUDA(c, stage_event(c, 0));
    if (P.mode == SPH_TABLE_GRID && counting_sort_enabled()) {
        // counting sort over the table (a one-pass radix sort whose digit is the whole key): tickets in the cell
        // counters, in-place scan -> prefix table, placement; the reorder restores the canonical (stable) order
        const TableLayout T = table_layout(P.ncell);
        if (c->table_ncell != P.ncell) { c->table_two_level = false; c->table_ncell = P.ncell; }
        c->table_seg_off = P.seg_off;
        if (!c->table_two_level) {                 // first build, or the allocation last held a flat / reference table
            SPH_CUDA(c, cudaMemsetAsync(c->tstart, 0, T.total * sizeof(uint32_t), st));
            if (!c->capturing) c->table_two_level = true;
        } else launch_table_clear(st, c->tstart, T, &c->launches);
        launch_predict_key(st, c->A_pos, c->A_vel, c->key_a, nullptr, P.n, false, P, dt, c->tstart, c->perm_b, &c->launches);
        if (timing) SPH_CUDA(c, stage_event(c, 1));
        // segment counts -> segment bases (out of place) and, independently of it, the in-segment prefixes of the occupied
        // segments: in a recording the two run side by side (a fork through a second stream; plain launches stay serial,
        // an event pair would cost what the overlap saves)
        const uint32_t* seg_cnt = c->tstart + T.cells_pad + T.nseg_pad;
        if (c->capturing && c->st_fork) {
            SPH_CUDA(c, cudaEventRecord(c->ev_fork, st));
            SPH_CUDA(c, cudaStreamWaitEvent(c->st_fork, c->ev_fork, 0));
            exclusive_scan_u32(c->st_fork, seg_cnt, c->tstart + T.cells_pad, T.nseg_pad, c->scan_tmp, &c->launches);
            SPH_CUDA(c, cudaEventRecord(c->ev_join, c->st_fork));
            launch_inseg_scan(st, c->tstart, T, &c->launches);
            SPH_CUDA(c, cudaStreamWaitEvent(st, c->ev_join, 0));
        } else {
            exclusive_scan_u32(st, seg_cnt, c->tstart + T.cells_pad, T.nseg_pad, c->scan_tmp, &c->launches);
            launch_inseg_scan(st, c->tstart, T, &c->launches);
        }
        launch_place(st, c->key_a, c->perm_b, c->tstart, c->perm_a, P.n, P, &c->launches);
        launch_reorder(st, c->perm_a, c->key_a, c->tstart, c->key_b, c->A_pos, c->A_vel, nullptr, c->S_pos, c->S_vel, c->pred,
                       c->predpk, P, dt, &c->launches);
        c->sorted_where = 1;
    } else {
        launch_predict_key(st, c->A_pos, c->A_vel, c->key_a, nullptr, P.n, false, P, dt, nullptr, nullptr, &c->launches);
        if (timing) SPH_CUDA(c, stage_event(c, 1));
        const int bits = ceil_log2(P.mode == SPH_TABLE_GRID ? (uint64_t)P.ncell : (uint64_t)P.n);
        c->sorted_where = radix_sort_pairs(st, c->key_a, c->key_b, c->perm_a, c->perm_b, true, P.n, bits, c->counts,
                                           &c->launches);
        const uint32_t* keys = c->sorted_where ? c->key_b : c->key_a;
        const uint32_t* perm = c->sorted_where ? c->perm_b : c->perm_a;
        launch_build_table(st, keys, c->tstart, c->tend, c->gap_list, P, &c->launches);
        c->table_two_level = false;
        c->table_seg_off = P.seg_off;
        launch_reorder(st, perm, nullptr, nullptr, nullptr, c->A_pos, c->A_vel, nullptr, c->S_pos, c->S_vel, c->pred, c->predpk, P, dt,
                       &c->launches);
    }
    if (timing) SPH_CUDA(c, stage_event(c, 2));
    NbrList L;
    rc = ensure_list(c, &L);
    if (rc != SPH_OK) return rc;
    launch_density(st, c->pred, c->predpk, c->tstart, c->tend, c->dens, L, P, &c->launches);
    c->ncount_valid = true;
    // (a recording forks the copy of the list words here, right behind the density pass; plain launches send it off at the end
    // of the step, so that no event record stands between two kernels of the chain)
    if (L.idx && c->capturing) { rc = copy_list_words(c, st); if (rc != SPH_OK) return rc; }
    if (timing) SPH_CUDA(c, stage_event(c, 3));
    if (advance) {
        launch_pressure(st, c->pred, c->dens, c->S_vel, c->tstart, c->tend, c->velp, L, P, dt, &c->launches);
        if (timing) SPH_CUDA(c, stage_event(c, 4));
        // v'' goes into S_vel: dead after the pressure pass read it, and never read by the viscosity pass
        launch_viscosity(st, c->pred, c->velp, c->tstart, c->tend, c->S_vel, L, P, dt, &c->launches);
        if (timing) SPH_CUDA(c, stage_event(c, 5));
        launch_integrate(st, c->S_pos, c->S_vel, c->A_pos, c->A_vel, P, dt, &c->launches);
        if (timing) SPH_CUDA(c, stage_event(c, 6));
        if (!c->capturing) c->ev_recorded = timing;       // (a recording runs nothing: the last plain step's events stay valid)
    } else {
        // keep "every per-particle array shares the device order": adopt the sorted order
        SPH_CUDA(c, cudaMemcpyAsync(c->A_pos, c->S_pos, (size_t)c->n * 16, cudaMemcpyDeviceToDevice, st));
        SPH_CUDA(c, cudaMemcpyAsync(c->A_vel, c->S_vel, (size_t)c->n * 16, cudaMemcpyDeviceToDevice, st));
    }
    if (L.idx && c->capturing && c->st_fork) SPH_CUDA(c, cudaStreamWaitEvent(st, c->ev_join, 0));      // a recording must join its branches
    else if (L.idx && !c->capturing) { rc = copy_list_words(c, st); if (rc != SPH_OK) return rc; }
    SPH_CUDA(c, cudaGetLastError());
    c->step_valid = true;
    return SPH_OK;
}

// ---- CUDA-graph replay inside sph_step_n ------------------------------------------------------------------------
// A step is ~10 launches; at the reference's own scene sizes (10 k - 100 k particles) each kernel runs for a few
// microseconds and the host's launch rate is the limit.  sph_step_n therefore records the launch sequence of one
// step once (stream capture) and replays it: one cudaGraphLaunch per step.  The recording is valid for one StepKey
// only; a changed particle count / dt / parameter / list capacity / buffer falls back to a plain step, which also
// does whatever allocation the change needs, and the next step records again.  The last step of a call runs plainly
// so that the stage timers (getElapsedTime*) describe a real step.  SPH_GRAPH=0 turns the replay off.
static bool graphs_enabled()
{
    static const bool on = [] { const char* e = getenv("SPH_GRAPH"); return !(e && e[0] == '0'); }();
    return on;
}

static SphContext::StepKey step_key(const SphContext* c, float dt)
{
    SphContext::StepKey k;
    memset(&k, 0, sizeof(k));                       // padding too: keys are compared with memcmp
    k.n = c->n; k.dt = dt; k.params = c->params; k.extras = c->extras; k.mode = c->mode; k.list_k = c->list_k; k.list_k_alloc = c->list_k_alloc; k.tile_capn = c->tile_capn; k.two_level = c->table_two_level ? 1 : 0; k.timing = c->deep_stack ? 1 : 0;
    k.nc_tap = c->nc_tap ? 1 : 0; k.nlist = c->nlist; k.tstart = c->tstart; k.scan_tmp = c->scan_tmp; k.tend = c->tend;
    return k;
}

static void drop_graph(SphContext* c)
{
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
    c->graph_valid = false;
}

// records one step; on any failure the context simply keeps stepping without graphs
static bool record_step(SphContext* c, float dt)
{
    drop_graph(c);
    if (cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->graph_disabled = true; return false; }
    const uint64_t l0 = c->launches;
    c->capturing = true;
    const int rc = run_step(c, dt, true, false);      // no stage timers inside the recording (stage_event)
    c->capturing = false;
    const uint64_t per_step = c->launches - l0;
    c->launches = l0;                               // nothing ran yet
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->st, &g);
    if (rc != SPH_OK || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        c->graph_disabled = true;
        return false;
    }
    cudaGraphExec_t x = nullptr;
    if (cudaGraphInstantiate(&x, g, 0) != cudaSuccess || !x) {
        cudaGraphDestroy(g);
        cudaGetLastError();
        c->graph_disabled = true;
        return false;
    }
    c->graph = g; c->graph_exec = x; c->graph_launches = per_step;
    c->graph_key = step_key(c, dt);
    c->graph_valid = true;
    return true;
}

// One step: replayed as a CUDA graph when a recording for exactly this configuration exists, recorded when the
// configuration has just been stepped plainly with the same key (so a frame loop replays from its third frame on),
// plain otherwise -- a list or staging buffer that has to grow, or any change of configuration, takes the plain step,
// which also does the reallocation.
static int step_once(SphContext* c, float dt)
{
    if (c->nranks > 1) return multi_step(c, dt);
    const bool can = graphs_enabled() && !c->graph_disabled && !c->graph_user_off && c->n > 0;
    if (can) {
        update_stack_depth(c);
        const bool grow = (c->list_auto && c->list_k && c->h_overflow && *c->h_overflow > c->list_k) ||
                          (c->h_tile_need && *c->h_tile_need > c->tile_capn);
        const SphContext::StepKey k = step_key(c, dt);
        bool match = c->graph_valid && memcmp(&k, &c->graph_key, sizeof(k)) == 0;
        if (!grow && !match && c->step_valid && c->have_last_key && memcmp(&k, &c->last_key, sizeof(k)) == 0) match = record_step(c, dt);
        // the six stage timers (getElapsedTime*) are refreshed by a plain step every 16 steps while replays run
        const bool refresh = c->timing && c->replays_since_timed >= 15;
        if (!grow && match && !refresh) {
            SPH_CUDA(c, cudaSetDevice(c->device));
            SPH_CUDA(c, cudaGraphLaunch(c->graph_exec, c->st));
            c->launches += c->graph_launches;
            c->graph_replays++;
            c->replays_since_timed++;
            c->ncount_valid = true;
            c->step_valid = true;
            // (the stage events of the last plain step stay recorded: sph_get_timings keeps reporting that step)
            return SPH_OK;
        }
    }
    const int rc = run_step(c, dt, true);
    c->replays_since_timed = 0;
    if (can && rc == SPH_OK) { c->last_key = step_key(c, dt); c->have_last_key = true; }
    return rc;
}

int sph_step(SphContext* c, float dt)
{
    if (!c) return SPH_ERR_INVALID;
    return step_once(c, dt);
}

int sph_step_n(SphContext* c, float dt, uint32_t nsteps)
{
    if (!c) return SPH_ERR_INVALID;
    for (uint32_t i = 0; i < nsteps; i++) {
        const int rc = step_once(c, dt);
        if (rc != SPH_OK) return rc;
    }
    return SPH_OK;
}

int sph_refresh_densities(SphContext* c)
{
    if (!c) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_UNSUPPORTED, "sph_refresh_densities: single-GPU contexts only");
    return run_step(c, 0.0f, false);
}

int sph_synchronize(SphContext* c)
{
    if (!c) return SPH_ERR_INVALID;
    SPH_CUDA(c, cudaSetDevice(c->device));
    SPH_CUDA(c, cudaStreamSynchronize(c->st));
    if (c->st_fork) SPH_CUDA(c, cudaStreamSynchronize(c->st_fork));     // the list words of the last step have reached the host too
    return SPH_OK;
}

int sph_spawn_grid(SphContext* c, uint32_t n)
{   // InitializeData (:112-147) + GridArrangement (:518-557); host-side lattice, same fp32 expressions
    if (!c) return SPH_ERR_INVALID;
    if (n > c->cap) return fail(c, SPH_ERR_CAPACITY, "sph_spawn_grid: n exceeds capacity");
    std::vector<float> pos((size_t)n * 3, 0.0f);
    const int per_axis = (int)ceil(powf((float)(int)n, (1.0f / 3.0f)));
    const float gap = 0.215f;
    const float total = per_axis * gap;
    uint32_t i = 0;
    for (int ly = 0; ly < per_axis && i < n; ly++)
        for (int lx = 0; lx < per_axis && i < n; lx++)
            for (int lz = 0; lz < per_axis && i < n; lz++) {
                const float xo = lx * gap, yo = ly * gap, zo = lz * gap;
                const float wx = (0 - ((total - gap) / 2.0f));
                const float wy = (0 + (total - gap) / 2.0f);
                const float wz = (0 - (total - gap) / 2.0f);
                pos[3 * (size_t)i] = wx + xo; pos[3 * (size_t)i + 1] = wy - yo; pos[3 * (size_t)i + 2] = wz + zo;
                i++;
            }
    int rc = sph_upload_state(c, n, pos.data(), nullptr);
    if (rc != SPH_OK) return rc;
    rc = sph_refresh_densities(c);                       // :144-145
    if (rc != SPH_OK) return rc;
    return sph_synchronize(c);                           // pos is a local: the H2D copy must finish first
}

int sph_spawn_block(SphContext* c, const SphBlockSpawn* b)
{
    if (!c || !b) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_UNSUPPORTED, "sph_spawn_block: single-GPU contexts only (slab mode: sph_upload_owned)");
    const uint64_t n64 = (uint64_t)b->nx * b->ny * b->nz;
    if (n64 > c->cap) return fail(c, SPH_ERR_CAPACITY, "sph_spawn_block: nx*ny*nz exceeds capacity");
    if (!std::isfinite(b->gap) || !(b->gap > 0.0)) return fail(c, SPH_ERR_INVALID, "sph_spawn_block: gap must be > 0");
    for (int a = 0; a < 3; a++)
        if (!std::isfinite(b->origin[a])) return fail(c, SPH_ERR_INVALID, "sph_spawn_block: origin must be finite");
    SPH_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n = (uint32_t)n64;
    launch_spawn_block(c->st, c->A_pos, c->A_vel, b->nx, b->ny, b->nz, b->gap, b->origin, b->jitter_amp, b->velocity_scale,
                       b->seed, n, &c->launches);
    SPH_CUDA(c, cudaGetLastError());
    c->n = n;
    c->step_valid = false;
    c->ncount_valid = false;
    return SPH_OK;
}

static size_t field_bytes(int field, size_t n)
{
    switch (field) {
    case SPH_FIELD_POSITIONS: case SPH_FIELD_VELOCITIES: case SPH_FIELD_PREDICTED:
    case SPH_FIELD_VEL_AFTER_PRESSURE: case SPH_FIELD_VEL_AFTER_VISCOSITY: return n * 12;
    case SPH_FIELD_OUT_POSITIONS: case SPH_FIELD_COLORS: return n * 16;
    case SPH_FIELD_DENSITIES: return n * 8;
    case SPH_FIELD_HASH: case SPH_FIELD_KEY: case SPH_FIELD_NEIGHBOUR_COUNT: case SPH_FIELD_SPEED_NORMALIZED: return n * 4;
    default: return 0;
    }
}

}  // extern "C"

int sphb200::export_field(SphContext* c, int field, void* dev_out, bool by_id, uint32_t n)
{
    const void* src = nullptr;
    bool needs_step = true;
    switch (field) {
    case SPH_FIELD_POSITIONS: case SPH_FIELD_OUT_POSITIONS: src = c->A_pos; needs_step = false; break;
    case SPH_FIELD_VELOCITIES: case SPH_FIELD_SPEED_NORMALIZED: case SPH_FIELD_COLORS: src = c->A_vel; needs_step = false; break;
    case SPH_FIELD_DENSITIES: src = c->dens; break;
    case SPH_FIELD_PREDICTED: case SPH_FIELD_HASH: case SPH_FIELD_KEY: src = c->pred; break;
    case SPH_FIELD_VEL_AFTER_PRESSURE: src = c->velp; break;
    case SPH_FIELD_VEL_AFTER_VISCOSITY: src = c->S_vel; break;
    case SPH_FIELD_NEIGHBOUR_COUNT:
        if (!c->ncount_valid) return fail(c, SPH_ERR_INVALID, "neighbour counts not recorded: enable the tap before the step");
        src = c->ncount; break;
    default: return fail(c, SPH_ERR_INVALID, "unknown field");
    }
    if (needs_step && !c->step_valid) return fail(c, SPH_ERR_INVALID, "field needs a step (or spawn/refresh) first");
    DevParams P;
    make_dev_params(c, c->n, &P);
    launch_export(c->st, field, c->A_pos, src, nullptr, dev_out, n, P, by_id, &c->launches);
    SPH_CUDA(c, cudaGetLastError());
    return SPH_OK;
}

extern "C" {

int sph_download(SphContext* c, int field, void* host, size_t host_bytes)
{
    if (!c) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_INVALID, "slab mode: use sph_download_owned");
    const size_t need = field_bytes(field, c->n);
    if (need == 0 && c->n) return fail(c, SPH_ERR_INVALID, "unknown field");
    if (host_bytes < need) return fail(c, SPH_ERR_INVALID, "sph_download: host buffer too small");
    if (c->n == 0) return SPH_OK;
    if (!host) return fail(c, SPH_ERR_INVALID, "sph_download: host is NULL");
    SPH_CUDA(c, cudaSetDevice(c->device));
    int rc = export_field(c, field, c->stage, true, c->n);
    if (rc != SPH_OK) return rc;
    SPH_CUDA(c, cudaMemcpyAsync(host, c->stage, need, cudaMemcpyDeviceToHost, c->st));
    SPH_CUDA(c, cudaStreamSynchronize(c->st));
    return SPH_OK;
}

int sph_download_table(SphContext* c, int table, void* host, size_t host_bytes, size_t* out_len)
{
    if (!c) return SPH_ERR_INVALID;
    if (!c->step_valid) return fail(c, SPH_ERR_INVALID, "tables need a step (or spawn/refresh) first");
    SPH_CUDA(c, cudaSetDevice(c->device));
    const void* src = nullptr;
    size_t len = c->n;
    switch (table) {
    case SPH_TABLE_SORTED_INDEX:
        launch_export_ids(c->st, c->A_pos, (uint32_t*)c->stage, c->n, &c->launches);
        src = c->stage; break;
    case SPH_TABLE_SORTED_KEY: src = c->sorted_where ? c->key_b : c->key_a; break;
    case SPH_TABLE_SORTED_HASH: {
        int rc = export_field(c, SPH_FIELD_HASH, c->stage, false, c->n);
        if (rc != SPH_OK) return rc;
        src = c->stage; } break;
    case SPH_TABLE_START_INDICES:
        src = c->tstart;
        len = (c->mode == SPH_TABLE_GRID && !c->grid_too_large) ? (size_t)c->ncell + 1 : (size_t)c->n;
        if (c->mode == SPH_TABLE_GRID && !c->grid_too_large && host) {            // the flat prefix table, from its two levels
            if (host_bytes < len * 4) return fail(c, SPH_ERR_INVALID, "sph_download_table: host buffer too small");
            DevParams P;
            int rc = make_dev_params(c, c->n, &P);
            if (rc != SPH_OK) return rc;
            if (len * 4 > (size_t)c->cap * 32) {            // larger than the staging buffer: a temporary
                uint32_t* tmp = nullptr;
                SPH_CUDA(c, cudaMalloc(&tmp, len * 4));
                launch_table_flatten(c->st, c->tstart, P, tmp, (uint32_t)len, &c->launches);
                cudaError_t e = cudaMemcpyAsync(host, tmp, len * 4, cudaMemcpyDeviceToHost, c->st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
                cudaFree(tmp);
                if (e != cudaSuccess) return cuda_fail(c, e, "sph_download_table");
                if (out_len) *out_len = len;
                return SPH_OK;
            }
            launch_table_flatten(c->st, c->tstart, P, (uint32_t*)c->stage, (uint32_t)len, &c->launches);
            src = c->stage;
        }
        break;
    default: return fail(c, SPH_ERR_INVALID, "unknown table");
    }
    if (out_len) *out_len = len;
    if (!host) return SPH_OK;                                  // length query
    if (host_bytes < len * 4) return fail(c, SPH_ERR_INVALID, "sph_download_table: host buffer too small");
    if (len) SPH_CUDA(c, cudaMemcpyAsync(host, src, len * 4, cudaMemcpyDeviceToHost, c->st));
    SPH_CUDA(c, cudaStreamSynchronize(c->st));
    return SPH_OK;
}

int sph_get_particle(SphContext* c, uint32_t index, float* out10)
{
    if (!c || !out10) return SPH_ERR_INVALID;
    for (int i = 0; i < 10; i++) out10[i] = 0.0f;
    if (index >= c->n || c->nranks > 1) return SPH_OK;       // bounds check returns zeros (:151,157,163,168,174,180)
    SPH_CUDA(c, cudaSetDevice(c->device));
    float* d = (float*)c->stage;
    SPH_CUDA(c, cudaMemsetAsync(d, 0, 10 * sizeof(float), c->st));
    if (!c->row_of) SPH_CUDA(c, cudaMalloc(&c->row_of, (size_t)c->cap * sizeof(uint32_t)));
    const bool rebuild = c->row_of_stamp != c->launches;
    launch_find_particle(c->st, c->A_pos, c->A_vel, c->step_valid ? c->dens : nullptr, c->n, index, d, c->row_of, rebuild, &c->launches);
    c->row_of_stamp = c->launches;
    SPH_CUDA(c, cudaMemcpyAsync(out10, d, 10 * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    SPH_CUDA(c, cudaStreamSynchronize(c->st));
    return SPH_OK;
}

int sph_save_state(SphContext* c, const char* path)
{
    if (!c || !path) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_UNSUPPORTED, "sph_save_state: single-GPU contexts only (gather with sph_download_owned)");
    std::vector<float> pos((size_t)c->n * 3), vel((size_t)c->n * 3);
    int rc = sph_download(c, SPH_FIELD_POSITIONS, pos.data(), pos.size() * 4);
    if (rc != SPH_OK) return rc;
    rc = sph_download(c, SPH_FIELD_VELOCITIES, vel.data(), vel.size() * 4);
    if (rc != SPH_OK) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(c, SPH_ERR_INVALID, std::string("sph_save_state: cannot open ") + path);
    const char magic[8] = {'S', 'P', 'H', 'B', '2', '0', '0', '2'};      // "...1" had no sizeof(SphParams) word
    const uint32_t n = c->n, psize = (uint32_t)sizeof(SphParams);
    bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&n, 4, 1, f) == 1 && fwrite(&psize, 4, 1, f) == 1 &&
              fwrite(&c->params, sizeof(SphParams), 1, f) == 1;
    ok = ok && (n == 0 || (fwrite(pos.data(), 12, n, f) == n && fwrite(vel.data(), 12, n, f) == n));
    ok = (fclose(f) == 0) && ok;
    return ok ? SPH_OK : fail(c, SPH_ERR_INVALID, "sph_save_state: short write");
}

int sph_load_state(SphContext* c, const char* path)
{
    if (!c || !path) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_UNSUPPORTED, "sph_load_state: single-GPU contexts only");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(c, SPH_ERR_INVALID, std::string("sph_load_state: cannot open ") + path);
    char magic[8];
    uint32_t n = 0;
    SphParams p;
    bool ok = fread(magic, 1, 8, f) == 8 && fread(&n, 4, 1, f) == 1;
    const bool v2 = ok && memcmp(magic, "SPHB2002", 8) == 0;
    ok = ok && (v2 || memcmp(magic, "SPHB2001", 8) == 0);
    uint32_t psize = (uint32_t)sizeof(SphParams);
    if (ok && v2) ok = fread(&psize, 4, 1, f) == 1 && psize == (uint32_t)sizeof(SphParams);
    ok = ok && fread(&p, sizeof(SphParams), 1, f) == 1;
    if (!ok) { fclose(f); return fail(c, SPH_ERR_INVALID, "sph_load_state: not a snapshot file (or one of another ABI version)"); }
    if (n > c->cap) { fclose(f); return fail(c, SPH_ERR_CAPACITY, "sph_load_state: snapshot exceeds capacity"); }
    std::vector<float> pos((size_t)n * 3), vel((size_t)n * 3);
    ok = n == 0 || (fread(pos.data(), 12, n, f) == n && fread(vel.data(), 12, n, f) == n);
    fclose(f);
    if (!ok) return fail(c, SPH_ERR_INVALID, "sph_load_state: truncated snapshot");
    int rc = sph_set_params(c, &p);
    if (rc != SPH_OK) return rc;
    rc = sph_upload_state(c, n, pos.data(), vel.data());
    if (rc != SPH_OK) return rc;
    return sph_synchronize(c);                          // the host vectors die here
}

int sph_host_register(void* ptr, size_t bytes)
{
    if (!ptr || !bytes) return SPH_ERR_INVALID;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped)   /* pinned for every device's context and mapped into their address spaces (multi-GPU hosts, sph_download_owned_scatter) */;
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, SPH_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
    return SPH_OK;
}

int sph_host_unregister(void* ptr)
{
    if (!ptr) return SPH_ERR_INVALID;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, SPH_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e)); }
    return SPH_OK;
}

int sph_get_timings(SphContext* c, double* out6)
{
    if (!c || !out6) return SPH_ERR_INVALID;
    if (c->ev_recorded) {
        SPH_CUDA(c, cudaSetDevice(c->device));
        SPH_CUDA(c, cudaEventSynchronize(c->ev[6]));
        for (int i = 0; i < 6; i++) {
            float ms = 0;
            SPH_CUDA(c, cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
            c->timings[i] = ms;
        }
    }
    for (int i = 0; i < 6; i++) out6[i] = c->timings[i];
    return SPH_OK;
}

}  // extern "C"

// ---- pipelined transfers ---------------------------------------------------------------------------
// The blocking pair sph_upload_state / sph_download serialises PCIe and compute on one stream: at 1 M particles a
// frame costs 0.44 ms (H2D 24 MB) + 0.37 ms (step) + 0.29 ms (D2H 16 MB).  With their own staging buffers and copy
// streams the three overlap -- the upload of frame k+1 and the download of frame k-1 travel while frame k is
// computed -- and the frame time drops to the longest of the three.  Orderings (events, no host waits except
// sph_download_wait):
//   st_in : [wait ev_pack: the previous pack has consumed stage_in]  H2D -> stage_in                    record ev_h2d
//   st    : [wait ev_h2d] pack stage_in -> state   record ev_pack   ...step(s)...  export -> stage_out  record ev_export
//   st_out: [wait ev_export] D2H stage_out -> host                                                      record ev_d2h
int sphb200::ensure_pipeline(SphContext* c)
{
    if (c->st_in) return SPH_OK;
    cudaStream_t si = nullptr, so = nullptr;
    SPH_CUDA(c, cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking));
    cudaError_t e = cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cudaStreamDestroy(si); return cuda_fail(c, e, "cudaStreamCreate (copy-out)"); }
    cudaEvent_t* evs[4] = {&c->ev_h2d, &c->ev_pack, &c->ev_export, &c->ev_d2h};
    for (cudaEvent_t* ev : evs) {
        e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e != cudaSuccess) { cudaStreamDestroy(si); cudaStreamDestroy(so); return cuda_fail(c, e, "cudaEventCreate (pipeline)"); }
    }
    c->st_in = si; c->st_out = so;
    return SPH_OK;
}

extern "C" {

int sph_upload_state_begin(SphContext* c, uint32_t n, const float* pos3, const float* vel3)
{
    if (!c) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_UNSUPPORTED, "sph_upload_state_begin: single-GPU contexts only");
    if (n > c->cap) return fail(c, SPH_ERR_CAPACITY, "sph_upload_state_begin: n exceeds capacity");
    if (n && !pos3) return fail(c, SPH_ERR_INVALID, "sph_upload_state_begin: pos3 is NULL");
    if (c->upload_pending) return fail(c, SPH_ERR_INVALID, "sph_upload_state_begin: an upload is already pending (commit it first)");
    SPH_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_pipeline(c);
    if (rc != SPH_OK) return rc;
    if (!c->stage_in) SPH_CUDA(c, cudaMalloc((void**)&c->stage_in, (size_t)c->cap * 28));
    // the previous upload's pack kernel reads stage_in on the solver stream: do not overwrite it before that ran
    if (c->pack_recorded) SPH_CUDA(c, cudaStreamWaitEvent(c->st_in, c->ev_pack, 0));
    if (n) {
        SPH_CUDA(c, cudaMemcpyAsync(c->stage_in, pos3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st_in));
        if (vel3) SPH_CUDA(c, cudaMemcpyAsync(c->stage_in + (size_t)c->cap * 12, vel3, (size_t)n * 12, cudaMemcpyHostToDevice, c->st_in));
    }
    SPH_CUDA(c, cudaEventRecord(c->ev_h2d, c->st_in));
    c->upload_pending = true;
    c->upload_has_vel = vel3 != nullptr;
    c->upload_has_ids = false;
    c->upload_n = n;
    return SPH_OK;
}

int sph_upload_state_commit(SphContext* c)
{
    if (!c) return SPH_ERR_INVALID;
    if (!c->upload_pending) return fail(c, SPH_ERR_INVALID, "sph_upload_state_commit: no upload pending");
    SPH_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n = c->upload_n;
    SPH_CUDA(c, cudaStreamWaitEvent(c->st, c->ev_h2d, 0));
    if (n) {
        const float* dpos = (const float*)c->stage_in;
        const float* dvel = (const float*)(c->stage_in + (size_t)c->cap * 12);
        const uint32_t* dids = (const uint32_t*)(c->stage_in + (size_t)c->cap * 24);
        launch_pack_state(c->st, dpos, c->upload_has_vel ? dvel : nullptr, c->upload_has_ids ? dids : nullptr, c->A_pos, c->A_vel, n,
                          &c->launches);
        SPH_CUDA(c, cudaGetLastError());
    }
    SPH_CUDA(c, cudaEventRecord(c->ev_pack, c->st));
    c->pack_recorded = true;
    c->upload_pending = false;
    c->n = n;
    c->step_valid = false;
    c->ncount_valid = false;
    multi_adopt_upload(c, n);
    return SPH_OK;
}

int sph_download_begin(SphContext* c, int field, void* host, size_t host_bytes)
{
    if (!c) return SPH_ERR_INVALID;
    if (c->nranks > 1) return fail(c, SPH_ERR_UNSUPPORTED, "sph_download_begin: single-GPU contexts only");
    if (c->download_pending) return fail(c, SPH_ERR_INVALID, "sph_download_begin: a download is already pending (wait for it first)");
    const size_t need = field_bytes(field, c->n);
    if (need == 0 && c->n) return fail(c, SPH_ERR_INVALID, "unknown field");
    if (host_bytes < need) return fail(c, SPH_ERR_INVALID, "sph_download_begin: host buffer too small");
    if (c->n && !host) return fail(c, SPH_ERR_INVALID, "sph_download_begin: host is NULL");
    SPH_CUDA(c, cudaSetDevice(c->device));
    int rc = ensure_pipeline(c);
    if (rc != SPH_OK) return rc;
    if (!c->stage_out) SPH_CUDA(c, cudaMalloc((void**)&c->stage_out, (size_t)c->cap * 20));
    if (c->n) {
        // the previous download was waited for (download_pending is false), so stage_out is free
        rc = export_field(c, field, c->stage_out, true, c->n);
        if (rc != SPH_OK) return rc;
        SPH_CUDA(c, cudaEventRecord(c->ev_export, c->st));
        SPH_CUDA(c, cudaStreamWaitEvent(c->st_out, c->ev_export, 0));
        SPH_CUDA(c, cudaMemcpyAsync(host, c->stage_out, need, cudaMemcpyDeviceToHost, c->st_out));
    }
    SPH_CUDA(c, cudaEventRecord(c->ev_d2h, c->st_out));
    c->download_pending = true;
    return SPH_OK;
}

int sph_download_wait(SphContext* c)
{
    if (!c) return SPH_ERR_INVALID;
    if (!c->download_pending) return fail(c, SPH_ERR_INVALID, "sph_download_wait: no download pending");
    SPH_CUDA(c, cudaSetDevice(c->device));
    c->download_pending = false;
    SPH_CUDA(c, cudaEventSynchronize(c->ev_d2h));
    return SPH_OK;
}

}  // extern "C"
