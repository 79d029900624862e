// sph_context.h -- the object behind the opaque SphContext handle.
#pragma once
#include <string>
#include <vector>
#include "sph_internal.h"

struct ncclComm;

struct SphContext {
    int device = 0;
    uint32_t cap = 0;            // row capacity of every per-particle array
    uint32_t n = 0;              // rows in use (single GPU: particles; slab mode: owned particles)
    SphParams params;
    SphExtras extras = {{0.0f, 0.0f, 0.0f, 1.0f}, 0.0f, 0.0f};
    float rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // rotation matrix of extras.bound_rotation (world = R * local)
    bool extras_on = false;
    int mode = SPH_TABLE_GRID;
    bool timing = true;
    bool nc_tap = false;
    cudaStream_t st = nullptr;

    // persistent state, device order
    float4 *A_pos = nullptr, *A_vel = nullptr;
    // per-step arrays, sorted order of the step
    float4 *S_pos = nullptr, *S_vel = nullptr, *pred = nullptr;
    float4* predpk = nullptr;        // predicted positions, pair-interleaved (sph_internal.h: PredPair), read by the density pass
    float4* velp = nullptr;          // v' = velocity after the pressure pass (snapshot read by the viscosity pass)
    sphb200::Rec8* dens = nullptr;   // density records (sph_internal.h: Rec8)
    uint32_t *key_a = nullptr, *key_b = nullptr, *perm_a = nullptr, *perm_b = nullptr;
    uint32_t* ncount = nullptr;  // neighbour count incl. self of every row (density pass); also the list lengths
    uint32_t* lcount = nullptr;  // list lengths
    uint32_t* nlist = nullptr;   // neighbour list, k-major: entry k of row i at nlist[k * cap + i]
    uint32_t list_k = 64;        // entries per row (0: lists off)
    uint32_t list_k_alloc = 0;
    bool list_auto = true;       // grow list_k when the density pass reports overflowing particles
    uint32_t* d_overflow = nullptr;  // device word written by the density kernel (longest list that did not fit)
    uint32_t* h_overflow = nullptr;  // pinned mirror (4 words, see ensure_list), refreshed asynchronously after every density pass
    uint64_t rows_warps_seen = 0;    // (warps << 32 | list rows) of the density passes already accounted for
    bool deep_stack = false;         // density pass: deep survivor stack (lists are long throughout)
    uint32_t* row_of = nullptr;           // id -> row of the device order, built on demand by sph_get_particle
    uint64_t row_of_stamp = ~0ull;        // c->launches when it was built: any kernel since then may have changed the order
    uint32_t* d_noncanonical = nullptr;   // device counter, see DevParams::noncanonical
    uint32_t tile_capn = 0;          // tile generation: staged candidates per warp (0: not initialised yet)
    uint32_t* d_tile_need = nullptr; // device word: largest single-cell neighbourhood that did not fit
    uint32_t* h_tile_need = nullptr; // pinned mirror, refreshed with h_overflow
    uint32_t *tstart = nullptr, *tend = nullptr;
    size_t table_cap = 0;        // entries allocated for tstart (tend has cap entries, hash mode only)
    uint32_t table_ncell = 0;     // cells the table in tstart was laid out for
    uint32_t table_seg_off = 0;   // ... and the word offset of its segment bases (tbl_at)
    bool table_two_level = false; // tstart holds a consistent two-level GRID table (cells of clean segments are zero): the next
                                  // counting-sort build clears only what the last one touched
    uint32_t* scan_tmp = nullptr; // block sums of the table scan (counting sort)
    size_t scan_cap = 0;
    uint32_t* gap_list = nullptr;
    size_t gap_cap = 0;
    uint32_t* counts = nullptr;  // radix sort digit matrix
    size_t counts_cap = 0;
    unsigned char* stage = nullptr;   // device staging for upload / export (cap * 32 B)

    // pipelined transfers (sph_upload_state_begin / _commit, sph_download_begin / _wait): their own staging
    // buffers and copy streams, so a PCIe copy in either direction overlaps the step on `st`
    cudaStream_t st_fork = nullptr;  // second branch of the recorded step (segment scan beside the in-segment scan)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t st_in = nullptr, st_out = nullptr;
    unsigned char* stage_in = nullptr;    // cap * 28 B: pos3 | vel3 | global ids (slab mode) of the pending upload
    unsigned char* stage_out = nullptr;   // cap * 20 B: the exported field (+ global ids, slab mode) of the pending download
    cudaEvent_t ev_h2d = nullptr, ev_pack = nullptr, ev_export = nullptr, ev_d2h = nullptr;
    bool upload_pending = false, upload_has_vel = false, upload_has_ids = false, pack_recorded = false, download_pending = false;
    uint32_t upload_n = 0;

    // CUDA-graph replay of the step inside sph_step_n (launch-bound at the reference's own scene sizes): the captured
    // launch sequence is valid for exactly one configuration (StepKey); anything that changes it falls back to a
    // plain step and re-captures
    struct StepKey {
        uint32_t n; float dt; SphParams params; SphExtras extras; int mode; uint32_t list_k, list_k_alloc, tile_capn; int nc_tap, two_level, timing;   /* `timing` carries the deep-stack flag (the recording holds no timers) */
        const void *nlist, *tstart, *scan_tmp, *tend;
    };
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    StepKey graph_key = {};
    StepKey last_key = {};           // configuration of the last plain step: the same key again records the graph
    bool have_last_key = false;
    bool graph_valid = false;
    bool graph_user_off = false;     // sph_set_graph_replay(ctx, 0)
    bool graph_disabled = false;     // SPH_GRAPH=0, or a capture failed once on this context
    bool capturing = false;          // run_step is being recorded: no allocation, no list growth
    uint64_t graph_launches = 0;     // kernels per replay
    uint64_t graph_replays = 0;
    uint32_t replays_since_timed = 0; // replays since the last plain step (which refreshes the stage timers)
    int sorted_where = 0;        // 0: sorted keys in key_a, 1: key_b
    bool step_valid = false;     // per-step arrays describe the current device order
    bool ncount_valid = false;

    cudaEvent_t ev[7] = {};
    bool ev_recorded = false;
    double timings[6] = {0, 0, 0, 0, 0, 0};
    uint64_t launches = 0;
    std::string err;

    // geometry of the GRID table, derived from params
    int gmin[3] = {0, 0, 0};
    int gdim[3] = {1, 1, 1};
    uint32_t ncell = 1;
    int xsub = 8;                // x subdivision of the GRID table cells in force (power of two)
    int xsub_pref = 8;           // ... the preferred one (SPH_XSUB overrides); coarsened when the grid would get too large
    bool grid_too_large = false; // even the unsubdivided grid exceeds the table limit: single-GPU steps run on the
                                 // REFERENCE_HASH table (whose size is the particle count) until the geometry fits again

    // slab-decomposed multi-GPU (sph_multi.cu)
    ncclComm* comm = nullptr;
    int rank = 0, nranks = 1;
    std::vector<float> planes;   // nranks + 1 z planes
    struct SlabState* slab = nullptr;
};

namespace sphb200 {

int fail(SphContext* c, int code, const std::string& msg);
int cuda_fail(SphContext* c, cudaError_t e, const char* what);
#define SPH_CUDA(c, call)                                                      \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return sphb200::cuda_fail((c), e__, #call);    \
    } while (0)

// derive kernel parameters (constants with the reference's fp32 expressions, grid geometry)
int make_dev_params(SphContext* c, uint32_t n, DevParams* P);
int ensure_tables(SphContext* c, const DevParams& P);
int export_field(SphContext* c, int field, void* dev_out, bool by_id, uint32_t n);
int ensure_list(SphContext* c, NbrList* L);
int copy_list_words(SphContext* c, cudaStream_t after);
bool counting_sort_enabled();
int ensure_pipeline(SphContext* c);      // copy streams + events of the pipelined transfers (sph_api.cu)

// sph_multi.cu
int multi_step(SphContext* c, float dt);
void multi_teardown(SphContext* c);
int multi_params_changed(SphContext* c);              // slab mode: re-derive the cell layers of the planes after sph_set_params
void multi_adopt_upload(SphContext* c, uint32_t n);   // slab mode: the owned rows are [0, n) again after an upload

}  // namespace sphb200
