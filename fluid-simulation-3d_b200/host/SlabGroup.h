// host/SlabGroup.h -- several GPUs of one box behind one solver object, in ONE process.
//
// The reference is a single-process application; its solver class cannot be spread over processes
// without changing the application.  SlabGroup keeps the process model: it owns one C-ABI context per GPU
// (include/sph_b200.h, slab mode), drives each from its own host thread -- the slab step exchanges ghosts and
// migrants with its neighbours over NCCL, so all ranks must step at the same time -- and presents the particles
// in the reference's index order again (every particle carries its index as a global id).  FluidSimulation uses it
// when setDevices() names more than one GPU.  Pure host C++ on top of the C ABI, like the rest of host/.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sph_b200.h"

namespace sphb200 {

class SlabGroup {
public:
    // one slab per device, in the order given (slab k is the k-th z range); `particles` sizes the per-rank capacity
    SlabGroup(const std::vector<int>& devices, uint32_t particles, const SphParams& params);
    ~SlabGroup();
    SlabGroup(const SlabGroup&) = delete;
    SlabGroup& operator=(const SlabGroup&) = delete;

    int ranks() const { return (int)rank_.size(); }
    uint32_t particles() const { return n_; }
    void setParams(const SphParams& p);
    void setExtras(const SphExtras& e);          // all ranks or none, like setParams
    // Replace the whole state (index order, velocities may be null = zero).  Cuts the slabs at the particle-count
    // quantiles of the z cell layers (sph_slab_balance_layers) and hands every rank its particles.
    void upload(uint32_t n, const float* pos3, const float* vel3);
    void step(float dt, uint32_t nsteps = 1);                     // all ranks in lockstep
    // COLLECTIVE sph_comm_rebalance on every rank; returns whether a plane moved
    bool rebalance(uint32_t max_shift = 1);
    // gather a field into index order; `out` holds particles() elements of the field's size (sph_b200.h: SPH_FIELD_*)
    void download(int field, void* out, size_t out_bytes);
    // Opt-in: downloads into page-locked, device-mapped arrays (sph_host_register) let every rank's export kernel write
    // its rows straight to out[id] (sph_download_owned_scatter): no staging copy, no host scatter.  An array that is
    // not mapped falls back to the staged path for that call.
    void setDirectScatter(bool on) { direct_ = on; }
    void timings(double out6[6]);                                 // per stage: the slowest rank
    std::vector<int32_t> layers() const;
    std::vector<uint32_t> ownedCounts() const;
    uint64_t launches() const;
    // Where the host time of the calls since the last resetBreakdown() went, in ms, slowest rank per phase:
    // [0] step (sph_step: enqueue, the slab exchanges' host synchronisations), [1] waiting for the device + export +
    // device-to-host copies of the downloads, [2] scattering the rows into index order, [3] re-balancing
    void breakdown(double out4[4]) const;
    void resetBreakdown();

private:
    struct Rank {
        SphContext* ctx = nullptr;
        int device = 0;
        std::vector<uint32_t> ids;          // ids of the owned rows, device order (page-locked; fetched once per state)
        std::vector<unsigned char> buf;     // field rows of the current download (page-locked)
        uint32_t owned = 0;                 // rows behind `ids`
        bool idsFresh = false;
        double ms[4] = {0, 0, 0, 0};        // host time per phase (breakdown())
        std::vector<uint32_t> upIds;        // scratch of upload()
        std::vector<float> pos, vel;
    };
    // f(rank) on every rank's own worker thread at once; returns when all are done and rethrows the first error.
    // The workers live as long as the group (a step is a few hundred microseconds: no thread start per call).
    void parallel(const std::function<void(int)>& f);
    void workerLoop(int k);
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cvWork_, cvDone_;
    const std::function<void(int)>* job_ = nullptr;
    uint64_t generation_ = 0;
    int pending_ = 0;
    bool quit_ = false;
    std::vector<std::string> err_;
    static void check(SphContext* c, int rc, const char* what);
    static size_t fieldBytes(int field);
    std::vector<Rank> rank_;
    SphParams params_;
    uint32_t n_ = 0;
    uint32_t cap_ = 0;
    bool direct_ = false;
};

}  // namespace sphb200
