// host/SlabGroup.cc -- see SlabGroup.h.  Pure host C++: talks to the GPUs only through the C ABI of libsph_b200.so.
#include "SlabGroup.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace sphb200 {

static double nowMs()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void SlabGroup::check(SphContext* c, int rc, const char* what)
{
    if (rc == SPH_OK) return;
    const char* msg = sph_last_error(c);
    throw std::runtime_error(std::string(what) + ": " + (msg && *msg ? msg : "error " + std::to_string(rc)));
}

size_t SlabGroup::fieldBytes(int field)
{
    switch (field) {
    case SPH_FIELD_POSITIONS: case SPH_FIELD_VELOCITIES: case SPH_FIELD_PREDICTED: case SPH_FIELD_VEL_AFTER_PRESSURE:
    case SPH_FIELD_VEL_AFTER_VISCOSITY: return 12;
    case SPH_FIELD_OUT_POSITIONS: case SPH_FIELD_COLORS: return 16;
    case SPH_FIELD_DENSITIES: return 8;
    case SPH_FIELD_HASH: case SPH_FIELD_KEY: case SPH_FIELD_NEIGHBOUR_COUNT: case SPH_FIELD_SPEED_NORMALIZED: return 4;
    default: return 0;
    }
}

void SlabGroup::workerLoop(int k)
{
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(int)>* job;
        {
            std::unique_lock<std::mutex> lock(mu_);
            cvWork_.wait(lock, [&] { return quit_ || generation_ != seen; });
            if (quit_) return;
            seen = generation_;
            job = job_;
        }
        std::string err;
        try { (*job)(k); } catch (const std::exception& e) { err = e.what()[0] ? e.what() : "error"; } catch (...) { err = "unknown exception"; }
        {
            std::lock_guard<std::mutex> lock(mu_);
            err_[(size_t)k] = err;
            if (--pending_ == 0) cvDone_.notify_all();
        }
    }
}

void SlabGroup::parallel(const std::function<void(int)>& f)
{
    const int R = ranks();
    {
        std::unique_lock<std::mutex> lock(mu_);
        job_ = &f;
        pending_ = R;
        generation_++;
        cvWork_.notify_all();
        cvDone_.wait(lock, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
    for (int k = 0; k < R; k++)
        if (!err_[(size_t)k].empty()) throw std::runtime_error("rank " + std::to_string(k) + ": " + err_[(size_t)k]);
}

SlabGroup::SlabGroup(const std::vector<int>& devices, uint32_t particles, const SphParams& params) : params_(params)
{
    const size_t R = devices.size();
    if (R == 0) throw std::runtime_error("SlabGroup: no devices");
    // a rank holds its owned rows, the rows arriving in a step and two ghost layers; slabs are cut at particle-count
    // quantiles, so 1.5x the even share plus slack for thin scenes is ample (re-balancing keeps it that way)
    const uint64_t share = (uint64_t)particles / R * 3 / 2 + 131072;
    const uint64_t all = (uint64_t)particles + 4096;
    cap_ = (uint32_t)(R == 1 ? (particles ? particles : 1) : (share < all ? share : all));
    rank_.resize(R);
    err_.resize(R);
    for (size_t k = 0; k < R; k++) workers_.emplace_back(&SlabGroup::workerLoop, this, (int)k);
    auto stopWorkers = [&] {
        { std::lock_guard<std::mutex> lock(mu_); quit_ = true; }
        cvWork_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
    };
    try {
        for (size_t k = 0; k < R; k++) {
            rank_[k].device = devices[k];
            const int rc = sph_create(&rank_[k].ctx, devices[k], cap_);
            if (rc != SPH_OK) {
                const char* msg = sph_last_error(nullptr);
                throw std::runtime_error(std::string("sph_create on device ") + std::to_string(devices[k]) + ": " + (msg ? msg : "?"));
            }
            check(rank_[k].ctx, sph_set_params(rank_[k].ctx, &params_), "sph_set_params");
            // download scratch, page-locked once (a pageable D2H copy runs at a tenth of the PCIe rate)
            rank_[k].ids.resize(cap_);
            rank_[k].buf.resize((size_t)cap_ * 16);
            sph_host_register(rank_[k].ids.data(), rank_[k].ids.size() * 4);
            sph_host_register(rank_[k].buf.data(), rank_[k].buf.size());
        }
        std::vector<unsigned char> id(sph_comm_id_bytes());
        check(nullptr, sph_comm_get_id(id.data(), id.size()), "sph_comm_get_id");
        // ncclCommInitRank blocks until every rank has joined: one thread per rank
        parallel([&](int k) { check(rank_[(size_t)k].ctx, sph_comm_init(rank_[(size_t)k].ctx, k, (int)R, id.data(), id.size()), "sph_comm_init"); });
    } catch (...) {
        stopWorkers();
        for (auto& r : rank_) {
            if (!r.ids.empty()) sph_host_unregister(r.ids.data());
            if (!r.buf.empty()) sph_host_unregister(r.buf.data());
            if (r.ctx) sph_destroy(r.ctx);
        }
        throw;
    }
}

SlabGroup::~SlabGroup()
{
    // communicators are torn down together (ncclCommDestroy may wait for the peers)
    for (auto& r : rank_) {
        if (!r.ids.empty()) sph_host_unregister(r.ids.data());
        if (!r.buf.empty()) sph_host_unregister(r.buf.data());
    }
    try { parallel([&](int k) { if (rank_[(size_t)k].ctx) sph_destroy(rank_[(size_t)k].ctx); }); } catch (...) {}
    { std::lock_guard<std::mutex> lock(mu_); quit_ = true; }
    cvWork_.notify_all();
    for (auto& t : workers_) t.join();
}

void SlabGroup::setParams(const SphParams& p)
{   // all ranks or none: ranks that disagree on the grid geometry would disagree on the ghost layers of the next step
    const SphParams old = params_;
    for (size_t k = 0; k < rank_.size(); k++) {
        if (sph_set_params(rank_[k].ctx, &p) == SPH_OK) continue;
        const char* msg = sph_last_error(rank_[k].ctx);
        const std::string why = std::string("sph_set_params (rank ") + std::to_string(k) + "): " + (msg ? msg : "unknown error");
        for (size_t j = 0; j < k; j++) sph_set_params(rank_[j].ctx, &old);      // the failing rank restored itself
        throw std::runtime_error(why);
    }
    params_ = p;
}

void SlabGroup::setExtras(const SphExtras& e)
{
    std::vector<SphExtras> old(rank_.size());
    for (size_t k = 0; k < rank_.size(); k++) {
        sph_get_extras(rank_[k].ctx, &old[k]);
        if (sph_set_extras(rank_[k].ctx, &e) == SPH_OK) continue;
        const char* msg = sph_last_error(rank_[k].ctx);
        const std::string why = std::string("sph_set_extras (rank ") + std::to_string(k) + "): " + (msg ? msg : "unknown error");
        for (size_t j = 0; j < k; j++) sph_set_extras(rank_[j].ctx, &old[j]);
        throw std::runtime_error(why);
    }
}

void SlabGroup::upload(uint32_t n, const float* pos3, const float* vel3)
{
    const int R = ranks();
    if (n && !pos3) throw std::runtime_error("SlabGroup::upload: positions missing");
    int32_t dims[3], origin[3];
    check(rank_[0].ctx, sph_get_grid(rank_[0].ctx, dims, origin), "sph_get_grid");
    const int gz = dims[2], gmin_z = origin[2];
    const float r = params_.interaction_radius;
    // z cell layer of every particle, with the device's arithmetic: floor(z / r) in fp32 (sph_device.cuh: cell_of)
    std::vector<int32_t> layer((size_t)n);
    std::vector<uint32_t> hist((size_t)gz, 0u);
    for (uint32_t i = 0; i < n; i++) {
        int l = (int)std::floor(pos3[3 * (size_t)i + 2] / r) - gmin_z;
        l = l < 0 ? 0 : (l > gz - 1 ? gz - 1 : l);
        layer[i] = l;
        hist[(size_t)l]++;
    }
    std::vector<int32_t> L((size_t)R + 1);
    if (sph_slab_balance_layers(hist.data(), gz, R, nullptr, 0, 0, L.data()) != SPH_OK)
        throw std::runtime_error("SlabGroup::upload: the box has fewer than three cell layers per GPU along z (" +
                                 std::to_string(gz) + " layers, " + std::to_string(R) + " GPUs)");
    std::vector<float> planes((size_t)R + 1);
    for (int k = 0; k <= R; k++) planes[(size_t)k] = ((float)(L[(size_t)k] + gmin_z) + 0.5f) * r;   // a z inside the slab's first layer
    std::vector<int> owner_of_layer((size_t)gz);
    for (int k = 0; k < R; k++)
        for (int l = L[(size_t)k]; l < L[(size_t)k + 1]; l++) owner_of_layer[(size_t)l] = k;
    for (auto& rk : rank_) { rk.upIds.clear(); rk.pos.clear(); rk.vel.clear(); rk.idsFresh = false; }
    for (uint32_t i = 0; i < n; i++) {
        Rank& rk = rank_[(size_t)owner_of_layer[(size_t)layer[i]]];
        rk.upIds.push_back(i);
        rk.pos.insert(rk.pos.end(), pos3 + 3 * (size_t)i, pos3 + 3 * (size_t)i + 3);
        if (vel3) rk.vel.insert(rk.vel.end(), vel3 + 3 * (size_t)i, vel3 + 3 * (size_t)i + 3);
    }
    for (int k = 0; k < R; k++)
        if (rank_[(size_t)k].upIds.size() > cap_)
            throw std::runtime_error("SlabGroup::upload: rank " + std::to_string(k) + " would own " +
                                     std::to_string(rank_[(size_t)k].upIds.size()) + " particles, capacity " + std::to_string(cap_));
    parallel([&](int k) {
        Rank& rk = rank_[(size_t)k];
        check(rk.ctx, sph_comm_set_planes(rk.ctx, planes.data()), "sph_comm_set_planes");
        check(rk.ctx, sph_upload_owned(rk.ctx, (uint32_t)rk.upIds.size(), rk.upIds.data(), rk.pos.data(), vel3 ? rk.vel.data() : nullptr),
              "sph_upload_owned");
        std::vector<uint32_t>().swap(rk.upIds);
        std::vector<float>().swap(rk.pos);
        std::vector<float>().swap(rk.vel);
    });
    n_ = n;
}

void SlabGroup::step(float dt, uint32_t nsteps)
{
    parallel([&](int k) {
        SphContext* c = rank_[(size_t)k].ctx;
        rank_[(size_t)k].idsFresh = false;               // the step re-sorts the rows and may migrate some
        const double t0 = nowMs();
        if (nsteps <= 1) check(c, sph_step(c, dt), "sph_step");
        else check(c, sph_step_n(c, dt, nsteps), "sph_step_n");
        rank_[(size_t)k].ms[0] += nowMs() - t0;
    });
}

bool SlabGroup::rebalance(uint32_t max_shift)
{
    std::vector<int> changed((size_t)ranks(), 0);
    parallel([&](int k) {
        SphContext* c = rank_[(size_t)k].ctx;
        const double t0 = nowMs();
        check(c, sph_comm_rebalance(c, max_shift, nullptr, nullptr, 0, &changed[(size_t)k]), "sph_comm_rebalance");
        rank_[(size_t)k].ms[3] += nowMs() - t0;
    });
    return changed[0] != 0;
}

void SlabGroup::download(int field, void* out, size_t out_bytes)
{
    const size_t per = fieldBytes(field);
    if (!per) throw std::runtime_error("SlabGroup::download: unknown field");
    if (out_bytes < per * (size_t)n_) throw std::runtime_error("SlabGroup::download: output buffer too small");
    std::vector<uint32_t> got((size_t)ranks(), 0u);
    unsigned char* dst = static_cast<unsigned char*>(out);
    if (direct_) {
        // every rank's export kernel writes its rows to out[id] itself; any failure (the array is not device-mapped)
        // sends this call down the staged path below, which overwrites whatever was written
        std::vector<int> rc((size_t)ranks(), SPH_OK);
        parallel([&](int k) {
            Rank& rk = rank_[(size_t)k];
            const double t0 = nowMs();
            rc[(size_t)k] = sph_download_owned_scatter(rk.ctx, field, out, n_, &got[(size_t)k]);
            rk.ms[1] += nowMs() - t0;
        });
        bool ok = true;
        uint64_t total = 0;
        for (int k = 0; k < ranks(); k++) { ok = ok && rc[(size_t)k] == SPH_OK; total += got[(size_t)k]; }
        if (ok && total == n_) return;
    }
    parallel([&](int k) {
        Rank& rk = rank_[(size_t)k];
        const uint32_t m = sph_num_particles(rk.ctx);
        if (m > cap_) throw std::runtime_error("sph_num_particles exceeds the rank's capacity");
        uint32_t cnt = 0;
        const double t0 = nowMs();
        // the ids travel once per state, every field after that is one export + one copy
        check(rk.ctx, sph_download_owned(rk.ctx, field, rk.idsFresh ? nullptr : rk.ids.data(), rk.buf.data(), (size_t)m * per, &cnt),
              "sph_download_owned");
        const double t1 = nowMs();
        rk.ms[1] += t1 - t0;
        if (cnt != m || (rk.idsFresh && rk.owned != m)) throw std::runtime_error("sph_download_owned: row count changed under the download");
        if (!rk.idsFresh) {
            for (uint32_t i = 0; i < m; i++)
                if (rk.ids[i] >= n_) throw std::runtime_error("sph_download_owned: particle id out of range");
            rk.owned = m;
            rk.idsFresh = true;
        }
        // scatter by particle index (the ranks own disjoint indices, so the threads never write the same element)
        const uint32_t* ids = rk.ids.data();
        if (per == 16) {
            struct R16 { uint64_t a, b; };
            const R16* src = reinterpret_cast<const R16*>(rk.buf.data());
            R16* o = reinterpret_cast<R16*>(dst);
            for (uint32_t i = 0; i < m; i++) o[ids[i]] = src[i];
        } else if (per == 12) {
            struct R12 { uint32_t a, b, c; };
            const R12* src = reinterpret_cast<const R12*>(rk.buf.data());
            R12* o = reinterpret_cast<R12*>(dst);
            for (uint32_t i = 0; i < m; i++) o[ids[i]] = src[i];
        } else if (per == 8) {
            const uint64_t* src = reinterpret_cast<const uint64_t*>(rk.buf.data());
            uint64_t* o = reinterpret_cast<uint64_t*>(dst);
            for (uint32_t i = 0; i < m; i++) o[ids[i]] = src[i];
        } else {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(rk.buf.data());
            uint32_t* o = reinterpret_cast<uint32_t*>(dst);
            for (uint32_t i = 0; i < m; i++) o[ids[i]] = src[i];
        }
        rk.ms[2] += nowMs() - t1;
        got[(size_t)k] = m;
    });
    uint64_t total = 0;
    for (uint32_t g : got) total += g;
    if (total != n_) throw std::runtime_error("SlabGroup::download: the ranks own " + std::to_string(total) + " particles, expected " + std::to_string(n_));
}

void SlabGroup::timings(double out6[6])
{
    for (int i = 0; i < 6; i++) out6[i] = 0.0;
    for (auto& r : rank_) {
        double t[6];
        check(r.ctx, sph_get_timings(r.ctx, t), "sph_get_timings");
        for (int i = 0; i < 6; i++) if (t[i] > out6[i]) out6[i] = t[i];
    }
}

void SlabGroup::breakdown(double out4[4]) const
{
    for (int i = 0; i < 4; i++) {
        out4[i] = 0.0;
        for (const auto& r : rank_) if (r.ms[i] > out4[i]) out4[i] = r.ms[i];
    }
}

void SlabGroup::resetBreakdown()
{
    for (auto& r : rank_) for (double& v : r.ms) v = 0.0;
}

std::vector<int32_t> SlabGroup::layers() const
{
    std::vector<int32_t> L((size_t)ranks() + 1, 0);
    if (sph_comm_get_layers(rank_[0].ctx, L.data()) != SPH_OK) L.clear();
    return L;
}

std::vector<uint32_t> SlabGroup::ownedCounts() const
{
    std::vector<uint32_t> c;
    for (const auto& r : rank_) c.push_back(sph_num_particles(r.ctx));
    return c;
}

uint64_t SlabGroup::launches() const
{
    uint64_t l = 0;
    for (const auto& r : rank_) l += sph_launch_count(r.ctx);
    return l;
}

}  // namespace sphb200
