// host/host_demo.cc -- headless driver of the host class, the way FluidSimCPU drives the reference
// (fluidSimCPU.cc:9-46): InitializeData(n), then Update(dt) per frame.  Prints one line per run that
// tests/test_variants_gpu.py / tests/test_host_gpu.py compare with the same scene run through the C ABI from Python.
//   host_demo n steps table_mode [class | adapter | getters | substeps | multi ndev | snapshot path [ndev] | snapshotio path | snapshotread path | slabgroup ndev | setters]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cmath>
#include <string>
#include <thread>
#include <vector>
#include "FluidSimB200.h"
#include "SlabGroup.h"

static unsigned long long fnv1a(const void* p, size_t n, unsigned long long h = 0xcbf29ce484222325ull)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

// FluidSimB200 driven the way GameApp::Run drives a FluidSimBase (gameApp.cc:112-113,231,238,272):
// initialize, then update + render per frame, one reset in between.
struct Captured { unsigned long long pos = 0, col = 0; int count = -1, frames = 0; };
static void capture(const FluidSimB200::Frame& f, Shader*, RenderUtils::Camera*, void* user)
{
    Captured* c = (Captured*)user;
    c->count = f.count;
    c->pos = fnv1a(f.positions, (size_t)f.count * 16);
    c->col = fnv1a(f.colors, (size_t)f.count * 16);
    c->frames++;
}

static int run_adapter(int n, int steps, int mode)
{
    auto& sim = Physics::Fluid::FluidSimulation::getInstance();
    sim.setTableMode(mode);
    sim.setGravity(true);
    FluidSimB200 backend;
    FluidSimBase* b = &backend;                          // through the base class, like the application
    Captured cap;
    backend.setPresenter(capture, &cap);
    alignas(16) static char shader_stub[16], camera_stub[16];   // render()'s arguments are only passed through
    Shader& sh = *reinterpret_cast<Shader*>(shader_stub);
    RenderUtils::Camera& cam = *reinterpret_cast<RenderUtils::Camera*>(camera_stub);
    b->initialize(n);
    b->render(sh, cam);
    printf("adapter init count=%d pos_fnv=%016llx col_fnv=%016llx\n", cap.count, cap.pos, cap.col);
    b->update(0.016667f);
    b->reset();                                          // back to the spawn state
    for (int s = 0; s < steps; s++) { b->update(0.016667f); b->render(sh, cam); }
    printf("adapter n=%d steps=%d frames=%d out_fnv=%016llx col_fnv=%016llx\n", n, steps, cap.frames, cap.pos, cap.col);
    b->cleanup();
    return 0;
}

// the per-particle getters: single device reads for the first few calls of a frame, then host mirrors that
// FluidSimCPU::updateColors-style parallel loops can hit from any thread
static int run_getters(int n, int steps, int mode)
{
    auto& sim = Physics::Fluid::FluidSimulation::getInstance();
    sim.setTableMode(mode);
    sim.setGravity(true);
    sim.setHostMirrors(true, true);
    sim.InitializeData(n);
    for (int s = 0; s < steps; s++) sim.Update(0.016667f);
    const uint32 probe[4] = {0u, (uint32)n / 3u, (uint32)n / 2u, (uint32)n - 1u};
    float single[4][8];
    for (int k = 0; k < 4; k++) {                        // <= kSingleReads distinct particles: device reads
        const uint32 i = probe[k];
        const auto p = sim.getPosition(i); const auto v = sim.getVelocity(i);
        const float row[8] = {p.x, p.y, p.z, v.x, v.y, v.z, sim.getDensity(i), sim.getSpeedNormalzied(i)};
        memcpy(single[k], row, sizeof(row));
    }
    std::vector<float> speedn((size_t)n), rho((size_t)n * 2), pos((size_t)n * 3);
    auto work = [&](int t, int nt) {
        for (int i = t; i < n; i += nt) {
            speedn[i] = sim.getSpeedNormalzied((uint32)i);
            rho[2 * (size_t)i] = sim.getDensity((uint32)i); rho[2 * (size_t)i + 1] = sim.getNearDensity((uint32)i);
            const auto p = sim.getPosition((uint32)i);
            pos[3 * (size_t)i] = p.x; pos[3 * (size_t)i + 1] = p.y; pos[3 * (size_t)i + 2] = p.z;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++) th.emplace_back(work, t, 4);
    for (auto& t : th) t.join();
    double worst = 0.0;
    for (int k = 0; k < 4; k++) {
        const uint32 i = probe[k];
        const auto p = sim.getPosition(i); const auto v = sim.getVelocity(i);
        const float row[8] = {p.x, p.y, p.z, v.x, v.y, v.z, sim.getDensity(i), sim.getSpeedNormalzied(i)};
        for (int c = 0; c < 8; c++) {
            const double d = std::fabs((double)row[c] - (double)single[k][c]);
            const double tol = (c == 7) ? 1e-6 : 0.0;    // the device's speed may be FMA-contracted
            if (d > tol && d > worst) worst = d;
        }
    }
    float sum = 0.0f;
    for (int i = 0; i < n; i++) sum += speedn[i];
    printf("getters n=%d steps=%d single_vs_bulk_worst=%g pos_fnv=%016llx mirror_fnv=%016llx dens_fnv=%016llx speed_sum=%.6f oob=%.1f\n",
           n, steps, worst, fnv1a(pos.data(), pos.size() * 4), fnv1a(sim.positions.data(), sim.positions.size() * 12),
           fnv1a(rho.data(), rho.size() * 4), sum, sim.getSpeedNormalzied((uint32)n));
    return 0;
}

static int run_substeps(int n, int steps, int mode)
{
    auto& sim = Physics::Fluid::FluidSimulation::getInstance();
    sim.setTableMode(mode);
    sim.setGravity(true);
    sim.setHostMirrors(true, true);
    sim.InitializeData(n);
    sim.setMaxTimestep(0.005f);
    for (int s = 0; s < steps; s++) sim.Update(0.016667f);     // 4 sub-steps of 0.016667f / 4 each
    printf("substeps n=%d steps=%d sub=%u pos_fnv=%016llx\n", n, steps, sim.lastSubsteps(),
           fnv1a(sim.positions.data(), sim.positions.size() * 12));
    return 0;
}

// the same scene on one GPU and slab-decomposed over ndev GPUs of the box, through the same class
static int run_multi(int n, int steps, int ndev)
{
    auto& sim = Physics::Fluid::FluidSimulation::getInstance();
    sim.setGravity(true);
    sim.setHostMirrors(true, true);
    sim.setDevice(0);
    sim.InitializeData(n);
    const float rho_spawn = sim.getDensity((uint32)n / 2u);
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t1 = 0.0, tn = 0.0;                            // wall time of the Updates after the warm-up ones (mirrors included)
    const int skip = steps > 6 ? 3 : 1;                   // (the first re-balance sets up NCCL's all-reduce: a one-off of tens of ms)
    for (int s = 0; s < steps; s++) { const double a = now(); sim.Update(0.016667f); if (s >= skip) t1 += now() - a; }
    const auto pos1 = sim.positions;
    const auto out1 = sim.OutPositions;
    std::vector<float> rho1;
    sim.downloadDensities(rho1);
    std::vector<int> devs;
    for (int d = 0; d < ndev; d++) devs.push_back(d);
    sim.setDevices(devs);
    sim.setDirectMirrors(getenv("SPH_DEMO_DIRECT_MIRRORS") != nullptr);     // A/B: GPUs write the mirrors themselves
    sim.setRebalanceInterval(2);
    sim.InitializeData(n);                               // the reset path: same lattice, now handed to the slabs
    const float rho_spawn_multi = sim.getDensity((uint32)n / 2u);      // valid before the first Update, like the reference
    double phase[4] = {0, 0, 0, 0};
    for (int s = 0; s < steps; s++) {
        if (s == skip) sim.multiGpuBreakdown(phase);      // drop the warm-up Updates from the per-phase sums
        const double a = now(); sim.Update(0.016667f); if (s >= skip) tn += now() - a;
    }
    sim.multiGpuBreakdown(phase);
    double dpos = 0.0, dout = 0.0, drho = 0.0;
    for (int i = 0; i < n; i++) {
        dpos = std::fmax(dpos, std::fmax(std::fabs((double)sim.positions[i].x - pos1[i].x),
                         std::fmax(std::fabs((double)sim.positions[i].y - pos1[i].y), std::fabs((double)sim.positions[i].z - pos1[i].z))));
        dout = std::fmax(dout, std::fabs((double)sim.OutPositions[i].x - out1[i].x) + std::fabs((double)sim.OutPositions[i].w - out1[i].w));
    }
    std::vector<float> rho2;
    sim.downloadDensities(rho2);
    for (size_t i = 0; i < rho2.size(); i++) drho = std::fmax(drho, std::fabs((double)rho2[i] - rho1[i]) / std::fmax(1e-30, std::fabs((double)rho1[i])));
    // getters: answered from the bulk mirrors, whichever slab owns the particle now
    int getters_ok = 1;
    const uint32 probe[3] = {0u, (uint32)n / 2u, (uint32)n - 1u};
    for (uint32 i : probe) {
        const auto p = sim.getPosition(i);
        if (p.x != sim.positions[i].x || p.y != sim.positions[i].y || p.z != sim.positions[i].z) getters_ok = 0;
        if (sim.getDensity(i) != rho2[2 * (size_t)i]) getters_ok = 0;
    }
    if (sim.getDensity((uint32)n) != 0.0f) getters_ok = 0;
    unsigned long long owned = 0;
    std::string per;
    for (uint32 c : sim.particlesPerDevice()) { owned += c; per += (per.empty() ? "" : ",") + std::to_string(c); }
    printf("multi n=%d ndev=%d steps=%d max_pos_diff=%.3g max_out_diff=%.3g max_rho_rel=%.3g spawn_rho=%.6f/%.6f getters_ok=%d owned=%llu per_device=[%s] density_ms=%.4f update_ms_1gpu=%.3f update_ms_ngpu=%.3f "
           "ngpu_phase_ms=step:%.3f,download:%.3f,scatter:%.3f,rebalance:%.3f\n",
           n, ndev, steps, dpos, dout, drho, rho_spawn, rho_spawn_multi, getters_ok, owned, per.c_str(), sim.getElapsedTimeDensity(),
           steps > skip ? t1 / (steps - skip) : 0.0, steps > skip ? tn / (steps - skip) : 0.0,
           steps > skip ? phase[0] / (steps - skip) : 0.0, steps > skip ? phase[1] / (steps - skip) : 0.0,
           steps > skip ? phase[2] / (steps - skip) : 0.0, steps > skip ? phase[3] / (steps - skip) : 0.0);
    sim.shutdown();
    return 0;
}

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 10000;
    int steps = argc > 2 ? atoi(argv[2]) : 3;
    int mode = argc > 3 ? atoi(argv[3]) : SPH_TABLE_GRID;
    const char* what = argc > 4 ? argv[4] : "class";
    auto& sim = Physics::Fluid::FluidSimulation::getInstance();
    try {
        if (!strcmp(what, "adapter")) return run_adapter(n, steps, mode);
        if (!strcmp(what, "getters")) return run_getters(n, steps, mode);
        if (!strcmp(what, "substeps")) return run_substeps(n, steps, mode);
        if (!strcmp(what, "multi")) return run_multi(n, steps, argc > 5 ? atoi(argv[5]) : 2);
        if (!strcmp(what, "snapshot")) {               // saveState / loadState through the class: resume == uninterrupted run
            const char* path = argc > 5 ? argv[5] : "/tmp/sph_snapshot.bin";
            const int ndev = argc > 6 ? atoi(argv[6]) : 1;
            if (ndev > 1) { std::vector<int> devs; for (int d = 0; d < ndev; d++) devs.push_back(d); sim.setDevices(devs); }
            sim.setTableMode(mode);
            sim.setGravity(true);
            sim.setViscosityStrength(0.6f);
            sim.setHostMirrors(true, true);
            sim.InitializeData(n);
            for (int s = 0; s < steps; s++) sim.Update(0.016667f);
            sim.saveState(path);
            for (int s = 0; s < 2; s++) sim.Update(0.016667f);
            const unsigned long long straight = fnv1a(sim.positions.data(), sim.positions.size() * 12);
            sim.setViscosityStrength(0.1f);             // loadState must bring the parameters back too
            sim.loadState(path);
            const float mu = sim.getViscosityStrength();
            for (int s = 0; s < 2; s++) sim.Update(0.016667f);
            const unsigned long long resumed = fnv1a(sim.positions.data(), sim.positions.size() * 12);
            printf("snapshot n=%d steps=%d ndev=%d straight_fnv=%016llx resumed_fnv=%016llx mu=%.2f\n", n, steps, ndev, straight, resumed, mu);
            sim.shutdown();
            return 0;
        }
        if (!strcmp(what, "snapshotio")) {             // the snapshot file format, host-side only: argv[5] = path
            const char* path = argc > 5 ? argv[5] : "/tmp/sph_snapshot.bin";
            SphParams p;
            sph_default_params(&p);
            p.gravity = 1; p.viscosity_strength = 0.75f; p.bound[2] = 7.5f;
            std::vector<float> pos((size_t)n * 3), vel((size_t)n * 3);
            for (size_t i = 0; i < pos.size(); i++) { pos[i] = 0.25f * (float)i - 3.0f; vel[i] = -0.5f * (float)i; }
            if (!sphb200::writeSnapshotFile(path, (uint32_t)n, p, pos.data(), vel.data())) { printf("snapshotio write failed\n"); return 3; }
            uint32_t n2 = 0; SphParams q; std::vector<float> pos2, vel2;
            const bool ok = sphb200::readSnapshotFile(path, n2, q, pos2, vel2) && n2 == (uint32_t)n && pos2 == pos && vel2 == vel &&
                            !memcmp(&p, &q, sizeof(p));
            std::vector<float> dummy;
            const bool rejects = !sphb200::readSnapshotFile(std::string(path) + ".missing", n2, q, dummy, dummy);
            printf("snapshotio n=%d roundtrip=%d rejects_missing=%d\n", n, (int)ok, (int)rejects);
            return ok && rejects ? 0 : 3;
        }
        if (!strcmp(what, "classbench")) {
            // the drop-in path timed the way the application drives it (fluidSimCPU.cc:35-40,58): per frame ONE
            // FluidSimulation::Update(dt) that returns with the OutPositions mirror refreshed in host memory.
            //   host_demo 0 frames 0 classbench state.bin [ndev [direct]]      (state.bin: a snapshot file, saveState's format)
            const char* path = argc > 5 ? argv[5] : "/tmp/sph_snapshot.bin";
            const int ndev = argc > 6 ? atoi(argv[6]) : 1;
            const bool direct = argc > 7 && atoi(argv[7]) != 0;
            if (ndev > 1) { std::vector<int> devs; for (int d = 0; d < ndev; d++) devs.push_back(d); sim.setDevices(devs); sim.setDirectMirrors(direct); sim.setRebalanceInterval(4); }
            sim.setHostMirrors(true, false);
            sim.loadState(path);
            const int count = (int)sim.positions.size();
            auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
            const int warm = 5;
            for (int s = 0; s < warm; s++) sim.Update(0.016667f);
            double phase[4] = {0, 0, 0, 0};
            if (ndev > 1) sim.multiGpuBreakdown(phase);
            const double a = now();
            for (int s = 0; s < steps; s++) sim.Update(0.016667f);
            const double ms = (now() - a) / (steps > 0 ? steps : 1);
            if (ndev > 1) sim.multiGpuBreakdown(phase);
            bool finite = true;
            for (int i = 0; i < count; i += 997) finite = finite && std::isfinite(sim.OutPositions[(size_t)i].x) && sim.OutPositions[(size_t)i].w == 0.34f;
            printf("classbench n=%d ndev=%d direct=%d frames=%d update_ms=%.4f updates_per_s=%.1f finite=%d phase_ms=step:%.3f,download:%.3f,scatter:%.3f,rebalance:%.3f\n",
                   count, ndev, (int)direct, steps, ms, count / (ms * 1e-3), (int)finite,
                   steps ? phase[0] / steps : 0.0, steps ? phase[1] / steps : 0.0, steps ? phase[2] / steps : 0.0, steps ? phase[3] / steps : 0.0);
            sim.shutdown();
            return finite ? 0 : 3;
        }
        if (!strcmp(what, "setters")) {                // UI sliders: setters never throw, a rejected value is not committed
            sim.setTableMode(mode);
            sim.setGravity(true);
            sim.setHostMirrors(true, true);
            sim.InitializeData(n);
            sim.Update(0.016667f);
            sim.setInteractionRadius(0.01f);            // gameApp.cc:371 slider minimum: 2000^3 cells -- the context falls back
            const float r_small = sim.getInteractionRadius();      // to the reference's own table instead of refusing
            sim.Update(0.016667f);
            const float rho_small = sim.getDensity(0);
            sim.setInteractionRadius(0.0f);             // nonsense: refused, recorded, not committed
            const float r_after_bad = sim.getInteractionRadius();
            const bool recorded = !sim.lastError().empty();
            sim.setBound(sphb200::vec3(30.0f, 30.0f, 30.0f));           // slider maximum (gameApp.cc:408)
            sim.setInteractionRadius(0.35f);
            sim.Update(0.016667f);
            bool finite = true;
            for (const auto& q : sim.positions) finite = finite && std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z);
            printf("setters r_small=%.3f rho_small=%.4f r_after_bad=%.3f recorded=%d r_final=%.3f bound=%.1f finite=%d\n", r_small, rho_small,
                   r_after_bad, (int)recorded, sim.getInteractionRadius(), sim.getBounds().x, (int)finite);
            sim.shutdown();
            return 0;
        }
        if (!strcmp(what, "snapshotread")) {           // read only: a malformed header must be refused (argv[5] = path)
            uint32_t n2 = 0; SphParams q; std::vector<float> pos2, vel2;
            const bool ok = sphb200::readSnapshotFile(argc > 5 ? argv[5] : "", n2, q, pos2, vel2);
            printf("snapshotread read=%d n=%u\n", (int)ok, ok ? n2 : 0u);
            return 0;
        }
        if (!strcmp(what, "slabgroup")) {              // the group on its own: construction (and its failure path), nothing else
            SphParams p;
            sph_default_params(&p);
            std::vector<int> devs;
            for (int d = 0; d < (argc > 5 ? atoi(argv[5]) : 2); d++) devs.push_back(d);
            sphb200::SlabGroup group(devs, (uint32_t)n, p);
            printf("slabgroup ranks=%d\n", group.ranks());
            return 0;
        }
        sim.setTableMode(mode);
        sim.setGravity(true);
        sim.setHostMirrors(true, true);
        sim.InitializeData(n);
        printf("init out0=(%.6f %.6f %.6f %.2f) rho0=%.6f\n", sim.OutPositions[0].x, sim.OutPositions[0].y,
               sim.OutPositions[0].z, sim.OutPositions[0].w, sim.getDensity(0));
        for (int s = 0; s < steps; s++) sim.Update(0.016667f);
        printf("n=%d steps=%d pos_fnv=%016llx out_fnv=%016llx\n", n, steps,
               fnv1a(sim.positions.data(), sim.positions.size() * 12),
               fnv1a(sim.OutPositions.data(), sim.OutPositions.size() * 16));
        printf("p0=(%.6f %.6f %.6f) v0=(%.6f %.6f %.6f) rho=%.6f nrho=%.6f oob=%.1f\n",
               sim.getPosition(0).x, sim.getPosition(0).y, sim.getPosition(0).z,
               sim.getVelocity(0).x, sim.getVelocity(0).y, sim.getVelocity(0).z,
               sim.getDensity(0), sim.getNearDensity(0), sim.getDensity((uint32)n));
        printf("timers_ms %.4f %.4f %.4f %.4f %.4f %.4f\n", sim.getElapsedTimeGravity(), sim.getElapsedTimeSpatial(),
               sim.getElapsedTimeDensity(), sim.getElapsedTimePressure(), sim.getElapsedTimeViscosity(),
               sim.getElapsedTimePosNColl());
    } catch (const std::exception& e) {
        fprintf(stderr, "host_demo failed: %s\n", e.what());
        return 2;
    }
    return 0;
}
