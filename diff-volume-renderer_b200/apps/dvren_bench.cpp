// dvren_bench -- the drop-in path a reference caller links: dvren::Context / Plan / DenseGridField /
// Renderer::Forward + Backward through libdvren.so with pageable std::vector inputs and results (the shapes the
// reference's API forces, reference src/render/renderer.cpp:376-386,441-444).  Two uses:
//
//   dvren_bench bench <grid n> <width> <steps> <stratified 0|1> <iters> <warmup> [pin 0|1] [staged 0|1]
//       times Forward+Backward per step on the host clock (every host<->device copy inside) and prints ONE JSON line:
//       bench.py reports it as e2e.renderer next to the pinned C-ABI end-to-end number.
//   dvren_bench shard <grid n> <width> <steps> <stratified 0|1> <iters> <warmup> <gpus> <groups> <reserve_sms> <max_ctas> [mode]
//       (mode 0: interleaved tile rows + slab all-reduces; 1: balanced bands + sparse exchange, every rank ends with the
//       whole gradient; 2: the same, every rank ends with the finished sum of the slabs it owns)
//       ONE frame rendered by <gpus> GPUs of this box with NO Python: one host thread per GPU, hpx_comm (NCCL) +
//       hpx_shard_step (interleaved tile rows, signalled backward, slab all-reduces behind it).  Rank 0 first checks
//       the reduced gradient against a plain single-GPU backward, then all ranks time <iters> steps (CUDA events, max
//       over ranks).  Prints ONE JSON line.
//   dvren_bench selftest
//       small end-to-end checks of the C++ surface that need a GPU: UpdateValues reaches the staged path, RenderStats
//       carry GPU (CUDA-event) stage times, objects outlive their context.  Exit code 0 = pass.
//
// The synthetic volume / camera / dL/dI are those of SURVEY 8(d) (same integer hash as python/synth.py).
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "dvren/core/context.hpp"
#include "dvren/core/plan.hpp"
#include "dvren/fields/dense_grid.hpp"
#include "dvren/render/renderer.hpp"

namespace {

using Clock = std::chrono::steady_clock;

uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

float hash_unit(uint64_t index, uint64_t seed) {   // synth.hash_unit
    const uint64_t z = mix64((index ^ seed) + 0x9E3779B97F4A7C15ULL);
    return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}

dvren::DenseGridConfig hashed_volume(int n, float scale, uint64_t seed = 1234) {
    dvren::DenseGridConfig c;
    c.resolution = {n, n, n};
    const size_t v = static_cast<size_t>(n) * n * n;
    c.sigma.resize(v);
    c.color.resize(v * 3);
    for (size_t i = 0; i < v; ++i) {
        c.sigma[i] = hash_unit(i, seed) * scale;
        for (int ch = 0; ch < 3; ++ch) c.color[3 * i + ch] = hash_unit(i, seed + 1 + ch);
    }
    return c;
}

dvren::PlanDescriptor bench_plan(uint32_t w, uint32_t h, uint32_t steps, bool stratified) {   // synth.bench_plan, view 0
    dvren::PlanDescriptor d;
    d.width = w; d.height = h;
    d.t_near = 0.9f; d.t_far = 4.0f;
    d.sampling.dt = static_cast<float>(1.5 / steps);
    d.sampling.max_steps = steps;
    d.sampling.mode = stratified ? dvren::SamplingMode::kStratified : dvren::SamplingMode::kFixed;
    d.seed = 42;
    d.camera.K = {1.2f * w, 0.f, w / 2.0f, 0.f, 1.2f * w, h / 2.0f, 0.f, 0.f, 1.f};
    d.camera.c2w = {1.f, 0.f, 0.f, 0.5f, 0.f, 1.f, 0.f, 0.5f, 0.f, 0.f, 1.f, -1.0f};
    return d;
}

std::vector<float> hashed_image_grad(size_t rays) {
    std::vector<float> g(rays * 3);
    for (size_t i = 0; i < g.size(); ++i) g[i] = hash_unit(i, 777) - 0.5f;
    return g;
}

#define CHECK(cond, ...)                                        \
    do {                                                        \
        if (!(cond)) {                                          \
            std::fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            std::fprintf(stderr, __VA_ARGS__);                  \
            std::fprintf(stderr, "\n");                         \
            return 1;                                           \
        }                                                       \
    } while (0)

int run_bench(int n, uint32_t w, uint32_t steps, bool strat, int iters, int warmup, bool pin, bool staged) {
    dvren::Context ctx;
    dvren::Status st = dvren::Context::Create({}, ctx);
    CHECK(st.ok(), "Context::Create: %s", st.ToString().c_str());
    dvren::Plan plan;
    st = dvren::Plan::Create(ctx, bench_plan(w, w, steps, strat), plan);
    CHECK(st.ok(), "Plan::Create: %s", st.ToString().c_str());
    dvren::DenseGridField field;
    {
        const dvren::DenseGridConfig cfg = hashed_volume(n, 2.0f);
        st = dvren::DenseGridField::Create(ctx, cfg, field);
        CHECK(st.ok(), "DenseGridField::Create: %s", st.ToString().c_str());
    }
    dvren::RenderOptions opt{};
    opt.pin_result_buffers = pin;   // one set of result objects, alive for the whole loop: the opt-in's contract
    opt.use_fused_path = !staged;   // staged: hp_ray -> hp_samp -> hp_int -> hp_img, hp_diff -> scatter on materialised samples
    dvren::Renderer renderer(ctx, plan, opt);
    const std::vector<float> dl = hashed_image_grad(static_cast<size_t>(w) * w);
    dvren::ForwardResult fwd;      // reused across steps, as a training loop does
    dvren::BackwardResult bwd;
    double fwd_kernel = 0, fwd_read = 0, fwd_integrate = 0, bwd_kernel = 0, bwd_read = 0, total = 0;
    for (int i = 0; i < warmup + iters; ++i) {
        const auto t0 = Clock::now();
        st = renderer.Forward(field, fwd);
        CHECK(st.ok(), "Forward: %s", st.ToString().c_str());
        st = renderer.Backward(field, dl, bwd);
        CHECK(st.ok(), "Backward: %s", st.ToString().c_str());
        const double ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
        if (i >= warmup) {
            total += ms;
            fwd_kernel += fwd.stats.sample_ms; fwd_read += fwd.stats.compose_ms;
            fwd_integrate += fwd.stats.integrate_ms;
            bwd_kernel += renderer.backward_stats().sample_ms; bwd_read += renderer.backward_stats().compose_ms;
        }
    }
    double mass = 0.0;
    for (size_t i = 0; i < bwd.sigma.size(); i += 4099) mass += std::fabs(bwd.sigma[i]);
    const size_t pixels = static_cast<size_t>(w) * w;
    std::printf("{\"pin_result_buffers\": %s, \"staged_path\": %s, \"ms_per_step\": %.4f, \"samples\": %zu, \"rays\": %zu, \"live_samples\": %zu, \"forward_kernel_ms\": %.4f, "
                "\"forward_integrate_ms\": %.4f, \"forward_readback_ms\": %.4f, \"backward_kernel_ms\": %.4f, \"backward_readback_ms\": %.4f, "
                "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu, \"iters\": %d, \"warmup\": %d, \"checksum\": %.6g}\n",
                pin ? "true" : "false", staged ? "true" : "false", total / iters, fwd.sample_count, fwd.ray_count, renderer.live_sample_count(), fwd_kernel / iters, fwd_integrate / iters, fwd_read / iters,
                bwd_kernel / iters, bwd_read / iters, dl.size() * 4, pixels * 28 + bwd.sigma.size() * 16 + 64, iters, warmup, mass);
    return 0;
}

int run_selftest() {
    dvren::Context* ctx = new dvren::Context();
    dvren::Status st = dvren::Context::Create({}, *ctx);
    CHECK(st.ok(), "Context::Create: %s", st.ToString().c_str());
    dvren::Plan plan;
    st = dvren::Plan::Create(*ctx, bench_plan(48, 40, 64, true), plan);
    CHECK(st.ok(), "Plan::Create");
    dvren::DenseGridConfig a = hashed_volume(16, 40.0f, 11), b = hashed_volume(16, 2.0f, 99);
    dvren::DenseGridField field, fresh;
    CHECK(dvren::DenseGridField::Create(*ctx, a, field).ok(), "field");
    CHECK(dvren::DenseGridField::Create(*ctx, b, fresh).ok(), "fresh field");
    dvren::RenderOptions staged_opt;
    staged_opt.use_fused_path = false;
    dvren::Renderer fused(*ctx, plan, dvren::RenderOptions{}), staged(*ctx, plan, staged_opt);
    dvren::ForwardResult f0, f1, s1, r1;
    CHECK(fused.Forward(field, f0).ok(), "fused forward");
    // (1) UpdateValues reaches BOTH paths: fused == staged == a field created from the new values
    CHECK(field.UpdateValues(b.sigma, b.color).ok(), "UpdateValues");
    CHECK(fused.Forward(field, f1).ok() && staged.Forward(field, s1).ok() && fused.Forward(fresh, r1).ok(), "forwards after update");
    CHECK(std::memcmp(f0.image.data(), f1.image.data(), f0.image.size() * 4) != 0, "the update did not change the image");
    CHECK(std::memcmp(f1.image.data(), r1.image.data(), f1.image.size() * 4) == 0, "updated field != fresh field (fused)");
    double worst = 0.0, peak = 0.0;
    for (size_t i = 0; i < s1.image.size(); ++i) {
        worst = std::fmax(worst, std::fabs(static_cast<double>(s1.image[i]) - f1.image[i]));
        peak = std::fmax(peak, std::fabs(static_cast<double>(f1.image[i])));
    }
    CHECK(worst <= 1e-5 * peak, "staged path still renders the old grid after UpdateValues: |diff| %.3g of %.3g", worst, peak);
    CHECK(s1.sample_count == f1.sample_count && s1.ray_count == f1.ray_count, "counts");
    // (2) RenderStats carry GPU stage times
    bool events = false;
    for (const auto& note : f1.stats.notes) events = events || note == "timing=cuda_events";
    CHECK(events && f1.stats.sample_ms > 0.0 && f1.stats.compose_ms > 0.0 && f1.stats.total_ms >= f1.stats.sample_ms, "fused RenderStats");
    CHECK(s1.stats.ray_ms > 0.0 && s1.stats.sample_ms > 0.0 && s1.stats.integrate_ms > 0.0 && s1.stats.compose_ms > 0.0, "staged RenderStats");
    // (3) backward through both paths agrees; stats present; mirrors are lazy but correct
    const std::vector<float> dl = hashed_image_grad(48 * 40);
    dvren::BackwardResult gb_f, gb_s;
    CHECK(fused.Backward(field, dl, gb_f).ok(), "fused backward");
    CHECK(fused.backward_stats().sample_ms > 0.0 && fused.backward_stats().compose_ms > 0.0, "backward stats");
    const std::vector<float> mirror = field.sigma_gradients();
    CHECK(mirror.size() == gb_f.sigma.size() && std::memcmp(mirror.data(), gb_f.sigma.data(), mirror.size() * 4) == 0, "mirror");
    CHECK(staged.Forward(field, s1).ok() && staged.Backward(field, dl, gb_s).ok(), "staged backward");
    double gworst = 0.0, gpeak = 0.0;
    for (size_t i = 0; i < gb_f.sigma.size(); ++i) {
        gworst = std::fmax(gworst, std::fabs(static_cast<double>(gb_f.sigma[i]) - gb_s.sigma[i]));
        gpeak = std::fmax(gpeak, std::fabs(static_cast<double>(gb_s.sigma[i])));
    }
    CHECK(gpeak > 0.0 && gworst <= 1e-4 * gpeak, "fused vs staged sigma gradients: %.3g of %.3g", gworst, gpeak);
    // (4) the context may go first (legal with the reference, whose fields never touch it)
    delete ctx;
    CHECK(fused.Forward(field, f1).ok(), "forward after the context was released");
    std::printf("selftest ok\n");
    return 0;
}

// ---- one frame over several GPUs, C ABI only ----------------------------------------------------------------------
struct ShardArgs {
    int n; uint32_t w, steps; bool strat; int iters, warmup, gpus; uint32_t groups, reserve; int max_ctas;
    int mode;   // 0: interleaved tile rows + slab all-reduces; 1: balanced bands, replicated result; 2: balanced bands, owned result
};

struct ShardShared {
    uint8_t id[HPX_COMM_ID_BYTES];
    const dvren::DenseGridConfig* volume;
    const std::vector<float>* dl;
    std::vector<double> ms, ms_no_reduce, send_mb, recv_mb, tile_order;
    std::vector<int> status;
    std::vector<uint32_t> band_row0, band_rows;
    std::vector<int32_t> wedges, cuts;
    int direct = 0;
    double verify = -1.0;
    uint64_t samples = 0;
    uint32_t usable_sms = 0, total_sms = 0;
    int nccl_version = 0;
};

#define SCHECK(expr)                                                                                     \
    do {                                                                                                 \
        const hp_status st__ = (expr);                                                                   \
        if (st__ != HP_STATUS_SUCCESS) {                                                                 \
            std::fprintf(stderr, "rank %d: %s -> %d (%s)\n", rank, #expr, static_cast<int>(st__), hpx_last_error()); \
            sh.status[rank] = 1;                                                                         \
            return;                                                                                      \
        }                                                                                                \
    } while (0)

void shard_rank(int rank, const ShardArgs& a, ShardShared& sh, std::barrier<>& sync) {
    hpx_ctx_ext2 ext{HPX_CTX_EXT2_MAGIC, rank, nullptr, a.reserve, 0u};
    hp_ctx_desc cd{};
    cd.reserved = &ext;
    hp_ctx* ctx = nullptr;
    SCHECK(hp_ctx_create(&cd, &ctx));
    hp_plan_desc pd{};
    {
        const dvren::PlanDescriptor d = bench_plan(a.w, a.w, a.steps, a.strat);
        pd.width = d.width; pd.height = d.height; pd.t_near = d.t_near; pd.t_far = d.t_far; pd.seed = d.seed;
        pd.sampling.dt = d.sampling.dt; pd.sampling.max_steps = d.sampling.max_steps;
        pd.sampling.mode = a.strat ? HP_SAMPLING_STRATIFIED : HP_SAMPLING_FIXED;
        pd.camera.model = HP_CAMERA_PINHOLE;
        for (int i = 0; i < 9; ++i) pd.camera.K[i] = d.camera.K[static_cast<size_t>(i)];
        for (int i = 0; i < 12; ++i) pd.camera.c2w[i] = d.camera.c2w[static_cast<size_t>(i)];
    }
    hp_plan* plan = nullptr;
    SCHECK(hp_plan_create(ctx, &pd, &plan));
    hpx_grid* grid = nullptr;
    SCHECK(hpx_grid_create_raw(ctx, a.n, a.n, a.n, sh.volume->sigma.data(), sh.volume->color.data(), HP_MEMSPACE_HOST, HP_INTERP_LINEAR,
                               HP_OOB_ZERO, nullptr, nullptr, &grid));
    hpx_comm* comm = nullptr;
    SCHECK(hpx_comm_create(ctx, sh.id, rank, a.gpus, a.max_ctas, &comm));
    std::vector<float> weights(a.groups);
    for (uint32_t g = 0; g < a.groups; ++g) weights[g] = a.groups == 1 ? 1.0f : std::pow(0.72f, static_cast<float>(g));   // small last group
    hpx_shard* shard = nullptr;
    if (a.mode == 0) SCHECK(hpx_shard_create(comm, plan, grid, weights.data(), a.groups, &shard));
    else SCHECK(hpx_shard_create_bands(comm, plan, grid, HPX_SHARD_RESULT_REPLICATED, &shard));   // verified replicated, timed as asked
    void* d_dl = nullptr;
    SCHECK(hpx_device_alloc(ctx, sh.dl->size() * 4, &d_dl));
    SCHECK(hpx_copy_to_device(ctx, d_dl, sh.dl->data(), sh.dl->size() * 4));
    const uint32_t flags = HPX_BACKWARD_GRID | HPX_BACKWARD_ZERO;

    if (a.mode != 0) {   // untimed start-up: tile dispatch order by measurement, then bands re-cut from the measured per-rank time until they stop moving
        SCHECK(hpx_shard_step(shard, static_cast<const float*>(d_dl), flags));
        if (std::getenv("DVREN_BENCH_NO_TUNE") == nullptr) SCHECK(hpx_shard_tune_order(shard, static_cast<const float*>(d_dl), flags, nullptr));
        for (int round = 0; round < 4; ++round) {
            SCHECK(hpx_shard_step(shard, static_cast<const float*>(d_dl), flags));
            SCHECK(hpx_shard_step(shard, static_cast<const float*>(d_dl), flags));
            int32_t changed = 0;
            SCHECK(hpx_shard_rebalance(shard, &changed));
            if (!changed) break;
        }
    }
    // correctness first: the reduced gradient of one sharded step against a plain single-GPU backward (rank 0)
    SCHECK(hpx_shard_step(shard, static_cast<const float*>(d_dl), flags));
    SCHECK(hpx_ctx_synchronize(ctx));
    hpx_frame* own = nullptr;
    SCHECK(hpx_shard_frame(shard, &own));
    hpx_counts counts{};
    SCHECK(hpx_frame_counts(own, &counts));
    static std::atomic<uint64_t> total_samples{0};
    total_samples += counts.samples;
    sync.arrive_and_wait();
    if (rank == 0) {
        const size_t v = static_cast<size_t>(a.n) * a.n * a.n;
        std::vector<float> sg(v), cg(v * 3), sg1(v), cg1(v * 3);
        SCHECK(hpx_grid_read_grad(grid, sg.data(), cg.data(), nullptr, HP_MEMSPACE_HOST));
        hpx_frame* full = nullptr;
        SCHECK(hpx_frame_create(plan, &full));
        SCHECK(hpx_forward(full, grid));
        SCHECK(hpx_backward(full, grid, static_cast<const float*>(d_dl), HP_MEMSPACE_DEVICE, flags));
        SCHECK(hpx_grid_read_grad(grid, sg1.data(), cg1.data(), nullptr, HP_MEMSPACE_HOST));
        hpx_frame_release(full);
        // Both sides are float32 red accumulations of the SAME contributions in different orders (each kernel has its own
        // oracle parity tests); this check is there to catch a slab summed twice or not at all, hence the 1e-2 floor.
        double peak_s = 0, peak_c = 0, worst = 0;
        for (size_t i = 0; i < v; ++i) peak_s = std::fmax(peak_s, std::fabs(sg1[i]));
        for (size_t i = 0; i < 3 * v; ++i) peak_c = std::fmax(peak_c, std::fabs(cg1[i]));
        for (size_t i = 0; i < v; ++i)
            worst = std::fmax(worst, std::fabs(static_cast<double>(sg[i]) - sg1[i]) / std::fmax(std::fabs(sg1[i]), 1e-2 * peak_s));
        for (size_t i = 0; i < 3 * v; ++i)
            worst = std::fmax(worst, std::fabs(static_cast<double>(cg[i]) - cg1[i]) / std::fmax(std::fabs(cg1[i]), 1e-2 * peak_c));
        sh.verify = worst;
        sh.samples = total_samples.load();
        hpx_ctx_sm_counts(ctx, &sh.usable_sms, &sh.total_sms);
        int32_t r = 0, w = 0, ver = 0;
        hpx_comm_info(comm, &r, &w, &ver);
        sh.nccl_version = ver;
    }
    if (a.mode != 0 && rank == 0) {
        sh.band_row0.resize(a.gpus); sh.band_rows.resize(a.gpus); sh.wedges.resize(2 * a.gpus); sh.cuts.resize(a.gpus + 1);
        hpx_shard_bands(shard, sh.band_row0.data(), sh.band_rows.data(), sh.wedges.data(), sh.cuts.data(), nullptr, nullptr);
    }
    if (a.mode != 0) {
        int32_t direct = 0;
        hpx_shard_exchange_is_direct(shard, &direct);
        if (rank == 0) sh.direct = direct;
        size_t out = 0, in = 0;
        hpx_shard_bands(shard, nullptr, nullptr, nullptr, nullptr, &out, &in);
        sh.send_mb[rank] = out * 4.0 / 1e6;
        sh.recv_mb[rank] = in * 4.0 / 1e6;
        int32_t order = 0;
        hpx_shard_tile_order(shard, &order);
        sh.tile_order[rank] = order;
    }
    if (a.mode == 2) SCHECK(hpx_shard_set_result(shard, HPX_SHARD_RESULT_OWNED));
    for (int pass = 0; pass < 2; ++pass) {   // pass 0: the real step; pass 1: the same without its collectives
        SCHECK(hpx_shard_set_reduce(shard, pass == 0 ? 1 : 0));
        for (int i = 0; i < a.warmup; ++i) SCHECK(hpx_shard_step(shard, static_cast<const float*>(d_dl), flags));
        SCHECK(hpx_ctx_synchronize(ctx));
        sync.arrive_and_wait();
        SCHECK(hpx_ctx_mark(ctx, 14));
        for (int i = 0; i < a.iters; ++i) SCHECK(hpx_shard_step(shard, static_cast<const float*>(d_dl), flags));
        SCHECK(hpx_ctx_mark(ctx, 15));
        float ms = 0.0f;
        SCHECK(hpx_ctx_elapsed_ms(ctx, 14, 15, &ms));
        (pass == 0 ? sh.ms : sh.ms_no_reduce)[rank] = ms / a.iters;
        sync.arrive_and_wait();
    }
    hpx_shard_release(shard);
    hpx_comm_release(comm);
    hpx_device_free(ctx, d_dl);
    hpx_grid_release(grid);
    hp_plan_release(plan);
    hp_ctx_release(ctx);
}

int run_shard(const ShardArgs& a) {
    ShardShared sh;
    if (a.gpus > 1) {
        const hp_status st = hpx_comm_unique_id(sh.id);
        CHECK(st == HP_STATUS_SUCCESS, "hpx_comm_unique_id: %s", hpx_last_error());
    }
    const dvren::DenseGridConfig volume = hashed_volume(a.n, 2.0f);
    const std::vector<float> dl = hashed_image_grad(static_cast<size_t>(a.w) * a.w);
    sh.volume = &volume;
    sh.dl = &dl;
    sh.ms.assign(a.gpus, 0.0); sh.ms_no_reduce.assign(a.gpus, 0.0); sh.status.assign(a.gpus, 0);
    sh.send_mb.assign(a.gpus, 0.0); sh.recv_mb.assign(a.gpus, 0.0); sh.tile_order.assign(a.gpus, 0.0);
    std::barrier<> sync(a.gpus);
    std::vector<std::thread> threads;
    for (int r = 0; r < a.gpus; ++r) threads.emplace_back([&, r] {
        shard_rank(r, a, sh, sync);
        if (sh.status[r] != 0) std::_Exit(3);   // a failed rank would leave the others waiting at the barrier
    });
    for (auto& t : threads) t.join();
    double ms = 0, ms0 = 0;
    for (int r = 0; r < a.gpus; ++r) { ms = std::fmax(ms, sh.ms[r]); ms0 = std::fmax(ms0, sh.ms_no_reduce[r]); }
    const double total = static_cast<double>(a.w) * a.w * a.steps;
    auto list = [](const char* key, const auto& v) {
        std::printf("\"%s\": [", key);
        for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.6g", i ? ", " : "", static_cast<double>(v[i]));
        std::printf("], ");
    };
    std::printf("{");
    list("ms_per_rank", sh.ms); list("ms_per_rank_without_collectives", sh.ms_no_reduce);
    if (a.mode != 0) {
        list("band_row0", sh.band_row0); list("band_rows", sh.band_rows); list("wedges", sh.wedges); list("owner_cuts", sh.cuts);
        list("send_mb", sh.send_mb); list("recv_mb", sh.recv_mb); list("tile_order", sh.tile_order);
    }
    if (a.mode != 0) std::printf("\"exchange\": \"%s\", ", sh.direct ? "own kernels over peer memory" : "nccl");
    std::printf("\"sharding\": \"%s\", ", a.mode == 0 ? "interleaved tile rows, slab all-reduces behind a signalled backward"
                                           : a.mode == 1 ? "balanced bands, sparse exchange, replicated result"
                                                         : "balanced bands, sparse exchange, owned result (reduce-scatter)");
    std::printf("\"mode\": \"shard\", \"gpus\": %d, \"groups\": %u, \"reserve_sms\": %u, \"usable_sms\": %u, \"total_sms\": %u, "
                "\"max_ctas\": %d, \"nccl_version\": %d, \"ms_per_step\": %.4f, \"ms_per_step_without_collectives\": %.4f, "
                "\"msamples_per_s\": %.1f, \"samples\": %.0f, \"verify_max_rel_err_vs_single_gpu\": %.3e}\n",
                a.gpus, a.groups, a.reserve, sh.usable_sms, sh.total_sms, a.max_ctas, sh.nccl_version, ms, ms0, total / (ms * 1e-3) / 1e6,
                total, sh.verify);
    CHECK(sh.verify >= 0.0 && sh.verify <= 1e-4, "sharded gradient differs from the single-GPU one: %.3e", sh.verify);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 12 && std::string(argv[1]) == "shard") {
        ShardArgs a{std::atoi(argv[2]), static_cast<uint32_t>(std::atoi(argv[3])), static_cast<uint32_t>(std::atoi(argv[4])),
                    std::atoi(argv[5]) != 0, std::atoi(argv[6]), std::atoi(argv[7]), std::atoi(argv[8]),
                    static_cast<uint32_t>(std::atoi(argv[9])), static_cast<uint32_t>(std::atoi(argv[10])), std::atoi(argv[11]),
                    argc >= 13 ? std::atoi(argv[12]) : 0};
        return run_shard(a);
    }
    if (argc >= 2 && std::string(argv[1]) == "selftest") return run_selftest();
    if (argc >= 8 && std::string(argv[1]) == "bench")
        return run_bench(std::atoi(argv[2]), static_cast<uint32_t>(std::atoi(argv[3])), static_cast<uint32_t>(std::atoi(argv[4])),
                         std::atoi(argv[5]) != 0, std::atoi(argv[6]), std::atoi(argv[7]), argc >= 9 && std::atoi(argv[8]) != 0,
                         argc >= 10 && std::atoi(argv[9]) != 0);
    std::fprintf(stderr, "usage: %s selftest | bench <grid n> <width> <steps> <stratified 0|1> <iters> <warmup> [pin 0|1] | "
                         "shard <grid n> <width> <steps> <stratified> <iters> <warmup> <gpus> <groups> <reserve_sms> <max_ctas> [mode 0|1|2]\n", argv[0]);
    return 2;
}
