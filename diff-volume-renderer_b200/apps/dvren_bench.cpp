// dvren_bench -- the drop-in path a reference caller links: dvren::Context / Plan / DenseGridField /
// Renderer::Forward + Backward through libdvren.so with pageable std::vector inputs and results (the shapes the
// reference's API forces, reference src/render/renderer.cpp:376-386,441-444).  Two uses:
//
//   dvren_bench bench <grid n> <width> <steps> <stratified 0|1> <iters> <warmup>
//       times Forward+Backward per step on the host clock (every host<->device copy inside) and prints ONE JSON line:
//       bench.py reports it as e2e.renderer next to the pinned C-ABI end-to-end number.
//   dvren_bench selftest
//       small end-to-end checks of the C++ surface that need a GPU: UpdateValues reaches the staged path, RenderStats
//       carry GPU (CUDA-event) stage times, objects outlive their context.  Exit code 0 = pass.
//
// The synthetic volume / camera / dL/dI are those of SURVEY 8(d) (same integer hash as python/synth.py).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

int run_bench(int n, uint32_t w, uint32_t steps, bool strat, int iters, int warmup) {
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
    dvren::Renderer renderer(ctx, plan, dvren::RenderOptions{});
    const std::vector<float> dl = hashed_image_grad(static_cast<size_t>(w) * w);
    dvren::ForwardResult fwd;      // reused across steps, as a training loop does
    dvren::BackwardResult bwd;
    double fwd_kernel = 0, fwd_read = 0, bwd_kernel = 0, bwd_read = 0, total = 0;
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
            bwd_kernel += renderer.backward_stats().sample_ms; bwd_read += renderer.backward_stats().compose_ms;
        }
    }
    double mass = 0.0;
    for (size_t i = 0; i < bwd.sigma.size(); i += 4099) mass += std::fabs(bwd.sigma[i]);
    const size_t pixels = static_cast<size_t>(w) * w;
    std::printf("{\"ms_per_step\": %.4f, \"samples\": %zu, \"rays\": %zu, \"live_samples\": %zu, \"forward_kernel_ms\": %.4f, "
                "\"forward_readback_ms\": %.4f, \"backward_kernel_ms\": %.4f, \"backward_readback_ms\": %.4f, "
                "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu, \"iters\": %d, \"warmup\": %d, \"checksum\": %.6g}\n",
                total / iters, fwd.sample_count, fwd.ray_count, renderer.live_sample_count(), fwd_kernel / iters, fwd_read / iters,
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

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "selftest") return run_selftest();
    if (argc >= 8 && std::string(argv[1]) == "bench")
        return run_bench(std::atoi(argv[2]), static_cast<uint32_t>(std::atoi(argv[3])), static_cast<uint32_t>(std::atoi(argv[4])),
                         std::atoi(argv[5]) != 0, std::atoi(argv[6]), std::atoi(argv[7]));
    std::fprintf(stderr, "usage: %s selftest | bench <grid n> <width> <steps> <stratified 0|1> <iters> <warmup>\n", argv[0]);
    return 2;
}
