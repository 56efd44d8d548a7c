// dv_graph.cu -- hp_graph_{create,capture,execute,release} of hp.h as a real CUDA graph.
//
// The reference's capture cannot succeed (it launches on the default stream, allocates and
// synchronises inside the capture and hands null rays to its own kernel; reference
// hotpath/src/cuda/graph_cuda.cu:142-159, SURVEY finding 5).  Here the graph body is
//   rays -> offsets -> sample+integrate -> compose [-> per-sample backward]
// enqueued on the context's stream with library-owned buffers sized from what the plan
// actually emits (rays x samples-per-ray), not from the ws_* hints and not from
// max_samples: the reference's own graph tests request less capacity than their plans
// need (hp_runner.cpp:2866-2904).  hp_graph_execute returns views of those buffers.
#include <new>
#include <vector>

#include "dv_objects.h"

using namespace dv;

#define DV_TRY(expr)                                     \
    do {                                                 \
        const hp_status dv_st__ = (expr);                \
        if (dv_st__ != HP_STATUS_SUCCESS) return dv_st__; \
    } while (0)

namespace {

struct GraphExec {
    const hp_ctx* ctx = nullptr;
    std::vector<void*> owned;
    hp_rays_t rays{};
    hp_samp_t samp{};
    hp_intl_t intl{};
    hp_img_t img{};
    hp_grads_t grads{};
    float* d_dense_grad = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool captured = false;
    size_t n_rays = 0, n_samples = 0, pixels = 0;
};

void* own(GraphExec* g, size_t bytes, hp_status* st) {
    if (*st != HP_STATUS_SUCCESS) return nullptr;
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        *st = cuda_fail(e, "cudaMalloc(graph)");
        return nullptr;
    }
    g->owned.push_back(p);
    return p;
}

void drop_graph(GraphExec* g) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    g->exec = nullptr;
    g->graph = nullptr;
    g->captured = false;
}

}  // namespace

extern "C" {

HP_API hp_status hp_graph_create(const hp_plan* plan, const hp_field* fs, const hp_field* fc, size_t, size_t, size_t,
                                 size_t, void** out_graph_handle) {
    DV_RANGE("hp_graph_create");
    if (plan == nullptr || fs == nullptr || fc == nullptr || out_graph_handle == nullptr)
        return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(plan->ctx);
    GraphExec* g = new (std::nothrow) GraphExec();
    if (g == nullptr) return HP_STATUS_OUT_OF_MEMORY;
    g->ctx = ctx_retain(plan->ctx);
    *out_graph_handle = g;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hp_graph_capture(void* handle, const hp_plan* plan, const hp_field* fs, const hp_field* fc,
                                  const hp_tensor* dL_dI) {
    DV_RANGE("hp_graph_capture");
    if (handle == nullptr || plan == nullptr || fs == nullptr || fc == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    if (fs->kind != FieldKind::kDenseSigma || fc->kind != FieldKind::kDenseColor) return HP_STATUS_INVALID_ARGUMENT;
    GraphExec* g = static_cast<GraphExec*>(handle);
    if (plan->ctx != g->ctx) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    cudaStream_t s = g->ctx->stream;
    DV_CUDA(cudaStreamSynchronize(s));
    drop_graph(g);
    for (void* p : g->owned) cudaFree(p);
    g->owned.clear();

    const hp_plan_desc& d = plan->desc;
    const size_t n = static_cast<size_t>(d.roi.width) * d.roi.height;
    const uint64_t m64 = static_cast<uint64_t>(n) * plan->uniform_count;
    if (m64 > UINT32_MAX) return HP_STATUS_INVALID_ARGUMENT;  // offsets are u32 in this ABI
    const size_t m = static_cast<size_t>(m64);
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    g->n_rays = n; g->n_samples = m; g->pixels = pixels;
    const bool backward = dL_dI != nullptr && dL_dI->data != nullptr;
    if (backward && (dL_dI->rank < 2 || dL_dI->shape[0] != static_cast<int64_t>(n) || dL_dI->shape[1] < 3))
        return HP_STATUS_INVALID_ARGUMENT;

    hp_status st = HP_STATUS_SUCCESS;
    const hp_memspace dev = HP_MEMSPACE_DEVICE;
    g->rays = hp_rays_t{};
    shape_tensor(g->rays.origins, HP_DTYPE_F32, dev, 2, n, 3);     g->rays.origins.data = own(g, n * 12, &st);
    shape_tensor(g->rays.directions, HP_DTYPE_F32, dev, 2, n, 3);  g->rays.directions.data = own(g, n * 12, &st);
    shape_tensor(g->rays.t_near, HP_DTYPE_F32, dev, 1, n);         g->rays.t_near.data = own(g, n * 4, &st);
    shape_tensor(g->rays.t_far, HP_DTYPE_F32, dev, 1, n);          g->rays.t_far.data = own(g, n * 4, &st);
    shape_tensor(g->rays.pixel_ids, HP_DTYPE_U32, dev, 1, n);      g->rays.pixel_ids.data = own(g, n * 4, &st);
    g->samp = hp_samp_t{};
    shape_tensor(g->samp.positions, HP_DTYPE_F32, dev, 2, m, 3);   g->samp.positions.data = own(g, m * 12, &st);
    shape_tensor(g->samp.dt, HP_DTYPE_F32, dev, 1, m);             g->samp.dt.data = own(g, m * 4, &st);
    shape_tensor(g->samp.sigma, HP_DTYPE_F32, dev, 1, m);          g->samp.sigma.data = own(g, m * 4, &st);
    shape_tensor(g->samp.color, HP_DTYPE_F32, dev, 2, m, 3);       g->samp.color.data = own(g, m * 12, &st);
    shape_tensor(g->samp.ray_offset, HP_DTYPE_U32, dev, 1, n + 1); g->samp.ray_offset.data = own(g, (n + 1) * 4, &st);
    g->intl = hp_intl_t{};
    shape_tensor(g->intl.radiance, HP_DTYPE_F32, dev, 2, n, 3);    g->intl.radiance.data = own(g, n * 12, &st);
    shape_tensor(g->intl.transmittance, HP_DTYPE_F32, dev, 1, n);  g->intl.transmittance.data = own(g, n * 4, &st);
    shape_tensor(g->intl.opacity, HP_DTYPE_F32, dev, 1, n);        g->intl.opacity.data = own(g, n * 4, &st);
    shape_tensor(g->intl.depth, HP_DTYPE_F32, dev, 1, n);          g->intl.depth.data = own(g, n * 4, &st);
    shape_tensor(g->intl.aux, HP_DTYPE_F32, dev, 2, m, 4);         g->intl.aux.data = own(g, m * 16, &st);
    g->img = hp_img_t{};
    shape_tensor(g->img.image, HP_DTYPE_F32, dev, 3, d.height, d.width, 3); g->img.image.data = own(g, pixels * 12, &st);
    shape_tensor(g->img.trans, HP_DTYPE_F32, dev, 2, d.height, d.width);    g->img.trans.data = own(g, pixels * 4, &st);
    shape_tensor(g->img.opacity, HP_DTYPE_F32, dev, 2, d.height, d.width);  g->img.opacity.data = own(g, pixels * 4, &st);
    shape_tensor(g->img.depth, HP_DTYPE_F32, dev, 2, d.height, d.width);    g->img.depth.data = own(g, pixels * 4, &st);
    shape_tensor(g->img.hitmask, HP_DTYPE_U32, dev, 2, d.height, d.width);  g->img.hitmask.data = own(g, pixels * 4, &st);
    g->grads = hp_grads_t{};
    const float* d_grad_in = nullptr;
    int64_t sr = 3, sc = 1;
    if (backward) {
        shape_tensor(g->grads.sigma, HP_DTYPE_F32, dev, 1, m);     g->grads.sigma.data = own(g, m * 4, &st);
        shape_tensor(g->grads.color, HP_DTYPE_F32, dev, 2, m, 3);  g->grads.color.data = own(g, m * 12, &st);
        shape_tensor(g->grads.camera, HP_DTYPE_F32, dev, 2, 3, 4); g->grads.camera.data = own(g, 48, &st);
        if (dL_dI->memspace == HP_MEMSPACE_DEVICE) {
            d_grad_in = static_cast<const float*>(dL_dI->data);   // stays the caller's, like the reference
            sr = dL_dI->stride[0];
            sc = dL_dI->stride[1];
        } else {
            g->d_dense_grad = static_cast<float*>(own(g, n * 12, &st));
            if (st == HP_STATUS_SUCCESS) {
                std::vector<float> dense(n * 3);
                const float* hg = static_cast<const float*>(dL_dI->data);
                for (size_t r = 0; r < n; ++r)
                    for (int c = 0; c < 3; ++c)
                        dense[3 * r + c] = hg[static_cast<int64_t>(r) * dL_dI->stride[0] + c * dL_dI->stride[1]];
                DV_CUDA(cudaMemcpyAsync(g->d_dense_grad, dense.data(), n * 12, cudaMemcpyHostToDevice, s));
                DV_CUDA(cudaStreamSynchronize(s));
            }
            d_grad_in = g->d_dense_grad;
        }
    }
    if (st != HP_STATUS_SUCCESS) return st;

    const FrameParams fp = frame_params_from_plan(*plan);
    RayArrays ra;
    ra.origins = static_cast<float*>(g->rays.origins.data);
    ra.directions = static_cast<float*>(g->rays.directions.data);
    ra.t_near = static_cast<float*>(g->rays.t_near.data);
    ra.t_far = static_cast<float*>(g->rays.t_far.data);
    ra.pixel_ids = static_cast<uint32_t*>(g->rays.pixel_ids.data);
    SampleArrays sa;
    sa.positions = static_cast<float*>(g->samp.positions.data);
    sa.dt = static_cast<float*>(g->samp.dt.data);
    sa.sigma = static_cast<float*>(g->samp.sigma.data);
    sa.color = static_cast<float*>(g->samp.color.data);
    sa.ray_offset = static_cast<uint32_t*>(g->samp.ray_offset.data);
    IntegralArrays ia;
    ia.radiance = static_cast<float*>(g->intl.radiance.data);
    ia.transmittance = static_cast<float*>(g->intl.transmittance.data);
    ia.opacity = static_cast<float*>(g->intl.opacity.data);
    ia.depth = static_cast<float*>(g->intl.depth.data);
    ia.aux = static_cast<float*>(g->intl.aux.data);
    ImagePlanes ip;
    ip.image = static_cast<float*>(g->img.image.data);
    ip.trans = static_cast<float*>(g->img.trans.data);
    ip.opacity = static_cast<float*>(g->img.opacity.data);
    ip.depth = static_cast<float*>(g->img.depth.data);
    ip.hitmask = static_cast<uint32_t*>(g->img.hitmask.data);

    DV_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = launch_rays(s, fp, ra, static_cast<uint32_t>(n));
    if (e == cudaSuccess) e = launch_uniform_offsets(s, sa.ray_offset, static_cast<uint32_t>(n), plan->uniform_count);
    if (e == cudaSuccess)
        e = launch_sample(s, fp.march, d.t_near, d.t_far, field_pair(fs, fc), ra, static_cast<uint32_t>(n), sa, true, ia);
    if (e == cudaSuccess) e = launch_background(s, ip, pixels, d.t_far);
    if (e == cudaSuccess) e = launch_compose(s, ip, pixels, ra.pixel_ids, ia, static_cast<uint32_t>(n), g->ctx->d_status);
    if (e == cudaSuccess && backward) {
        e = cudaMemsetAsync(g->grads.camera.data, 0, 48, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->grads.sigma.data, 0, m * 4, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(g->grads.color.data, 0, m * 12, s);
        if (e == cudaSuccess)
            e = launch_diff(s, d_grad_in, sr, sc, sa, ia.aux, static_cast<uint32_t>(n), static_cast<uint32_t>(m),
                            static_cast<float*>(g->grads.sigma.data), static_cast<float*>(g->grads.color.data),
                            g->ctx->d_status);
    }
    const cudaError_t end = cudaStreamEndCapture(s, &g->graph);
    if (e != cudaSuccess) return cuda_fail(e, "graph body");
    if (end != cudaSuccess) return cuda_fail(end, "cudaStreamEndCapture");
    DV_CUDA(cudaGraphInstantiate(&g->exec, g->graph, nullptr, nullptr, 0));
    g->captured = true;
    return HP_STATUS_SUCCESS;
}

HP_API hp_status hp_graph_execute(void* handle, hp_rays_t* out_rays, hp_samp_t* out_samp, hp_intl_t* out_intl,
                                  hp_img_t* out_img, hp_grads_t* out_grads) {
    DV_RANGE("hp_graph_execute");
    if (handle == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    GraphExec* g = static_cast<GraphExec*>(handle);
    if (!g->captured || g->exec == nullptr) return HP_STATUS_INVALID_ARGUMENT;
    DV_ENTER(g->ctx);
    DV_CUDA(cudaGraphLaunch(g->exec, g->ctx->stream));
    DV_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (out_rays) *out_rays = g->rays;
    if (out_samp) *out_samp = g->samp;
    if (out_intl) *out_intl = g->intl;
    if (out_img) *out_img = g->img;
    if (out_grads) *out_grads = g->grads;
    return HP_STATUS_SUCCESS;
}

HP_API void hp_graph_release(void* handle) {
    DV_RANGE("hp_graph_release");
    if (handle == nullptr) return;
    GraphExec* g = static_cast<GraphExec*>(handle);
    if (g->ctx != nullptr && g->ctx->ready) {
        DeviceScope scope;
        scope.enter(g->ctx);
        cudaStreamSynchronize(g->ctx->stream);
        drop_graph(g);
        for (void* p : g->owned) cudaFree(p);
    }
    ctx_unref(g->ctx);
    delete g;
}

}  // extern "C"
