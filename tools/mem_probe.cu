// mem_probe.cu -- one-off B200 microbenchmarks behind the roofline denominators in DESIGN.md.
//
//   gather   : random 16-byte (float4) loads from a footprint of F MB, 8 independent loads per
//              thread per iteration (the shape of one trilinear sample).  F <= ~100 MB stays in L2,
//              larger footprints spill to HBM.  Reported as GB/s of gathered bytes.
//   red      : random red.global.add.v4.f32 into a footprint of F MB (the backward scatter primitive).
//   stream   : float4 copy (HBM STREAM figure to compare with MEASURED_PEAKS.json).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mem_probe tools/mem_probe.cu
// Run (on the GPU box): tools/mem_probe > gpurun_out/mem_probe.json
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

__global__ void gather_kernel(const float4* __restrict__ src, uint32_t mask, int iters, float* __restrict__ out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(src + ((lcg(s) >> 4) & mask));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// "local" variant: the 32 lanes of a warp gather from a window of `win` consecutive voxels around a
// random base (what a pixel tile does: many lanes, few distinct lines) -> L1-resident gather rate.
__global__ void gather_local_kernel(const float4* __restrict__ src, uint32_t mask, uint32_t win, int iters,
                                    float* __restrict__ out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t ws = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 40503u + 7u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const uint32_t base = (lcg(ws) >> 4) & mask;
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(src + ((base + ((lcg(s) >> 8) % win)) & mask));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// 256-bit variant (sm_100 LDG.E.ENL2.256): 4 loads of 32-byte x-pair records per "sample" instead of 8 x 16 B.
struct alignas(32) f8 { float v[8]; };
__device__ __forceinline__ f8 ldg256(const void* p) {
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__global__ void gather_local256_kernel(const f8* __restrict__ src, uint32_t mask, uint32_t win, int iters,
                                       float* __restrict__ out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t ws = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 40503u + 7u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const uint32_t base = (lcg(ws) >> 4) & mask;
        f8 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = ldg256(src + ((base + ((lcg(s) >> 8) % win)) & mask));
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[k].v[j];
    }
    if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ void red_add4(float4* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void red_kernel(float4* __restrict__ dst, uint32_t mask, uint32_t win, int iters) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t ws = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 40503u + 7u;
    const float4 v = make_float4(1e-6f, 2e-6f, 3e-6f, 4e-6f);
    for (int it = 0; it < iters; ++it) {
        const uint32_t base = win ? ((lcg(ws) >> 4) & mask) : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t idx = win ? ((base + ((lcg(s) >> 8) % win)) & mask) : ((lcg(s) >> 4) & mask);
            red_add4(dst + idx, v);
        }
    }
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F>
static float time_ms(F&& launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const size_t max_bytes = size_t(1) << 30;
    float4 *buf, *buf2; float* out;
    CK(cudaMalloc(&buf, max_bytes)); CK(cudaMalloc(&buf2, max_bytes)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(buf, 0, max_bytes)); CK(cudaMemset(buf2, 0, max_bytes));
    const int blocks = p.multiProcessorCount * 16, threads = 256, iters = 256;
    const double ops = double(blocks) * threads * iters * 8;
    printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, p.multiProcessorCount);
    printf(" \"gather_random_16B\": [");
    const int mbs[] = {4, 16, 64, 96, 256, 1024};
    for (int i = 0; i < 6; ++i) {
        const uint32_t mask = uint32_t((size_t(mbs[i]) << 20) / 16 - 1);
        float ms = time_ms([&] { gather_kernel<<<blocks, threads>>>(buf, mask, iters, out); });
        printf("%s{\"footprint_mb\": %d, \"gbs\": %.1f}", i ? ", " : "", mbs[i], ops * 16 / (ms * 1e-3) / 1e9);
    }
    printf("],\n \"gather_warp_local_16B\": [");
    const int wins[] = {8, 32, 128};
    for (int i = 0; i < 3; ++i) {
        const uint32_t mask = uint32_t((size_t(256) << 20) / 16 - 1);
        float ms = time_ms([&] { gather_local_kernel<<<blocks, threads>>>(buf, mask, wins[i], iters, out); });
        printf("%s{\"window_voxels\": %d, \"footprint_mb\": 256, \"gbs\": %.1f}", i ? ", " : "", wins[i], ops * 16 / (ms * 1e-3) / 1e9);
    }
    printf("],\n \"gather_warp_local_32B_ldg256\": [");
    for (int i = 0; i < 3; ++i) {
        const uint32_t mask = uint32_t((size_t(512) << 20) / 32 - 1);
        float ms = time_ms([&] { gather_local256_kernel<<<blocks, threads>>>(reinterpret_cast<const f8*>(buf), mask, wins[i], iters, out); });
        printf("%s{\"window_records\": %d, \"footprint_mb\": 512, \"gbs\": %.1f}", i ? ", " : "", wins[i], ops * 16 / (ms * 1e-3) / 1e9);
    }
    printf("],\n \"red_v4_f32\": [");
    const int rmb[] = {16, 64, 256, 256, 256};
    const int rwin[] = {0, 0, 0, 8, 32};
    for (int i = 0; i < 5; ++i) {
        const uint32_t mask = uint32_t((size_t(rmb[i]) << 20) / 16 - 1);
        float ms = time_ms([&] { red_kernel<<<blocks, threads>>>(buf2, mask, rwin[i], iters); });
        printf("%s{\"footprint_mb\": %d, \"warp_window_voxels\": %d, \"gops\": %.2f, \"gbs\": %.1f}", i ? ", " : "", rmb[i], rwin[i],
               ops / (ms * 1e-3) / 1e9, ops * 16 / (ms * 1e-3) / 1e9);
    }
    {
        const size_t n = max_bytes / 16;
        float ms = time_ms([&] { copy_kernel<<<p.multiProcessorCount * 8, 512>>>(buf, buf2, n); });
        printf("],\n \"stream_copy_gbs\": %.1f}\n", 2.0 * max_bytes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
