// ubench_red.cu -- micro-benchmark: scalar red.global.add.f32 against the vector form
// red.global.add.v4.f32 (PTX 8.1, sm_90+) on records of 12 contiguous floats at scattered
// addresses (the access pattern of the per-face gradient statistics of both renderers).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_red tools/ubench_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_v4(float* a, float x, float y, float z, float w)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void red_v2(float* a, float x, float y)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(x), "f"(y) : "memory");
}

__device__ __forceinline__ uint32_t hash(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// MODE 0: 12 scalar reds / record, 1: 3 x v4, 2: 6 x v2, 3: 10 scalar reds (reference tet pattern)
// `share`: number of adjacent lanes that target the same record (1 = every lane its own record)
template <int MODE>
__global__ void k(float* buf, uint32_t nrec, int iters, int share)
{
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        uint32_t r = hash((tid / share) * 977u + it * 0x9e3779b9u) % nrec;
        float* a = buf + (size_t)r * 12;
        float v = 1.0f + it;
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < 12; q++) atomicAdd(a + q, v);
        } else if (MODE == 1) {
            red_v4(a, v, v, v, v); red_v4(a + 4, v, v, v, v); red_v4(a + 8, v, v, v, v);
        } else if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 6; q++) red_v2(a + 2 * q, v, v);
        } else {
#pragma unroll
            for (int q = 0; q < 10; q++) atomicAdd(a + q, v);
        }
    }
}

int main()
{
    const uint32_t nrec = 3170304;   // C3 face count
    float* buf;
    cudaMalloc(&buf, (size_t)nrec * 48);
    cudaMemset(buf, 0, (size_t)nrec * 48);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 4096, threads = 64, iters = 180;
    for (int share = 1; share <= 8; share *= 2) {
        for (int mode = 0; mode < 4; mode++) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<blocks, threads>>>(buf, nrec, iters, share);
                if (mode == 1) k<1><<<blocks, threads>>>(buf, nrec, iters, share);
                if (mode == 2) k<2><<<blocks, threads>>>(buf, nrec, iters, share);
                if (mode == 3) k<3><<<blocks, threads>>>(buf, nrec, iters, share);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double recs = (double)blocks * threads * iters;
            const char* names[4] = { "12 x red.f32", "3 x red.v4.f32", "6 x red.v2.f32", "10 x red.f32" };
            printf("share=%d  %-16s  %8.3f ms   %7.2f G records/s   %7.2f G floats/s\n", share, names[mode], best,
                   recs / best * 1e-6, recs * (mode == 3 ? 10 : 12) / best * 1e-6);
        }
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
