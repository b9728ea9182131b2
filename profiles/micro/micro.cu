// Micro-benchmarks that decide kernel design on B200 (run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro micro.cu):
//   1. FP32 issue rate: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) per SM
//   2. shared-memory LDS.128 rate per SM
//   3. PCIe: pinned H2D, D2H, both at once; cudaHostRegister'ed malloc memory
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

template <int PACKED>
__global__ void __launch_bounds__(512) k_fma(float *out, float a, float b, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
        if (PACKED) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, aa, bb;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(v), "l"(aa), "l"(bb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) k_lds(float *out, int iters) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 v = sm[(idx + j * 72) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

static float timeit(cudaStream_t s, void (*fn)(void *), void *arg) { return 0; }

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, dev));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    printf("device %s, %d SMs, max clock %d MHz\n", pr.name, pr.multiProcessorCount, clk_khz / 1000);
    float *out;
    CK(cudaMalloc(&out, 148 * 8 * 512 * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (packed) k_fma<1><<<148 * 2, 512>>>(out, 1.0001f, 0.5f, iters);
            else k_fma<0><<<148 * 2, 512>>>(out, 1.0001f, 0.5f, iters);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fmas = 148.0 * 2 * 512 * 16.0 * iters;
            if (rep == 2) printf("%s: %.3f ms, %.1f GFMA/s = %.1f FMA/clk/SM at max clock (%.1f at 1.5 GHz)\n", packed ? "FFMA2" : "FFMA ", ms,
                                 fmas / ms * 1e-6, fmas / (ms * 1e-3) / 148 / (clk_khz * 1e3), fmas / (ms * 1e-3) / 148 / 1.5e9);
        }
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_lds<<<148 * 2, 512>>>(out, 20000);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 148.0 * 2 * 512 * 8.0 * 20000 * 16;
        if (rep == 2) printf("LDS.128: %.3f ms, %.1f B/clk/SM at max clock\n", ms, bytes / (ms * 1e-3) / 148 / (clk_khz * 1e3));
    }

    // ---- PCIe ----
    const size_t N = (size_t)4 << 30;
    void *d0, *d1, *h0, *h1;
    CK(cudaMalloc(&d0, N)); CK(cudaMalloc(&d1, N));
    CK(cudaMallocHost(&h0, N)); CK(cudaMallocHost(&h1, N));
    memset(h0, 1, N); memset(h1, 2, N);
    cudaStream_t s0, s1;
    cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    auto run = [&](const char *name, bool up, bool down, size_t chunk) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0, s0);
            cudaStreamWaitEvent(s1, e0, 0);
            for (size_t o = 0; o < N; o += chunk) {
                if (up) CK(cudaMemcpyAsync((char *)d0 + o, (char *)h0 + o, chunk, cudaMemcpyHostToDevice, s0));
                if (down) CK(cudaMemcpyAsync((char *)h1 + o, (char *)d1 + o, chunk, cudaMemcpyDeviceToHost, s1));
            }
            cudaEventRecord(e1, s1);
            cudaStreamWaitEvent(s0, e1, 0);
            cudaEventRecord(e1, s0);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1) printf("%s (chunk %zu MiB): %.1f ms, %.1f GB/s per direction\n", name, chunk >> 20, ms, N / (ms * 1e-3) * 1e-9);
        }
    };
    run("pinned H2D", true, false, N);
    run("pinned D2H", false, true, N);
    run("pinned H2D+D2H", true, true, N);
    run("pinned H2D", true, false, (size_t)64 << 20);
    run("pinned H2D+D2H", true, true, (size_t)64 << 20);
    // registered malloc memory
    void *m0 = nullptr;
    posix_memalign(&m0, 4096, N);
    memset(m0, 3, N);
    {
        cudaEventRecord(e0, s0);
        CK(cudaMemcpyAsync(d0, m0, N, cudaMemcpyHostToDevice, s0));
        cudaEventRecord(e1, s0); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("pageable H2D: %.1f ms, %.1f GB/s\n", ms, N / (ms * 1e-3) * 1e-9);
    }
    {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        CK(cudaHostRegister(m0, N, cudaHostRegisterDefault));
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("cudaHostRegister of 4 GiB: %.1f ms\n", (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
        cudaEventRecord(e0, s0);
        CK(cudaMemcpyAsync(d0, m0, N, cudaMemcpyHostToDevice, s0));
        cudaEventRecord(e1, s0); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("registered H2D: %.1f ms, %.1f GB/s\n", ms, N / (ms * 1e-3) * 1e-9);
        clock_gettime(CLOCK_MONOTONIC, &t0);
        cudaHostUnregister(m0);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("cudaHostUnregister: %.1f ms\n", (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    }
    return 0;
}
