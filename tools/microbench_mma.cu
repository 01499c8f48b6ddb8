// mma.sync throughput on sm_100a (legacy tensor path): m16n8k8 tf32 and m16n8k16 bf16, plus a mixed
// FFMA + MMA loop to see whether they overlap.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int NACC, int NFMA>
__global__ void __launch_bounds__(128) k(float* out, int iters, float fb) {
    float d[NACC][4]; uint32_t a[4], b[2]; float f[8];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    for (int j = 0; j < 4; ++j) a[j] = 0x3f800000u + threadIdx.x + j;
    b[0] = 0x3f000000u + threadIdx.x; b[1] = 0x3e800000u;
    for (int j = 0; j < 8; ++j) f[j] = threadIdx.x + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) { if (KIND == 0) mma_tf32(d[i], a, b); else mma_bf16(d[i], a, b); }
#pragma unroll
        for (int r = 0; r < NFMA; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], fb, 1e-3f);
    }
    float s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += d[i][j]; for (int j = 0; j < 8; ++j) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double timeit(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best * 1e-3;
}

int main() {
    const int blocks = 148 * 8, threads = 128, iters = 4000; const double clk = 1.965e9;
    float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
#define RUN(KIND, NACC, NFMA, MACS, label) { double t = timeit([&] { k<KIND, NACC, NFMA><<<blocks, threads>>>(out, iters, 0.999f); }); \
        double mmas = (double)blocks * (threads / 32) * iters * NACC; \
        printf("%-34s %8.3f ms  %8.1f MAC/clk/SM  %6.3f mma/clk/SM  (%.0f TFLOP/s)  ffma/clk/SM=%.1f\n", label, t * 1e3, mmas * MACS / (t * clk * 148), mmas / (t * clk * 148), 2 * mmas * MACS / t / 1e12, \
               (double)blocks * threads * iters * NFMA * 8 / (t * clk * 148)); }
    RUN(0, 4, 0, 1024, "tf32 m16n8k8, 4 indep acc") RUN(0, 8, 0, 1024, "tf32 m16n8k8, 8 indep acc") RUN(1, 4, 0, 2048, "bf16 m16n8k16, 4 indep acc") RUN(1, 8, 0, 2048, "bf16 m16n8k16, 8 indep acc")
    RUN(0, 4, 2, 1024, "tf32 4 mma + 16 ffma / iter") RUN(0, 4, 4, 1024, "tf32 4 mma + 32 ffma / iter") RUN(0, 4, 8, 1024, "tf32 4 mma + 64 ffma / iter") RUN(0, 0 + 1, 8, 1024, "tf32 1 mma + 64 ffma / iter")
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
