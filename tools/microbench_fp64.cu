// fp64 latency / throughput microbenchmark for sm_100a (tuning aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64.bin tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma(double* out, int iters, double b, double c) {
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void lds_dfma(double* out, int iters, double b) {
    __shared__ double sm[256];
    sm[threadIdx.x] = threadIdx.x * b;
    __syncthreads();
    double a = 1.0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        a = fma(a, b, sm[idx]);
        idx = (idx + 1) & 255;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void sync_only(double* out, int iters) {
    double a = threadIdx.x;
    for (int it = 0; it < iters; ++it) { __syncthreads(); a += 1.0; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
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
    double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    const double clk = 1.965e9;
    const int iters = 20000;
    { double t = timeit([&] { dfma<1><<<1, 32>>>(out, iters, 0.999, 1e-3); });
      printf("DFMA dependent chain, 1 warp: %.1f cycles per DFMA (latency)\n", t * clk / iters); }
    { double t = timeit([&] { dfma<8><<<1, 32>>>(out, iters, 0.999, 1e-3); });
      printf("DFMA 8 chains, 1 warp: %.1f cycles per DFMA issue\n", t * clk / iters / 8); }
    { double t = timeit([&] { dfma<8><<<148 * 8, 256>>>(out, iters, 0.999, 1e-3); });
      printf("DFMA throughput full chip: %.2f DFMA/clk/SM\n", (double)148 * 8 * 256 * iters * 8 / (t * clk * 148)); }
    { double t = timeit([&] { dfma<2><<<1, 256>>>(out, iters, 0.999, 1e-3); });
      printf("DFMA 2 chains, 8 warps on one SM: %.1f cycles per (2 DFMA x 8 warps)\n", t * clk / iters); }
    { double t = timeit([&] { lds_dfma<<<1, 32>>>(out, iters, 0.999); });
      printf("LDS + DFMA dependent chain (address independent): %.1f cycles per iteration\n", t * clk / iters); }
    { double t = timeit([&] { sync_only<<<1, 256>>>(out, iters); });
      printf("__syncthreads + DADD, 256 threads: %.1f cycles per iteration\n", t * clk / iters); }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
