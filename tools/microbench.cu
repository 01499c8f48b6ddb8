// Instruction-throughput microbenchmarks for sm_100a (tuning aid; not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, int iters, float fb, float fc, uint32_t ub) {
    float a[CHAINS]; uint32_t u[CHAINS]; float2 p[CHAINS / 2];
    for (int i = 0; i < CHAINS; ++i) { a[i] = threadIdx.x + i; u[i] = threadIdx.x * 7 + i; }
    for (int i = 0; i < CHAINS / 2; ++i) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], fb, fc);
            if (OP == 1) { if (i < CHAINS / 2) { uint64_t x = *(uint64_t*)&p[i], y, b2, c2; float2 bb = make_float2(fb, fb), cc = make_float2(fc, fc);
                               b2 = *(uint64_t*)&bb; c2 = *(uint64_t*)&cc; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(y) : "l"(x), "l"(b2), "l"(c2)); p[i] = *(float2*)&y; } }
            if (OP == 2) u[i] = u[i] * ub + 12345u;
            if (OP == 3) u[i] = __umulhi(u[i], ub) ^ (uint32_t)it;
            if (OP == 4) { uint64_t w = (uint64_t)u[i] * ub; u[i] = (uint32_t)(w >> 32) ^ (uint32_t)w; }
            if (OP == 5) a[i] = __sinf(a[i]);
            if (OP == 6) a[i] = __log2f(a[i]);
            if (OP == 7) { float r; asm volatile("sqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 8) { float r; asm volatile("rcp.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 9) u[i] = (u[i] ^ ub) + (u[i] >> 3);
            if (OP == 10) a[i] = a[i] * fb;
            if (OP == 11) a[i] = a[i] + fc;
            if (OP == 12) { float r; asm volatile("ex2.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 13) { float r; asm volatile("lg2.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 14) { float r; asm volatile("sin.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 15) { float r; asm volatile("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 16) { float r; asm volatile("cos.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(a[i])); a[i] = r; }
            if (OP == 17) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0,%1,%2;" : "=r"(r) : "f"(a[i]), "f"(fb)); a[i] = __uint_as_float(r); }
            if (OP == 18) { double dd = (double)a[i] * (double)fc; float r; asm volatile("cvt.rn.f32.f64 %0,%1;" : "=f"(r) : "d"(dd)); a[i] = r; }
            if (OP == 19) { uint32_t r; asm volatile("prmt.b32 %0,%1,%2,0x7632;" : "=r"(r) : "r"(u[i]), "r"(ub)); u[i] = r + 1u; }
            if (OP == 20) { uint32_t r; asm volatile("cvt.rna.tf32.f32 %0,%1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r); }
            if (OP == 21) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0,%1,%2;" : "=r"(r) : "f"(a[i]), "f"(fb)); a[i] = __uint_as_float(r); }
        }
    }
    float s = 0; for (int i = 0; i < CHAINS; ++i) s += a[i] + (float)u[i]; for (int i = 0; i < CHAINS / 2; ++i) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: FFMA + independent MUFU / IMAD streams to see co-issue
template <int MIX>
__global__ void __launch_bounds__(256) mix(float* out, int iters, float fb, float fc, uint32_t ub) {
    float a[8], m[4]; uint32_t u[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    for (int i = 0; i < 4; ++i) { m[i] = 0.5f + i; u[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], fb, fc);
        if (MIX == 1) {
#pragma unroll
            for (int i = 0; i < 2; ++i) { float r; asm volatile("ex2.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(m[i])); m[i] = r; }
        }
        if (MIX == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { uint64_t w = (uint64_t)u[i] * ub; u[i] = (uint32_t)(w >> 32) ^ (uint32_t)w; }
        }
        if (MIX == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) u[i] = (u[i] ^ ub) + (u[i] >> 3);
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i]; for (int i = 0; i < 4; ++i) s += m[i] + (float)u[i];
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
    const int blocks = 148 * 8, threads = 256, iters = 4000;
    float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
    const double clk = 1.965e9;
    const char* names[] = {"FFMA", "FFMA2(instr)", "IMAD.lo", "IMAD.HI", "IMAD.WIDE+xor", "sinf(MUFU+FMUL)", "__log2f", "sqrt.approx", "rcp.approx", "LOP3+SHF+IADD(3 ops)", "FMUL", "FADD", "ex2.approx", "lg2.approx.ftz", "sin.approx.ftz", "rsqrt.approx", "cos.approx.ftz", "cvt.bf16x2(F2FP)", "DMUL+cvt.f32.f64", "PRMT+IADD", "cvt.rna.tf32", "cvt.f16x2(F2FP)"};
#define RUN(OP, PER) { double t = timeit([&] { k<OP><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f, 2654435761u); }); \
        double ops = (double)blocks * threads * iters * PER; printf("%-24s %8.3f ms  %7.2f thread-ops/clk/SM\n", names[OP], t * 1e3, ops / (t * clk * 148)); }
    RUN(0, 8) RUN(1, 4) RUN(2, 8) RUN(3, 8) RUN(4, 8) RUN(5, 8) RUN(6, 8) RUN(7, 8) RUN(8, 8) RUN(9, 8) RUN(10, 8) RUN(11, 8) RUN(12, 8) RUN(13, 8) RUN(14, 8) RUN(15, 8) RUN(16, 8) RUN(17, 8) RUN(18, 8) RUN(19, 8) RUN(20, 8) RUN(21, 8)
    const char* mn[] = {"8 FFMA", "8 FFMA + 2 EX2", "8 FFMA + 4 IMAD.WIDE", "8 FFMA + 4x(3 ALU)"};
#define RUNM(M) { double t = timeit([&] { mix<M><<<blocks, threads>>>(out, iters, 0.999f, 1e-3f, 2654435761u); }); \
        printf("%-24s %8.3f ms  (FFMA-only equivalent %.2f FFMA/clk/SM)\n", mn[M], t * 1e3, (double)blocks * threads * iters * 8 / (t * clk * 148)); }
    RUNM(0) RUNM(1) RUNM(2) RUNM(3)
    cudaError_t e = cudaGetLastError(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
