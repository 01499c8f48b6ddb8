// Validation of the UMMA shape used by the learned-dynamics kernel (smooth_mlp.cuh): D[128 x 112] (TMEM, fp32) =
// A[128 x 112] * B[112 x 112]^T, bf16 kind::f16, BOTH operands K-major, no swizzle, seven MMAs of K = 16.
//   addr(row r, k) = (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2,  SBO = 128, LBO = (rows/8)*128
// Sweeps which descriptor field takes which stride.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
constexpr int M = 128, N = 112, K = 112;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__global__ void __launch_bounds__(128) k(const float* A /*[M][K]*/, const float* B /*[N][K]*/, float* Dout /*[M][N]*/, int swap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t SBO = 128, LBO_A = (M / 8) * 128, LBO_B = (N / 8) * 128;
    unsigned char* sA = smem;
    unsigned char* sB = smem + M * K * 2;
    for (int idx = tid; idx < M * K; idx += 128) {
        const int r = idx / K, kk = idx % K;
        *reinterpret_cast<__nv_bfloat16*>(sA + (r / 8) * SBO + (kk / 8) * LBO_A + (r % 8) * 16 + (kk % 8) * 2) = __float2bfloat16(A[idx]);
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, kk = idx % K;
        *reinterpret_cast<__nv_bfloat16*>(sB + (r / 8) * SBO + (kk / 8) * LBO_B + (r % 8) * 16 + (kk % 8) * 2) = __float2bfloat16(B[idx]);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < K / 16; ++kb) {
            const uint32_t sa = smem_u32(sA) + kb * 2 * LBO_A, sb = smem_u32(sB) + kb * 2 * LBO_B;
            const uint64_t adesc = swap ? make_desc(sa, SBO, LBO_A) : make_desc(sa, LBO_A, SBO);
            const uint64_t bdesc = swap ? make_desc(sb, SBO, LBO_B) : make_desc(sb, LBO_B, SBO);
            const uint32_t acc = kb > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int q = 0; q < 16; ++q) Dout[tid * N + c0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128));
}
// Rate of the operand shape: `reps` rounds of the 3 * K/16 UMMAs the learned-dynamics kernel issues per tile
// (a_hi b_hi + a_hi b_lo + a_lo b_hi share the same smem tiles here), commit + wait per round.
__global__ void __launch_bounds__(128) rate(int reps, long long* cycles_out, int blocks_share) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t SBO = 128, LBO_A = (M / 8) * 128, LBO_B = (N / 8) * 128;
    for (int i = tid; i < (M + N) * K * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + M * K * 2);
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int kb = 0; kb < K / 16; ++kb) {
                const uint64_t adesc = make_desc(sA + kb * 2 * LBO_A, LBO_A, SBO);
                const uint64_t bdesc = make_desc(sB + kb * 2 * LBO_B, LBO_B, SBO);
                for (int piece = 0; piece < 3; ++piece) {
                    const uint32_t acc = (kb > 0 || piece > 0) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 :: "r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase));
            phase ^= 1u;
        }
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128));
}
int main() {
    static float hA[M * K], hB[N * K], hD[M * N];
    static double ref[M][N];
    for (int i = 0; i < M * K; ++i) hA[i] = (float)(((i * 7 + (i / K) * 3) % 13) - 6) * 0.25f;
    for (int i = 0; i < N * K; ++i) hB[i] = (float)(((i * 5 + (i / K) * 11) % 9) - 4) * 0.5f;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)hA[m * K + kk] * hB[n * K + kk]; ref[m][n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    const int smem_bytes = (M + N) * K * 2;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    for (int swap = 0; swap < 1; ++swap) {      // the swapped assignment faults (illegal address): not run
        cudaMemset(dD, 0, sizeof(hD));
        k<<<1, 128, smem_bytes>>>(dA, dB, dD, swap);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        int match = 0; double maxerr = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double d = fabs(hD[m * N + n] - ref[m][n]); if (d < 1e-3) ++match; if (d > maxerr) maxerr = d; }
        printf("K-major M=128 N=112 K=112 (LBO,SBO)=%s : match %d/%d maxerr %.3g | D[5][0..3] %g %g %g %g (ref %g %g %g %g)\n",
               swap ? "(row-group stride, k-group stride)" : "(k-group stride, row-group stride)", match, M * N, maxerr,
               hD[5 * N], hD[5 * N + 1], hD[5 * N + 2], hD[5 * N + 3], ref[5][0], ref[5][1], ref[5][2], ref[5][3]);
    }
    {
        long long* dc; cudaMalloc(&dc, sizeof(long long) * 296);
        cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        const int reps = 200;
        for (int blocks = 148; blocks <= 296; blocks += 148) {
            rate<<<blocks, 128, smem_bytes>>>(reps, dc, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            static long long hc[296]; cudaMemcpy(hc, dc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)hc[i] / blocks;
            printf("%d blocks (%d per SM): %.0f cycles per round of %d UMMAs (128 x 112 x 16) = %.1f cycles per UMMA per block\n",
                   blocks, blocks / 148, avg / reps, 3 * K / 16, avg / reps / (3 * K / 16));
        }
    }
    return 0;
}
