// Variant sweep for the tcgen05 tf32 SS MMA: operand major (K / MN), M (64/128), to find the working
// descriptor conventions.  D = A(MxK) * B(KxN), K = 32 (4 MMAs), N = 32; B aliases the first N rows of A.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
constexpr int N = 32, KTOT = 32, KB = KTOT / 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
struct Cfg { int mn_major; int M; int swap; };
__global__ void __launch_bounds__(128) k(const float* Aglob /*[KTOT][128]*/, float* Dout /*[128][N] raw lanes*/, Cfg c) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = c.M;
    // layouts (bytes):
    //  MN-major: addr(f,k) = (f/4)*SBO + (k/8)*LBO + (k%8)*16 + (f%4)*4,  LBO = 128, SBO = KB*128
    //  K-major : addr(f,k) = (f/8)*SBO + (k/4)*LBO + (f%8)*16 + (k%4)*4,  LBO = 128, SBO = (KTOT/4)*128
    const uint32_t LBO = 128;
    const uint32_t SBO = c.mn_major ? KB * 128 : (KTOT / 4) * 128;
    for (int idx = tid; idx < KTOT * M; idx += 128) {
        const int kk = idx / M, f = idx % M;
        const float v = Aglob[kk * 128 + f];
        uint32_t off = c.mn_major ? (f / 4) * SBO + (kk / 8) * LBO + (kk % 8) * 16 + (f % 4) * 4
                                  : (f / 8) * SBO + (kk / 4) * LBO + (f % 8) * 16 + (kk % 4) * 4;
        *reinterpret_cast<float*>(smem + off) = v;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    // zero the accumulator region first so that "nothing written" is distinguishable
    {
        const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * warp) << 16);
        for (int col = 0; col < 32; ++col) {
            const uint32_t val = __float_as_uint(-777.0f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr0 + col), "r"(val));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.mn_major << 15) | ((uint32_t)c.mn_major << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        for (int kb = 0; kb < KB; ++kb) {
            // one MMA consumes K = 8: MN-major -> one k-group of 8 (advance by LBO); K-major -> two k-chunks of 4 (advance by 2*LBO)
            const uint32_t start = base + (c.mn_major ? kb * LBO : kb * 2 * LBO);
            const uint64_t adesc = c.swap ? make_desc(start, SBO, LBO) : make_desc(start, LBO, SBO);
            const uint32_t acc = kb > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem_base), "l"(adesc), "l"(adesc), "r"(idesc), "r"(acc));
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
    for (int col = 0; col < N; ++col) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + col));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        Dout[tid * N + col] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(32));
}
int main() {
    float* hA = (float*)malloc(sizeof(float) * KTOT * 128);
    for (int kk = 0; kk < KTOT; ++kk) for (int f = 0; f < 128; ++f) hA[kk * 128 + f] = (float)(((kk * 7 + f * 3) % 11) - 5) * 0.25f;
    static double ref[128][N];
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < KTOT; ++kk) s += (double)hA[kk * 128 + m] * hA[kk * 128 + n]; ref[m][n] = s; }
    float *dA, *dD; cudaMalloc(&dA, sizeof(float) * KTOT * 128); cudaMalloc(&dD, sizeof(float) * 128 * N);
    cudaMemcpy(dA, hA, sizeof(float) * KTOT * 128, cudaMemcpyHostToDevice);
    const int smem_bytes = 64 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    static float hD[128 * N];
    for (int mn = 0; mn < 2; ++mn) for (int M = 64; M <= 128; M += 64) for (int swap = 0; swap < 2; ++swap) {
        Cfg c{mn, M, swap};
        cudaMemset(dD, 0, sizeof(float) * 128 * N);
        k<<<1, 128, smem_bytes>>>(dA, dD, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("major=%s M=%d swap=%d: CUDA error %s\n", mn ? "MN" : "K", M, swap, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        int untouched = 0, match = 0, total = 0; double maxerr = 0;
        for (int m = 0; m < M; ++m) {
            const int lane = M == 128 ? m : (m % 16) + 32 * (m / 16);
            for (int n = 0; n < N; ++n) { float v = hD[lane * N + n]; ++total; if (v == -777.0f) ++untouched; double d = fabs(v - ref[m][n]); if (d < 1e-3) ++match; if (d > maxerr) maxerr = d; }
        }
        printf("major=%s M=%3d swap=%d : match %4d/%4d untouched %4d maxerr %.3g | lane0: %g %g %g %g (ref %g %g %g %g)\n", mn ? "MN" : "K ", M, swap, match, total, untouched, maxerr,
               hD[0], hD[1], hD[2], hD[3], ref[0][0], ref[0][1], ref[0][2], ref[0][3]);
    }
    return 0;
}
