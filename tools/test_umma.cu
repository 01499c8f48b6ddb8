// Stand-alone validation of the tcgen05 (UMMA) building block used by the smoothing kernel:
// D[64 x N] (TMEM, fp32) += A[64 x K] * B[K x N], tf32, both operands MN-major, no swizzle,
// operands written by threads (lane = k) into the canonical interleaved layout
//     addr(feature f, sample k) = (f/4)*SBO + (k/8)*128 + (k%8)*16 + (f%4)*4   bytes
// B aliases the first N "features" of A.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 64, N = 32, KTOT = 128;          // 128 samples = 16 MMAs of K = 8
constexpr int KB = KTOT / 8;
constexpr int SBO = KB * 128;                      // bytes between 4-feature groups
constexpr int STAGE_BYTES = (M / 4) * SBO;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    return d;                                       // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) k(const float* Aglob /*[KTOT][M]*/, float* Dout /*[M][N]*/, int variant_flags) {
    const int lbo_variant = variant_flags & 1;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    // thread tid = sample k writes its 64 features as 16 float4
    {
        const int kidx = tid;
        for (int g = 0; g < M / 4; ++g) {
            float4 v = *reinterpret_cast<const float4*>(Aglob + kidx * M + 4 * g);
            *reinterpret_cast<float4*>(smem + g * SBO + (kidx / 8) * 128 + (kidx % 8) * 16) = v;
        }
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
    if (variant_flags & 2) {
        // readout self-test: every thread stores (row-ish id * 100 + column) into its TMEM lane
        const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < 32; ++c) {
            const uint32_t val = __float_as_uint((float)(tid * 100 + c));
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr0 + c), "r"(val));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (tid == 0 && !(variant_flags & 2)) {
        // instruction descriptor: c=f32 (1<<4), a=tf32 (2<<7), b=tf32 (2<<10), a_major=MN (1<<15), b_major=MN (1<<16), N>>3 <<17, M>>4 <<24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        for (int kb = 0; kb < KB; ++kb) {
            const uint32_t lbo = lbo_variant == 0 ? 128u : (uint32_t)SBO;
            const uint32_t sbo = lbo_variant == 0 ? (uint32_t)SBO : 128u;
            const uint64_t adesc = make_desc(base + kb * 128, lbo, sbo);
            const uint64_t bdesc = adesc;                   // B = first N features of the same array
            const uint32_t acc = kb > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
    }
    // everyone waits for the MMAs
    if (!(variant_flags & 2)) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0));
        }
    }
    if (variant_flags & 4) __nanosleep(200000);
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    const int lane = tid & 31;
    if (variant_flags & 2) {
        if (tid == 37 || tid == 0 || tid == 127) printf("tid %d reads %g %g ... %g\n", tid, __uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[31]));
    }
    if (variant_flags & 16) {      // dump raw lanes: Dout[tid][n] for the first 64 threads' worth
        if (tid < 64) for (int n = 0; n < N; ++n) Dout[tid * N + n] = __uint_as_float(v[n]);
    } else if (lane < 16) {
        const int row = 16 * warp + lane;               // M = 64: row r -> TMEM lane (r%16) + 32*(r/16)
        for (int n = 0; n < N; ++n) Dout[row * N + n] = __uint_as_float(v[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(32));
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    float* hA = (float*)malloc(sizeof(float) * KTOT * M);
    for (int kk = 0; kk < KTOT; ++kk) for (int f = 0; f < M; ++f) hA[kk * M + f] = (float)(((kk * 7 + f * 3) % 11) - 5) * 0.25f;   // exact in tf32
    double ref[M][N];
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < KTOT; ++kk) s += (double)hA[kk * M + m] * hA[kk * M + n]; ref[m][n] = s; }
    float *dA, *dD; cudaMalloc(&dA, sizeof(float) * KTOT * M); cudaMalloc(&dD, sizeof(float) * M * N);
    cudaMemcpy(dA, hA, sizeof(float) * KTOT * M, cudaMemcpyHostToDevice); cudaMemset(dD, 0, sizeof(float) * M * N);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES);
    k<<<1, 128, STAGE_BYTES>>>(dA, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d status: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    float hD[M * N]; cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double d = fabs(hD[m * N + n] - ref[m][n]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
    { int nz = 0; for (int i = 0; i < M * N; ++i) if (hD[i] != 0.f) ++nz; printf("nonzero outputs: %d\n", nz); }
    printf("max abs err %.3g, mismatches %d / %d ; D[0][0..3] = %g %g %g %g  ref %g %g %g %g ; D[37][5]=%g ref %g\n", maxerr, bad, M * N,
           hD[0], hD[1], hD[2], hD[3], ref[0][0], ref[0][1], ref[0][2], ref[0][3], hD[37 * N + 5], ref[37][5]);
    return 0;
}
