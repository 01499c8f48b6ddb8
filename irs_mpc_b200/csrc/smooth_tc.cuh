// Zero-order smoothing with the Gram accumulation on the 5th-generation tensor cores (tcgen05).
//
// Per sample the fit needs the rank-1 update  [dx du]^T [dx du | dF]  (d x (d+n) outputs; 16 x 28
// for the quadrotor: 78 % of the algorithmic flops of that kernel).  On the CUDA cores this costs
// 164 FFMA2 issue slots and 84 accumulator registers per thread; here it is one UMMA per 16 samples,
// issued by a single thread, accumulating in TMEM:
//
//     D[64 x N] (TMEM, fp32)  +=  A[64 x 16] (smem)  *  B[16 x N] (smem),     kind::f16 (bf16), MN-major
//
// A rows ("features") = [ z_1 | z_2 | dF_1 | dF_2 ],  B = the first N = 2*dq rows of the SAME smem
// array = [ z_1 | z_2 ].  Every fp32 value x is split into two bf16 pieces, x ~ x_1 + x_2 with
// x_1 = bf16_rn(x), x_2 = bf16_rn(x - x_1)  (|x - x_1 - x_2| <= 2^-18 |x|, round-to-nearest, so the
// residual is zero-mean and averages out over the samples); D then holds all four partial products
// and their sum reproduces the fp32 product to ~2^-17 relative — far inside the 1e-4 budget, and the
// bf16 operands halve the shared-memory traffic of a tf32 split (the UMMA operand reads were the
// bottleneck of the tf32 variant: 3 KB per 8 samples against 128 B/clk of smem bandwidth).
//
// MN-major canonical layout without swizzle (conventions verified on hardware by
// tools/test_umma3.cu):
//     byte address of (feature f, sample k) = (f/8)*SBO + (k/8)*LBO + (k%8)*16 + (f%8)*2
// with LBO = 128 B, SBO = (samples per stage / 8) * 128 B: a thread (lane = sample) stores 8
// consecutive features with one 16-byte STS and a warp's store covers 512 contiguous bytes.
#pragma once
#include "smooth.cuh"

namespace irs {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, Blackwell version bit).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

template <class Sys>
struct TcCfg {
    static constexpr int n = Sys::N, m = Sys::M, d = Sys::D;
    static constexpr int W = d + n;
    static constexpr int dq = (d + 7) / 8 * 8;           // regressor rows per piece (8-feature groups)
    static constexpr int nq = (n + 7) / 8 * 8;           // response rows per piece
    static constexpr int kN = 2 * dq;                    // UMMA N (columns of D): [z_1 | z_2]
    static constexpr int kRows = 2 * dq + 2 * nq;        // used rows of A (<= 64)
    static constexpr int kM = 64;
    static constexpr int kThreads = 128;                 // 4 warps, lane = sample
    static constexpr int kTile = 128;                    // samples per stage
    static constexpr int kLBO = 128;                     // bytes between k-groups (8 samples)
    static constexpr int kSBO = (kTile / 8) * kLBO;      // bytes between 8-feature groups (2048)
    static constexpr int kGroups = kM / 8;
    static constexpr int kStageBytes = kGroups * kSBO;   // 16,384
    static constexpr int kTmemCols = kN < 32 ? 32 : kN;
    static constexpr int NACC = gram_nacc(n, m);
    static constexpr int RS = (W + 1) / 2 * 2;
    static_assert(kRows <= kM, "operand rows must fit one M = 64 UMMA");
    static_assert(kN % 8 == 0 && kN >= 8 && kN <= 256, "invalid UMMA N");
    static_assert(kM * kN * 4 <= kStageBytes, "read-back scratch must fit a stage");
    // instruction descriptor: D fp32, A/B bf16, both MN-major, N >> 3, M >> 4
    static constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    // feature-group index (8 rows each) of the four pieces
    static constexpr int grp_z1 = 0, grp_z2 = dq / 8, grp_f1 = 2 * dq / 8, grp_f2 = (2 * dq + nq) / 8;
    __host__ __device__ static constexpr int row_1(int j) { return j < d ? j : 2 * dq + (j - d); }
    __host__ __device__ static constexpr int row_2(int j) { return j < d ? dq + j : 2 * dq + nq + (j - d); }
};

// fp32 pair -> two packed bf16x2 words: first pieces and second pieces (low half = v0, high half = v1)
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& p1, uint32_t& p2) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(v1), "f"(v0));
    const float r0 = v0 - __uint_as_float(p1 << 16);
    const float r1 = v1 - __uint_as_float(p1 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
}

template <class Sys, int NSTAGE>
__global__ void __launch_bounds__(128) smooth_zero_order_tc_kernel(const SmoothArgs a) {
    using C = TcCfg<Sys>;
    constexpr int n = Sys::N, m = Sys::M, d = Sys::D;
    extern __shared__ __align__(1024) unsigned char stage_mem[];
    __shared__ uint64_t mbar_empty[NSTAGE];
    __shared__ uint64_t mbar_done;
    __shared__ uint32_t tmem_base_s;

    const Sys sys(a.prm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = blockIdx.x / a.C;
    const int c = blockIdx.x % a.C;
    const long long s_begin = (long long)c * a.S;
    const long long s_end = s_begin + a.S < a.N ? s_begin + a.S : a.N;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&mbar_empty[s], 1);
        mbar_init(&mbar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(C::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    float xbar[n], ubar[m], fbar[n];
#pragma unroll
    for (int q = 0; q < n; ++q) xbar[q] = (float)a.x_nom[(long long)p * n + q];
#pragma unroll
    for (int q = 0; q < m; ++q) ubar[q] = (float)a.u_nom[(long long)p * m + q];
    sys.template step<false>(xbar, ubar, fbar);      // scalar dynamics at the nominal (…zero_order.py:52)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t stage_base = smem_u32(stage_mem);
    const bool batch = (a.flags & kFlagSamplesBatchVariant) != 0;

    // byte offset of this thread's sample inside a stage (feature group 0): (k/8)*LBO + (k%8)*16
    const uint32_t my_off = (uint32_t)(tid >> 3) * C::kLBO + (uint32_t)(tid & 7) * 16;

    int round = 0;
    for (long long base = s_begin; base < s_end; base += C::kTile, ++round) {
        const int stage = round % NSTAGE;
        if (round >= NSTAGE) mbar_wait(&mbar_empty[stage], (uint32_t)((round / NSTAGE - 1) & 1));
        unsigned char* sm = stage_mem + stage * C::kStageBytes + my_off;
        {
            float w[C::RS];
#pragma unroll
            for (int q = 0; q < C::RS; ++q) w[q] = 0.f;
            const long long s = base + tid;
            if (s < s_end) {
                if (batch) make_sample<Sys, true, C::RS>(sys, a, p, s, xbar, ubar, fbar, w);
                else make_sample<Sys, false, C::RS>(sys, a, p, s, xbar, ubar, fbar, w);
            }
            // regressors: dq/8 groups of 8 features, first and second bf16 pieces
#pragma unroll
            for (int g = 0; g < C::dq / 8; ++g) {
                uint32_t p1[4], p2[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c0 = 8 * g + 2 * q, c1 = c0 + 1;
                    split_bf16x2(c0 < d ? w[c0] : 0.f, c1 < d ? w[c1] : 0.f, p1[q], p2[q]);
                }
                *reinterpret_cast<uint4*>(sm + (C::grp_z1 + g) * C::kSBO) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
                *reinterpret_cast<uint4*>(sm + (C::grp_z2 + g) * C::kSBO) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
            }
            // responses
#pragma unroll
            for (int g = 0; g < C::nq / 8; ++g) {
                uint32_t p1[4], p2[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c0 = 8 * g + 2 * q, c1 = c0 + 1;
                    split_bf16x2(c0 < n ? w[d + c0] : 0.f, c1 < n ? w[d + c1] : 0.f, p1[q], p2[q]);
                }
                *reinterpret_cast<uint4*>(sm + (C::grp_f1 + g) * C::kSBO) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
                *reinterpret_cast<uint4*>(sm + (C::grp_f2 + g) * C::kSBO) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (UMMA)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t sbase = stage_base + stage * C::kStageBytes;
#pragma unroll 4
            for (int kb = 0; kb < C::kTile / 16; ++kb) {       // one UMMA = K 16 = two k-groups of 8 samples
                const uint64_t desc = umma_smem_desc(sbase + kb * 2 * C::kLBO, C::kLBO, C::kSBO);
                const uint32_t acc = (round > 0 || kb > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base),
                    "l"(desc), "l"(desc), "r"(C::kIdesc), "r"(acc)
                    : "memory");
            }
            // frees the stage for the round that reuses it (tcgen05.commit implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&mbar_empty[stage]))
                         : "memory");
        }
    }
    if (tid == 0)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&mbar_done))
                     : "memory");
    mbar_wait(&mbar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- read the accumulator back: row r lives in TMEM lane (r % 16) + 32 * (r / 16) (M = 64) ----
    float* scratch = reinterpret_cast<float*>(stage_mem);            // [64][kN], all UMMAs are done
    {
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
        const int row = 16 * warp + lane;
#pragma unroll
        for (int c0 = 0; c0 < C::kN; c0 += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (lane < 16) {
#pragma unroll
                for (int q = 0; q < 8; ++q) scratch[row * C::kN + c0 + q] = __uint_as_float(v[q]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols));
    // G[i][j] = sum_k z_i w_j = D[w_j rows][z_i cols], hi and lo blocks added
    float* out = a.partials + ((long long)p * a.C + c) * C::NACC;
    for (int e = tid; e < C::NACC; e += C::kThreads) {
        int i = 0;
        while (i + 1 < d && gram_row_offset(i + 1, C::W) <= e) ++i;
        const int j = i + (e - gram_row_offset(i, C::W));
        const int rh = C::row_1(j), rl = C::row_2(j);
        const int ch = i, cl = C::dq + i;
        out[e] = (scratch[rh * C::kN + ch] + scratch[rl * C::kN + cl]) +
                 (scratch[rh * C::kN + cl] + scratch[rl * C::kN + ch]);
    }
}

}  // namespace irs
