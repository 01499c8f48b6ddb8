// Zero-order smoothing of LEARNED dynamics (systems.cuh: Mlp; the reference's examples/pendulum/pendulum_nn.py)
// with the first and the hidden layer on the 5th-generation tensor cores.
//
// Per sample the network costs d H1 + H1 H2 + H2 n multiply-adds (3 x 100 + 100 x 100 + 100 x 2 for the
// reference's network): 98 % of them are the first two layers, dense GEMMs over the samples of a tile:
//
//     D1[128 x Kp] (TMEM, fp32) = A1[128 x 16] (smem) * B1[16 x Kp] (smem)         first layer, pre-activations
//     D [128 x Np] (TMEM, fp32) = A [128 x Kp] (smem) * B [Kp x Np] (smem)         hidden layer
//                                                                   kind::f16 (bf16), all operands K-major
// A1 row = one sample [x; u; 1 | 0 ..];  B1 = [W1 | b1]^T plus one constant unit (column H1 = 1)
// A  row = relu(D1 row): the sample's hidden activations, the constant 1 in column H1 carries b2
// B      = [W2 | b2]^T;  B1, B prepared once per registered network (api.cu: irs_mlp_register)
// Kp = ceil16(H1 + 1), Np = ceil16(H2).  Every operand is split into two bf16 pieces (x ~ x_hi + x_lo,
// |x - x_hi - x_lo| <= 2^-17 |x|) and three products are accumulated, a_hi w_hi + a_hi w_lo + a_lo w_hi: the
// float32 product to ~2^-16 relative, inside the 1e-4 budget with margin (the reference evaluates the torch
// module in float32, pendulum_nn.py:72-81).  Residual pieces of A1 / A are staged NEGATED (split_bf16x2) and the
// third product is issued with the negate-A bit of the instruction descriptor.
//
// K-major canonical layout without swizzle (verified on hardware by tools/test_umma_mlp.cu):
//     byte address of (row r, k) = (r/8)*128 + (k/8)*LBO + (r%8)*16 + (k%8)*2,   LBO = (rows/8)*128
// A thread (= sample row) stores eight consecutive k with one 16-byte STS; a warp's store is 512 contiguous bytes.
//
// A GROUP of 128 threads works on one tile of 128 samples at a time: every thread draws its sample and stages its
// A1 row (and the group B1) in the part of the operand tile that is still free, one elected thread issues the three
// UMMAs of the first layer, every thread reads its own row of D1 back (TMEM lane = sample), applies ReLU, splits
// and stores its A row; the elected thread issues the 3 Kp/16 UMMAs of the hidden layer into the same TMEM
// columns; every thread reads its row of D, applies ReLU and the last layer (CUDA cores) and updates its Gram
// registers (smooth.cuh: gram_update — the same packed block the generic kernel writes, so the fp64 fit is shared).
#pragma once
#include "smooth_tc.cuh"

namespace irs {

// 16 accumulator columns of this thread's TMEM lane -> registers (asynchronous), and the wait that makes them
// readable.  The wait names the registers as in / out operands, so the compiler cannot move a use above it; loads
// are issued one chunk ahead of the chunk being processed (the TMEM load latency hides behind the arithmetic).
#define IRS_TMEM_LD16(v, taddr)                                                                                             \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"    \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),           \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])      \
                 : "r"(taddr))
#define IRS_TMEM_WAIT16(v)                                                                                                  \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                           \
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),           \
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])      \
                 :                                                                                                          \
                 : "memory")

struct MlpTcLayout {
    int H1, H2, Kp, Np;
    __host__ __device__ MlpTcLayout(int h1, int h2)
        : H1(h1), H2(h2), Kp((h1 + 1 + 15) / 16 * 16), Np((h2 + 15) / 16 * 16) {}
    static constexpr int kRows = 128;                                        // samples per tile (UMMA M)
    __host__ __device__ int lbo_a() const { return kRows / 8 * 128; }
    __host__ __device__ int lbo_b() const { return Np / 8 * 128; }
    __host__ __device__ int a_piece_bytes() const { return kRows * Kp * 2; }
    __host__ __device__ int b_piece_bytes() const { return Np * Kp * 2; }
    // byte offset of element (row j, k) inside a B piece
    __host__ __device__ int b_offset(int j, int k) const { return (j / 8) * 128 + (k / 8) * lbo_b() + (j % 8) * 16 + (k % 8) * 2; }
    // first layer as a UMMA too: D1[128 x Kp] = [x; u; 1 | 0..] (K = 16) * B1^T, B1 row j = (W1[j], b1[j], 0..) and row H1 =
    // (0, .., 0, 1) — the constant unit that carries b2 through the hidden layer
    static constexpr int kK1 = 16;
    __host__ __device__ int lbo_b1() const { return Kp / 8 * 128; }
    __host__ __device__ int b1_piece_bytes() const { return Kp * kK1 * 2; }
    __host__ __device__ int a1_piece_bytes() const { return kRows * kK1 * 2; }
    __host__ __device__ int b1_offset(int j, int k) const { return (j / 8) * 128 + (k / 8) * lbo_b1() + (j % 8) * 16 + (k % 8) * 2; }
    // TMEM columns of a tile group: the accumulator holds the first layer's Kp, then the hidden layer's Np columns
    __host__ __device__ int tmem_cols_per_group() const { return (Kp > 128 || Np > 128) ? 256 : 128; }
    static __host__ __device__ int tmem_alloc_cols(int cols) { return cols <= 128 ? 128 : (cols <= 256 ? 256 : 512); }
    // bytes the first-layer operands of a tile need: A1 hi | A1 lo | B1 hi | B1 lo.  They live in the residual piece of
    // the operand tile (free until the activations are written) when it is large enough, else in a region of their own
    __host__ __device__ int stage_bytes() const { return 2 * a1_piece_bytes() + 2 * b1_piece_bytes(); }
    __host__ __device__ bool stage_in_tile() const { return a_piece_bytes() >= stage_bytes(); }
    // device blob of operand pieces: [W2 hi | W2 lo | B1 hi | B1 lo]
    __host__ __device__ int blob_bytes() const { return 2 * b_piece_bytes() + 2 * b1_piece_bytes(); }
};

template <class Sys>
struct MlpTcSmem {
    static constexpr int n = Sys::N, d = Sys::D;
    static constexpr int WIDTH = gram_width_of<Sys>();
    // byte offsets of the regions behind the operand tiles
    static constexpr int kMaxGroups = 3;
    MlpTcLayout L;
    int G;
    int a_hi, a_lo, a_stage, a_stride, b_hi, b_lo, w1b, w3, act1, act2, fbar, slabs, bar, total;
    __host__ __device__ MlpTcSmem(const MlpTcLayout& l, int groups) : L(l), G(groups) {
        int o = 0;
        // per group: [hi piece | lo piece (| first-layer staging, when the lo piece is too small for it)]
        a_stride = 2 * L.a_piece_bytes() + (L.stage_in_tile() ? 0 : L.stage_bytes());
        a_hi = o;
        a_lo = o + L.a_piece_bytes();
        a_stage = L.stage_in_tile() ? a_lo : o + 2 * L.a_piece_bytes();
        o += G * a_stride;
        b_hi = o;  o += L.b_piece_bytes();
        b_lo = o;  o += L.b_piece_bytes();
        w1b = o;   o += L.Kp * (d + 1) * 4;          // [Kp][d + 1]: first-layer row and bias; row H1 = (0, .., 0, 1)
        w3 = o;    o += (n * L.Np + n) * 4;          // [n][Np] last layer (zero padded) | b3[n]
        act1 = o;  o += G * kMlpMaxHidden * 4;       // per group — nominal point: hidden activations
        act2 = o;  o += G * kMlpMaxHidden * 4;
        fbar = o;  o += G * ((n + 3) / 4 * 4) * 4;
        slabs = o; o += 4 * G * WIDTH * 4;
        o = (o + 7) / 8 * 8;
        bar = o;   o += 8 * kMaxGroups + 8;          // one mbarrier per group | TMEM base address
        total = o;
    }
    // groups per block that fit `limit` bytes of shared memory (0: not even one)
    static __host__ int groups_for(const MlpTcLayout& l, int limit) {
        for (int g = kMaxGroups; g >= 1; --g)
            if (MlpTcSmem(l, g).total <= limit && g * l.tmem_cols_per_group() <= 512) return g;
        return 0;
    }
};

// (Measured and dropped: the first / last layer weights as a kernel parameter — constant-bank loads — with the
// layers unrolled for the reference's 100 / 100 network: LDC of the first layer is slower than the broadcast LDS,
// uniform-register operands for the last layer gain 3 %.)
template <class Sys>
__global__ void __launch_bounds__(128 * MlpTcSmem<Sys>::kMaxGroups) smooth_zero_order_mlp_kernel(const SmoothArgs a) {
    static_assert(Sys::kIsMlp, "learned dynamics only");
    using C = ZeroOrderCfg<Sys, 1>;
    constexpr int n = Sys::N, m = Sys::M, d = Sys::D;
    constexpr int NPAIR = role_pairs(d, C::Wp, 1, 0);
    constexpr int WIDTH = gram_width_of<Sys>();
    static_assert(WIDTH == C::NACC, "no first moments for learned dynamics");
    const Sys sys(a.prm);
    const MlpView& net = sys.net;
    const MlpTcLayout L(net.H1, net.H2);
    const int G = blockDim.x >> 7;                      // tile groups of this block
    const MlpTcSmem<Sys> sm(L, G);
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = tid >> 7, gtid = tid & 127, gwarp = warp & 3;
    float* w1b = reinterpret_cast<float*>(smem + sm.w1b);
    float* w3s = reinterpret_cast<float*>(smem + sm.w3);
    // per group: scratch of the nominal point, warp slabs of the Gram flush, mbarrier
    float* act1 = reinterpret_cast<float*>(smem + sm.act1) + group * kMlpMaxHidden;
    float* act2 = reinterpret_cast<float*>(smem + sm.act2) + group * kMlpMaxHidden;
    float* fbar_s = reinterpret_cast<float*>(smem + sm.fbar) + group * ((n + 3) / 4 * 4);
    float* slabs = reinterpret_cast<float*>(smem + sm.slabs) + group * 4 * WIDTH;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + sm.bar) + group;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + sm.bar + 8 * MlpTcSmem<Sys>::kMaxGroups);
    const int nthreads = blockDim.x;
    const uint32_t tmem_cols = (uint32_t)MlpTcLayout::tmem_alloc_cols(G * L.tmem_cols_per_group());

    // ---- once per block: the network ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.prm.mlp_w2);
        uint4* dst = reinterpret_cast<uint4*>(smem + sm.b_hi);
        for (int e = tid; e < 2 * L.b_piece_bytes() / 16; e += nthreads) dst[e] = __ldg(src + e);
        for (int e = tid; e < L.Kp * (d + 1); e += nthreads) {
            const int j = e / (d + 1), q = e % (d + 1);
            float v = 0.f;
            if (j < L.H1) v = q < d ? __ldg(net.w1 + j * d + q) : __ldg(net.b1 + j);
            else if (j == L.H1 && q == d) v = 1.f;                    // the constant column that carries b2
            w1b[e] = v;
        }
        for (int e = tid; e < n * L.Np; e += nthreads) {
            const int k = e / L.Np, j = e % L.Np;
            w3s[e] = j < L.H2 ? __ldg(net.w3 + k * L.H2 + j) : 0.f;
        }
        if (tid < n) w3s[n * L.Np + tid] = __ldg(net.b3 + tid);
    }
    if (gtid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_s;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(L.tmem_cols_per_group() * group);   // this group's accumulator columns
    const uint32_t tmem_row = tmem_acc + ((uint32_t)(32 * gwarp) << 16);           // this warp's lane quarter
    // instruction descriptor: D fp32, A / B bf16, both K-major, N >> 3, M >> 4; bit 13 negates A
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_neg_a = idesc | (1u << 13);
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.Kp >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    unsigned char* a_tile = smem + group * sm.a_stride;
    const uint32_t sa_hi = smem_u32(a_tile + sm.a_hi), sa_lo = smem_u32(a_tile + sm.a_lo);
    const uint32_t sb_hi = smem_u32(smem + sm.b_hi), sb_lo = smem_u32(smem + sm.b_lo);
    uint32_t phase = 0;

    // the groups of a block are independent workers (they share the network in shared memory and nothing else):
    // each walks its own items and synchronises on its own named barrier, so their phases drift apart and one
    // group's UMMA / TMEM latency is covered by the others' CUDA-core work
    auto group_sync = [&] { asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory"); };
    const int items = a.P * a.C;
    for (int item = blockIdx.x * G + group; item < items; item += gridDim.x * G) {
        const int p = item / a.C, c = item % a.C;
        const long long s_begin = (long long)c * a.S;
        const long long s_end = s_begin + a.S < a.N ? s_begin + a.S : a.N;
        float nom[d];
#pragma unroll
        for (int q = 0; q < n; ++q) nom[q] = (float)a.x_nom[(long long)p * n + q];
#pragma unroll
        for (int q = 0; q < m; ++q) nom[n + q] = (float)a.u_nom[(long long)p * m + q];
        // ---- f(xbar, ubar) in float32 with the exact weights, the group's threads over the units ----
        if (gtid < L.H1) {
            float s = w1b[gtid * (d + 1) + d];
#pragma unroll
            for (int q = 0; q < d; ++q) s = fmaf(w1b[gtid * (d + 1) + q], nom[q], s);
            act1[gtid] = fmaxf(s, 0.f);
        }
        group_sync();
        // W2^T: the threads of a warp read consecutive words (a row of W2 per thread is a cache line per load)
        if (gtid < L.H2) act2[gtid] = fmaxf(Sys::dot_row(net.w2t + gtid, act1, L.H1, __ldg(net.b2 + gtid), L.H2), 0.f);
        group_sync();
        if (gtid < n) fbar_s[gtid] = Sys::dot_row(net.w3 + (long long)gtid * L.H2, act2, L.H2, __ldg(net.b3 + gtid));
        group_sync();
        float fbar[n];
#pragma unroll
        for (int k = 0; k < n; ++k) fbar[k] = fbar_s[k];

        float2 acc[NPAIR];
#pragma unroll
        for (int k = 0; k < NPAIR; ++k) acc[k] = make_float2(0.f, 0.f);
        for (long long s0 = s_begin; s0 < s_end; s0 += 128) {
            const long long s = s0 + gtid;
            const bool valid = s < s_end;
            float w[C::RS];
#pragma unroll
            for (int q = 0; q < C::RS; ++q) w[q] = 0.f;
            if (valid) draw_deltas<Sys, C::RS>(a, p, s, w);
            float in[d];
#pragma unroll
            for (int q = 0; q < d; ++q) in[q] = nom[q] + w[q];
            // ---- first layer on the tensor cores: stage [x; u; 1] (two bf16 pieces, K = 16) and B1 in the part of the
            //      operand tile that is free until the activations are written (the residual piece) ----
            {
                static_assert(d + 1 <= 8, "inputs and the constant fit the first k-group");
                unsigned char* st = a_tile + sm.a_stage;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = q < d ? in[q] : (q == d ? 1.f : 0.f);
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split_bf16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
                const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(st + gtid * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(st + L.lbo_a() + gtid * 16) = zero;
                *reinterpret_cast<uint4*>(st + L.a1_piece_bytes() + gtid * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(st + L.a1_piece_bytes() + L.lbo_a() + gtid * 16) = zero;
                const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.prm.mlp_w2) + 2 * L.b_piece_bytes());
                uint4* dst = reinterpret_cast<uint4*>(st + 2 * L.a1_piece_bytes());
                for (int e = gtid; e < 2 * L.b1_piece_bytes() / 16; e += 128) dst[e] = __ldg(src + e);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");       // this thread's accumulator reads of the last tile
            group_sync();
            if (gwarp == 0) {
                if (elect_one()) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t base = smem_u32(a_tile + sm.a_stage);
                    const uint32_t la = (uint32_t)L.lbo_a(), l1 = (uint32_t)L.lbo_b1();
                    const uint64_t ah = umma_smem_desc(base, la, 128);
                    const uint64_t al = umma_smem_desc(base + L.a1_piece_bytes(), la, 128);
                    const uint64_t bh = umma_smem_desc(base + 2 * L.a1_piece_bytes(), l1, 128);
                    const uint64_t bl = umma_smem_desc(base + 2 * L.a1_piece_bytes() + L.b1_piece_bytes(), l1, 128);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem_acc), "l"(ah), "l"(bh), "r"(idesc1), "r"(0u));
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem_acc), "l"(ah), "l"(bl), "r"(idesc1), "r"(1u));
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem_acc), "l"(al), "l"(bh), "r"(idesc1 | (1u << 13)), "r"(1u));
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                                 : "memory");
                }
                __syncwarp();
            }
            mbar_wait(mbar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- own row of pre-activations: ReLU, two bf16 pieces -> operand row of the hidden layer ----
            {
                unsigned char* row_hi = a_tile + sm.a_hi + gtid * 16;
                unsigned char* row_lo = a_tile + sm.a_lo + gtid * 16;
                auto to_operand = [&](const uint32_t (&v)[16], int c0) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        split_bf16x2(fmaxf(__uint_as_float(v[2 * q]), 0.f), fmaxf(__uint_as_float(v[2 * q + 1]), 0.f), hi[q], lo[q]);
                    *reinterpret_cast<uint4*>(row_hi + (c0 >> 3) * L.lbo_a()) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(row_hi + ((c0 >> 3) + 1) * L.lbo_a()) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    *reinterpret_cast<uint4*>(row_lo + (c0 >> 3) * L.lbo_a()) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4*>(row_lo + ((c0 >> 3) + 1) * L.lbo_a()) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                };
                uint32_t va[16], vb[16];
                IRS_TMEM_LD16(va, tmem_row);
                for (int c0 = 0; c0 < L.Kp; c0 += 32) {
                    IRS_TMEM_WAIT16(va);
                    if (c0 + 16 < L.Kp) IRS_TMEM_LD16(vb, tmem_row + (uint32_t)(c0 + 16));
                    to_operand(va, c0);
                    if (c0 + 16 < L.Kp) {
                        IRS_TMEM_WAIT16(vb);
                        if (c0 + 32 < L.Kp) IRS_TMEM_LD16(va, tmem_row + (uint32_t)(c0 + 32));
                        to_operand(vb, c0 + 16);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");       // this thread's accumulator reads of the last tile
            group_sync();                                                          // the group's operand tile is complete
            // ---- hidden layer: 3 Kp / 16 UMMAs, one issuing thread per group ----
            if (gwarp == 0) {
                if (elect_one()) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t la = (uint32_t)L.lbo_a(), lb = (uint32_t)L.lbo_b();
                    for (int kb = 0; kb < L.Kp / 16; ++kb) {
                        const uint64_t ah = umma_smem_desc(sa_hi + kb * 2 * la, la, 128);
                        const uint64_t al = umma_smem_desc(sa_lo + kb * 2 * la, la, 128);
                        const uint64_t bh = umma_smem_desc(sb_hi + kb * 2 * lb, lb, 128);
                        const uint64_t bl = umma_smem_desc(sb_lo + kb * 2 * lb, lb, 128);
                        const uint32_t first = kb > 0 ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem_acc), "l"(ah), "l"(bh), "r"(idesc), "r"(first));
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem_acc), "l"(ah), "l"(bl), "r"(idesc), "r"(1u));
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem_acc), "l"(al), "l"(bh), "r"(idesc_neg_a), "r"(1u));
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                                 : "memory");
                }
                __syncwarp();
            }
            mbar_wait(mbar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- own accumulator row: ReLU, last layer ----
            float o[n];
#pragma unroll
            for (int k = 0; k < n; ++k) o[k] = w3s[n * L.Np + k];
            {
                auto last_layer = [&](const uint32_t (&v)[16], int c0) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float h = fmaxf(__uint_as_float(v[q]), 0.f);
#pragma unroll
                        for (int k = 0; k < n; ++k) o[k] = fmaf(w3s[k * L.Np + c0 + q], h, o[k]);
                    }
                };
                uint32_t va[16], vb[16];
                IRS_TMEM_LD16(va, tmem_row);
                for (int c0 = 0; c0 < L.Np; c0 += 32) {
                    IRS_TMEM_WAIT16(va);
                    if (c0 + 16 < L.Np) IRS_TMEM_LD16(vb, tmem_row + (uint32_t)(c0 + 16));
                    last_layer(va, c0);
                    if (c0 + 16 < L.Np) {
                        IRS_TMEM_WAIT16(vb);
                        if (c0 + 32 < L.Np) IRS_TMEM_LD16(va, tmem_row + (uint32_t)(c0 + 32));
                        last_layer(vb, c0 + 16);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < n; ++k) w[d + k] = valid ? o[k] - fbar[k] : 0.f;
            gram_update<Sys, 1, 0, NPAIR>(w, acc);
        }
        // ---- packed Gram block of the item (as zero_order_registers) ----
        float* slab = slabs + gwarp * WIDTH;
        gram_flush<Sys, 1, 0, NPAIR>(acc, slab, lane);
        group_sync();
        float* out = a.partials + (long long)item * WIDTH;
        for (int e = gtid; e < WIDTH; e += 128)
            out[e] = (slabs[e] + slabs[WIDTH + e]) + (slabs[2 * WIDTH + e] + slabs[3 * WIDTH + e]);
        group_sync();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
}

}  // namespace irs
