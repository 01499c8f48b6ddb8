// Shared device helpers for the iRS-MPC sm_100a kernels: error plumbing, Philox4x32,
// Box-Muller, packed-FP32 FMA (FFMA2), scalar-type math wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace irs {

// ---------------------------------------------------------------------------------------------
// Error plumbing (C-ABI returns int status; message retrievable through irs_last_error()).
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define IRS_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            irs::set_error(__VA_ARGS__);  \
            return 1;                     \
        }                                 \
    } while (0)

// System parameters travel by value into the kernels (<= 10 doubles).  `f` mirrors `v` in fp32
// followed by the derived loop invariants of the system (filled on the host in double precision,
// load_params in api.cu), so that the fp32 sample kernels read them as constant-bank operands
// instead of converting doubles (F2F runs on the XU pipe, the busiest one in the sample loop).
struct SysParams {
    double v[10];
    float f[20];
    // learned dynamics (systems.cuh: Mlp): device blob of the registered network, hidden widths
    const float* mlp;
    const unsigned short* mlp_w2;      // hidden layer as tensor-core operand: bf16 pieces [hi | lo] (smooth_mlp.cuh)
    int h1, h2;
};

// ---------------------------------------------------------------------------------------------
// Scalar math wrappers.  float: MUFU-backed fast intrinsics (the sample path is throughput
// bound and the least-squares fit averages their ~1e-7 absolute error away); double: libdevice.
// ---------------------------------------------------------------------------------------------
template <typename R>
struct Math;

template <>
struct Math<float> {
    static __device__ __forceinline__ void sincos(float a, float& s, float& c) {
        __sincosf(a, &s, &c);
    }
    static __device__ __forceinline__ float rcp(float a) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));      // one MUFU.RCP, no slow path
        return r;
    }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
};

template <>
struct Math<double> {
    static __device__ __forceinline__ void sincos(double a, double& s, double& c) {
        ::sincos(a, &s, &c);
    }
    static __device__ __forceinline__ double rcp(double a) { return 1.0 / a; }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
};

// ---------------------------------------------------------------------------------------------
// Philox4x32 (Salmon et al., SC'11).  Spec of the counter layout: oracle/philox_ref.py.
// The sample stream uses kPhiloxRounds = 7 rounds: the smallest round count of Philox4x32 that
// the paper reports as Crush-resistant (10 is its default safety margin).  On sm_100a the two
// 32x32->64 multiplies per round issue at a quarter of the FP32 rate (measured: 32
// IMAD.WIDE/clk/SM vs 128 FFMA/clk/SM), so the rounds are the single largest item of the
// sample-generation cost.  The round function itself is pinned on the Random123 10-round
// known-answer vectors (tests/test_oracle_philox.py).
// ---------------------------------------------------------------------------------------------
constexpr int kPhiloxRounds = 7;
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

template <int ROUNDS = kPhiloxRounds>
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0);
        const uint32_t lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2);
        const uint32_t lo1 = kPhiloxM1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// float in [1,2) from the top 23 bits of w (exact).
__device__ __forceinline__ float unit_float(uint32_t w) {
    return __uint_as_float((w >> 9) | 0x3f800000u);
}

// Box-Muller on two 32-bit words, returned UNSCALED:  (g0, g1) = sqrt(-log2 u) * (cos, sin)(2 pi f)
// with u = 2 - unit_float(wa) in (0,1] and f = unit_float(wb) in [1,2).  The standard normals of
// the spec (oracle/philox_ref.py: angle 2 pi (f - 1.5), radius sqrt(-2 ln u)) are
//     e = -sqrt(2 ln 2) * g        (cos/sin(x - 3 pi) = -cos/sin(x))
// so callers fold kBoxMullerScale into sigma once instead of spending FMULs per sample.
constexpr float kBoxMullerScale = -1.1774100225154747f;     // -sqrt(2 ln 2)

__device__ __forceinline__ void box_muller_raw(uint32_t wa, uint32_t wb, float& g0, float& g1) {
    const float u = 2.0f - unit_float(wa);                      // (0, 1]
    float l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));      // MUFU.LG2 (u >= 2^-23: never denormal)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-l2));    // MUFU.SQRT
    float s, c;
    __sincosf(unit_float(wb) * 6.283185307179586f, &s, &c);     // argument in [2 pi, 4 pi)
    g0 = r * c;
    g1 = r * s;
}

// ---------------------------------------------------------------------------------------------
// Packed FP32 FMA (Blackwell FFMA2): d = a * b + c on two lanes, one issue slot.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t ua, ub, uc, ud;
    ua = *reinterpret_cast<uint64_t*>(&a);
    ub = *reinterpret_cast<uint64_t*>(&b);
    uc = *reinterpret_cast<uint64_t*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
    return *reinterpret_cast<float2*>(&ud);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace irs
