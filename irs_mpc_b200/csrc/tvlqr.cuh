// Time-varying LQR: backward affine Riccati recursion + closed-loop rollout on the true
// dynamics, one warp per independent MPC instance, fp64 (sequential and latency bound; the
// B200 has full-rate-enough FP64 FMA for this and it keeps the cost curves at round-off of the
// reference's float64).
//
// Replaces the T re-solved QPs of irs_lqr/irs_lqr.py:169-184 + irs_lqr/tv_lqr.py:69-137:
//   cost  sum_{s<T} (x_s-xd_s)'Q(x_s-xd_s) + 1/2 u_s'R u_s + (x_T-xd_T)'Qd(x_T-xd_T)
//   s.t.  x_{s+1} = A_s x_s + B_s u_s + c_s
// (Drake's AddQuadraticCost(R,0,u) is 1/2 u'Ru, tv_lqr.py:110).  With inactive bounds the first
// input of every re-solved QP is u_t = K_t x_t + k_t from ONE backward pass (Bellman).
#pragma once
#include "systems.cuh"

namespace irs {

struct TvlqrArgs {
    const double* At;    // [I, T, n, n]
    const double* Bt;    // [I, T, n, m]
    const double* ct;    // [I, T, n]
    const double* Q;     // [n, n]
    const double* Qd;    // [n, n]
    const double* R;     // [m, m]  (full R; halved inside, tv_lqr.py:110)
    const double* xd;    // [I or 1, T+1, n]
    long long xd_stride; // 0 when one desired trajectory is shared by all instances
    double* K;           // [I, T, m, n]
    double* k;           // [I, T, m]
    int* status;         // [I] 0 ok, 1 H not SPD / NaN
    int I, T;
};

// Solve H y = b for m x m SPD H given its Cholesky factor (registers, fully unrolled).
template <int m>
__device__ __forceinline__ bool cholesky_inplace(double (&H)[m][m]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < m; ++j) {
        double dj = H[j][j];
#pragma unroll
        for (int q = 0; q < j; ++q) dj -= H[j][q] * H[j][q];
        if (!(dj > 0.0)) { ok = false; dj = 1.0; }
        const double l = sqrt(dj);
        H[j][j] = l;
        const double il = 1.0 / l;
#pragma unroll
        for (int i = j + 1; i < m; ++i) {
            double s = H[i][j];
#pragma unroll
            for (int q = 0; q < j; ++q) s -= H[i][q] * H[j][q];
            H[i][j] = s * il;
        }
    }
    return ok;
}
template <int m>
__device__ __forceinline__ void cholesky_solve(const double (&L)[m][m], double (&b)[m]) {
#pragma unroll
    for (int i = 0; i < m; ++i) {
        double s = b[i];
#pragma unroll
        for (int q = 0; q < i; ++q) s -= L[i][q] * b[q];
        b[i] = s / L[i][i];
    }
#pragma unroll
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
#pragma unroll
        for (int q = i + 1; q < m; ++q) s -= L[q][i] * b[q];
        b[i] = s / L[i][i];
    }
}

template <int n, int m>
struct TvlqrSmem {
    double P[n * n], A[n * n], PA[n * n], Pn[n * n];
    double B[n * m], PB[n * m], G[m * n], Kt[m * n];
    double H[m * m];
    double p[n], w[n], c[n], xd[n], g[m], kt[m];
};

constexpr int kTvlqrWarps = 4;

template <int n, int m>
__global__ void __launch_bounds__(32 * kTvlqrWarps) tvlqr_riccati_kernel(const TvlqrArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x * kTvlqrWarps + warp;
    if (inst >= a.I) return;
    TvlqrSmem<n, m>& s = reinterpret_cast<TvlqrSmem<n, m>*>(smem_raw)[warp];
    const double* xd_i = a.xd + inst * a.xd_stride;
    bool ok = true;
    // terminal condition: P_T = Qd, p_T = -Qd xd_T
    for (int e = lane; e < n * n; e += 32) s.P[e] = a.Qd[e];
    for (int i = lane; i < n; i += 32) {
        double acc = 0.0;
        for (int q = 0; q < n; ++q) acc -= a.Qd[i * n + q] * xd_i[(long long)a.T * n + q];
        s.p[i] = acc;
    }
    __syncwarp();
    for (int t = a.T - 1; t >= 0; --t) {
        const long long it = (long long)inst * a.T + t;
        for (int e = lane; e < n * n; e += 32) s.A[e] = a.At[it * n * n + e];
        for (int e = lane; e < n * m; e += 32) s.B[e] = a.Bt[it * n * m + e];
        for (int e = lane; e < n; e += 32) {
            s.c[e] = a.ct[it * n + e];
            s.xd[e] = xd_i[(long long)t * n + e];
        }
        __syncwarp();
        // PA = P A, PB = P B, w = P c + p
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e % n;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.P[i * n + q] * s.A[q * n + j];
            s.PA[e] = acc;
        }
        for (int e = lane; e < n * m; e += 32) {
            const int i = e / m, j = e % m;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.P[i * n + q] * s.B[q * m + j];
            s.PB[e] = acc;
        }
        for (int i = lane; i < n; i += 32) {
            double acc = s.p[i];
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.P[i * n + q] * s.c[q];
            s.w[i] = acc;
        }
        __syncwarp();
        // H = R/2 + B'PB, G = B'PA, g = B'w
        for (int e = lane; e < m * m; e += 32) {
            const int i = e / m, j = e % m;
            double acc = 0.5 * a.R[e];
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.B[q * m + i] * s.PB[q * m + j];
            s.H[e] = acc;
        }
        for (int e = lane; e < m * n; e += 32) {
            const int i = e / n, j = e % n;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.B[q * m + i] * s.PA[q * n + j];
            s.G[e] = acc;
        }
        for (int i = lane; i < m; i += 32) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.B[q * m + i] * s.w[q];
            s.g[i] = acc;
        }
        __syncwarp();
        // every lane factors H (m <= 4: registers); lanes 0..n-1 solve a column of K, lane n solves k
        {
            double L[m][m];
#pragma unroll
            for (int i = 0; i < m; ++i)
#pragma unroll
                for (int j = 0; j < m; ++j) L[i][j] = 0.5 * (s.H[i * m + j] + s.H[j * m + i]);
            ok = cholesky_inplace<m>(L) && ok;
            for (int col = lane; col <= n; col += 32) {
                double b[m];
#pragma unroll
                for (int i = 0; i < m; ++i) b[i] = col < n ? s.G[i * n + col] : s.g[i];
                cholesky_solve<m>(L, b);
                if (col < n) {
#pragma unroll
                    for (int i = 0; i < m; ++i) {
                        s.Kt[i * n + col] = -b[i];
                        a.K[(it * m + i) * n + col] = -b[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < m; ++i) {
                        s.kt[i] = -b[i];
                        a.k[it * m + i] = -b[i];
                    }
                }
            }
        }
        __syncwarp();
        // P <- Q + A'PA + G'K,  p <- -Q xd_t + A'w + G'k
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e % n;
            double acc = a.Q[e];
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.A[q * n + i] * s.PA[q * n + j];
#pragma unroll
            for (int q = 0; q < m; ++q) acc += s.G[q * n + i] * s.Kt[q * n + j];
            s.Pn[e] = acc;
        }
        for (int i = lane; i < n; i += 32) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) acc -= a.Q[i * n + q] * s.xd[q];
#pragma unroll
            for (int q = 0; q < n; ++q) acc += s.A[q * n + i] * s.w[q];
#pragma unroll
            for (int q = 0; q < m; ++q) acc += s.G[q * n + i] * s.kt[q];
            s.p[i] = acc;
        }
        __syncwarp();
        for (int e = lane; e < n * n; e += 32) {
            const int i = e / n, j = e % n;
            s.P[e] = 0.5 * (s.Pn[i * n + j] + s.Pn[j * n + i]);
        }
        __syncwarp();
    }
    // NaN guard on the final value function
    for (int e = lane; e < n * n; e += 32)
        if (!(s.P[e] == s.P[e])) ok = false;
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) a.status[inst] = ok ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// Rollouts (thread per instance; x lives in registers).
//   closed loop: u_t = K_t x_t + k_t, x_{t+1} = f(x_t, u_t)      (irs_lqr.py:183-184)
//   open loop  : x_{t+1} = f(x_t, u_t) for given u               (irs_lqr.py:105-119)
// Both also return the cost of irs_lqr.py:121-137 (terminal term uses Q, :135-136).
// ---------------------------------------------------------------------------------------------
struct RolloutArgs {
    const double* K;      // [I, T, m, n] or nullptr (open loop)
    const double* k;      // [I, T, m]
    const double* x0;     // [I, n]
    const double* u_in;   // [I, T, m] (open loop)
    const double* xd;     // [I or 1, T+1, n]
    long long xd_stride;
    const double* Q;      // [n, n]
    const double* R;      // [m, m]
    double* x_trj;        // [I, T+1, n]
    double* u_trj;        // [I, T, m]
    double* cost;         // [I]
    int I, T;
    SysParams prm;
};

template <class Sys, bool CLOSED>
__global__ void __launch_bounds__(64) rollout_kernel(const RolloutArgs a) {
    constexpr int n = Sys::N, m = Sys::M;
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= a.I) return;
    const Sys sys(a.prm);
    const double* xd_i = a.xd + inst * a.xd_stride;
    double x[n], u[m], xn[n];
#pragma unroll
    for (int q = 0; q < n; ++q) {
        x[q] = a.x0[(long long)inst * n + q];
        a.x_trj[((long long)inst * (a.T + 1)) * n + q] = x[q];
    }
    double cost = 0.0;
    for (int t = 0; t < a.T; ++t) {
        const long long it = (long long)inst * a.T + t;
        if (CLOSED) {
#pragma unroll
            for (int i = 0; i < m; ++i) {
                double acc = a.k[it * m + i];
#pragma unroll
                for (int q = 0; q < n; ++q) acc += a.K[(it * m + i) * n + q] * x[q];
                u[i] = acc;
            }
        } else {
#pragma unroll
            for (int i = 0; i < m; ++i) u[i] = a.u_in[it * m + i];
        }
        // stage cost (x_t - xd_t)'Q(x_t - xd_t) + u_t'R u_t
        double e[n];
#pragma unroll
        for (int q = 0; q < n; ++q) e[q] = x[q] - xd_i[(long long)t * n + q];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) acc += a.Q[i * n + q] * e[q];
            cost += e[i] * acc;
        }
#pragma unroll
        for (int i = 0; i < m; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < m; ++q) acc += a.R[i * m + q] * u[q];
            cost += u[i] * acc;
        }
        sys.template step<false>(x, u, xn);
        if (CLOSED) {
#pragma unroll
            for (int i = 0; i < m; ++i) a.u_trj[it * m + i] = u[i];
        }
#pragma unroll
        for (int q = 0; q < n; ++q) {
            x[q] = xn[q];
            a.x_trj[((long long)inst * (a.T + 1) + t + 1) * n + q] = x[q];
        }
    }
    double e[n];
#pragma unroll
    for (int q = 0; q < n; ++q) e[q] = x[q] - xd_i[(long long)a.T * n + q];
#pragma unroll
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < n; ++q) acc += a.Q[i * n + q] * e[q];
        cost += e[i] * acc;
    }
    a.cost[inst] = cost;
}

// evaluate_cost for given trajectories (irs_lqr.py:121-137); one warp per instance, lanes over t.
template <int n, int m>
__global__ void __launch_bounds__(128) evaluate_cost_kernel(const double* x_trj, const double* u_trj,
                                                          const double* xd, long long xd_stride,
                                                          const double* Q, const double* R, int I, int T,
                                                          double* cost) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x * 4 + warp;
    if (inst >= I) return;
    const double* xd_i = xd + inst * xd_stride;
    double acc = 0.0;
    for (int t = lane; t <= T; t += 32) {
        double e[n];
#pragma unroll
        for (int q = 0; q < n; ++q) e[q] = x_trj[((long long)inst * (T + 1) + t) * n + q] - xd_i[(long long)t * n + q];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < n; ++q) s += Q[i * n + q] * e[q];
            acc += e[i] * s;
        }
        if (t < T) {
            double u[m];
#pragma unroll
            for (int q = 0; q < m; ++q) u[q] = u_trj[((long long)inst * T + t) * m + q];
#pragma unroll
            for (int i = 0; i < m; ++i) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < m; ++q) s += R[i * m + q] * u[q];
                acc += u[i] * s;
            }
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) cost[inst] = acc;
}

// Linear-model rollout for solve_tvlqr's return value (x*, u*) (tv_lqr.py:142-145):
// x_{t+1} = A_t x_t + B_t u_t + c_t with u_t = K_t x_t + k_t.
template <int n, int m>
__global__ void __launch_bounds__(64) linear_rollout_kernel(const TvlqrArgs a, const double* x0,
                                                           double* xs, double* us) {
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= a.I) return;
    double x[n], u[m], xn[n];
#pragma unroll
    for (int q = 0; q < n; ++q) {
        x[q] = x0[(long long)inst * n + q];
        xs[((long long)inst * (a.T + 1)) * n + q] = x[q];
    }
    for (int t = 0; t < a.T; ++t) {
        const long long it = (long long)inst * a.T + t;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            double acc = a.k[it * m + i];
#pragma unroll
            for (int q = 0; q < n; ++q) acc += a.K[(it * m + i) * n + q] * x[q];
            u[i] = acc;
            us[it * m + i] = acc;
        }
#pragma unroll
        for (int i = 0; i < n; ++i) {
            double acc = a.ct[it * n + i];
#pragma unroll
            for (int q = 0; q < n; ++q) acc += a.At[(it * n + i) * n + q] * x[q];
#pragma unroll
            for (int q = 0; q < m; ++q) acc += a.Bt[(it * n + i) * m + q] * u[q];
            xn[i] = acc;
        }
#pragma unroll
        for (int q = 0; q < n; ++q) {
            x[q] = xn[q];
            xs[((long long)inst * (a.T + 1) + t + 1) * n + q] = x[q];
        }
    }
}

}  // namespace irs
