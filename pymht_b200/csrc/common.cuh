// Shared helpers for libmht_b200 (sm_100a only).  Compiled with --fmad=false: every fused
// multiply-add below is explicit, so the float32 covariance chain reproduces the reference's
// ascending-k FMA accumulation (NumPy/OpenBLAS sgemm) and nothing else gets contracted.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/mht_b200.h"

namespace mht {

void set_error(const char *fmt, ...);
extern long long g_launches;   // kernels launched by this library (mht_launch_count)
static inline void count_launch() { ++g_launches; }
int check_device();  // MHT_OK or MHT_E_NODEVICE

#define MHT_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            mht::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                          \
            return MHT_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

constexpr int kSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kTile = 256;         // hypotheses per CTA tile
constexpr int kGridMaxCells = 1 << 18;

// Uniform measurement grid of one scan (built on device by grid_build_kernel).
struct GridDesc {
    double x0, y0, inv_cell;
    int nx, ny, n_meas, pad;
};

// ------------------------------------------------------------------------------------------------
// Per-leaf Kalman quantities (reference pymht/utils/kalman.py:55-64 predict, :82-101 precalc).
// ------------------------------------------------------------------------------------------------
struct LeafKF {
    double xbar[4];
    double zhat[2];
    double si[4];      // S^-1 (float32 values, upcast like np.matmul(f64, f32) does)
    double hx, hy;     // half extents of the gate ellipse's bounding box
    double logterm;    // ln(lambda_ex sqrt(det(2 pi S)) / P_d), float32 chain like kalman.py:19
    float Pbar[16];
    float Phat[16];
    float K[8];
};

// c[i][j] = sum_k a[i][k] b[k][j] as an ascending-k FMA chain starting from +0 (sgemm order).
template <int M, int K, int N, bool TRANSB>
__device__ __forceinline__ void mm_f32(const float *a, const float *b, float *c) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < K; ++k) acc = fmaf(a[i * K + k], TRANSB ? b[j * K + k] : b[k * N + j], acc);
            c[i * N + j] = acc;
        }
}

// State part (float64): x_bar = A x, z_hat = C x_bar  (kalman.py:55-59, :88)
__device__ __forceinline__ void state_kf(const mht_model &m, const double x0[4], LeafKF &o) {
    // x_bar = A x  (A float32 upcast; dgemm-order FMA chain)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma((double)m.A[i * 4 + k], x0[k], acc);
        o.xbar[i] = acc;
    }
    // z_hat = C x_bar
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma((double)m.C[i * 4 + k], o.xbar[k], acc);
        o.zhat[i] = acc;
    }
}

// Covariance part (float32 chain): P_bar, S, S^-1, gate box, NLLR log term, K, P_hat.  Depends on
// (P0, P_d) only -- never on the state -- which is what lets the forest evaluate it once per
// hit/miss pattern of a tree instead of once per leaf (forest.cu, pat_table_kernel).
template <bool WANT_UPDATE>
__device__ __forceinline__ void cov_kf(const mht_model &m, const float P0[16], double Pd, LeafKF &o) {
    // P_bar = (A P) A^T + Q
    float AP[16];
    mm_f32<4, 4, 4, false>(m.A, P0, AP);
    mm_f32<4, 4, 4, true>(AP, m.A, o.Pbar);
#pragma unroll
    for (int i = 0; i < 16; ++i) o.Pbar[i] = o.Pbar[i] + m.Q[i];
    // S = (C P_bar) C^T + R
    float CP[8], S[4];
    mm_f32<2, 4, 4, false>(m.C, o.Pbar, CP);
    mm_f32<2, 4, 2, true>(CP, m.C, S);
#pragma unroll
    for (int i = 0; i < 4; ++i) S[i] = S[i] + m.R[i];
    // S^-1 by LU without pivoting (exact 1/s for the diagonal S of the CV model)
    const float a = S[0], b = S[1], c = S[2], d = S[3];
    const float l = c / a;
    const float u22 = d - l * b;
    float Si[4];
    Si[3] = 1.0f / u22;
    Si[2] = (-l) / u22;
    Si[0] = (1.0f - b * Si[2]) / a;
    Si[1] = (-(b * Si[3])) / a;
#pragma unroll
    for (int i = 0; i < 4; ++i) o.si[i] = (double)Si[i];
    // bounding box of { v : v^T Si v <= eta2 }
    {
        const double qa = o.si[0], qb = 0.5 * (o.si[1] + o.si[2]), qc = o.si[3];
        const double det = qa * qc - qb * qb;
        o.hx = sqrt(m.eta2 * qc / det) * (1.0 + 1e-9) + 1e-9;
        o.hy = sqrt(m.eta2 * qa / det) * (1.0 + 1e-9) + 1e-9;
    }
    // ln(lambda_ex * sqrt(det(2 pi S)) / P_d): every step float32 (NumPy weak-scalar promotion)
    {
        const float two_pi = (float)6.283185307179586;
        const float sa = two_pi * S[0], sb = two_pi * S[1], sc = two_pi * S[2], sd = two_pi * S[3];
        const float ll = sc / sa;
        const float det = sa * (sd - ll * sb);
        const float v = ((float)m.lambda_ex * sqrtf(det)) / (float)Pd;
        o.logterm = (double)logf(v);
    }
    if (WANT_UPDATE) {
        // K = (P_bar C^T) S^-1 ; P_hat = P_bar - (K C) P_bar
        float PCt[8], KC[16], KCP[16];
        mm_f32<4, 4, 2, true>(o.Pbar, m.C, PCt);
        mm_f32<4, 2, 2, false>(PCt, Si, o.K);
        mm_f32<4, 2, 4, false>(o.K, m.C, KC);
        mm_f32<4, 4, 4, false>(KC, o.Pbar, KCP);
#pragma unroll
        for (int i = 0; i < 16; ++i) o.Phat[i] = o.Pbar[i] - KCP[i];
    }
}

template <bool WANT_UPDATE>
__device__ __forceinline__ void leaf_kf(const mht_model &m, const double x0[4], const float P0[16], double Pd,
                                        LeafKF &o) {
    state_kf(m, x0, o);
    cov_kf<WANT_UPDATE>(m, P0, Pd, o);
}

// d2 = sum((z~ @ S^-1) * z~)  (kalman.py:25-28): two FMA-chain products, then mul, mul, add.
__device__ __forceinline__ double nis_f64(const double si[4], double v0, double v1) {
    const double t0 = fma(v1, si[2], v0 * si[0]);
    const double t1 = fma(v1, si[3], v0 * si[1]);
    return t0 * v0 + t1 * v1;
}

// x_hat = x_bar + K z~   (kalman.py:43-52)
__device__ __forceinline__ void filter_f64(const LeafKF &kf, double v0, double v1, double xh[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) xh[i] = kf.xbar[i] + fma((double)kf.K[i * 2 + 1], v1, (double)kf.K[i * 2] * v0);
}

// Grid cells under the bounding box of the leaf's gate ellipse; false = nothing to visit.
__device__ __forceinline__ bool gate_box(const GridDesc &g, const LeafKF &kf, int &cx0, int &cx1, int &cy0, int &cy1) {
    if (g.n_meas == 0) return false;
    cx0 = (int)floor((kf.zhat[0] - kf.hx - g.x0) * g.inv_cell);
    cx1 = (int)floor((kf.zhat[0] + kf.hx - g.x0) * g.inv_cell);
    cy0 = (int)floor((kf.zhat[1] - kf.hy - g.y0) * g.inv_cell);
    cy1 = (int)floor((kf.zhat[1] + kf.hy - g.y0) * g.inv_cell);
    cx0 = max(cx0, 0);
    cy0 = max(cy0, 0);
    cx1 = min(cx1, g.nx - 1);
    cy1 = min(cy1, g.ny - 1);
    return cx0 <= cx1 && cy0 <= cy1;
}

// Visit every measurement inside the leaf's gate among the cells [cx0,cx1] x [cy0,cy1].
// f(original measurement index, d2, v0, v1); the index travels with the point (loaded before the test).
template <class F>
__device__ __forceinline__ void for_each_gated_box(const GridDesc &g, const int *__restrict__ cell_start,
                                                   const double2 *__restrict__ gz, const int *__restrict__ gidx,
                                                   const LeafKF &kf, double eta2, int cx0, int cx1, int cy0, int cy1,
                                                   F f) {
    for (int cy = cy0; cy <= cy1; ++cy) {
        const int beg = cell_start[cy * g.nx + cx0];
        const int end = cell_start[cy * g.nx + cx1 + 1];  // cells of one row are contiguous
        for (int p = beg; p < end; ++p) {
            const double2 z = gz[p];
            const int m = gidx[p];
            const double v0 = z.x - kf.zhat[0], v1 = z.y - kf.zhat[1];
            const double d2 = nis_f64(kf.si, v0, v1);
            if (d2 <= eta2) f(m, d2, v0, v1);
        }
    }
}

template <class F>
__device__ __forceinline__ void for_each_gated(const GridDesc &g, const int *__restrict__ cell_start,
                                               const double2 *__restrict__ gz, const int *__restrict__ gidx,
                                               const LeafKF &kf, double eta2, F f) {
    int cx0, cx1, cy0, cy1;
    if (!gate_box(g, kf, cx0, cx1, cy0, cy1)) return;
    for_each_gated_box(g, cell_start, gz, gidx, kf, eta2, cx0, cx1, cy0, cy1, f);
}

// ordered-uint64 encoding of a double for atomicMin/Max
__device__ __forceinline__ unsigned long long f64_key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// launchers implemented in gate.cu (shared by the stateless operator and the forest)
int launch_grid_build(const double *d_z, int M, GridDesc *d_grid, int *d_cell_start, int *d_cell_fill,
                      double2 *d_gz, int *d_gidx, cudaStream_t s);
int64_t grid_workspace_bytes(int64_t M);

}  // namespace mht
