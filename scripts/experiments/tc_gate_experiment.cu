// north_star asks for evidence, not an argument: "tensor cores only if padding the tiny state blocks to a dense 8x8 bf16
// contraction actually wins, shown by ncu tensor-pipe utilisation vs the memory-bound path's achieved HBM GB/s".
// This stand-alone experiment (NOT part of libmht_b200.so) computes the gate's normalised innovation squared
//      d2[l][c] = v^T S^-1 v,   v = z[l][c] - zhat[l]          (kalman.py:25-28, tracker.py:826-829)
// for L leaves x 16 candidate measurements each (what the measurement grid hands a leaf), two ways:
//   (a) the product path's arithmetic: float64 on the CUDA cores, 4 FMAs per pair;
//   (b) the innovation block padded to a 16x8 bf16 tile times S^-1 padded to 8x8 bf16 on the tensor cores
//       (mma.sync.m16n8k8 -- tcgen05 cannot even express it: every leaf has its own 8x8 B operand and the smallest
//       tcgen05 tile is M = 64 rows of ONE B), float32 accumulate, final dot with v on the CUDA cores.
// It prints time, achieved GB/s of the same algorithmic bytes, the worst relative error of d2 and how many of the gate
// decisions d2 <= 5.99 flip.  Run under ncu for sm__pipe_tensor_cycles_active / dram throughput (profiles/tc_experiment_r2.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_gate_experiment tc_gate_experiment.cu && ./tc_gate_experiment
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int kCand = 16;

__global__ void __launch_bounds__(256) nis_fp64_kernel(int L, const double2 *__restrict__ zhat, const float4 *__restrict__ si,
                                                       const double2 *__restrict__ z, double *__restrict__ d2) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // pair index
    if (i >= (long long)L * kCand) return;
    const int l = (int)(i / kCand);
    const double2 zh = zhat[l];
    const float4 s = si[l];
    const double2 m = z[i];
    const double v0 = m.x - zh.x, v1 = m.y - zh.y;
    const double t0 = fma(v1, (double)s.z, v0 * (double)s.x);
    const double t1 = fma(v1, (double)s.w, v0 * (double)s.y);
    d2[i] = t0 * v0 + t1 * v1;
}

// one warp per leaf: A = 16 candidates x 8 (innovation padded with zeros), B = 8 x 8 (S^-1 padded)
__global__ void __launch_bounds__(256) nis_bf16_mma_kernel(int L, const double2 *__restrict__ zhat, const float4 *__restrict__ si,
                                                           const double2 *__restrict__ z, float *__restrict__ d2) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int l = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (l >= L) return;
    const double2 zh = zhat[l];
    const float4 s = si[l];
    // innovations of candidates g and g + 8 (float64 subtraction: positions are ~1e3 m, innovations ~1e1 m)
    float va0 = 0.f, va1 = 0.f, vb0 = 0.f, vb1 = 0.f;
    if (q == 0) {
        const double2 ma = z[(long long)l * kCand + g], mb = z[(long long)l * kCand + g + 8];
        va0 = (float)(ma.x - zh.x); va1 = (float)(ma.y - zh.y);
        vb0 = (float)(mb.x - zh.x); vb1 = (float)(mb.y - zh.y);
    }
    __nv_bfloat162 a01 = __floats2bfloat162_rn(va0, va1), a23 = __floats2bfloat162_rn(vb0, vb1);
    // B[k][n]: k = 2q + {0,1}, n = g ; only k, n < 2 are non-zero: S^-1[k][n]
    float b0 = 0.f, b1 = 0.f;
    if (q == 0 && g < 2) { b0 = g == 0 ? s.x : s.y; b1 = g == 0 ? s.z : s.w; }
    __nv_bfloat162 b01 = __floats2bfloat162_rn(b0, b1);
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    const unsigned ra0 = *(unsigned *)&a01, ra1 = *(unsigned *)&a23, rb0 = *(unsigned *)&b01;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(ra0), "r"(ra1), "r"(rb0));
    if (q == 0) {   // (v S^-1) . v with the unrounded innovation
        d2[(long long)l * kCand + g] = c0 * va0 + c1 * va1;
        d2[(long long)l * kCand + g + 8] = c2 * vb0 + c3 * vb1;
    }
}

int main(int argc, char **argv) {
    const int L = argc > 1 ? atoi(argv[1]) : 4 << 20;
    const long long P = (long long)L * kCand;
    std::vector<double2> zhat(L), z(P);
    std::vector<float4> si(L);
    srand(7);
    auto u = []() { return rand() / (double)RAND_MAX; };
    for (int l = 0; l < L; ++l) {
        zhat[l] = make_double2(2000.0 * u() - 1000.0, 2000.0 * u() - 1000.0);
        const double sxx = 40.0 + 60.0 * u(), syy = 40.0 + 60.0 * u(), sxy = 10.0 * (u() - 0.5);   // S ~ 63 I (CV steady state)
        const double det = sxx * syy - sxy * sxy;
        si[l] = make_float4((float)(syy / det), (float)(-sxy / det), (float)(-sxy / det), (float)(sxx / det));
        for (int c = 0; c < kCand; ++c)   // candidates inside ~1.6 gate radii
            z[(long long)l * kCand + c] = make_double2(zhat[l].x + 60.0 * (u() - 0.5), zhat[l].y + 60.0 * (u() - 0.5));
    }
    double2 *d_zhat, *d_z; float4 *d_si; double *d_a; float *d_b;
    cudaMalloc(&d_zhat, L * sizeof(double2)); cudaMalloc(&d_z, P * sizeof(double2)); cudaMalloc(&d_si, L * sizeof(float4));
    cudaMalloc(&d_a, P * sizeof(double)); cudaMalloc(&d_b, P * sizeof(float));
    cudaMemcpy(d_zhat, zhat.data(), L * sizeof(double2), cudaMemcpyHostToDevice);
    cudaMemcpy(d_z, z.data(), P * sizeof(double2), cudaMemcpyHostToDevice);
    cudaMemcpy(d_si, si.data(), L * sizeof(float4), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_a = 0, ms_b = 0;
    for (int rep = 0; rep < 4; ++rep) {   // warm-up + 3 timed (inputs 1.2 GB > L2)
        cudaEventRecord(e0);
        nis_fp64_kernel<<<(unsigned)((P + 255) / 256), 256>>>(L, d_zhat, d_si, d_z, d_a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1); if (rep) ms_a += t / 3;
        cudaEventRecord(e0);
        nis_bf16_mma_kernel<<<(unsigned)(((long long)L * 32 + 255) / 256), 256>>>(L, d_zhat, d_si, d_z, d_b);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&t, e0, e1); if (rep) ms_b += t / 3;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    std::vector<double> a(P); std::vector<float> b(P);
    cudaMemcpy(a.data(), d_a, P * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), d_b, P * sizeof(float), cudaMemcpyDeviceToHost);
    double worst = 0.0; long long flips = 0, inside = 0;
    for (long long i = 0; i < P; ++i) {
        const double rel = fabs((double)b[i] - a[i]) / fmax(a[i], 1e-9);
        if (rel > worst) worst = rel;
        inside += a[i] <= 5.99;
        flips += (a[i] <= 5.99) != ((double)b[i] <= 5.99);
    }
    const double bytes_a = (double)L * 32 + (double)P * (16 + 8), bytes_b = (double)L * 32 + (double)P * (16 + 4);
    printf("L = %d leaves x %d candidates = %lld pairs (%lld inside the gate)\n", L, kCand, P, inside);
    printf("(a) float64 CUDA cores : %.3f ms  %.0f GB/s of %.2f GB\n", ms_a, bytes_a / ms_a * 1e-6, bytes_a * 1e-9);
    printf("(b) bf16 mma.sync 16x8x8: %.3f ms  %.0f GB/s of %.2f GB   (useful flops per MMA: 64 of 2048)\n", ms_b,
           bytes_b / ms_b * 1e-6, bytes_b * 1e-9);
    printf("(b) vs (a): worst relative error of d2 %.3e (contract: 1e-5), gate decisions flipped %lld of %lld (%.4f %%)\n",
           worst, flips, P, 100.0 * flips / P);
    return 0;
}
