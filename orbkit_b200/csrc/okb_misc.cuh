// okb_misc.cuh -- FP64 GEMM behind the cy_core.mocreator drop-in and the FP64 peak microbenchmarks.
#pragma once
#include "okb_common.cuh"

namespace okb {

// ---- separable exponentials of a regular grid ----------------------------------------------------------------
// out[p][i] = scale_p * exp(-alpha_p (ax[i] - centre_p)^2) for every primitive p and axis point i; prim5 holds
// (alpha, cN, X, Y, Z) per primitive, `axis` selects the centre component, the x table carries cN.
__global__ void __launch_bounds__(256) okb_axis_table_kernel(const double *__restrict__ prim5, int nprim,
                                                             const double *__restrict__ ax, int n, int axis,
                                                             double *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)nprim * n) return;
    const int pr = (int)(idx / n), i = (int)(idx - (long long)pr * n);
    const double *q = prim5 + 5 * (size_t)pr;
    const double d = ax[i] - q[2 + axis];
    const double e = exp(-(q[0] * (d * d)));
    out[idx] = axis == 0 ? q[1] * e : e;
}

// ---- FP64 peak microbenchmarks (roofline denominators measured on the box) ------------
// kind 0: DFMA issue-bound (8 independent chains per thread)
// kind 1: DMMA mma.sync.m8n8k4.f64 (8 independent accumulator tiles per warp)
// kind 2: even warps DFMA, odd warps DMMA (do the two share the FP64 datapath?)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) okb_fp64_peak_kernel(double *sink, int iters, int kind) {
    const int warp = threadIdx.x >> 5;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 1e-3 * i;
    const bool use_mma = (kind == 1) || (kind == 2 && (warp & 1));
    if (!use_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma884(acc[i], acc[i + 1], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) sink[0] = s;   // keep the chains alive
}

}  // namespace okb
