// okb_overlap.cuh -- analytic overlap matrix of contracted Cartesian Gaussians (replaces cy_overlap.aooverlap,
// orbkit/cy_overlap.pyx:75-156, and get_overlap / s of orbkit/c_non-grid-based.c:9-52).
//
//   aoom[i][j] = sum over the primitives p of function i and q of function j of
//                c_p c_q N_p N_q  E_AB (pi / (a_p + b_q))^(3/2)  prod_axis s_axis(la, lb)          (drv 0)
//   drv 1..3:    <i| d/dx_drv j> through the ket exponents: lb -> lb - 1 (factor lb) and lb + 1 (factor -2 b_q)
//
// One thread owns one matrix element and walks its primitive pairs in the reference's order (p outer, q inner), with the
// reference's factor order; s() is the same recursion (initial conditions, recurrence in a, transfer equation) evaluated
// bottom-up in a 16-entry table instead of recursively, without fused multiply-add.  The only difference to the reference is
// libm: exp / pow of the device, i.e. agreement to a few ulp per term (tests: 1e-12 of the largest element).  Tiny O(n_ao^2 n_prim^2) work that exists so that
// main_read (Molden renormalisation, check_norm) needs no reference code.
#pragma once
#include <cuda_runtime.h>

namespace okb {

struct OvParams {
    const double *geo_a, *geo_b;      // [n_atom][3]
    const int *la, *lb;               // [n_fn][3] exponents of the bra / ket functions
    const int *fn_atom, *fn_e0, *fn_ne;   // per function: atom, first entry, number of entries (primitives)
    const double *e_alpha, *e_c, *e_n;    // per entry: exponent, contraction coefficient, primitive norm
    int n_fn, drv;
    double *aoom;                     // [n_fn][n_fn]
};

// NO fused multiply-add anywhere in the recursion: the transfer equation cancels heavily for high angular momenta (g
// functions: 1e-10 absolute between fused and unfused evaluation), and the reference's C code is compiled without FMA
// contraction -- __dmul_rn / __dadd_rn keep its roundings.
__device__ __forceinline__ double ov_s(int a, int b, double RA, double RB, double alpha, double beta) {
    double t[16];
    const double p = __dadd_rn(alpha, beta);
    const double PA = -__dadd_rn(RA, -(__dadd_rn(__dmul_rn(alpha, RA), __dmul_rn(beta, RB)) / p));
    t[0] = 1.;
    t[1] = PA;
    for (int k = 2; k <= a + b; ++k)
        t[k] = __dadd_rn(__dmul_rn(PA, t[k - 1]), __dmul_rn(((k - 1) / __dmul_rn(2., p)), t[k - 2]));
    const double AB = __dadd_rn(RA, -RB);
    for (int q = 1; q <= b; ++q)
        for (int k = 0; k <= a + b - q; ++k) t[k] = __dadd_rn(t[k + 1], __dmul_rn(AB, t[k]));
    return t[a];
}

__device__ __forceinline__ double ov_prim(const double (&RA)[3], const double (&RB)[3], const int (&la)[3], const int (&lb)[3],
                                          double alpha, double beta) {
    double rr = 0.;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double d = __dadd_rn(RA[i], -RB[i]);
        rr = __dadd_rn(rr, __dmul_rn(d, d));
    }
    const double p = __dadd_rn(alpha, beta);
    const double EAB = exp(__dmul_rn(-(__dmul_rn(alpha, beta) / p), rr));
    double ov = __dmul_rn(EAB, pow((3.14159265358979323846 / p), 3. / 2.));
#pragma unroll
    for (int i = 0; i < 3; ++i) ov = __dmul_rn(ov, ov_s(la[i], lb[i], RA[i], RB[i], alpha, beta));
    return ov;
}

__global__ void __launch_bounds__(128) okb_overlap_kernel(const OvParams p) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.n_fn * p.n_fn) return;
    const int i = (int)(idx / p.n_fn), j = (int)(idx - (long long)i * p.n_fn);
    double RA[3], RB[3];
    int la[3], lb0[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        RA[r] = p.geo_a[3 * p.fn_atom[i] + r];
        RB[r] = p.geo_b[3 * p.fn_atom[j] + r];
        la[r] = p.la[3 * i + r];
        lb0[r] = p.lb[3 * j + r];
    }
    double acc = 0.;
    for (int ei = p.fn_e0[i]; ei < p.fn_e0[i] + p.fn_ne[i]; ++ei)
        for (int ej = p.fn_e0[j]; ej < p.fn_e0[j] + p.fn_ne[j]; ++ej) {
            const double alpha = p.e_alpha[ei], beta = p.e_alpha[ej];
            const double w = __dmul_rn(__dmul_rn(__dmul_rn(p.e_c[ei], p.e_c[ej]), p.e_n[ei]), p.e_n[ej]);   // the reference's factor order
            int lb[3] = {lb0[0], lb0[1], lb0[2]};
            if (p.drv <= 0) {
                acc = __dadd_rn(acc, __dmul_rn(w, ov_prim(RA, RB, la, lb, alpha, beta)));
            } else if (lb0[p.drv - 1] == 0) {
                lb[p.drv - 1] = 1;
                acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(-2 * beta, p.e_c[ei]), p.e_c[ej]), p.e_n[ei]), p.e_n[ej]),
                                                ov_prim(RA, RB, la, lb, alpha, beta)));
            } else {
                const int l0 = lb0[p.drv - 1];
                lb[p.drv - 1] = l0 - 1;
                acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn((double)l0, p.e_c[ei]), p.e_c[ej]), p.e_n[ei]), p.e_n[ej]),
                                                ov_prim(RA, RB, la, lb, alpha, beta)));
                lb[p.drv - 1] = l0 + 1;
                acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(-2 * beta, p.e_c[ei]), p.e_c[ej]), p.e_n[ei]), p.e_n[ej]),
                                                ov_prim(RA, RB, la, lb, alpha, beta)));
            }
        }
    p.aoom[idx] = acc;
}

}  // namespace okb
