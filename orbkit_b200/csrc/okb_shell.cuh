// okb_shell.cuh -- fast AO generation for shells in the standard (Molden) Cartesian order.
//
// The generic gen_shell (okb_common.cuh) decodes (lx,ly,lz) at run time and rebuilds every power per
// function: ncu showed ~390 warp instructions per (shell, 32 points) item, 70% of them integer /
// branch overhead.  Almost every shell of every input uses the default order of tools.exp[L]
// (orbkit/tools.py:118-135); for those the exponents are compile-time constants and the item becomes
// straight-line code: powers by multiplication once per shell, per-axis factor tables
//     g0[l] = r^l
//     g1[l] = l r^(l-1) R0 - 2 r^(l+1) R1                           (d/dr of r^l e^{-a r^2}, summed over primitives)
//     g2[l] = r^l (4 r^2 R2 - (4l+2) R1) + l(l-1) r^(l-2) R0        (d2/dr2)
// and 3..9 multiplications per function.  Shells with any other ordering (wfn f order, explicit
// lxlylz) and the SET_ALL / SET_ONE requests keep the generic path.
//
// exp(-t) is evaluated by exp_neg(): k = round(-t*64/ln2), Cody-Waite reduction to |r| <= ln2/128,
// degree-5 Taylor polynomial, 2^((k mod 64)/64) from a 64-entry table + exponent add (10 FP64 instructions instead of the
// 15 of the earlier degree-9 / 2^(k/4)-by-selects form, kept under OKB_EXP_POLY9 for A/B builds).  Relative error
// ~3e-16; results below the normal range (t > 708, values < 3.3e-308) are flushed to 0.
#pragma once
#include <type_traits>

#include "okb_common.cuh"

namespace okb {

// 2^(j/64), j = 0..63, correctly rounded
static __device__ const double OKB_EXP2_64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

#ifndef OKB_EXP_POLY9
__device__ __forceinline__ double exp_neg(double t) {
    // e^{-t}, 0 <= t < 708:  k = round(-t 64/ln2), r = -t - k ln2/64 (Cody-Waite, |r| <= ln2/128), degree-5 Taylor
    // polynomial (truncation r^6/720 < 3.6e-17), times 2^((k & 63)/64) from a 512-byte table (L1 resident), exponent add
    const double x = -t;
    const double MAGIC = 6755399441055744.0;                 // 2^52 + 2^51: round-to-nearest-integer
    double kd = fma(x, 92.33248261689366, MAGIC);            // 64/ln2
    const int ki = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -0x1.62e42fee00000p-7, x);            // ln2/64, leading 32 bits (k < 2^17: the product is exact)
    r = fma(kd, -0x1.a39ef35793c76p-39, r);                  // ln2/64, rest
    double p = 8.3333333333333332e-03;                       // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= __ldg(&OKB_EXP2_64[ki & 63]);
    const int m = ki >> 6;                                   // arithmetic shift: floor(k/64)
    return __hiloint2double(__double2hiint(p) + (m << 20), __double2loint(p));
}
#else
__device__ __forceinline__ double exp_neg(double t) {
    // e^{-t}, 0 <= t < 708
    const double x = -t;
    const double MAGIC = 6755399441055744.0;                 // 2^52 + 2^51: round-to-nearest-integer
    double kd = fma(x, 5.7707801635558535, MAGIC);           // 4/ln2
    const int ki = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -1.7328679509228095e-01, x);          // ln2/4 high part (fdlibm ln2_hi / 4)
    r = fma(kd, -4.7705373231764692e-11, r);                 // ln2/4 low part
    double p = 2.7557319223985893e-06;                       // 1/9!
    p = fma(p, r, 2.4801587301587302e-05);
    p = fma(p, r, 1.9841269841269841e-04);
    p = fma(p, r, 1.3888888888888889e-03);
    p = fma(p, r, 8.3333333333333332e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int j = ki & 3;
    const double s = (j & 2) ? ((j & 1) ? 1.6817928305074290 : 1.4142135623730951)
                             : ((j & 1) ? 1.1892071150027210 : 1.0);
    p *= s;
    const int m = ki >> 2;                                   // arithmetic shift: floor(k/4)
    return __hiloint2double(__double2hiint(p) + (m << 20), __double2loint(p));
}
#endif

// Molden order of the Cartesian exponents, packed lx | ly<<4 | lz<<8 (tools.py:118-135)
__host__ __device__ constexpr int std_lxyz(int L, int j) {
    constexpr int T0[1] = {0x000};
    constexpr int T1[3] = {0x001, 0x010, 0x100};
    constexpr int T2[6] = {0x002, 0x020, 0x200, 0x011, 0x101, 0x110};
    constexpr int T3[10] = {0x003, 0x030, 0x300, 0x021, 0x012, 0x102, 0x201, 0x210, 0x120, 0x111};
    constexpr int T4[15] = {0x004, 0x040, 0x400, 0x013, 0x103, 0x031, 0x130, 0x301, 0x310,
                            0x022, 0x202, 0x220, 0x112, 0x121, 0x211};
    return L == 0 ? T0[j] : L == 1 ? T1[j] : L == 2 ? T2[j] : L == 3 ? T3[j] : T4[j];
}
__host__ __device__ constexpr int std_nfn(int L) { return (L + 1) * (L + 2) / 2; }

// Sparsity pattern of the Cartesian -> real-spherical rows in the canonical order m = -L..L of the
// reference table (orbkit/tools.py:155-191): sph_nterm(L, r) terms, term t combines the Cartesian
// function number sph_cart(L, r, t) of the standard order above.  Only the PATTERN is compiled in; the
// coefficient values travel in the chunk tables (the host takes them from the caller's CSR and falls back
// to folding the transform into the MO coefficients whenever a shell does not match).  The l=4 rows
// reproduce the reference's table as it is, including the duplicated yyyz term of (4,-1).
__host__ __device__ constexpr int sph_nterm(int L, int r) {
    constexpr int N2[5] = {1, 1, 3, 1, 2};
    constexpr int N3[7] = {2, 1, 3, 3, 3, 2, 2};
    constexpr int N4[9] = {2, 2, 3, 3, 6, 3, 4, 2, 3};
    return L == 2 ? N2[r] : L == 3 ? N3[r] : N4[r];
}
__host__ __device__ constexpr int sph_cart(int L, int r, int t) {
    constexpr int C2[5][3] = {{3}, {5}, {2, 0, 1}, {4}, {0, 1}};
    constexpr int C3[7][3] = {{1, 4}, {9}, {7, 1, 4}, {2, 5, 8}, {6, 0, 3}, {5, 8}, {0, 3}};
    constexpr int C4[9][6] = {{3, 5}, {6, 12}, {14, 3, 5}, {6, 6, 12}, {2, 0, 1, 10, 11, 9},
                              {7, 4, 13}, {10, 11, 0, 1}, {4, 13}, {0, 1, 9}};
    return L == 2 ? C2[r][t] : L == 3 ? C3[r][t] : C4[r][t];
}
__host__ __device__ constexpr int sph_nnz(int L) { return L == 2 ? 8 : L == 3 ? 16 : 28; }
// aux record of a kind-2 shell (doubles): f[std_nfn(L)], then per canonical row r: tile-row position
// (stored as a double), then its sph_nterm(L, r) coefficients
__host__ __device__ constexpr int sph_aux_doubles(int L) { return std_nfn(L) + (2 * L + 1) + sph_nnz(L); }
__host__ __device__ constexpr int sph_aux_row(int L, int r) {      // offset of canonical row r in the record
    int a = std_nfn(L);
    for (int q = 0; q < r; ++q) a += 1 + sph_nterm(L, q);
    return a;
}
// compile-time loop: f(integral_constant<int, I>) for I in [I0, N)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// Radial sums over the primitives of one shell for NP points of the same thread:
// R0 = sum cN e, R1 = sum cN a e, R2 = sum cN a^2 e.
// G primitives x NP points at once: the exp chains are independent, so the ~15 dependent FP64 operations
// of one exp_neg overlap with those of the others (the producer warps are latency bound: ~10 cycles per
// dependent FP64 instruction and only two or three warps per sub-partition to hide it).
template <int G, int NP, bool N1, bool N2>
__device__ __forceinline__ void radial_group(const double2 *__restrict__ pp, const double (&rr)[NP], double (&R0)[NP],
                                             double (&R1)[NP], double (&R2)[NP]) {
    double2 ac[G];
    double arg[G][NP];
    bool live = false;
#pragma unroll
    for (int u = 0; u < G; ++u) {
        ac[u] = pp[u];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            arg[u][q] = ac[u].x * rr[q];
            live |= (arg[u][q] < 708.0);
        }
    }
    if (!__any_sync(0xffffffffu, live)) return;              // every exp of the group underflows
    double t[G][NP];
#pragma unroll
    for (int u = 0; u < G; ++u)
#pragma unroll
        for (int q = 0; q < NP; ++q) t[u][q] = exp_neg(fmin(arg[u][q], 708.0));
#pragma unroll
    for (int u = 0; u < G; ++u)
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const double tu = (arg[u][q] < 708.0) ? ac[u].y * t[u][q] : 0.0;
            R0[q] += tu;
            if (N1) {
                const double ta = tu * ac[u].x;
                R1[q] += ta;
                if (N2) R2[q] = fma(ta, ac[u].x, R2[q]);
            }
        }
}

template <int NP, bool N1, bool N2>
__device__ __forceinline__ void radial_sums(const ShellMeta &sh, const double2 *__restrict__ prims,
                                            const double (&rr)[NP], double (&R0)[NP], double (&R1)[NP],
                                            double (&R2)[NP]) {
#pragma unroll
    for (int q = 0; q < NP; ++q) R0[q] = R1[q] = R2[q] = 0.0;
    const double2 *pp = prims + sh.prim_off;
    constexpr int G = (NP >= 2) ? 2 : 4;                     // keep G*NP exp chains in flight
    int i = 0;
    for (; i + G <= sh.nprim; i += G) radial_group<G, NP, N1, N2>(pp + i, rr, R0, R1, R2);
    if (G == 4 && i + 2 <= sh.nprim) {
        radial_group<2, NP, N1, N2>(pp + i, rr, R0, R1, R2);
        i += 2;
    }
    if (i < sh.nprim) radial_group<1, NP, N1, N2>(pp + i, rr, R0, R1, R2);
}

// Same sums on a regular grid from the separable axis tables (okb_axis_table_kernel): per primitive and point
// three cached loads and two multiplications replace the exponential,
//     cN exp(-a r^2) = [cN exp(-a X^2)] [exp(-a Y^2)] [exp(-a Z^2)].
// ox/oy/oz: table offsets of the shell's first primitive at each of the thread's points.
template <int G, int NP, bool N1, bool N2>
__device__ __forceinline__ void radial_group_tab(const double2 *__restrict__ pp, const AxTab &tab, int u0,
                                                 const int (&ox)[NP], const int (&oy)[NP], const int (&oz)[NP],
                                                 double (&R0)[NP], double (&R1)[NP], double (&R2)[NP]) {
    double t[G][NP];
#pragma unroll
    for (int u = 0; u < G; ++u)
#pragma unroll
        for (int q = 0; q < NP; ++q)
            t[u][q] = (__ldg(tab.ex + ox[q] + (u0 + u) * tab.nx) * __ldg(tab.ey + oy[q] + (u0 + u) * tab.ny)) *
                      __ldg(tab.ez + oz[q] + (u0 + u) * tab.nz);
#pragma unroll
    for (int u = 0; u < G; ++u) {
        const double a = N1 ? pp[u0 + u].x : 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            R0[q] += t[u][q];
            if (N1) {
                const double ta = t[u][q] * a;
                R1[q] += ta;
                if (N2) R2[q] = fma(ta, a, R2[q]);
            }
        }
    }
}

template <int NP, bool N1, bool N2>
__device__ __forceinline__ void radial_sums_tab(const ShellMeta &sh, const double2 *__restrict__ prims, const AxTab &tab,
                                                double (&R0)[NP], double (&R1)[NP], double (&R2)[NP]) {
    int ox[NP], oy[NP], oz[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        R0[q] = R1[q] = R2[q] = 0.0;
        ox[q] = sh.gprim * tab.nx + tab.ii[32 * q];
        oy[q] = sh.gprim * tab.ny + tab.jj[32 * q];
        oz[q] = sh.gprim * tab.nz + tab.kk[32 * q];
    }
    const double2 *pp = prims + sh.prim_off;
    constexpr int G = (NP >= 2) ? 2 : 4;
    int i = 0;
    for (; i + G <= sh.nprim; i += G) radial_group_tab<G, NP, N1, N2>(pp, tab, i, ox, oy, oz, R0, R1, R2);
    if (G == 4 && i + 2 <= sh.nprim) {
        radial_group_tab<2, NP, N1, N2>(pp, tab, i, ox, oy, oz, R0, R1, R2);
        i += 2;
    }
    if (i < sh.nprim) radial_group_tab<1, NP, N1, N2>(pp, tab, i, ox, oy, oz, R0, R1, R2);
}

// Hook of the generators for the remainder orbitals of okb_ws.cuh (REM > 0): ra.row(k, q, v) is called once per tile row k
// (chunk-local) and point q of the thread with the D values the generator has just written; NoRem compiles to nothing.
struct NoRem {
    static constexpr bool on = false;
    template <int D>
    __device__ __forceinline__ void row(int, int, const double (&)[D]) const {}
};

// One standard shell for NP points of the same thread (points pt, pt+32, ...: tile columns tp[32*q]).
// ONLY (SET_ONE requests): 1..6 = write just that derivative code (d/dx, d/dy, d/dz, d2/dx2, d2/dy2, d2/dz2) as the single
// set of the tile; 0 = the sets of SET.
template <int SET, int L, int STRIDE, bool SPH, int NP, int ONLY = 0, class RA = NoRem>
__device__ __forceinline__ void gen_shell_std(const ShellMeta &sh, const double2 *__restrict__ prims,
                                              const FnMeta *__restrict__ fns, const double *__restrict__ aux,
                                              const double *__restrict__ xs, const double *__restrict__ ys,
                                              const double *__restrict__ zs, double *__restrict__ tp, const AxTab &tab,
                                              const RA &ra = RA()) {
    static_assert(SET == SET_VAL || SET == SET_GRAD || SET == SET_LAP || SET == SET_D2 || SET == SET_D2P ||
                      (SET == SET_ONE && ONLY >= 1 && ONLY <= 6), "specialised sets");
    static_assert(ONLY == 0 || SET == SET_ONE, "ONLY selects the code of a SET_ONE request");
    // N1 / N2: radial sums R1 / R2 needed; W0 / W1 / W2: value / first / pure second derivative rows written; O2: tile set
    // of d2/dx2
    constexpr bool N1 = (SET != SET_VAL), N2 = (SET == SET_LAP || SET == SET_D2 || SET == SET_D2P || ONLY >= 4);
    constexpr bool W0 = (SET != SET_D2P && ONLY == 0), W1 = (SET == SET_GRAD || SET == SET_LAP);
    constexpr bool W2 = (SET == SET_LAP || SET == SET_D2 || SET == SET_D2P);
    constexpr int O2 = (SET == SET_D2) ? 1 : (SET == SET_D2P) ? 0 : 4;
    constexpr int D = set_ncodes(SET);
    double r[NP][3], rr[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        r[q][0] = xs[32 * q] - sh.cx;
        r[q][1] = ys[32 * q] - sh.cy;
        r[q][2] = zs[32 * q] - sh.cz;
        rr[q] = r[q][0] * r[q][0] + r[q][1] * r[q][1] + r[q][2] * r[q][2];
    }
    double R0[NP], R1[NP], R2[NP];
    if (tab.ex != nullptr) radial_sums_tab<NP, N1, N2>(sh, prims, tab, R0, R1, R2);      // warp-uniform
    else radial_sums<NP, N1, N2>(sh, prims, rr, R0, R1, R2);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        // per-axis factor tables (all indices are compile-time constants after unrolling)
        double g0[3][L + 2], g1[3][L + 1], g2[3][L + 1];
        const double m2R1 = -2.0 * R1[q];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            g0[a][0] = 1.0;
#pragma unroll
            for (int l = 1; l <= L + 1; ++l) g0[a][l] = g0[a][l - 1] * r[q][a];
            if (W1 || ONLY == a + 1) {
#pragma unroll
                for (int l = 0; l <= L; ++l)
                    g1[a][l] = (l == 0) ? g0[a][1] * m2R1
                                        : fma(g0[a][l + 1], m2R1, g0[a][l - 1] * ((double)l * R0[q]));
            }
            if (W2 || ONLY == a + 4) {
                const double r2R2 = 4.0 * (r[q][a] * r[q][a]) * R2[q];
#pragma unroll
                for (int l = 0; l <= L; ++l) {
                    const double t = g0[a][l] * fma((double)(-(4 * l + 2)), R1[q], r2R2);
                    g2[a][l] = (l < 2) ? t : fma(g0[a][l - 2], (double)(l * (l - 1)) * R0[q], t);
                }
            }
        }
        using TL = TileLay<STRIDE>;
        double *o = tp + TL::PQ * q + TL::off(0, sh.fn_off);
        if constexpr (!SPH) {
            const FnMeta *ff = fns + sh.fn_off;
#pragma unroll
            for (int j = 0; j < std_nfn(L); ++j) {
                const int e = std_lxyz(L, j);
                const int lx = e & 15, ly = (e >> 4) & 15, lz = (e >> 8) & 15;
                const double f = ff[j].f;
                const double fz = f * g0[2][lz];
                const double fyz = fz * g0[1][ly];
                double w[D];                               // the row's values, tile set by tile set
                if (W0) w[0] = fyz * (R0[q] * g0[0][lx]);
                if (N1) {
                    const double fxz = fz * g0[0][lx];
                    const double fxy = f * (g0[0][lx] * g0[1][ly]);
                    if (W1) {
                        w[1] = fyz * g1[0][lx];
                        w[2] = fxz * g1[1][ly];
                        w[3] = fxy * g1[2][lz];
                    }
                    if (W2) {
                        w[O2 + 0] = fyz * g2[0][lx];
                        w[O2 + 1] = fxz * g2[1][ly];
                        w[O2 + 2] = fxy * g2[2][lz];
                    }
                    if (ONLY == 1) w[0] = fyz * g1[0][lx];
                    if (ONLY == 2) w[0] = fxz * g1[1][ly];
                    if (ONLY == 3) w[0] = fxy * g1[2][lz];
                    if (ONLY == 4) w[0] = fyz * g2[0][lx];
                    if (ONLY == 5) w[0] = fxz * g2[1][ly];
                    if (ONLY == 6) w[0] = fxy * g2[2][lz];
                }
                TL::template store<D>(o, j, w);
                if (RA::on) ra.row(sh.fn_off + j, q, w);
            }
        } else {
            // Cartesian values stay in registers; the 2L+1 real-spherical rows are the only ones written
            // (core.cartesian2spherical, core.py:135-176, applied per point instead of being folded into the
            // coefficients: the contraction then runs over n_sph instead of n_cart functions).
            const double *ax = aux + sh.aux_off;
            double v[std_nfn(L)][D];
#pragma unroll
            for (int j = 0; j < std_nfn(L); ++j) {
                const int e = std_lxyz(L, j);
                const int lx = e & 15, ly = (e >> 4) & 15, lz = (e >> 8) & 15;
                const double f = ax[j];
                const double fz = f * g0[2][lz];
                const double fyz = fz * g0[1][ly];
                if (W0) v[j][0] = fyz * (R0[q] * g0[0][lx]);
                if (N1) {
                    const double fxz = fz * g0[0][lx];
                    const double fxy = f * (g0[0][lx] * g0[1][ly]);
                    if (W1) {
                        v[j][1] = fyz * g1[0][lx];
                        v[j][2] = fxz * g1[1][ly];
                        v[j][3] = fxy * g1[2][lz];
                    }
                    if (W2) {
                        v[j][O2 + 0] = fyz * g2[0][lx];
                        v[j][O2 + 1] = fxz * g2[1][ly];
                        v[j][O2 + 2] = fxy * g2[2][lz];
                    }
                    if (ONLY == 1) v[j][0] = fyz * g1[0][lx];
                    if (ONLY == 2) v[j][0] = fxz * g1[1][ly];
                    if (ONLY == 3) v[j][0] = fxy * g1[2][lz];
                    if (ONLY == 4) v[j][0] = fyz * g2[0][lx];
                    if (ONLY == 5) v[j][0] = fxz * g2[1][ly];
                    if (ONLY == 6) v[j][0] = fxy * g2[2][lz];
                }
            }
            static_for<0, 2 * L + 1>([&](auto rc) {
                constexpr int rw = decltype(rc)::value;
                constexpr int a = sph_aux_row(L, rw);
                const int pos = (int)ax[a];
                double sacc[D];
                static_for<0, sph_nterm(L, rw)>([&](auto tc) {
                    constexpr int t = decltype(tc)::value;
                    constexpr int cj = sph_cart(L, rw, t);
                    const double c = ax[a + 1 + t];
#pragma unroll
                    for (int d = 0; d < D; ++d) sacc[d] = (t == 0) ? c * v[cj][d] : fma(c, v[cj][d], sacc[d]);
                });
                TL::template store<D>(o, pos, sacc);
                if (RA::on) ra.row(sh.fn_off + pos, q, sacc);
            });
        }
    }
}

// standard shells of L <= 4 (kind 1: Cartesian rows, kind 2: spherical rows): the straight-line code; returns false for
// every other shell
template <int S, int STRIDE, int NP, int ONLY, class RA = NoRem>
__device__ __forceinline__ bool gen_shell_try_std(const ShellMeta &sh, const double2 *__restrict__ prims,
                                                  const FnMeta *__restrict__ fns, const double *__restrict__ aux,
                                                  const double *__restrict__ xs, const double *__restrict__ ys,
                                                  const double *__restrict__ zs, double *__restrict__ tp, const AxTab &tab,
                                                  const RA &ra = RA()) {
    if (sh.kind == 1) {                  // warp-uniform
        switch (sh.L) {
            case 0: gen_shell_std<S, 0, STRIDE, false, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            case 1: gen_shell_std<S, 1, STRIDE, false, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            case 2: gen_shell_std<S, 2, STRIDE, false, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            case 3: gen_shell_std<S, 3, STRIDE, false, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            case 4: gen_shell_std<S, 4, STRIDE, false, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            default: return false;
        }
    } else if (sh.kind == 2) {           // spherical output rows (host guarantees 2 <= L <= 4)
        switch (sh.L) {
            case 2: gen_shell_std<S, 2, STRIDE, true, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            case 3: gen_shell_std<S, 3, STRIDE, true, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
            default: gen_shell_std<S, 4, STRIDE, true, NP, ONLY, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra); return true;
        }
    }
    return false;
}

// dispatcher: standard shells of L <= 4 take the specialised code, everything else the generic one.
// xs/ys/zs point at the coordinates of the thread's first point; its NP points are 32 apart.
template <int SET, int STRIDE, int NP, class RA = NoRem>
__device__ __forceinline__ void gen_shell_any(const ShellMeta &sh, const double2 *__restrict__ prims,
                                              const FnMeta *__restrict__ fns, const double *__restrict__ aux,
                                              const double *__restrict__ xs, const double *__restrict__ ys,
                                              const double *__restrict__ zs, double *__restrict__ tp,
                                              int one_code, int exact, const AxTab &tab, const RA &ra = RA()) {
    if constexpr (SET == SET_VAL || SET == SET_GRAD || SET == SET_LAP || SET == SET_D2 || SET == SET_D2P) {
        if (gen_shell_try_std<SET, STRIDE, NP, 0, RA>(sh, prims, fns, aux, xs, ys, zs, tp, tab, ra)) return;
    }
    if constexpr (SET == SET_ONE) {
        // single first or pure second derivatives (ao_creator / mo_creator with drv = x .. zz, cy_core.aocreator with
        // drv = 1..6): the straight-line code of that one code; the mixed derivatives 7..9 keep the generic path
        // (they reproduce the reference's incomplete formulas, c_support.c:121-168)
        bool done = false;
        switch (one_code) {              // warp-uniform
            case 1: done = gen_shell_try_std<SET_ONE, STRIDE, NP, 1>(sh, prims, fns, aux, xs, ys, zs, tp, tab); break;
            case 2: done = gen_shell_try_std<SET_ONE, STRIDE, NP, 2>(sh, prims, fns, aux, xs, ys, zs, tp, tab); break;
            case 3: done = gen_shell_try_std<SET_ONE, STRIDE, NP, 3>(sh, prims, fns, aux, xs, ys, zs, tp, tab); break;
            case 4: done = gen_shell_try_std<SET_ONE, STRIDE, NP, 4>(sh, prims, fns, aux, xs, ys, zs, tp, tab); break;
            case 5: done = gen_shell_try_std<SET_ONE, STRIDE, NP, 5>(sh, prims, fns, aux, xs, ys, zs, tp, tab); break;
            case 6: done = gen_shell_try_std<SET_ONE, STRIDE, NP, 6>(sh, prims, fns, aux, xs, ys, zs, tp, tab); break;
            default: break;
        }
        if (done) return;
    }
#pragma unroll 1
    for (int q = 0; q < NP; ++q)
        gen_shell<SET, STRIDE>(sh, prims, fns, xs[32 * q], ys[32 * q], zs[32 * q], tp + TileLay<STRIDE>::PQ * q, one_code,
                               exact);
    if constexpr (RA::on) {
        // generic shells: read the rows back (the thread's own columns of the tile, no synchronisation needed)
        constexpr int D = set_ncodes(SET);
#pragma unroll 1
        for (int k = sh.fn_off; k < sh.fn_off + sh.nfn; ++k)
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                double w[D];
#pragma unroll
                for (int d = 0; d < D; ++d) w[d] = tp[TileLay<STRIDE>::off(d, k) + TileLay<STRIDE>::PQ * q];
                ra.row(k, q, w);
            }
    }
}

}  // namespace okb
