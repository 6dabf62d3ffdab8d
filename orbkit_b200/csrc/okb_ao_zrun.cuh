// okb_ao_zrun.cuh -- SINK_AO on REGULAR grids: the z-run kernel (calc_ao / core.ao_creator values and single
// derivative-free requests; replaces c_lcreator + cy_core.aocreator + core.cartesian2spherical,
// c_grid-based.c:40-69, cy_core.pyx:51-78, core.py:135-176).
//
// Why (ncu, profiles/r01_ao_tile_summary.txt): the tile kernels spend ~64 thread instructions per AO value, two thirds
// of them in the 784 exponentials per point, and reach 0.39 of the HBM peak although DRAM traffic = algorithmic bytes.
// On a regular grid everything but ONE factor of an AO is constant along a run of consecutive z points at fixed (x, y):
//
//   chi(x_i, y_j, z_k) = f X^lx Y^ly Z_k^lz  sum_p [cN_p ex_p(i) ey_p(j)] ez_p(k)          X = x_i - cx, ...
//                      = ( sum_l P_l Z_k^l ) * R0(k),      R0(k) = sum_p w_p ez_p(k),  w_p = tabx[p][i] taby[p][j]
//
// with the separable-exponential tables tabx/taby/tabz of okb_axis_table_kernel ([n_prim][n_axis]; tabx carries c N).
// P_l collects, per OUTPUT row (Cartesian function, or real-spherical combination of the shell's Cartesian functions:
// the chunk's RowMeta / TermMeta records), the coefficients coef * f * X^lx * Y^ly of the terms with lz = l.  w_p and
// P_l are uniform over the run: they are computed once per (row group, chunk) by the CTA into shared memory.
//
//   CTA         = RG groups of J consecutive (x, y) rows of the grid  x  one block of blockDim.x consecutive z points
//   thread      = one z index k, all J rows:  per shell  R0[j] = sum_p w[p][j] * tabz[p][k]   (1 coalesced cached load
//                 per primitive, J FMAs), then per output row of the shell a Horner polynomial in Z_k (L FMAs), one
//                 multiplication by R0[j] and ONE coalesced 8-byte streaming store per AO value -- lanes are consecutive
//                 k, so a warp writes 256 contiguous bytes of the row; no shared-memory staging of the AO values.
//   per AO value ~7 thread instructions instead of ~64; no exponential is evaluated in this kernel.
//
// Derivatives (round 2; codes 1..6 = d/dx, d/dy, d/dz, d2/dx2, d2/dy2, d2/dz2 -- one launch per code): with
//   R1(k) = sum_p alpha_p w_p ez_p(k),  R2(k) = sum_p alpha_p^2 w_p ez_p(k)
// every derivative of x^l exp(-a x^2) is a polynomial times the same exponential,
//   d/dx:   l x^(l-1) - 2a x^(l+1)                        d2/dx2:  l(l-1) x^(l-2) - 2a (2l+1) x^l + 4a^2 x^(l+2),
// so an output row is  A(Z) R0 + B(Z) R1 [+ C(Z) R2]  with NPOLY = 2 (first) or 3 (second derivatives) polynomials in Z
// per row whose coefficients collect the x / y factors (derivatives along x, y) or shift the power of Z (along z):
// NPOLY Horner chains of degree L + NPOLY - 1 and NPOLY multiply-adds per AO value, still ONE streaming store.
// The mixed second derivatives (codes 7..9, whose reference formulas are incomplete) stay with the tile kernels.
//
// Any shell is handled (standard order or explicit lxlylz, any L <= 6, spherical rows of any term pattern): the
// exponents travel in the chunk tables.  Results agree with the exponential-per-point kernels to a few ulp (products
// re-associated; same tolerance as the axis-table path of the fused kernel, tests/test_gpu_parity.py).
#pragma once
#include "okb_common.cuh"

namespace okb {

constexpr int ZR_MAXP = 96;      // primitives per chunk (host: MAXP)
constexpr int ZR_MAXS = 32;      // shells per chunk (host: MAXS)
constexpr int ZR_MAXL = 6;       // highest angular momentum (28 functions <= KC)

struct ZShell { double cz; int gprim, prim_off, nprim, L, row_off, nrow; };

// NPOLY: 1 values, 2 first derivatives (R0, R1), 3 second derivatives (R0, R1, R2); polynomial degree <= ZR_MAXL + NPOLY - 1
template <int J, int NPOLY>
struct ZrunSmem {
    static constexpr int ND = ZR_MAXL + NPOLY;             // coefficients per polynomial
    double w[NPOLY][ZR_MAXP][J];                           // w_p, alpha_p w_p, alpha_p^2 w_p of the J rows
    double P[NPOLY][KC][ND][J];                            // polynomial coefficients per output row
    long long rowoff[KC];                                  // element offset of the output row (slot and row stride applied), -1: skip
    ZShell sh[ZR_MAXS];
    int nshell;
};

template <int J, int L, int NPOLY>
__device__ __forceinline__ void zrun_rows(const ZrunSmem<J, NPOLY> &S, const ZShell &sh, const double Z,
                                          const double (&R)[NPOLY][J], double *__restrict__ out, const long long off0,
                                          const int nz, const unsigned act) {
    constexpr int DEG = L + NPOLY - 1;
    for (int r = sh.row_off; r < sh.row_off + sh.nrow; ++r) {
        const long long ro = S.rowoff[r];
        if (ro < 0) continue;                                   // uniform
        double val[J];
#pragma unroll
        for (int n = 0; n < NPOLY; ++n) {
            double poly[J];
#pragma unroll
            for (int j = 0; j < J; ++j) poly[j] = S.P[n][r][DEG][j];
#pragma unroll
            for (int l = DEG - 1; l >= 0; --l)
#pragma unroll
                for (int j = 0; j < J; ++j) poly[j] = fma(poly[j], Z, S.P[n][r][l][j]);
#pragma unroll
            for (int j = 0; j < J; ++j) val[j] = (n == 0) ? poly[j] * R[0][j] : fma(poly[j], R[n][j], val[j]);
        }
        double *o = out + ro + off0;
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (act & (1u << j)) __stcs(o + (long long)j * nz, val[j]);
    }
}

// Grid: blockIdx.x = (block of RG row groups) * nzb + z block.  A CTA walks  chunk (outer) x its RG row groups (inner):
// the tabz slices of a chunk's primitives (~25 x blockDim.x doubles) are then re-read RG times in a row and stay in L1,
// instead of streaming the whole table (1.25 MB for the benchmark molecule) from L2 once per row group.  The uniform
// quantities of iteration i + 1 are prepared (phase U) into the other half of a two-deep shared-memory ring before
// iteration i is evaluated (phase T): one CTA barrier per iteration.
template <int J, int PF, int NPOLY = 1>
__global__ void __launch_bounds__(256) okb_ao_zrun_kernel(const KParams p, long long row_first, long long row_last,
                                                          int nzb, int RG) {
    extern __shared__ __align__(16) unsigned char zrun_smem[];
    ZrunSmem<J, NPOLY> *S2 = reinterpret_cast<ZrunSmem<J, NPOLY> *>(zrun_smem);      // two-deep ring
    constexpr int ND = ZrunSmem<J, NPOLY>::ND;
    const int code = p.one_code;                                 // 0 values, 1..6 derivative
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const long long rgb = blockIdx.x / nzb;                      // block of RG row groups
    const int zb = (int)(blockIdx.x - rgb * nzb);
    const int k = zb * nt + tid;
    const bool kvalid = k < p.nz;
    const int kc = kvalid ? k : p.nz - 1;
    const double zk = __ldg(p.gz + kc);
    const int sl = p.slot[code];
    const long long slot_off = (long long)sl * p.slot_stride;
    const long long ngroups = (row_last - row_first + J) / J;
    long long g0 = rgb * RG;
    const int nrg = (int)((ngroups - g0) < RG ? (ngroups - g0) : RG);
    const int niter = p.nchunk * nrg;

    // phase U of iteration `it` (chunk it / nrg, row group g0 + it % nrg) into S2[it & 1]
    auto phase_u = [&](int it) {
        ZrunSmem<J, NPOLY> &S = S2[it & 1];
        const int c = it / nrg;
        const long long row0 = row_first + (g0 + (it - c * nrg)) * J;
        const unsigned char *mb = p.meta + (size_t)c * p.lay.stride;
        const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
        const ShellMeta *shells = reinterpret_cast<const ShellMeta *>(mb + p.lay.off_shell);
        const FnMeta *fns = reinterpret_cast<const FnMeta *>(mb + p.lay.off_fn);
        const double2 *prims = reinterpret_cast<const double2 *>(mb + p.lay.off_prim);     // (alpha, c N)
        const RowMeta *rows = reinterpret_cast<const RowMeta *>(mb + p.lay.off_row);
        const TermMeta *terms = reinterpret_cast<const TermMeta *>(mb + p.lay.off_term);
        // (a) w[p][j]: one warp per shell, lanes over (primitive, row)
        for (int s = warp; s < hdr.nshell; s += nwarp) {
            const int np = __ldg(&shells[s].nprim), po = __ldg(&shells[s].prim_off), gp = __ldg(&shells[s].gprim);
            for (int e = lane; e < np * J; e += 32) {
                const int q = e / J, j = e - q * J;
                const long long row = row0 + j;
                double w = 0.0;
                if (row <= row_last) {
                    const int i = (int)(row / p.ny), jj = (int)(row - (long long)i * p.ny);
                    w = __ldg(p.tabx + (size_t)(gp + q) * p.nx + i) * __ldg(p.taby + (size_t)(gp + q) * p.ny + jj);
                }
                S.w[0][po + q][j] = w;
                if (NPOLY > 1) {
                    const double al = prims[po + q].x;
                    S.w[NPOLY > 1 ? 1 : 0][po + q][j] = al * w;
                    if (NPOLY > 2) S.w[NPOLY > 2 ? 2 : 0][po + q][j] = al * (al * w);
                }
            }
            if (lane == 0) {
                ZShell z;
                z.cz = __ldg(&shells[s].cz);
                z.gprim = gp; z.prim_off = po; z.nprim = np;
                z.L = __ldg(&shells[s].L);
                z.row_off = __ldg(&shells[s].row_off);
                z.nrow = __ldg(&shells[s].nrow);
                S.sh[s] = z;
            }
        }
        // (b) P[r][l][j]: one thread per (output row, grid row)
        for (int e = tid; e < hdr.nrow * J; e += nt) {
            const int r = e / J, j = e - r * J;
            const RowMeta rm = rows[r];
            // the thread owns the coefficients of (output row r, grid row j): accumulated in shared memory (run-time degree)
#pragma unroll
            for (int n = 0; n < NPOLY; ++n)
#pragma unroll
                for (int l = 0; l < ND; ++l) S.P[n][r][l][j] = 0.0;
            auto put = [&](int n, int l, double v) { S.P[n][r][l][j] += v; };
            const long long row = row0 + j;
            if (row <= row_last) {
                const int i = (int)(row / p.ny), jj = (int)(row - (long long)i * p.ny);
                const ShellMeta *sh = shells + rm.shell;
                const double X = __ldg(p.gx + i) - __ldg(&sh->cx), Y = __ldg(p.gy + jj) - __ldg(&sh->cy);
                for (int t = 0; t < rm.nterm; ++t) {
                    const TermMeta tm = terms[rm.term_off + t];
                    const FnMeta fm = fns[tm.k];
                    const int lx = fm.lxyz & 0xff, ly = (fm.lxyz >> 8) & 0xff, lz = (fm.lxyz >> 16) & 0xff;
                    const double cf = tm.coef * fm.f;
                    if (NPOLY == 1 || code == 0) {
                        put(0, lz, cf * (upow(X, lx) * upow(Y, ly)));
                    } else if (code == 3 || code == 6) {
                        // along z: the factor of x and y is constant, the power of Z shifts
                        const double a = cf * (upow(X, lx) * upow(Y, ly));
                        if (code == 3) {
                            if (lz > 0) put(0, lz - 1, a * (double)lz);
                            put(1, lz + 1, -2.0 * a);
                        } else {
                            if (lz > 1) put(0, lz - 2, a * (double)(lz * (lz - 1)));
                            put(1, lz, -(double)(4 * lz + 2) * a);
                            put(NPOLY > 2 ? 2 : 0, lz + 2, 4.0 * a);
                        }
                    } else {
                        // along x (codes 1, 4) or y (codes 2, 5): the derivative acts on one factor, Z^lz stays
                        const bool alongx = (code == 1 || code == 4);
                        const int l = alongx ? lx : ly;
                        const double U = alongx ? X : Y;
                        const double other = cf * (alongx ? upow(Y, ly) : upow(X, lx));
                        if (code <= 2) {
                            if (l > 0) put(0, lz, other * ((double)l * upow(U, l - 1)));
                            put(1, lz, other * (-2.0 * upow(U, l + 1)));
                        } else {
                            if (l > 1) put(0, lz, other * ((double)(l * (l - 1)) * upow(U, l - 2)));
                            put(1, lz, other * (-(double)(4 * l + 2) * upow(U, l)));
                            put(NPOLY > 2 ? 2 : 0, lz, other * (4.0 * upow(U, l + 2)));
                        }
                    }
                }
            }
            if (j == 0) S.rowoff[r] = sl < 0 ? -1 : slot_off + (long long)rm.out_row * p.ld;
        }
        if (tid == 0) S.nshell = hdr.nshell;
    };

    if (niter > 0) phase_u(0);
    __syncthreads();
    for (int it = 0; it < niter; ++it) {
        if (it + 1 < niter) phase_u(it + 1);
        // ---- phase T: one z point per thread, J rows --------------------------------------------------------------------
        const ZrunSmem<J, NPOLY> &S = S2[it & 1];
        const int c = it / nrg;
        const long long row0 = row_first + (g0 + (it - c * nrg)) * J;
        const long long off0 = row0 * p.nz + k - p.p0;
        unsigned act = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const long long o = off0 + (long long)j * p.nz;
            if (kvalid && row0 + j <= row_last && o >= 0 && o < (long long)p.npts) act |= 1u << j;
        }
        const int nshell = S.nshell;
        // the tabz values of the first PF primitives of shell s + 1 are fetched before the rows of shell s are written:
        // most shells have one to three primitives, and a load issued right in front of its use costs a full L2 round trip
        double tn[PF > 0 ? PF : 1];
        if (PF > 0 && nshell > 0) {
            const ZShell s0 = S.sh[0];
            const double *tz0 = p.tabz + (size_t)s0.gprim * p.nz + kc;
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (u < s0.nprim) tn[u] = __ldg(tz0 + (size_t)u * p.nz);
        }
        for (int s = 0; s < nshell; ++s) {
            const ZShell sh = S.sh[s];
            double R[NPOLY][J];
#pragma unroll
            for (int n = 0; n < NPOLY; ++n)
#pragma unroll
                for (int j = 0; j < J; ++j) R[n][j] = 0.0;
            // R[n][j] += w[n][prim][j] * t for every radial sum the code needs
            auto radial = [&](int prim, double t) {
#pragma unroll
                for (int n = 0; n < NPOLY; ++n)
#pragma unroll
                    for (int j = 0; j < J; ++j) R[n][j] = fma(S.w[n][prim][j], t, R[n][j]);
            };
            const double *tz = p.tabz + (size_t)sh.gprim * p.nz + kc;
            int q = 0;
            if (PF > 0) {
                double t[PF > 0 ? PF : 1];
#pragma unroll
                for (int u = 0; u < PF; ++u) t[u] = tn[u];
                if (s + 1 < nshell) {
                    const ZShell sn = S.sh[s + 1];
                    const double *tzn = p.tabz + (size_t)sn.gprim * p.nz + kc;
#pragma unroll
                    for (int u = 0; u < PF; ++u)
                        if (u < sn.nprim) tn[u] = __ldg(tzn + (size_t)u * p.nz);
                }
#pragma unroll
                for (int u = 0; u < PF; ++u)
                    if (u < sh.nprim) radial(sh.prim_off + u, t[u]);
                q = sh.nprim < PF ? sh.nprim : PF;
            }
            for (; q + 4 <= sh.nprim; q += 4) {
                double t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = __ldg(tz + (size_t)(q + u) * p.nz);
#pragma unroll
                for (int u = 0; u < 4; ++u) radial(sh.prim_off + q + u, t[u]);
            }
            for (; q < sh.nprim; ++q) {
                radial(sh.prim_off + q, __ldg(tz + (size_t)q * p.nz));
            }
            const double Z = zk - sh.cz;
            switch (sh.L) {                                       // uniform
                case 0: zrun_rows<J, 0, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
                case 1: zrun_rows<J, 1, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
                case 2: zrun_rows<J, 2, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
                case 3: zrun_rows<J, 3, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
                case 4: zrun_rows<J, 4, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
                case 5: zrun_rows<J, 5, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
                default: zrun_rows<J, 6, NPOLY>(S, sh, Z, R, p.out, off0, p.nz, act); break;
            }
        }
        __syncthreads();
    }
}

}  // namespace okb
