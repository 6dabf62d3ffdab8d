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

template <int J>
struct ZrunSmem {
    double w[ZR_MAXP][J];                    // w_p of the J rows
    double P[KC][ZR_MAXL + 1][J];            // polynomial coefficients per output row
    long long rowoff[KC];                    // element offset of the output row (slot and row stride applied), -1: skip
    ZShell sh[ZR_MAXS];
    int nshell;
};

template <int J, int L>
__device__ __forceinline__ void zrun_rows(const ZrunSmem<J> &S, const ZShell &sh, const double Z, const double (&R0)[J],
                                          double *__restrict__ out, const long long off0, const int nz, const unsigned act) {
    for (int r = sh.row_off; r < sh.row_off + sh.nrow; ++r) {
        const long long ro = S.rowoff[r];
        if (ro < 0) continue;                                   // uniform
        double poly[J];
#pragma unroll
        for (int j = 0; j < J; ++j) poly[j] = S.P[r][L][j];
#pragma unroll
        for (int l = L - 1; l >= 0; --l)
#pragma unroll
            for (int j = 0; j < J; ++j) poly[j] = fma(poly[j], Z, S.P[r][l][j]);
        double *o = out + ro + off0;
#pragma unroll
        for (int j = 0; j < J; ++j)
            if (act & (1u << j)) __stcs(o + (long long)j * nz, poly[j] * R0[j]);
    }
}

// Grid: blockIdx.x = (block of RG row groups) * nzb + z block.  A CTA walks  chunk (outer) x its RG row groups (inner):
// the tabz slices of a chunk's primitives (~25 x blockDim.x doubles) are then re-read RG times in a row and stay in L1,
// instead of streaming the whole table (1.25 MB for the benchmark molecule) from L2 once per row group.  The uniform
// quantities of iteration i + 1 are prepared (phase U) into the other half of a two-deep shared-memory ring before
// iteration i is evaluated (phase T): one CTA barrier per iteration.
template <int J, int PF>
__global__ void __launch_bounds__(256) okb_ao_zrun_kernel(const KParams p, long long row_first, long long row_last,
                                                          int nzb, int RG) {
    __shared__ ZrunSmem<J> S2[2];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const long long rgb = blockIdx.x / nzb;                      // block of RG row groups
    const int zb = (int)(blockIdx.x - rgb * nzb);
    const int k = zb * nt + tid;
    const bool kvalid = k < p.nz;
    const int kc = kvalid ? k : p.nz - 1;
    const double zk = __ldg(p.gz + kc);
    const int sl = p.slot[p.one_code];                           // SET_VAL: one_code = 0
    const long long slot_off = (long long)sl * p.slot_stride;
    const long long ngroups = (row_last - row_first + J) / J;
    long long g0 = rgb * RG;
    const int nrg = (int)((ngroups - g0) < RG ? (ngroups - g0) : RG);
    const int niter = p.nchunk * nrg;

    // phase U of iteration `it` (chunk it / nrg, row group g0 + it % nrg) into S2[it & 1]
    auto phase_u = [&](int it) {
        ZrunSmem<J> &S = S2[it & 1];
        const int c = it / nrg;
        const long long row0 = row_first + (g0 + (it - c * nrg)) * J;
        const unsigned char *mb = p.meta + (size_t)c * p.lay.stride;
        const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
        const ShellMeta *shells = reinterpret_cast<const ShellMeta *>(mb + p.lay.off_shell);
        const FnMeta *fns = reinterpret_cast<const FnMeta *>(mb + p.lay.off_fn);
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
                S.w[po + q][j] = w;
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
            double acc[ZR_MAXL + 1];
#pragma unroll
            for (int l = 0; l <= ZR_MAXL; ++l) acc[l] = 0.0;
            const long long row = row0 + j;
            if (row <= row_last) {
                const int i = (int)(row / p.ny), jj = (int)(row - (long long)i * p.ny);
                const ShellMeta *sh = shells + rm.shell;
                const double X = __ldg(p.gx + i) - __ldg(&sh->cx), Y = __ldg(p.gy + jj) - __ldg(&sh->cy);
                for (int t = 0; t < rm.nterm; ++t) {
                    const TermMeta tm = terms[rm.term_off + t];
                    const FnMeta fm = fns[tm.k];
                    const int lx = fm.lxyz & 0xff, ly = (fm.lxyz >> 8) & 0xff, lz = (fm.lxyz >> 16) & 0xff;
                    const double a = tm.coef * (fm.f * (upow(X, lx) * upow(Y, ly)));
#pragma unroll
                    for (int l = 0; l <= ZR_MAXL; ++l)
                        if (l == lz) acc[l] += a;
                }
            }
#pragma unroll
            for (int l = 0; l <= ZR_MAXL; ++l) S.P[r][l][j] = acc[l];
            if (j == 0) S.rowoff[r] = sl < 0 ? -1 : slot_off + (long long)rm.out_row * p.ld;
        }
        if (tid == 0) S.nshell = hdr.nshell;
    };

    if (niter > 0) phase_u(0);
    __syncthreads();
    for (int it = 0; it < niter; ++it) {
        if (it + 1 < niter) phase_u(it + 1);
        // ---- phase T: one z point per thread, J rows --------------------------------------------------------------------
        const ZrunSmem<J> &S = S2[it & 1];
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
            double R0[J];
#pragma unroll
            for (int j = 0; j < J; ++j) R0[j] = 0.0;
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
                    if (u < sh.nprim) {
#pragma unroll
                        for (int j = 0; j < J; ++j) R0[j] = fma(S.w[sh.prim_off + u][j], t[u], R0[j]);
                    }
                q = sh.nprim < PF ? sh.nprim : PF;
            }
            for (; q + 4 <= sh.nprim; q += 4) {
                double t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = __ldg(tz + (size_t)(q + u) * p.nz);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int j = 0; j < J; ++j) R0[j] = fma(S.w[sh.prim_off + q + u][j], t[u], R0[j]);
            }
            for (; q < sh.nprim; ++q) {
                const double t = __ldg(tz + (size_t)q * p.nz);
#pragma unroll
                for (int j = 0; j < J; ++j) R0[j] = fma(S.w[sh.prim_off + q][j], t, R0[j]);
            }
            const double Z = zk - sh.cz;
            switch (sh.L) {                                       // uniform
                case 0: zrun_rows<J, 0>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
                case 1: zrun_rows<J, 1>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
                case 2: zrun_rows<J, 2>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
                case 3: zrun_rows<J, 3>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
                case 4: zrun_rows<J, 4>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
                case 5: zrun_rows<J, 5>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
                default: zrun_rows<J, 6>(S, sh, Z, R0, p.out, off0, p.nz, act); break;
            }
        }
        __syncthreads();
    }
}

}  // namespace okb
