// okb_common.cuh -- sm_100a kernels of the ORBKIT grid path (AO -> MO -> rho and derivatives).
//
// One persistent kernel template does all three jobs; what differs is the SINK:
//   SINK_AO   evaluate the AO tile of a chunk of shells in shared memory, apply the (optional)
//             Cartesian -> spherical rows, store [n_drv][n_ao][pts]        (replaces c_lcreator +
//             cy_core.aocreator + core.cartesian2spherical: c_grid-based.c:9-79,
//             cy_core.pyx:51-78, core.py:135-176).  HBM-store bound.
//   SINK_MO   same AO tile, contracted against the MO coefficient tile that a bulk-async (TMA)
//             copy lands in shared memory; accumulators live in registers; stores
//             [n_drv][n_mo][pts]                                             (cy_core.pyx:82-101)
//   SINK_RHO  same contraction; epilogue squares / cross-multiplies, weights by occupation and
//             reduces over MOs: rho, delta_rho, mo_norm -- MO values never reach HBM
//                                                                            (core.py:265-304)
//
// Work decomposition (FP64 DFMA tile GEMM, no tensor cores: tcgen05 has no FP64 kind):
//   CTA tile  = P = 32*PT grid points x MC = NW*MW molecular orbitals, K looped in chunks of
//               <= KC Cartesian functions (whole shells).
//   phase A   every warp evaluates (shell, 32 points) items of chunk c+1 into tile[(c+1)&1]
//             (one thread = one point: exp once per primitive, powers by multiplication,
//             all D derivative sets from the same radial sums R0,R1,R2).
//   phase B   warp w owns MOs [w*MW, (w+1)*MW); lane owns points lane+32*j.  Per k: D*PT
//             conflict-free LDS.64 of AO values + MW/2 broadcast LDS.128 of coefficients feed
//             MW*PT*D DFMAs.
//   One __syncthreads per chunk; AO tiles double-buffered; coefficient tiles (2 buffers) and
//   chunk tables (3 buffers) arrive by cp.async.bulk + mbarrier issued by thread 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace okb {

constexpr int KC = 32;            // max Cartesian functions per chunk
constexpr int NMETA = 3;          // chunk-table buffers in flight
enum { SET_VAL = 0, SET_GRAD = 1, SET_LAP = 2, SET_ALL = 3, SET_ONE = 4,
       SET_D2 = 5,     // SET_D2: value + the three pure second derivatives (codes 0,4,5,6): second pass of rho + laplacian
       SET_D2P = 6 };  // SET_D2P: the three pure second derivatives only (codes 4,5,6): second pass when the first one
                       // left the MO values in HBM (KParams::phi)
enum { SINK_AO = 0, SINK_MO = 1, SINK_RHO = 2 };

__host__ __device__ constexpr int set_ncodes(int set) {
    return set == SET_VAL ? 1 : (set == SET_GRAD || set == SET_D2) ? 4 : set == SET_D2P ? 3 : set == SET_LAP ? 7
           : set == SET_ALL ? 10 : 1;
}

// ---- chunk tables (one fixed-stride blob per chunk, 16-byte aligned sections) ---------------
struct ChunkHdr { int nshell, nprim, nfn, nrow; };
struct ShellMeta {                 // 64 B
    double cx, cy, cz;
    int prim_off, nprim, fn_off, nfn;   // offsets are chunk-local; nfn = rows this shell writes
    int L, kind;                        // kind 1: Cartesian functions in the standard order of std_lxyz(L, .);
                                        // kind 2: same, but the shell writes its 2L+1 real-spherical rows
    int aux_off;                        // kind 2: offset (doubles) of [f[ncart] | per row: position, coefs] in aux
    int gprim;                          // index of the shell's first primitive in the basis-wide axis tables
    int row_off, nrow;                  // the chunk's output rows (RowMeta) built from this shell: [row_off, row_off + nrow)
};
struct FnMeta { int lxyz; int pad; double f; };          // lx | ly<<8 | lz<<16 ; f = angular norm * renorm
struct RowMeta { int out_row, term_off, nterm, shell; }; // SINK_AO output rows of this chunk, sorted by `shell` (chunk-local)
struct TermMeta { int k; int pad; double coef; };

struct BlobLayout { int off_shell, off_prim, off_fn, off_row, off_term, off_aux, stride; };

struct KParams {
    // grid
    int grid_kind;                 // 0 regular (axes), 1 vector (coordinates), 2 spherical / 3 cylindrical product grid
    const double *gx, *gy, *gz;    // kinds 0,1: axes / coordinates.  kinds 2,3: gx = r[nx]; gy = two tables [2][ny]; gz = [2][nz]
                                   //   spherical   gy = sin(theta), cos(theta); gz = cos(phi), sin(phi)
                                   //   cylindrical gy = cos(phi),   sin(phi);   gz = zed, (unused)
    int nx, ny, nz;
    int has_aff;                   // product grids: (x,y,z) <- A (x,y,z) + t  (grid_sym_op / grid_translate)
    double aff[12];                // A row-major, then t
    // regular grids: separable exponentials  c N exp(-a (x_i-X)^2), exp(-a (y_j-Y)^2), exp(-a (z_k-Z)^2)
    // per (primitive, axis point), [n_prim][n_axis] each; null -> evaluate exp(-a r^2) per point
    const double *tabx, *taby, *tabz;
    long long p0;                  // global index of the first point of this launch
    int npts;                      // points in this launch
    int ntiles;
    // basis
    const unsigned char *meta;
    BlobLayout lay;
    int nchunk;
    // MO
    const double *cblob;           // [n_mtile][nchunk][KC][MC]
    const double *occ;             // [n_mtile*MC], zero padded
    int n_mtile, n_mo;
    // outputs
    double *out;                   // AO/MO: out[slot][row][ld]
    double *rho, *delta;           // RHO: rho[ld-indexed], delta[slot][ld]
    double *mo_norm;               // RHO: [n_mo] or null (device, atomically accumulated)
    long long ld;                  // leading dimension (points) of an output row
    long long slot_stride;         // n_rows * ld
    int slot[10];                  // code -> output slot, -1 = not requested
    int one_code, exact_mixed;
    int epi;                       // SINK_RHO epilogue: 0 standard; 1 (SET_GRAD) rho and sum 2 occ (d phi)^2 into the slots of
                                   // codes 4..6; 2 (SET_D2 / SET_D2P) ADD sum 2 occ phi d2 phi to those slots (two-pass rho +
                                   // laplacian); 3 = 1 + the MO values are left in `phi` for a SET_D2P second pass
    double *phi;                   // [n_mtile*MC][ldp] MO values of the launch's points (epi 3 writes, SET_D2P reads)
    long long ldp;
    const double *crem;            // [n_mtile][nchunk][KC][REM] coefficients of the remainder orbitals of each MO tile
                                   // (okb_ws.cuh, REM > 0: contracted by the producer warps), else null
};

// Axis tables of a regular grid as the AO generators see them; ii/jj/kk point at the axis indices of the
// thread's first point in shared memory (its other points are 32 entries apart).
struct AxTab {
    const double *ex, *ey, *ez;    // null ex -> no tables (vector grid)
    int nx, ny, nz;
    const int *ii, *jj, *kk;
};

// Coordinates (and, for regular grids, axis indices) of point number n of the grid.  The index runs first axis
// slowest, last axis fastest (cy_grid.pyx:22-29, 67-75, 88-96).  Product grids reproduce the reference's expressions
// r*sin(theta)*cos(phi), ... (cy_grid.pyx:70-72, 91-93) on host-computed sin/cos tables, multiplication order included.
__device__ __forceinline__ void grid_point(const KParams &p, long long n, double &x, double &y, double &z, int &i, int &j,
                                           int &k) {
    if (p.grid_kind == 1) {
        x = p.gx[n]; y = p.gy[n]; z = p.gz[n];
        i = j = k = 0;
        return;
    }
    const long long nyz = (long long)p.ny * p.nz;
    const long long ii = n / nyz, rem = n - ii * nyz;
    i = (int)ii;
    j = (int)(rem / p.nz);
    k = (int)(rem - (long long)j * p.nz);
    if (p.grid_kind == 0) {
        x = p.gx[i]; y = p.gy[j]; z = p.gz[k];
        return;
    }
    const double r = p.gx[i];
    if (p.grid_kind == 2) {              // x = r sin(theta) cos(phi), y = r sin(theta) sin(phi), z = r cos(theta)
        const double rs = __dmul_rn(r, p.gy[j]);
        x = __dmul_rn(rs, p.gz[k]);
        y = __dmul_rn(rs, p.gz[p.nz + k]);
        z = __dmul_rn(r, p.gy[p.ny + j]);
    } else {                             // x = r cos(phi), y = r sin(phi), z = zed
        x = __dmul_rn(r, p.gy[j]);
        y = __dmul_rn(r, p.gy[p.ny + j]);
        z = p.gz[k];
    }
    if (p.has_aff) {
        const double X = x, Y = y, Z = z;
        x = fma(p.aff[2], Z, fma(p.aff[1], Y, p.aff[0] * X)) + p.aff[9];
        y = fma(p.aff[5], Z, fma(p.aff[4], Y, p.aff[3] * X)) + p.aff[10];
        z = fma(p.aff[8], Z, fma(p.aff[7], Y, p.aff[6] * X)) + p.aff[11];
    }
}

// ---- PTX helpers: mbarrier + bulk async copy (TMA, non-tensor form) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OKB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OKB_DONE_%=;\n"
        "bra OKB_WAIT_%=;\n"
        "OKB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- angular part -----------------------------------------------------------------------------
// r^l by square-and-multiply; l is warp-uniform (all lanes evaluate the same function).
__device__ __forceinline__ double upow(double r, int l) {
    double acc = 1.0;
    while (l > 0) {
        if (l & 1) acc *= r;
        r *= r;
        l >>= 1;
    }
    return acc;
}
// Per axis:  q0 = r^l, qm1 = l r^(l-1), qp1 = r^(l+1), qm2 = l(l-1) r^(l-2)
struct AxisQ { double q0, qm1, qp1, qm2; };
template <int LEVEL>
__device__ __forceinline__ AxisQ axis_q(double r, int l) {
    AxisQ a;
    if (LEVEL == 0) {
        a.q0 = upow(r, l);
        a.qm1 = a.qp1 = a.qm2 = 0.0;
        return a;
    }
    double pm2 = 0.0, pm1 = 0.0, p0 = 1.0;
    if (l >= 2) {
        pm2 = upow(r, l - 2);
        pm1 = pm2 * r;
        p0 = pm1 * r;
    } else if (l == 1) {
        pm1 = 1.0;
        p0 = r;
    }
    a.q0 = p0;
    a.qm1 = (double)l * pm1;
    a.qp1 = p0 * r;
    a.qm2 = (LEVEL >= 2) ? (double)(l * (l - 1)) * pm2 : 0.0;
    return a;
}

// Mixed second derivative d_a d_b of one Cartesian Gaussian shell function (without the common
// angular norm f and the third axis factor).  exact=0 reproduces c_support.c:121-168, which drops
// the -2*alpha cross terms; exact=1 is the analytic form.
__device__ __forceinline__ double mixed2(const AxisQ &a, const AxisQ &b, int la, int lb, double R0,
                                         double R1, double R2, int exact) {
    if (exact)
        return a.qm1 * b.qm1 * R0 - 2.0 * (a.qm1 * b.qp1 + a.qp1 * b.qm1) * R1 +
               4.0 * a.qp1 * b.qp1 * R2;
    double B = 0.0;
    if (la > 0 || lb > 0) B = (la > 0 ? a.qm1 : 1.0) * (lb > 0 ? b.qm1 : 1.0);
    return 4.0 * R2 * a.qp1 * b.qp1 + R0 * B;
}

// ---- AO tile layouts ------------------------------------------------------------------------------------
// STRIDE > 0: tile[d][k][pt], row stride STRIDE doubles; the points of a thread are 32 doubles apart.
// STRIDE < 0: derivative sets interleaved in pairs, tile[d/2][k][pt][d%2], row stride -STRIDE doubles (okb_ws.cuh: the
//             consumer fetches the B fragments of two sets with one 16-byte load, the producers store two sets with one
//             16-byte store); the points of a thread are 64 doubles apart and tp points at tile[0][0][pt][0].
template <int STRIDE>
struct TileLay {
    static constexpr bool IL = STRIDE < 0;
    static constexpr int RS = IL ? -STRIDE : STRIDE;
    static constexpr int PQ = IL ? 64 : 32;
    __host__ __device__ static constexpr size_t off(int d, int k) {
        return IL ? ((size_t)(d >> 1) * KC + k) * RS + (d & 1) : ((size_t)d * KC + k) * RS;
    }
    // store the D values of row k (relative to o = tile row 0 of the thread's point)
    template <int D>
    __device__ __forceinline__ static void store(double *__restrict__ o, int k, const double (&w)[D]) {
        if constexpr (IL) {
#pragma unroll
            for (int dp = 0; 2 * dp < D; ++dp) {
                if (2 * dp + 1 < D) *reinterpret_cast<double2 *>(o + off(2 * dp, k)) = make_double2(w[2 * dp], w[2 * dp + 1]);
                else o[off(2 * dp, k)] = w[2 * dp];
            }
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) o[off(d, k)] = w[d];
        }
    }
};

// ---- phase A: one (shell, point) item -------------------------------------------------------------
// tp points at the thread's point of tile row 0, set 0; element (d, k) lives at tp[TileLay<P>::off(d, k)].
template <int SET, int P>
__device__ __forceinline__ void gen_shell(const ShellMeta &sh, const double2 *__restrict__ prims,
                                          const FnMeta *__restrict__ fns, double x, double y, double z,
                                          double *__restrict__ tp, int one_code, int exact) {
    constexpr bool N1 = (SET != SET_VAL);
    constexpr bool N2 = (SET == SET_LAP || SET == SET_ALL || SET == SET_ONE || SET == SET_D2 || SET == SET_D2P);
    constexpr int LEVEL = N2 ? 2 : (N1 ? 1 : 0);
    const double X = x - sh.cx, Y = y - sh.cy, Z = z - sh.cz;
    const double rr = X * X + Y * Y + Z * Z;
    double R0 = 0.0, R1 = 0.0, R2 = 0.0;   // sum cN e, sum cN alpha e, sum cN alpha^2 e
    const double2 *pp = prims + sh.prim_off;
    for (int i = 0; i < sh.nprim; ++i) {
        const double2 ac = pp[i];
        const double arg = ac.x * rr;
        // exp(-arg) == 0.0 exactly in binary64 for arg > 745.14: skipping is bit-exact w.r.t. libm.
        if (__any_sync(0xffffffffu, arg < 746.0)) {
            const double t = ac.y * exp(-arg);
            R0 += t;
            if (N1) {
                const double ta = t * ac.x;
                R1 += ta;
                if (N2) R2 += ta * ac.x;
            }
        }
    }
    const FnMeta *ff = fns + sh.fn_off;
    for (int j = 0; j < sh.nfn; ++j) {
        const FnMeta fm = ff[j];
        const int lx = fm.lxyz & 0xff, ly = (fm.lxyz >> 8) & 0xff, lz = (fm.lxyz >> 16) & 0xff;
        const AxisQ ax = axis_q<LEVEL>(X, lx), ay = axis_q<LEVEL>(Y, ly), az = axis_q<LEVEL>(Z, lz);
        double *o = tp + TileLay<P>::off(0, sh.fn_off + j);
        const double f = fm.f;
        if (SET == SET_ONE) {
            double v;
            switch (one_code) {
                case 0: v = R0 * ax.q0 * ay.q0 * az.q0; break;
                case 1: v = ay.q0 * az.q0 * (ax.qm1 * R0 - 2.0 * ax.qp1 * R1); break;
                case 2: v = ax.q0 * az.q0 * (ay.qm1 * R0 - 2.0 * ay.qp1 * R1); break;
                case 3: v = ax.q0 * ay.q0 * (az.qm1 * R0 - 2.0 * az.qp1 * R1); break;
                case 4: v = ay.q0 * az.q0 * (ax.q0 * (4.0 * X * X * R2 - (double)(4 * lx + 2) * R1) + ax.qm2 * R0); break;
                case 5: v = ax.q0 * az.q0 * (ay.q0 * (4.0 * Y * Y * R2 - (double)(4 * ly + 2) * R1) + ay.qm2 * R0); break;
                case 6: v = ax.q0 * ay.q0 * (az.q0 * (4.0 * Z * Z * R2 - (double)(4 * lz + 2) * R1) + az.qm2 * R0); break;
                case 7: v = az.q0 * mixed2(ax, ay, lx, ly, R0, R1, R2, exact); break;
                case 8: v = ay.q0 * mixed2(ax, az, lx, lz, R0, R1, R2, exact); break;
                default: v = ax.q0 * mixed2(ay, az, ly, lz, R0, R1, R2, exact); break;
            }
            o[0] = f * v;
            continue;
        }
        const double yz = ay.q0 * az.q0, xz = ax.q0 * az.q0, xy = ax.q0 * ay.q0;
        if (SET == SET_D2P) {
            o[0] = f * (yz * (ax.q0 * (4.0 * X * X * R2 - (double)(4 * lx + 2) * R1) + ax.qm2 * R0));
            o[TileLay<P>::off(1, 0)] = f * (xz * (ay.q0 * (4.0 * Y * Y * R2 - (double)(4 * ly + 2) * R1) + ay.qm2 * R0));
            o[TileLay<P>::off(2, 0)] = f * (xy * (az.q0 * (4.0 * Z * Z * R2 - (double)(4 * lz + 2) * R1) + az.qm2 * R0));
            continue;
        }
        o[0] = f * (R0 * ax.q0 * yz);
        if (SET == SET_D2) {
            o[TileLay<P>::off(1, 0)] = f * (yz * (ax.q0 * (4.0 * X * X * R2 - (double)(4 * lx + 2) * R1) + ax.qm2 * R0));
            o[TileLay<P>::off(2, 0)] = f * (xz * (ay.q0 * (4.0 * Y * Y * R2 - (double)(4 * ly + 2) * R1) + ay.qm2 * R0));
            o[TileLay<P>::off(3, 0)] = f * (xy * (az.q0 * (4.0 * Z * Z * R2 - (double)(4 * lz + 2) * R1) + az.qm2 * R0));
            continue;
        }
        if (N1) {
            o[TileLay<P>::off(1, 0)] = f * (yz * (ax.qm1 * R0 - 2.0 * ax.qp1 * R1));
            o[TileLay<P>::off(2, 0)] = f * (xz * (ay.qm1 * R0 - 2.0 * ay.qp1 * R1));
            o[TileLay<P>::off(3, 0)] = f * (xy * (az.qm1 * R0 - 2.0 * az.qp1 * R1));
        }
        if (SET == SET_LAP || SET == SET_ALL) {
            o[TileLay<P>::off(4, 0)] = f * (yz * (ax.q0 * (4.0 * X * X * R2 - (double)(4 * lx + 2) * R1) + ax.qm2 * R0));
            o[TileLay<P>::off(5, 0)] = f * (xz * (ay.q0 * (4.0 * Y * Y * R2 - (double)(4 * ly + 2) * R1) + ay.qm2 * R0));
            o[TileLay<P>::off(6, 0)] = f * (xy * (az.q0 * (4.0 * Z * Z * R2 - (double)(4 * lz + 2) * R1) + az.qm2 * R0));
        }
        if (SET == SET_ALL) {
            o[TileLay<P>::off(7, 0)] = f * (az.q0 * mixed2(ax, ay, lx, ly, R0, R1, R2, exact));
            o[TileLay<P>::off(8, 0)] = f * (ay.q0 * mixed2(ax, az, lx, lz, R0, R1, R2, exact));
            o[TileLay<P>::off(9, 0)] = f * (ax.q0 * mixed2(ay, az, ly, lz, R0, R1, R2, exact));
        }
    }
}

}  // namespace okb
