/*
 * okb200.h -- C ABI of libokb200.so: the B200 (sm_100a) implementation of ORBKIT's grid-based
 * hot path  AO -> MO -> rho / grad rho / second derivatives of rho.
 *
 * Plain C, pointers and sizes only; no exceptions cross this boundary.  Every function returns an
 * int status (OKB_OK == 0) and leaves a message retrievable with okb_last_error() on failure --
 * the reference's C layer has no error channel at all (c_support.c:169-172 just prints), its
 * Python layer raises ValueError; the Python shim in orbkit_b200/ turns non-zero statuses back
 * into the reference's exceptions.
 *
 * Two levels are exported:
 *
 *  (1) drop-ins for the reference's compiled module `orbkit.cy_core` (cy_core.pyx:21-101) and the C
 *      routine under it (c_grid-based.h:1-3), HOST buffers in, HOST buffers out:
 *        okb_aocreator  <- cy_core.aocreator   (cy_core.pyx:51-78)  -> c_lcreator per contraction
 *        okb_lcreator   <- cy_core.lcreator / c_lcreator (cy_core.pyx:29-47, c_grid-based.c:9-79)
 *        okb_mocreator  <- cy_core.mocreator   (cy_core.pyx:82-101)
 *        okb_aonorm     <- cy_core.aonorm / ao_norm     (cy_core.pyx:21, c_support.c:177-188)
 *        okb_aoxyz      <- cy_core.aoxyz  / get_ao_xyz  (cy_core.pyx:24, c_support.c:28-175)
 *
 *  (2) the fused path behind core.rho_compute / rho_compute_no_slice / slice_rho
 *      (core.py:179-308, 314-605, 607-839): persistent handles for the basis tables, the MO
 *      coefficients and the grid, and one evaluation call per request that never materialises AO
 *      or MO arrays in HBM unless they are the requested output.
 *
 * All arrays are C-contiguous; doubles are IEEE binary64; index arrays are C int -- exactly the
 * dtypes tools.require() enforces in the reference (tools.py:290-295).  Pointers are borrowed for
 * the duration of the call.  A context is bound to one CUDA device and owns one stream; it is
 * thread-compatible (use one context per host thread), not re-entrant.
 */
#ifndef OKB200_H
#define OKB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct okb_ctx   okb_ctx;
typedef struct okb_basis okb_basis;
typedef struct okb_mo    okb_mo;
typedef struct okb_grid  okb_grid;

enum {
    OKB_OK = 0,
    OKB_ERR_ARG = 1,     /* invalid argument (maps to ValueError) */
    OKB_ERR_CUDA = 2,    /* CUDA runtime / launch failure, or no usable device */
    OKB_ERR_NOMEM = 3,   /* host or device allocation failed */
    OKB_ERR_UNSUPPORTED = 4
};

/* flags */
#define OKB_FLAG_EXACT_MIXED 1u  /* analytically correct xy/xz/yz AO derivatives instead of the
                                    reference's incomplete ones (c_support.c:121-168) */
#define OKB_FLAG_OUT_DEVICE  2u  /* output pointers are DEVICE pointers; the call is asynchronous
                                    on the context stream (okb_ctx_sync to wait) */

/* ---- library / context ------------------------------------------------------------------ */
const char *okb_last_error(void);
int  okb_version(void);
int  okb_device_count(int *n);
int  okb_ctx_create(int device, okb_ctx **out);
int  okb_ctx_destroy(okb_ctx *ctx);
int  okb_ctx_sync(okb_ctx *ctx);
void *okb_ctx_stream(okb_ctx *ctx);                       /* cudaStream_t of the context */
int  okb_ctx_launch_count(okb_ctx *ctx, long long *n);    /* kernels launched by this context */
int  okb_ctx_traffic(okb_ctx *ctx, long long *h2d_bytes, long long *d2h_bytes); /* PCIe bytes so far */
int  okb_ctx_last_kernel(okb_ctx *ctx, char *buf, int buflen); /* name of the last variant launched */

/* FP64 roofline denominator measured on the device: kind 0 = DFMA issue-bound loop, 1 = DMMA
 * (mma.sync.m8n8k4.f64), 2 = half the warps each.  min_seconds <= 0: burst (best single launch of 4);
 * > 0: launches back to back for at least that long (sustained, under the power cap).  Returns dense
 * TFLOP/s and the measured time. */
int  okb_measure_fp64(okb_ctx *ctx, int kind, double min_seconds, double *tflops, double *ms);

/* ---- (1) cy_core drop-ins, host buffers ------------------------------------------------- */
/* out[n_cart][npts] (row-major).  Arguments as cy_core.aocreator (cy_core.pyx:51-61):
 * lxlylz[n_cart][3], assign[n_cont] (functions per contraction), ao_coeffs[n_prim][2] (alpha,c),
 * pnum_list[n_cont], geo_spec[n_atoms][3], atom_indices[n_cont], x/y/z[npts], drv 0..9. */
int okb_aocreator(okb_ctx *ctx, const int *lxlylz, const int *assign, const double *ao_coeffs,
                  const int *pnum_list, const double *geo_spec, const int *atom_indices,
                  int n_cont, int n_cart, int n_prim, int n_atoms,
                  const double *x, const double *y, const double *z, long long npts,
                  int drv, int is_normalized, unsigned flags, double *out);
/* one contraction, written in place at row stride `row_stride` (>= npts) -- c_lcreator's layout
 * (c_grid-based.c:69). */
int okb_lcreator(okb_ctx *ctx, double *ao_list, long long row_stride, const int *lxlylz,
                 const double *coeff_list, const double *at_pos,
                 const double *x, const double *y, const double *z, long long npts,
                 int ao_num, int pnum, int drv, int is_normalized, unsigned flags);
/* mo[n_mo][npts] = coeffs[n_mo][n_ao] * ao[n_ao][npts]  (cy_core.pyx:82-101) */
int okb_mocreator(okb_ctx *ctx, const double *ao, const double *coeffs,
                  int n_ao, long long npts, int n_mo, double *mo);
double okb_aonorm(int lx, int ly, int lz, double alpha, int is_normalized);
double okb_aoxyz(double x, double y, double z, int lx, int ly, int lz, double alpha, int drv);

/* ---- (2) fused path: handles ------------------------------------------------------------- */
/* Basis tables.  Same arrays as okb_aocreator plus an optional per-Cartesian-function factor
 * `renorm[n_cart]` (ao_spec[0]['N'], core.py:98-100; NULL for none). */
int okb_basis_create(okb_ctx *ctx, const int *lxlylz, const int *assign, const double *ao_coeffs,
                     const int *pnum_list, const double *geo_spec, const int *atom_indices,
                     int n_cont, int n_cart, int n_prim, int n_atoms, int is_normalized,
                     const double *renorm, okb_basis **out);
/* Cartesian -> real-spherical transform as CSR over Cartesian rows (core.py:135-176):
 * sph[j] = sum_{t in [row_ptr[j],row_ptr[j+1])} val[t] * cart[col[t]].  After this call the AO
 * basis of `basis` has n_sph functions. */
int okb_basis_set_cart2sph(okb_basis *basis, int n_sph, const int *row_ptr, const int *col,
                           const double *val);
int okb_basis_info(okb_basis *basis, int *n_cart, int *n_ao, int *n_dev_shells, int *n_chunks);
int okb_basis_destroy(okb_basis *basis);

/* MO coefficients coeffs[n_mo][n_ao] (MOClass.get_coeffs, orbitals.py:760-771) and occupations
 * occ[n_mo] (get_occ :786) in the AO basis of `basis`. */
int okb_mo_create(okb_ctx *ctx, okb_basis *basis, int n_mo, const double *coeffs,
                  const double *occ, okb_mo **out);
int okb_mo_destroy(okb_mo *mo);

/* Grids.  Regular: axis vectors; the point index runs x slowest, z fastest (cy_grid.pyx:22-29) and
 * coordinates are generated in-kernel.  Vector: explicit coordinates (host, or device when
 * coords_on_device != 0; device arrays must stay alive while the grid is used). */
int okb_grid_regular(okb_ctx *ctx, const double *x, int nx, const double *y, int ny,
                     const double *z, int nz, okb_grid **out);
int okb_grid_vector(okb_ctx *ctx, const double *x, const double *y, const double *z,
                    long long npts, int coords_on_device, okb_grid **out);
/* Product grids in spherical (kind 2: a0 = r, a1 = theta, a2 = phi; cy_grid.sph2cart, cy_grid.pyx:58-75) or
 * cylindrical coordinates (kind 3: a0 = r, a1 = phi, a2 = zed; cy_grid.cyl2cart, cy_grid.pyx:79-97), first axis
 * slowest.  The Cartesian coordinates are generated in the kernels (same expressions and multiplication order as
 * the reference, sin/cos from the host libm): no coordinate upload.  affine (may be NULL) = 12 doubles, a row-major
 * 3x3 matrix A and a translation t applied as A (x,y,z) + t (grid.grid_sym_op / grid_translate, grid.py:293-321). */
#define OKB_GRID_SPHERICAL 2
#define OKB_GRID_CYLINDRICAL 3
int okb_grid_product(okb_ctx *ctx, int kind, const double *a0, int n0, const double *a1, int n1,
                     const double *a2, int n2, const double *affine, okb_grid **out);
int okb_grid_size(okb_grid *grid, long long *npts);
int okb_grid_destroy(okb_grid *grid);

/* ---- (2) fused path: evaluation over the point range [p0, p1) ---------------------------- */
/* AOs (calc_ao): out[n_drv][n_ao][p1-p0]; drv_codes[i] in 0..9 (tools.validate_drv). */
int okb_eval_ao(okb_ctx *ctx, okb_basis *basis, okb_grid *grid, long long p0, long long p1,
                const int *drv_codes, int n_drv, double *out, unsigned flags);
/* MOs (calc_mo): out[n_drv][n_mo][p1-p0]. */
int okb_eval_mo(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                const int *drv_codes, int n_drv, double *out, unsigned flags);
/* Density: rho[p1-p0]; delta_rho[n_drv][p1-p0] with, per code d,
 *   sum_i occ_i * 2 * d_d(phi_i) * phi_i  (+ sum_i occ_i * 2 * d_a(phi_i) d_b(phi_i) for the
 *   second-derivative codes 4..9, (a,b) the two letters) -- core.py:265-304.
 * mo_norm[n_mo] (may be NULL) receives sum_points phi_i^2 (core.py:273); always a HOST pointer. */
int okb_eval_rho(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                 const int *drv_codes, int n_drv, double *rho, double *delta_rho,
                 double *mo_norm, unsigned flags);

/* The same three calls writing into rows of a LARGER array: ld_out = row stride of the output arrays in
 * points (>= p1-p0; 0 means dense).  out / rho / delta_rho point at the first element of the point range
 * inside its row.  This is how the ranks of a multi-GPU job write their point shards side by side into one
 * shared host array (each GPU copies into its disjoint slice, no gather). */
int okb_eval_ao_ld(okb_ctx *ctx, okb_basis *basis, okb_grid *grid, long long p0, long long p1,
                   const int *drv_codes, int n_drv, double *out, long long ld_out, unsigned flags);
int okb_eval_mo_ld(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                   const int *drv_codes, int n_drv, double *out, long long ld_out, unsigned flags);
int okb_eval_rho_ld(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                    const int *drv_codes, int n_drv, double *rho, double *delta_rho, long long ld_out,
                    double *mo_norm, unsigned flags);

/* ---- (3) detCI grid contractions --------------------------------------------
 * Replace the per-slice loops cy_ci.get_rho / get_jab / get_a_nabla_b (orbkit/detci/cy_ci.pyx:70-97,
 * 156-186, 211-240) behind detci.ci_core.rho / jab / a_nabla_b (ci_core.py:85-267).  The Python lists
 * `zero` and `sing` arrive flattened into n_terms terms (coef[t], ia[t], ib[t]) in list order; for
 * OKB_CI_RHO the entries of `zero` are terms with ia == ib.  Results are bit-identical to the
 * reference loops for identical MO arrays (same expression order, no FMA contraction).
 *   OKB_CI_RHO        out[npts]          = sum_t coef mo[ia] mo[ib]
 *   OKB_CI_JAB        out[3][npts]       = sum_t -1/2 coef (mo[ia] dmo[d][ib] - mo[ib] dmo[d][ia])
 *   OKB_CI_A_NABLA_B  out[3][npts]       = sum_t coef mo[ia] dmo[d][ib]
 *   OKB_CI_PAIRS      out[n_terms][npts] = mo[ia] ket[ib]      (per-pair products, coef ignored; ket = molistdrv[n_mo][.]
 *                     if given, else mo: the products of core.calc_mo_matrix, core.py:925-941)
 *   OKB_CI_JAB_PAIRS  out[3][n_terms][npts] = -1/2 (mo[ia] dmo[d][ib] - mo[ib] dmo[d][ia]) per pair (extras.calc_jmo,
 *                     extras.py:441-493, restricted to the requested pairs; coef ignored) */
#define OKB_CI_RHO 0
#define OKB_CI_JAB 1
#define OKB_CI_A_NABLA_B 2
#define OKB_CI_PAIRS 3
#define OKB_CI_JAB_PAIRS 4
#define OKB_FLAG_IN_DEVICE 4u    /* okb_ci_contract: molist / molistdrv are DEVICE pointers */
#define OKB_FLAG_CI_FAST   8u    /* okb_ci_contract / okb_eval_ci / okb_ci_jab_full, summed modes: re-ordered sums.  Many terms
                                  * per orbital pair (r n_terms >= n_act^2): the terms are summed on the host into the dense
                                  * orbital-pair matrix and the grid work runs on the FP64 tensor path (n_act^2 fused
                                  * multiply-adds per point instead of n_terms gathers); otherwise the term list is split
                                  * over the warps of a CTA and every MO row is staged in shared memory once (DRAM traffic =
                                  * algorithmic instead of 3.5x).  Deterministic, equal to the sequential sum to rounding
                                  * (~1e-15 of the largest value) -- NOT bit for bit with the reference's loops; without the
                                  * flag the sums keep the reference's order. */
/* Level 1: given MO arrays molist[n_mo][ld_in] and (JAB, A_NABLA_B) molistdrv[3][n_mo][ld_in]; the first
 * npts points of every row are contracted into the first npts entries of the rows of out[.][ld_out]
 * (ld_* = row strides in points, >= npts). */
int okb_ci_contract(okb_ctx *ctx, int mode, int n_mo, long long npts, long long ld_in,
                    const double *molist, const double *molistdrv, int n_terms, const double *coef,
                    const int *ia, const int *ib, double *out, long long ld_out, unsigned flags);
/* Level 2 (fused): the MOs of `mo` (and, for JAB / A_NABLA_B / JAB_PAIRS, the three derivative sets drv_codes[3],
 * e.g. {1,2,3} or {4,5,6}; for PAIRS optionally ONE set drv_codes[0] for the second factor) are evaluated slab by
 * slab on the device over the points [p0, p1) of `grid` and contracted there; MO values never cross PCIe.
 * out as above with npts = p1 - p0. */
int okb_eval_ci(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1, int mode,
                const int *drv_codes, int n_terms, const double *coef, const int *ia,
                const int *ib, double *out, unsigned flags);

/* Time-dependent detCI contractions (orbkit/detci/cy_ci.pyx:101-122 get_rho_full, 126-151 get_j_full -- the reference's
 * OpenMP prange loops):  out[t][x] = sum_k w[t][k] * in[k][x],  t < nt, k < nk, x < n.
 *   get_rho_full(ReS[nt][ns][ns], rho[npair][npts]): in = rho, w[t][count] = ReS[t,m,m] (m == n) or 2 ReS[t,m,n] for the
 *     pairs count = (n, m >= n) in the reference's order;
 *   get_j_full(ImS[nt][ns][ns], j[npair][3][npts]): in = j read as rows of n = 3 npts entries, w[t][count] = -2 ImS[t,n,m]
 *     (0 on the diagonal pairs); out = tdj[nt][3][npts].
 * w is a HOST array [nt][nk]; in / out are host rows (staged in slabs) or device rows (OKB_FLAG_IN_DEVICE /
 * OKB_FLAG_OUT_DEVICE) of stride ld_in / ld_out >= n.  FP64 tensor cores (DMMA), k in the reference's order: agrees with
 * the reference to rounding (a few ulp of the largest term), not bit for bit. */
int okb_ci_td(okb_ctx *ctx, int nt, int nk, long long n, const double *w, const double *in, long long ld_in,
              double *out, long long ld_out, unsigned flags);
/* cy_ci.get_jab_full (cy_ci.pyx:186-202):
 *   out[c][x] = sum_n sum_{m<n} ImS[n][m] / mu * (chi[n][x] dchi[c][m][x] - chi[m][x] dchi[c][n][x]),  c < ncomp <= 3
 * ImS: HOST [nbasis][nbasis]; chi[nbasis][ld_in], dchi[ncomp][nbasis][ld_in], out[ncomp][ld_out] host or device as above.
 * Sequential sums in the reference's order without fused multiply-add: bit-identical to the reference. */
int okb_ci_jab_full(okb_ctx *ctx, int nbasis, int ncomp, long long npts, long long ld_in, const double *ImS,
                    const double *chi, const double *dchi, double mu, double *out, long long ld_out, unsigned flags);

/* ---- analytic overlap matrix (orbkit/cy_overlap.pyx:75-156 aooverlap; c_non-grid-based.c:9-52) --------------------
 * The argument list of cy_overlap.aooverlap: contraction i has assign[i] Cartesian functions (rows of lxlylz_a / _b, bra /
 * ket exponents) x pnum_list[i] primitives (rows of ao_coeffs: exponent, coefficient) on atom atom_indices[i] of geo_a
 * (bra) / geo_b (ket), [n_atom][3].  drv 0: overlap; 1..3: <a| d/dx b>, d/dy, d/dz.  aoom: HOST [ao_num][ao_num].
 * Used by main_read (Molden renormalisation, check_norm) through orbkit_b200.analytical_integrals.get_ao_overlap. */
int okb_aooverlap(okb_ctx *ctx, const double *geo_a, const double *geo_b, int n_atom, const int *lxlylz_a,
                  const int *lxlylz_b, int ao_num, const int *assign, const double *ao_coeffs, const int *pnum_list,
                  const int *atom_indices, int n_cont, int drv, int is_normalized, double *aoom);

/* ---- output sink: Gaussian cube text (orbkit/output/cube.py:5-101, cube_creator) ------------------------------
 * The data loop of cube_creator (cube.py:86-96) on the device: data[n_sets][nx][ny][nz] (C order, float64) becomes
 * the text the reference writes -- per (x, y) row the nz * n_sets values ('%.5E' right-justified in 13 columns, the
 * sets of one point next to each other), a newline behind every 6th value of the row and one at its end.  '%.5E' is
 * correctly rounded (round-half-even on the exact binary value) like Python's float formatting, so the bytes are
 * identical to the reference's.  The few header lines (cube.py:47-84) stay on the host.
 * okb_cube_body_bytes: size of that text, nx * ny * (13 n + n/6 + 1) with n = nz * n_sets (-1 for bad extents).
 * okb_format_cube: `data` host (default) or device (OKB_FLAG_IN_DEVICE) pointer; `text` host (default) or device
 * (OKB_FLAG_OUT_DEVICE) buffer of `capacity` >= okb_cube_body_bytes(...) bytes; host buffers are staged in slabs. */
long long okb_cube_body_bytes(int n_sets, long long nx, long long ny, long long nz);
int okb_format_cube(okb_ctx *ctx, const double *data, int n_sets, long long nx, long long ny, long long nz,
                    char *text, long long capacity, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* OKB200_H */
