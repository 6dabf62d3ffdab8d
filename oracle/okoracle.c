/*
 * okoracle.c -- CPU ORACLE for the ORBKIT grid-based hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * as the timed CPU baseline.  The product path (orbkit_b200/) never links,
 * imports or calls it and fails loudly without its CUDA library.
 *
 * This is a plain-C restatement (not a copy) of the arithmetic the reference
 * performs in
 *     orbkit/c_support.c:8-195      (ipow, xyz, get_ao_xyz, ao_norm, doublefactorial)
 *     orbkit/c_grid-based.c:9-79    (c_lcreator: one contraction on all points)
 *     orbkit/cy_core.pyx:51-101     (aocreator shell loop, mocreator triple loop)
 *     orbkit/cy_grid.pyx:14-55      (grid2vector / vector2grid)
 * The operation ORDER of every floating-point expression follows the
 * reference so that this port is bit-identical to the reference's own
 * objects compiled with the same compiler (checked by
 * tests/test_oracle.py against oracle/_ref/libokref.so, which is built from
 * the reference's untouched sources).  Parity is pinned: see
 * tests/golden/ (reference golden .npz + Gaussian cubegen KATs).
 *
 * Known reference bugs are reproduced on purpose (parity mode):
 *   - mixed second derivatives (codes 7,8,9) drop the -2*alpha cross terms
 *     (c_support.c:121-168).  okor_poly(..., exact_mixed=1) gives the
 *     analytically correct value for documentation/tests of the opt-in flag.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- scalar helpers ------------------------------------------------- */

/* base**e by binary exponentiation, e >= 0 (c_support.c:8-20). Negative
 * exponents return 1.0 exactly like the reference's while(exp>0) loop. */
double okor_ipow(double base, int e)
{
    double acc = 1.0;
    for (; e > 0; e >>= 1) {
        if (e & 1) acc *= base;
        base *= base;
    }
    return acc;
}

/* x^lx * y^ly * z^lz (c_support.c:22-26) */
double okor_mono(double x, double y, double z, int lx, int ly, int lz)
{
    return okor_ipow(x, lx) * okor_ipow(y, ly) * okor_ipow(z, lz);
}

static int okor_dfact(int n) /* n!! with (<=0)!! = 1 (c_support.c:190-195) */
{
    int r = 1;
    for (; n > 0; n -= 2) r *= n;
    return r;
}

/* primitive normalisation (c_support.c:177-188) */
double okor_ao_norm(int l, int m, int n, double alpha, int is_normalized)
{
    if (is_normalized > 0) return 1.0;
    double pref = pow(2. / M_PI, 3. / 4.);
    double rad  = pow(2., (double)(l + m + n)) *
                  pow(alpha, (2. * l + 2. * m + 2. * n + 3.) / 4.);
    double den  = pow((double)(okor_dfact(2 * l - 1) * okor_dfact(2 * m - 1) *
                               okor_dfact(2 * n - 1)), 0.5);
    return pref * rad / den;
}

/*
 * Polynomial prefactor of d^drv [x^lx y^ly z^lz exp(-alpha r^2)] / exp(-alpha r^2)
 * (c_support.c:28-175).  drv: 0 value; 1,2,3 = d/dx,d/dy,d/dz;
 * 4,5,6 = d2/dx2,d2/dy2,d2/dz2; 7,8,9 = dxdy, dxdz, dydz.
 * exact_mixed=0 reproduces the reference's (incomplete) codes 7-9.
 */
double okor_poly(double X, double Y, double Z, int lx, int ly, int lz,
                 double alpha, int drv, int exact_mixed)
{
    double r[3] = {X, Y, Z};
    int    l[3] = {lx, ly, lz};
    if (drv == 0) return okor_mono(X, Y, Z, lx, ly, lz);

    if (drv >= 1 && drv <= 3) {            /* first derivatives (:58-93) */
        int a = drv - 1;
        int up[3] = {lx, ly, lz}, dn[3] = {lx, ly, lz};
        up[a] += 1; dn[a] -= 1;
        if (l[a] == 0)
            return -2 * alpha * okor_mono(X, Y, Z, up[0], up[1], up[2]);
        return l[a] * okor_mono(X, Y, Z, dn[0], dn[1], dn[2])
             - 2 * alpha * okor_mono(X, Y, Z, up[0], up[1], up[2]);
    }
    if (drv >= 4 && drv <= 6) {            /* pure second derivatives (:94-120) */
        int a = drv - 4;
        double v = 2 * alpha * okor_mono(X, Y, Z, lx, ly, lz)
                 * (2 * alpha * (r[a] * r[a]) - (2 * l[a] + 1));
        if (l[a] >= 2) {
            int dn[3] = {lx, ly, lz};
            dn[a] -= 2;
            v += (double)(l[a] * l[a] - l[a]) * okor_mono(X, Y, Z, dn[0], dn[1], dn[2]);
        }
        return v;
    }
    if (drv >= 7 && drv <= 9) {            /* mixed second derivatives (:121-168) */
        static const int pa[3] = {0, 0, 1}, pb[3] = {1, 2, 2};
        int a = pa[drv - 7], b = pb[drv - 7];
        int up[3] = {lx, ly, lz};
        up[a] += 1; up[b] += 1;
        if (exact_mixed) {
            /* (l_a r^{l_a-1} - 2 alpha r^{l_a+1}) (l_b r^{l_b-1} - 2 alpha r^{l_b+1}) * rest */
            int c = 3 - a - b;
            double fa = (l[a] ? l[a] * okor_ipow(r[a], l[a] - 1) : 0.0)
                      - 2 * alpha * okor_ipow(r[a], l[a] + 1);
            double fb = (l[b] ? l[b] * okor_ipow(r[b], l[b] - 1) : 0.0)
                      - 2 * alpha * okor_ipow(r[b], l[b] + 1);
            return fa * fb * okor_ipow(r[c], l[c]);
        }
        double v = 4 * (alpha * alpha) * okor_mono(X, Y, Z, up[0], up[1], up[2]);
        int dn[3] = {lx, ly, lz};
        if (l[a] == 0 && l[b] > 0) {
            dn[b] -= 1;
            v += l[b] * okor_mono(X, Y, Z, dn[0], dn[1], dn[2]);
        } else if (l[a] > 0 && l[b] == 0) {
            dn[a] -= 1;
            v += l[a] * okor_mono(X, Y, Z, dn[0], dn[1], dn[2]);
        } else if (l[a] > 0 && l[b] > 0) {
            dn[a] -= 1; dn[b] -= 1;
            v += l[a] * l[b] * okor_mono(X, Y, Z, dn[0], dn[1], dn[2]);
        }
        return v;
    }
    return 0.0;                            /* invalid code: reference yields 0 (:169-172) */
}

/* ---- one contraction on npts points (c_grid-based.c:9-79) ------------ */
/*
 * out      [ao_num][row_stride]  (row_stride >= npts; the reference passes a
 *                                 row-offset view of the big array, stride npts)
 * lxlylz   [ao_num][3] int
 * coeffs   [pnum][2]   (alpha, c)
 */
void okor_lcreator(double *out, long row_stride, const int *lxlylz,
                   const double *coeffs, const double *centre,
                   const double *x, const double *y, const double *z,
                   long npts, int ao_num, int pnum, int drv,
                   int is_normalized, int exact_mixed)
{
    double *nrm = (double *)malloc(sizeof(double) * (size_t)ao_num * (size_t)pnum);
    double *rad = (double *)malloc(sizeof(double) * (size_t)pnum);
    for (int f = 0; f < ao_num; ++f)
        for (int p = 0; p < pnum; ++p)
            nrm[f * pnum + p] = okor_ao_norm(lxlylz[3 * f], lxlylz[3 * f + 1],
                                             lxlylz[3 * f + 2], coeffs[2 * p],
                                             is_normalized);
    for (long i = 0; i < npts; ++i) {
        double X = x[i] - centre[0], Y = y[i] - centre[1], Z = z[i] - centre[2];
        double rr = X * X + Y * Y + Z * Z;
        for (int p = 0; p < pnum; ++p)
            rad[p] = coeffs[2 * p + 1] * exp(-coeffs[2 * p] * rr);
        for (int f = 0; f < ao_num; ++f) {
            const int lx = lxlylz[3 * f], ly = lxlylz[3 * f + 1], lz = lxlylz[3 * f + 2];
            double s = 0.;
            if (drv == 0) {
                for (int p = 0; p < pnum; ++p) s += nrm[f * pnum + p] * rad[p];
                s *= okor_mono(X, Y, Z, lx, ly, lz);
            } else {
                for (int p = 0; p < pnum; ++p)
                    s += nrm[f * pnum + p] * rad[p] *
                         okor_poly(X, Y, Z, lx, ly, lz, coeffs[2 * p], drv, exact_mixed);
            }
            out[(long)f * row_stride + i] = s;
        }
    }
    free(nrm);
    free(rad);
}

/* ---- all contractions (cy_core.pyx:51-78) ---------------------------- */
/*
 * out [n_cart][npts] zero-initialised by the caller (numpy.zeros in the
 * reference, :67).  assign[s] = number of Cartesian functions of contraction
 * s; pnum_list[s] = number of primitives; atom_idx[s] indexes geo[n_at][3].
 */
void okor_aocreator(double *out, const int *lxlylz, const int *assign,
                    const double *ao_coeffs, const int *pnum_list,
                    const double *geo, const int *atom_idx, int n_cont,
                    const double *x, const double *y, const double *z,
                    long npts, int drv, int is_normalized, int exact_mixed)
{
    long c_ao = 0, c_p = 0;
    for (int s = 0; s < n_cont; ++s) {
        okor_lcreator(out + c_ao * npts, npts, lxlylz + 3 * c_ao,
                      ao_coeffs + 2 * c_p, geo + 3 * atom_idx[s], x, y, z,
                      npts, assign[s], pnum_list[s], drv, is_normalized,
                      exact_mixed);
        c_ao += assign[s];
        c_p  += pnum_list[s];
    }
}

/* ---- MO contraction (cy_core.pyx:82-101): mo[i,j] = sum_k C[i,k] ao[k,j],
 * k ascending, plain += starting from 0.0 ------------------------------ */
void okor_mocreator(double *mo, const double *ao, const double *C,
                    int n_mo, int n_ao, long npts)
{
    for (int i = 0; i < n_mo; ++i)
        for (long j = 0; j < npts; ++j) {
            double v = 0.0;
            for (int k = 0; k < n_ao; ++k) v += C[(long)i * n_ao + k] * ao[(long)k * npts + j];
            mo[(long)i * npts + j] = v;
        }
}

/* ---- regular grid <-> vector grid (cy_grid.pyx:14-55): x slowest, z fastest */
void okor_grid2vector(double *xyz, const double *x, const double *y, const double *z,
                      long nx, long ny, long nz)
{
    long n = nx * ny * nz, c = 0;
    for (long i = 0; i < nx; ++i)
        for (long j = 0; j < ny; ++j)
            for (long k = 0; k < nz; ++k, ++c) {
                xyz[c] = x[i]; xyz[n + c] = y[j]; xyz[2 * n + c] = z[k];
            }
}

void okor_vector2grid(double *xn, double *yn, double *zn,
                      const double *x, const double *y, const double *z,
                      long nx, long ny, long nz)
{
    for (long i = 0; i < nx; ++i) xn[i] = x[i * ny * nz];
    for (long j = 0; j < ny; ++j) yn[j] = y[j * nz];
    for (long k = 0; k < nz; ++k) zn[k] = z[k];
}

/* ---- detCI grid contractions (orbkit/detci/cy_ci.pyx) --------------------
 * The reference walks Python lists `zero = [coefficients, orbital indices]`
 * (identical determinants) and `sing = [coefficients, orbital pairs]`
 * (effective single excitations); here they arrive flattened:
 *   zc[nz], zi[nz]           every (prefactor, orbital) entry of zero, in list order
 *   sc[ns], sa[ns], sb[ns]   every (product of CI coefficients, orbital a, orbital b) of sing
 * molist is [n_mo][npts], molistdrv [3][n_mo][npts]; the slice [i0, i1) of the
 * points is written to out (rho: [i1-i0]; jab / a_nabla_b: [3][i1-i0]).
 * Operation order follows the reference expression by expression. */

/* get_rho, cy_ci.pyx:70-97:  rho[x] += c*mo[a,x]*mo[a,x]  then  rho[x] += c*mo[a,x]*mo[b,x] */
void okor_ci_rho(double *out, long i0, long i1, long npts, const double *molist,
                 long nz, const double *zc, const int *zi,
                 long ns, const double *sc, const int *sa, const int *sb)
{
    long slen = i1 - i0, t, x;
    for (x = 0; x < slen; ++x) out[x] = 0.0;
    for (t = 0; t < nz; ++t) {
        const double c = zc[t];
        const double *m = molist + (long)zi[t] * npts + i0;
        for (x = 0; x < slen; ++x) out[x] += c * m[x] * m[x];
    }
    for (t = 0; t < ns; ++t) {
        const double c = sc[t];
        const double *ma = molist + (long)sa[t] * npts + i0;
        const double *mb = molist + (long)sb[t] * npts + i0;
        for (x = 0; x < slen; ++x) out[x] += c * ma[x] * mb[x];
    }
}

/* get_jab, cy_ci.pyx:156-186:  jab[d,x] -= 0.5*(c*(mo[a,x]*dmo[d,b,x] - mo[b,x]*dmo[d,a,x])) */
void okor_ci_jab(double *out, long i0, long i1, long npts, long n_mo, const double *molist,
                 const double *molistdrv, long ns, const double *sc, const int *sa, const int *sb)
{
    long slen = i1 - i0, t, x;
    int d;
    for (x = 0; x < 3 * slen; ++x) out[x] = 0.0;
    for (t = 0; t < ns; ++t) {
        const double c = sc[t];
        const double *ma = molist + (long)sa[t] * npts + i0;
        const double *mb = molist + (long)sb[t] * npts + i0;
        for (d = 0; d < 3; ++d) {
            const double *da = molistdrv + ((long)d * n_mo + sa[t]) * npts + i0;
            const double *db = molistdrv + ((long)d * n_mo + sb[t]) * npts + i0;
            double *o = out + (long)d * slen;
            for (x = 0; x < slen; ++x) o[x] -= 0.5 * (c * (ma[x] * db[x] - mb[x] * da[x]));
        }
    }
}

/* get_a_nabla_b, cy_ci.pyx:211-240:  out[d,x] += c*(mo[a,x]*dmo[d,b,x]) */
void okor_ci_a_nabla_b(double *out, long i0, long i1, long npts, long n_mo, const double *molist,
                       const double *molistdrv, long ns, const double *sc, const int *sa, const int *sb)
{
    long slen = i1 - i0, t, x;
    int d;
    for (x = 0; x < 3 * slen; ++x) out[x] = 0.0;
    for (t = 0; t < ns; ++t) {
        const double c = sc[t];
        const double *ma = molist + (long)sa[t] * npts + i0;
        for (d = 0; d < 3; ++d) {
            const double *db = molistdrv + ((long)d * n_mo + sb[t]) * npts + i0;
            double *o = out + (long)d * slen;
            for (x = 0; x < slen; ++x) o[x] += c * (ma[x] * db[x]);
        }
    }
}

/* ---- time-dependent detCI contractions (orbkit/detci/cy_ci.pyx:101-151, 186-202) ---------------------------
 * get_rho_full: tdrho[t,r] = sum over the state pairs count = (n, m >= n) of ReS[t,m,m] rho[count,r] (m == n) or
 * 2 ReS[t,m,n] rho[count,r]; the sum runs n outer, m inner, one rounding per product and per addition. */
void okor_ci_rho_full(double *tdrho, const double *ReS, const double *rho, long nt, long nstate, long npts)
{
    long r, t, n, m, count;
    for (r = 0; r < npts; ++r)
        for (t = 0; t < nt; ++t) {
            double tmp = 0.0;
            count = 0;
            for (n = 0; n < nstate; ++n)
                for (m = n; m < nstate; ++m) {
                    if (m == n) tmp = tmp + ReS[(t * nstate + m) * nstate + m] * rho[count * npts + r];
                    else tmp = tmp + 2.0 * ReS[(t * nstate + m) * nstate + n] * rho[count * npts + r];
                    ++count;
                }
            tdrho[t * npts + r] = tmp;
        }
}

/* get_j_full, cy_ci.pyx:126-151: tdj[t,d,r] = - sum over the pairs m > n of 2 ImS[t,n,m] j[count,d,r]
 * (count also runs over the diagonal pairs, which are skipped) */
void okor_ci_j_full(double *tdj, const double *ImS, const double *j, long nt, long nstate, long npts)
{
    long r, t, n, m, count;
    for (r = 0; r < npts; ++r)
        for (t = 0; t < nt; ++t) {
            double tx = 0.0, ty = 0.0, tz = 0.0;
            count = 0;
            for (n = 0; n < nstate; ++n)
                for (m = n; m < nstate; ++m) {
                    if (m != n) {
                        const double s = ImS[(t * nstate + n) * nstate + m];
                        tx = tx - (2.0 * s * j[(count * 3 + 0) * npts + r]);
                        ty = ty - (2.0 * s * j[(count * 3 + 1) * npts + r]);
                        tz = tz - (2.0 * s * j[(count * 3 + 2) * npts + r]);
                    }
                    ++count;
                }
            tdj[(t * 3 + 0) * npts + r] = tx;
            tdj[(t * 3 + 1) * npts + r] = ty;
            tdj[(t * 3 + 2) * npts + r] = tz;
        }
}

/* get_jab_full, cy_ci.pyx:186-202: j[c,r] = sum_n sum_{m<n} f ImS[n,m] (chi[n,r] dchi[c,m,r] - chi[m,r] dchi[c,n,r]),
 * f = 1/mu */
void okor_ci_jab_full(double *out, const double *ImS, const double *chi, const double *dchi, double mu,
                      long nbasis, long ncomp, long npts)
{
    long c, r, n, m;
    const double f = 1. / mu;
    for (c = 0; c < ncomp; ++c)
        for (r = 0; r < npts; ++r) {
            double tmp = 0.0;
            for (n = 0; n < nbasis; ++n)
                for (m = 0; m < n; ++m)
                    tmp = tmp + f * ImS[n * nbasis + m] *
                                    (chi[n * npts + r] * dchi[(c * nbasis + m) * npts + r] -
                                     chi[m * npts + r] * dchi[(c * nbasis + n) * npts + r]);
            out[c * npts + r] = tmp;
        }
}

/* ---- analytic overlap integrals (orbkit/c_non-grid-based.c:9-52, orbkit/cy_overlap.pyx:24-156) -----------------
 * Primitive Cartesian Gaussians: S = E_AB (pi/(a+b))^(3/2) prod_i s_i(la_i, lb_i) with the Obara-Saika-type recursion of
 * c_non-grid-based.c:36-52 (initial conditions, recurrence in a, transfer equation). */
typedef struct { double alpha; int l[3]; double R[3]; } okor_prim;

static double okor_s(int i, int a, int b, const okor_prim *A, const okor_prim *B)
{
    if (a == 0 && b == 0) return 1.;
    else if (a == 1 && b == 0)
        return -(A->R[i] - ((A->alpha * A->R[i] + B->alpha * B->R[i]) / (A->alpha + B->alpha)));
    else if (b == 0)
        return -(A->R[i] - (A->alpha * A->R[i] + B->alpha * B->R[i]) / (A->alpha + B->alpha)) * okor_s(i, a - 1, 0, A, B) +
               ((a - 1) / (2. * (A->alpha + B->alpha))) * okor_s(i, a - 2, 0, A, B);
    else
        return okor_s(i, a + 1, b - 1, A, B) + (A->R[i] - B->R[i]) * okor_s(i, a, b - 1, A, B);
}

static double okor_prim_overlap(const okor_prim *A, const okor_prim *B)
{
    double rr = 0., EAB, ov;
    int i;
    for (i = 0; i < 3; ++i) rr += (A->R[i] - B->R[i]) * (A->R[i] - B->R[i]);
    EAB = exp(-((A->alpha * B->alpha) / (A->alpha + B->alpha)) * rr);
    ov = EAB * pow((M_PI / (A->alpha + B->alpha)), 3. / 2.);
    for (i = 0; i < 3; ++i) ov *= okor_s(i, A->l[i], B->l[i], A, B);
    return ov;
}

/* cy_overlap.aooverlap (cy_overlap.pyx:75-156).  The contractions are expanded to (primitive, function) entries in the
 * order contraction -> function -> primitive; aoom[fn_i][fn_j] accumulates over all entry pairs, i outer, j inner.
 * drv 0: overlap; 1..3: <a| d/dx_drv b> through the exponents of the ket (lxlylz_b). */
void okor_aooverlap(double *aoom, const double *geo_a, const double *geo_b, const int *lxlylz_a, const int *lxlylz_b,
                    long ao_num, const int *assign, const double *ao_coeffs, const int *pnum_list,
                    const int *atom_indices, long ncont, int drv, int is_normalized)
{
    long nindex = 0, i, j, c = 0, c_ao = 0, c_p = 0;
    int i_ao, i_p, rr;
    for (i = 0; i < ncont; ++i) nindex += (long)pnum_list[i] * assign[i];
    double *norm = (double *)malloc(sizeof(double) * (nindex > 0 ? nindex : 1));
    long *ip = (long *)malloc(sizeof(long) * 3 * (nindex > 0 ? nindex : 1));
    for (i = 0; i < ao_num * ao_num; ++i) aoom[i] = 0.;
    for (i = 0; i < ncont; ++i) {
        for (i_ao = 0; i_ao < assign[i]; ++i_ao)
            for (i_p = 0; i_p < pnum_list[i]; ++i_p) {
                const int *l = lxlylz_a + 3 * (c_ao + i_ao);
                norm[c] = okor_ao_norm(l[0], l[1], l[2], ao_coeffs[2 * (c_p + i_p)], is_normalized);
                ip[3 * c] = c_p + i_p;
                ip[3 * c + 1] = c_ao + i_ao;
                ip[3 * c + 2] = atom_indices[i];
                ++c;
            }
        c_ao += assign[i];
        c_p += pnum_list[i];
    }
    for (i = 0; i < nindex; ++i)
        for (j = 0; j < nindex; ++j) {
            const long i_l = ip[3 * i + 1], j_l = ip[3 * j + 1];
            const double *pa = ao_coeffs + 2 * ip[3 * i], *pb = ao_coeffs + 2 * ip[3 * j];
            okor_prim A, B;
            double *dst = aoom + i_l * ao_num + j_l;
            for (rr = 0; rr < 3; ++rr) {
                A.R[rr] = geo_a[3 * ip[3 * i + 2] + rr];
                B.R[rr] = geo_b[3 * ip[3 * j + 2] + rr];
                A.l[rr] = lxlylz_a[3 * i_l + rr];
                B.l[rr] = lxlylz_b[3 * j_l + rr];
            }
            A.alpha = pa[0];
            B.alpha = pb[0];
            if (drv <= 0) {
                *dst += (pa[1] * pb[1] * norm[i] * norm[j] * okor_prim_overlap(&A, &B));
            } else if (B.l[drv - 1] == 0) {
                B.l[drv - 1] = lxlylz_b[3 * j_l + drv - 1] + 1;
                *dst += ((-2 * B.alpha) * pa[1] * pb[1] * norm[i] * norm[j] * okor_prim_overlap(&A, &B));
            } else {
                const int lb = lxlylz_b[3 * j_l + drv - 1];
                B.l[drv - 1] = lb - 1;
                *dst += (lb * pa[1] * pb[1] * norm[i] * norm[j] * okor_prim_overlap(&A, &B));
                B.l[drv - 1] = lb + 1;
                *dst += ((-2 * B.alpha) * pa[1] * pb[1] * norm[i] * norm[j] * okor_prim_overlap(&A, &B));
            }
        }
    free(norm);
    free(ip);
}

/* cy_overlap.ommited_cca_norm (cy_overlap.pyx:24-52) / tmol_aomix_norm (54-72) */
void okor_cca_norm(double *norm, const int *lxlylz, long ao_num, int with_divisor)
{
    long i;
    for (i = 0; i < ao_num; ++i) {
        const int *l = lxlylz + 3 * i;
        const double num = (double)(okor_dfact(2 * l[0] - 1) * okor_dfact(2 * l[1] - 1) * okor_dfact(2 * l[2] - 1));
        norm[i] = with_divisor ? sqrt(num / (double)okor_dfact(2 * (l[0] + l[1] + l[2]) - 1)) : sqrt(num);
    }
}

/* cy_overlap.mooverlapmatrix (cy_overlap.pyx:177-205): moom[i][j] = sum_k sum_l mo_a[i,k] mo_b[j,l] aoom[k,l] */
void okor_mooverlapmatrix(double *moom, const double *mo_a, const double *mo_b, const double *aoom, long nmo_a,
                          long nmo_b, long nao)
{
    long i, j, k, l;
    for (i = 0; i < nmo_a; ++i)
        for (j = 0; j < nmo_b; ++j) {
            double t = 0.0;
            for (k = 0; k < nao; ++k)
                for (l = 0; l < nao; ++l) t += mo_a[i * nao + k] * mo_b[j * nao + l] * aoom[k * nao + l];
            moom[i * nmo_b + j] = t;
        }
}

/* ---- non-Cartesian product grids (orbkit/cy_grid.pyx:58-97) ------------------------------
 * xyz is [3][n0*n1*n2], first axis slowest; same expressions, same multiplication order. */
void okor_sph2cart(double *xyz, const double *r, long nr, const double *theta, long nt,
                   const double *phi, long np)
{
    long i, j, k, c = 0, n = nr * nt * np;
    for (i = 0; i < nr; ++i)
        for (j = 0; j < nt; ++j)
            for (k = 0; k < np; ++k) {
                xyz[c] = r[i] * sin(theta[j]) * cos(phi[k]);
                xyz[n + c] = r[i] * sin(theta[j]) * sin(phi[k]);
                xyz[2 * n + c] = r[i] * cos(theta[j]);
                ++c;
            }
}

void okor_cyl2cart(double *xyz, const double *r, long nr, const double *phi, long np,
                   const double *zed, long nz)
{
    long i, j, k, c = 0, n = nr * np * nz;
    for (i = 0; i < nr; ++i)
        for (j = 0; j < np; ++j)
            for (k = 0; k < nz; ++k) {
                xyz[c] = r[i] * cos(phi[j]);
                xyz[n + c] = r[i] * sin(phi[j]);
                xyz[2 * n + c] = zed[k];
                ++c;
            }
}
